// Integer-pipe microbenchmarks used by bench.py / profiles: how many 128-bit field multiplies, NTT butterflies and
// SHA-256 compressions per second the SMs sustain when nothing touches HBM.  These are the "speed of light"
// denominators for the integer-bound kernels (the HBM roofline is reported separately).
#include "field.cuh"
#include "internal.h"
#include "sha256.cuh"
#include <string>

namespace mlb {

template <int MODE>  // 0: multiply chains, 1: butterfly (mul + add + sub) chains
__global__ void __launch_bounds__(256) mb_field_kernel(fe* out, int iters, fe seed) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a[4], w = seed;
    w.v[0] ^= tid;
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = seed; a[i].v[1] ^= tid * 4 + i; a[i] = fe_new(a[i]); }
    w = fe_new(w);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i += (MODE == 0 ? 1 : 2)) {
            if (MODE == 0) a[i] = fe_mul(a[i], w);
            else {
                fe v = fe_mul(a[i + 1], w), u = a[i];
                a[i] = fe_add(u, v);
                a[i + 1] = fe_sub(u, v);
            }
        }
    }
    fe r = fe_add(fe_add(a[0], a[1]), fe_add(a[2], a[3]));
    if (r.v[0] == 0x12345678u && r.v[3] == 0x9abcdef0u) fe_store(out + tid, r);  // keep the chain live
}
template <int MODE>  // 0: 32-byte leaf hash, 1: 64-byte node hash
__global__ void __launch_bounds__(128) mb_sha_kernel(uint32_t* out, int iters, uint32_t seed) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { h[i] = seed + tid * 8 + i; g[i] = seed ^ (tid + i); }
    for (int it = 0; it < iters; it++) {
        uint32_t o[8];
        if (MODE == 0) sha256_leaf32(h, o);
        else sha256_node64(h, g, o);
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = o[i];
    }
    if (h[0] == 0x12345678u && h[7] == 0x9abcdef0u) out[tid] = h[3];
}
__global__ void __launch_bounds__(256) mb_copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __ldg(in + i);
}

}  // namespace mlb

using namespace mlb;

// what: "modmul" | "butterfly" | "sha_leaf" | "sha_node" (n = threads, iters = chain length per thread; work = ops)
//       "copy" (n = bytes; work = bytes read + written).  ms_out = device time of one launch (best of 5).
extern "C" int ml_microbench(const char* what, size_t n, int iters, double* ms_out, double* work_out) {
    Ctx* ctx;
    MLB_TRY(get_ctx(&ctx));
    cudaStream_t s = ctx->stream;
    cudaEvent_t e0, e1;
    MLB_CUDA(cudaEventCreate(&e0));
    MLB_CUDA(cudaEventCreate(&e1));
    void* buf = nullptr;
    const std::string w(what);
    size_t bytes = w == "copy" ? 2 * n : (n + 1024) * 16;
    MLB_CUDA(cudaMalloc(&buf, bytes));
    MLB_CUDA(cudaMemsetAsync(buf, 1, bytes, s));
    fe seed = fe{{0x9e3779b9u, 0x7f4a7c15u, 0xbf58476du, 0x14057b7eu}};
    double best = 1e30;
    for (int rep = 0; rep < 6; rep++) {
        MLB_CUDA(cudaEventRecord(e0, s));
        if (w == "modmul") mb_field_kernel<0><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "butterfly") mb_field_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "sha_leaf") mb_sha_kernel<0><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
        else if (w == "sha_node") mb_sha_kernel<1><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
        else if (w == "copy") mb_copy_kernel<<<148 * 16, 256, 0, s>>>((const uint4*)buf, (uint4*)((uint8_t*)buf + n), n / 16);
        else { cudaFree(buf); set_error("ml_microbench: unknown benchmark '%s'", what); return ML_ERR_ARG; }
        MLB_KERNEL_CHECK();
        MLB_CUDA(cudaEventRecord(e1, s));
        MLB_CUDA(cudaEventSynchronize(e1));
        float ms;
        MLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(buf);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = best;
    if (w == "modmul") *work_out = (double)n * iters * 4;
    else if (w == "butterfly") *work_out = (double)n * iters * 2;
    else if (w == "copy") *work_out = 2.0 * (double)n;
    else *work_out = (double)n * iters;
    return ML_OK;
}
