// Integer-pipe microbenchmarks used by bench.py / profiles: how many 128-bit field multiplies, NTT butterflies and
// SHA-256 compressions per second the SMs sustain when nothing touches HBM.  These are the "speed of light"
// denominators for the integer-bound kernels (the HBM roofline is reported separately).
#include "field.cuh"
#include "internal.h"
#include "sha256.cuh"
#include <cstdlib>
#include <string>

namespace mlb {

template <int MODE, int V = MLB_REDUCE_V>  // 0: multiply chains, 1: butterfly (mul + add + sub) chains; V = reduction variant
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
            if (MODE == 0) a[i] = fe_mul_t<V>(a[i], w);
            else {
                fe v = fe_mul_t<V>(a[i + 1], w), u = a[i];
                a[i] = fe_add(u, v);
                a[i + 1] = fe_sub(u, v);
            }
        }
    }
    fe r = fe_add(fe_add(a[0], a[1]), fe_add(a[2], a[3]));
    if (r.v[0] == 0x12345678u && r.v[3] == 0x9abcdef0u) fe_store(out + tid, r);  // keep the chain live
}
template <int MODE, int FMA_ADD, int ROT = 0, int MINB = 1>  // 0: 32-byte leaf hash, 1: 64-byte node hash; FMA_ADD = add-routing mask; ROT = rotation routing;
// MINB = min CTAs/SM (forcing 32 registers / full occupancy was measured slower: profiles/r1_sha_add_routing.txt)
__global__ void __launch_bounds__(128, MINB) mb_sha_kernel(uint32_t* out, int iters, uint32_t seed) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { h[i] = seed + tid * 8 + i; g[i] = seed ^ (tid + i); }
    for (int it = 0; it < iters; it++) {
        uint32_t o[8];
        sha_iv(o);
        if (MODE == 0) {
            uint32_t w[16] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 0x80000000u, 0u, 0u, 0u, 0u, 0u, 0u, 256u};
            sha_compress_t<FMA_ADD, ROT>(o, w);
        } else {
            uint32_t w[16] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7]};
            sha_compress_t<FMA_ADD, ROT>(o, w);
            sha_compress_pad512_t<FMA_ADD, ROT>(o);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = o[i];
    }
    if (h[0] == 0x12345678u && h[7] == 0x9abcdef0u) out[tid] = h[3];
}
// raw pipe throughput: MODE 0 IADD3, 1 IMAD (32-bit), 2 IMAD.WIDE, 3 SHF (funnel shift), 4 LOP3, 5 IADD3+IMAD 1:1,
// 6 SHF+IMAD 1:1, 7 SHF+IMAD.WIDE 1:1, 8 SHF+LOP3 1:1, 9 SHF+LOP3+IMAD 1:1:1, 10 IMAD.HI, 11 IMAD.HI+SHF,
// 12 DFMA (fp64 pipe), 13 DFMA+IMAD.WIDE 1:1, 14 DFMA+SHF 1:1
template <int MODE>
__global__ void __launch_bounds__(256) mb_pipe_kernel(uint32_t* out, int iters, uint32_t seed) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8];
    unsigned long long wd[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + tid * 8 + i; b[i] = seed ^ (tid + 77 * i); wd[i] = a[i]; }
    uint32_t m = seed | 1u;
    double fd[8];
#pragma unroll
    for (int i = 0; i < 8; i++) fd[i] = 1.0 + 1e-9 * (double)(tid + i);
    const double fm = 1.0000001, fa = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (MODE == 1 || MODE == 5 || MODE == 6 || MODE == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(a[i]));
            if (MODE == 12 || MODE == 13 || MODE == 14) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(fd[i]) : "d"(fm), "d"(fa));
            if (MODE == 14) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
            if (MODE == 2 || MODE == 7 || MODE == 13) asm volatile("{\n\t.reg .u32 lo;\n\tcvt.u32.u64 lo, %0;\n\tmad.wide.u32 %0, lo, %1, %0;\n\t}" : "+l"(wd[i]) : "r"(m));
            if (MODE == 10 || MODE == 11) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m));
            if (MODE == 11) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
            if (MODE == 3 || MODE == 6 || MODE == 7 || MODE == 8 || MODE == 9) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
            if (MODE == 4 || MODE == 8 || MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(m));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i] + b[i] + (uint32_t)wd[i] + (uint32_t)(wd[i] >> 32) + (uint32_t)__double2uint_rz(fd[i]);
    if (r == 0x12345678u) out[tid] = r;
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
        else if (w == "modmul_v1") mb_field_kernel<0, 1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "modmul_v2") mb_field_kernel<0, 2><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "butterfly_v1") mb_field_kernel<1, 1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "butterfly_v2") mb_field_kernel<1, 2><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((fe*)buf, iters, seed);
        else if (w == "sha_leaf") mb_sha_kernel<0, MLB_SHA_ADD_MASK><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
        else if (w == "sha_node") mb_sha_kernel<1, MLB_SHA_ADD_MASK><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
#define MB_SHA_VARIANT(M)                                                                                                              \
        else if (w == "sha_leaf_m" #M) mb_sha_kernel<0, M><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); \
        else if (w == "sha_node_m" #M) mb_sha_kernel<1, M><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
#define MB_SHA_ROT(M, R)                                                                                                               \
        else if (w == "sha_leaf_m" #M "_r" #R) mb_sha_kernel<0, M, R><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); \
        else if (w == "sha_node_m" #M "_r" #R) mb_sha_kernel<1, M, R><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((uint32_t*)buf, iters, 0x1234567u);
        MB_SHA_ROT(2, 1) MB_SHA_ROT(2, 11) MB_SHA_ROT(2, 111) MB_SHA_ROT(2, 101) MB_SHA_ROT(2, 21) MB_SHA_ROT(2, 121) MB_SHA_ROT(2, 211)
        MB_SHA_ROT(0, 1) MB_SHA_ROT(0, 11) MB_SHA_ROT(0, 111) MB_SHA_ROT(0, 121) MB_SHA_ROT(2, 221) MB_SHA_ROT(0, 211)
        MB_SHA_VARIANT(0) MB_SHA_VARIANT(63) MB_SHA_VARIANT(3) MB_SHA_VARIANT(1) MB_SHA_VARIANT(2) MB_SHA_VARIANT(19) MB_SHA_VARIANT(11)
        MB_SHA_VARIANT(27) MB_SHA_VARIANT(59) MB_SHA_VARIANT(43)
        else if (w.rfind("pipe", 0) == 0) {
            const int mode = atoi(what + 4);
            const unsigned gb = (unsigned)((n + 255) / 256);
            switch (mode) {
                case 0: mb_pipe_kernel<0><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 1: mb_pipe_kernel<1><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 2: mb_pipe_kernel<2><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 3: mb_pipe_kernel<3><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 4: mb_pipe_kernel<4><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 5: mb_pipe_kernel<5><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 6: mb_pipe_kernel<6><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 7: mb_pipe_kernel<7><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 8: mb_pipe_kernel<8><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 10: mb_pipe_kernel<10><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 11: mb_pipe_kernel<11><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 12: mb_pipe_kernel<12><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 13: mb_pipe_kernel<13><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                case 14: mb_pipe_kernel<14><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
                default: mb_pipe_kernel<9><<<gb, 256, 0, s>>>((uint32_t*)buf, iters, 0x1234567u); break;
            }
        }
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
    if (w.rfind("modmul", 0) == 0) *work_out = (double)n * iters * 4;
    else if (w.rfind("butterfly", 0) == 0) *work_out = (double)n * iters * 2;
    else if (w == "copy") *work_out = 2.0 * (double)n;
    else if (w.rfind("pipe", 0) == 0) {
        const int mode = atoi(what + 4);
        const int per = (mode <= 4 || mode == 10 || mode == 12) ? 1 : (mode == 9 ? 3 : 2);
        *work_out = (double)n * iters * 8 * per;  // thread-instructions
    }
    else *work_out = (double)n * iters;
    return ML_OK;
}
