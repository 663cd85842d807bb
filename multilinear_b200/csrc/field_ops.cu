// Element-wise Field128 kernels (src/field.rs:66-154) and the synthetic input generator used by bench.py.
#include "field.cuh"
#include "internal.h"

namespace mlb {

static inline unsigned grid_for(size_t n, int threads = 256, size_t cap = 148 * 32) {
    size_t b = (n + threads - 1) / threads;
    if (b > cap) b = cap;
    if (b == 0) b = 1;
    return (unsigned)b;
}

// op: 0 add, 1 sub, 2 mul, 3 inv (b unused; inv(0) = 0 as winter-math), 4 scale by b[0]
__global__ void fe_vec_kernel(int op, const fe* a, const fe* b, size_t n, fe* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        fe x = fe_load_nc(a + i), r;
        if (op == 3) {
            // x^(M-2), M-2 = 2^128 - 45*2^40 - 1 : limbs {0xFFFFFFFF, 0xFFFFD2FF, 0xFFFFFFFF, 0xFFFFFFFF}
            const uint32_t e[4] = {0xFFFFFFFFu, 0xFFFFD2FFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
            r = fe_one();
            for (int w = 3; w >= 0; w--)
                for (int bit = 31; bit >= 0; bit--) {
                    r = fe_sqr(r);
                    if ((e[w] >> bit) & 1u) r = fe_mul(r, x);
                }
        } else {
            fe y = fe_load_nc(b + (op == 4 ? 0 : i));
            r = op == 0 ? fe_add(x, y) : (op == 1 ? fe_sub(x, y) : fe_mul(x, y));
        }
        fe_store(out + i, r);
    }
}
int fe_vec_launch(int op, const fe* a, const fe* b, size_t n, fe* out, cudaStream_t s) {
    if (n == 0) return ML_OK;
    fe_vec_kernel<<<grid_for(n), 256, 0, s>>>(op, a, b, n, out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

__global__ void fe_pow_kernel(const fe* a, uint4 e, size_t n, fe* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
    for (; i < n; i += stride) {
        fe x = fe_load_nc(a + i), r = fe_one();
        for (int w = 3; w >= 0; w--)
            for (int bit = 31; bit >= 0; bit--) {
                r = fe_sqr(r);
                if ((ew[w] >> bit) & 1u) r = fe_mul(r, x);
            }
        fe_store(out + i, r);
    }
}
int fe_pow_vec_launch(const fe* a, hfe e, size_t n, fe* out, cudaStream_t s) {
    if (n == 0) return ML_OK;
    uint4 ev = make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)(e >> 64), (uint32_t)(e >> 96));
    fe_pow_kernel<<<grid_for(n), 256, 0, s>>>(a, ev, n, out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

// From<i64> (src/field.rs:150-154): sign-extend to u128, then BaseElement::new (one conditional subtract)
__global__ void fe_from_i64_kernel(const long long* v, size_t n, fe* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        long long x = v[i];
        uint32_t ext = x < 0 ? 0xFFFFFFFFu : 0u;
        fe_store(out + i, fe_new(fe{{(uint32_t)x, (uint32_t)((unsigned long long)x >> 32), ext, ext}}));
    }
}
int fe_from_i64_launch(const int64_t* v, size_t n, fe* out, cudaStream_t s) {
    if (n == 0) return ML_OK;
    fe_from_i64_kernel<<<grid_for(n), 256, 0, s>>>((const long long*)v, n, out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

// 256-bit little-endian integers -> canonical elements (v mod M), with both reduction variants side by side: the entry point
// behind ml_fe_from_wide_vec.  Used as hash-to-field for 32-byte digests and, in the tests, to drive fe_reduce_wide through
// adversarial 256-bit patterns that products of canonical operands reach only with negligible probability.
__global__ void fe_from_wide_kernel(const uint4* v, size_t n, int variant, fe* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint4 lo = v[2 * i], hi = v[2 * i + 1];
        const uint32_t p[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        fe_store(out + i, variant == 1 ? fe_reduce_wide_v1(p) : fe_reduce_wide_v2(p));
    }
}
int fe_from_wide_launch(const void* v, size_t n, int variant, fe* out, cudaStream_t s) {
    if (n == 0) return ML_OK;
    fe_from_wide_kernel<<<grid_for(n), 256, 0, s>>>((const uint4*)v, n, variant, out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// element i = new(lo | hi << 64), lo = splitmix64(seed + 2i), hi = splitmix64(seed + 2i + 1)
__global__ void synthetic_kernel(unsigned long long seed, size_t n, fe* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long lo = splitmix64(seed + 2ull * i), hi = splitmix64(seed + 2ull * i + 1ull);
        fe_store(out + i, fe_new(fe{{(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)}}));
    }
}
int synthetic_launch(uint64_t seed, size_t n, fe* out, cudaStream_t s) {
    if (n == 0) return ML_OK;
    synthetic_kernel<<<grid_for(n), 256, 0, s>>>(seed, n, out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

}  // namespace mlb
