// FRI folding (src/fri/mod.rs:79-114, src/fri/batched_fri.rs:101-150) as streaming kernels.
//
// Reference, per output i < n/2:  a = cur[i], b = cur[i + n/2],
//     next[i] = ((a + b) + r * ((a - b) * w_N0^(-i * 2^k))) * 2^-1
// with the inverse twiddle read from a 512 MiB gen_pows table as gen_pows[N0 - i*2^k] (:106-110).
// Here the twiddle comes from the two-level root tables already resident for the NTT (one extra multiply,
// no large-table gather), the challenge is pre-halved on the host (r/2), and the halving of (a + b) is a
// shift: 3 multiplies, 48 bytes of HBM traffic per output.
#include "field.cuh"
#include "internal.h"

namespace mlb {

__device__ __forceinline__ fe root_pow(const fe* __restrict__ lo, const fe* __restrict__ hi, size_t e) {
    fe w = fe_load_nc(lo + (e & (((size_t)1 << LO_BITS) - 1)));
    if (e >> LO_BITS) w = fe_mul(w, fe_load_nc(hi + (e >> LO_BITS)));
    return w;
}

// next[i] = half(a + b) + (r/2) * ((a - b) * w^-(i << k))
__global__ void __launch_bounds__(256) fri_fold_kernel(const fe* __restrict__ cur, size_t half_n, fe* __restrict__ next, fe r_half,
                                                       const fe* __restrict__ r_dev, int k, int log_n0, const fe* __restrict__ lo,
                                                       const fe* __restrict__ hi) {
    if (r_dev) r_half = fe_load(r_dev + 1);  // {r, r/2} left in HBM by the transcript kernel
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n0 = (size_t)1 << log_n0;
    for (; i < half_n; i += stride) {
        fe a = fe_load_nc(cur + i), b = fe_load_nc(cur + i + half_n);
        fe even = fe_half(fe_add(a, b));
        fe d = fe_sub(a, b);
        if (i != 0) d = fe_mul(d, root_pow(lo, hi, n0 - (i << k)));  // i = 0: gen_pows[0] = 1 (:92-95)
        fe_store(next + i, fe_add(even, fe_mul(r_half, d)));
    }
}

static fe to_dev_fe(hfe x) {
    fe r;
    r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); r.v[2] = (uint32_t)(x >> 64); r.v[3] = (uint32_t)(x >> 96);
    return r;
}
static inline unsigned grid_for(size_t n, size_t cap = 148 * 16) {
    size_t b = (n + 255) / 256;
    if (b > cap) b = cap;
    if (b == 0) b = 1;
    return (unsigned)b;
}

int fri_fold_launch(Ctx* ctx, const fe* cur, size_t n_cur, fe* next, hfe r, const fe* r_dev, size_t k, int log_n0, cudaStream_t s) {
    const RootTables* rt;
    MLB_TRY(get_root_tables(ctx, log_n0, s, &rt));
    const size_t half_n = n_cur / 2;
    ProfScope prof(PROF_FRI_FOLD, 24.0 * (double)n_cur, s);  // read n elements, write n/2
    fri_fold_kernel<<<grid_for(half_n), 256, 0, s>>>(cur, half_n, next, to_dev_fe(hfe_half(r)), r_dev, (int)k, log_n0, rt->lo, rt->hi);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

// batched first fold (batched_fri.rs:124-150): a, b are Horner fingerprints over the batch, acc = acc*rho + c_j
__global__ void __launch_bounds__(256) fri_batched_fold_kernel(const fe* const* __restrict__ codes, int n_codes, size_t half_n,
                                                               fe* __restrict__ next, fe rho, fe r_half, const fe* __restrict__ r_dev,
                                                               int log_n0, const fe* __restrict__ lo, const fe* __restrict__ hi) {
    if (r_dev) r_half = fe_load(r_dev + 1);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n0 = (size_t)1 << log_n0;
    for (; i < half_n; i += stride) {
        fe a = fe_zero(), b = fe_zero();
        for (int j = 0; j < n_codes; j++) {
            const fe* c = codes[j];
            a = fe_add(fe_mul(a, rho), fe_load_nc(c + i));
            b = fe_add(fe_mul(b, rho), fe_load_nc(c + i + half_n));
        }
        fe even = fe_half(fe_add(a, b));
        fe d = fe_sub(a, b);
        if (i != 0) d = fe_mul(d, root_pow(lo, hi, n0 - i));
        fe_store(next + i, fe_add(even, fe_mul(r_half, d)));
    }
}
int fri_batched_fold_launch(Ctx* ctx, const fe* const* codes, size_t n_codes, size_t n, fe* next, hfe fingerprint_r, hfe r,
                            const fe* r_dev, int log_n0, cudaStream_t s) {
    const RootTables* rt;
    MLB_TRY(get_root_tables(ctx, log_n0, s, &rt));
    const size_t half_n = n / 2;
    ProfScope prof(PROF_FRI_FOLD, 16.0 * (double)n * (double)n_codes + 8.0 * (double)n, s);
    fri_batched_fold_kernel<<<grid_for(half_n), 256, 0, s>>>(codes, (int)n_codes, half_n, next, to_dev_fe(fingerprint_r),
                                                              to_dev_fe(hfe_half(r)), r_dev, log_n0, rt->lo, rt->hi);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

// fingerprinted evaluation table (batched_pcs.rs:55-63): out[i] = Horner_rho(polys[0][i], ..., polys[B-1][i])
__global__ void __launch_bounds__(256) fingerprint_rows_kernel(const fe* const* __restrict__ polys, int n_polys, size_t n, fe rho,
                                                               fe* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        fe a = fe_zero();
        for (int j = 0; j < n_polys; j++) a = fe_add(fe_mul(a, rho), fe_load_nc(polys[j] + i));
        fe_store(out + i, a);
    }
}
int fingerprint_rows_launch(const fe* const* polys, size_t n_polys, size_t n, hfe r, fe* out, cudaStream_t s) {
    fingerprint_rows_kernel<<<grid_for(n), 256, 0, s>>>(polys, (int)n_polys, n, to_dev_fe(r), out);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

}  // namespace mlb
