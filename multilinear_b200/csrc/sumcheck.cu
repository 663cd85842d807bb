// Sumcheck tables for the PCS (src/constraint_system/sumcheck.rs:127-277), width 1, composition |x| x[0]
// (src/fri/multilinear_pcs.rs:56-57): per round the prover needs
//     e1 = sum_i m[i+off] * d[i+off]                                  (partial_sum at r = 1, :208-218)
//     e2 = sum_i (2 m[i+off] - m[i]) * (2 d[i+off] - d[i])            (partial_sum at r = 2, :219-231; s = 1 - r = -1)
// then folds both tables x[i] <- (1-r) x[i] + r x[i+off] (:234-247), computed as x[i] + r (x[i+off] - x[i]).
// Streaming kernels: each product is accumulated unreduced in a 288-bit register accumulator and reduced once
// per thread; the block results are added mod M (integer-exact, order-independent).
#include "field.cuh"
#include "internal.h"
#include "reduce.cuh"

namespace mlb {

static const int SC_THREADS = 256;
static const int SC_MAX_BLOCKS = 148 * 4;

// partials[2*blockIdx + {0,1}] = (e1, e2) of this CTA's slice
__global__ void __launch_bounds__(SC_THREADS) sumcheck_sums_kernel(const fe* __restrict__ m, const fe* __restrict__ d, size_t off,
                                                                   fe* __restrict__ partials) {
    __shared__ fe scratch[32];
    fe_acc a1, a2;
    acc_zero(a1);
    acc_zero(a2);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < off; i += stride) {
        fe m0 = fe_load_nc(m + i), m1 = fe_load_nc(m + i + off), d0 = fe_load_nc(d + i), d1 = fe_load_nc(d + i + off);
        acc_mul_add(a1, m1, d1);
        fe mm = fe_sub(fe_add(m1, m1), m0), dd = fe_sub(fe_add(d1, d1), d0);
        acc_mul_add(a2, mm, dd);
    }
    fe s1 = block_sum(acc_reduce(a1), scratch);
    fe s2 = block_sum(acc_reduce(a2), scratch);
    if (threadIdx.x == 0) {
        fe_store(partials + 2 * blockIdx.x, s1);
        fe_store(partials + 2 * blockIdx.x + 1, s2);
    }
}
// general partial_sum(r): r == 1 -> sum m1*d1 ; else sum (s m0 + r m1)(s d0 + r d1), s = 1 - r
__global__ void __launch_bounds__(SC_THREADS) sumcheck_partial_kernel(const fe* __restrict__ m, const fe* __restrict__ d, size_t off, fe r,
                                                                      fe sm1, int is_one, fe* __restrict__ partials) {
    __shared__ fe scratch[32];
    fe_acc a;
    acc_zero(a);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < off; i += stride) {
        fe m1 = fe_load_nc(m + i + off), d1 = fe_load_nc(d + i + off);
        if (is_one) acc_mul_add(a, m1, d1);
        else {
            fe m0 = fe_load_nc(m + i), d0 = fe_load_nc(d + i);
            fe mm = fe_add(fe_mul(sm1, m0), fe_mul(r, m1)), dd = fe_add(fe_mul(sm1, d0), fe_mul(r, d1));
            acc_mul_add(a, mm, dd);
        }
    }
    fe s = block_sum(acc_reduce(a), scratch);
    if (threadIdx.x == 0) fe_store(partials + blockIdx.x, s);
}
// out[c] = sum_b partials[b*ncols + c]
__global__ void __launch_bounds__(256) reduce_partials_kernel(const fe* __restrict__ partials, int count, int ncols, fe* __restrict__ out) {
    __shared__ fe scratch[32];
    for (int c = 0; c < ncols; c++) {
        fe a = fe_zero();
        for (int b = threadIdx.x; b < count; b += blockDim.x) a = fe_add(a, partials[(size_t)b * ncols + c]);
        a = block_sum(a, scratch);
        if (threadIdx.x == 0) fe_store(out + c, a);
    }
}
__global__ void __launch_bounds__(256) sumcheck_fold_kernel(fe* __restrict__ m, fe* __restrict__ d, size_t off, fe r, const fe* __restrict__ r_dev) {
    if (r_dev) r = fe_load(r_dev);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < off; i += stride) {
        fe m0 = fe_load(m + i), m1 = fe_load(m + i + off), d0 = fe_load(d + i), d1 = fe_load(d + i + off);
        fe_store(m + i, fe_add(m0, fe_mul(r, fe_sub(m1, m0))));
        fe_store(d + i, fe_add(d0, fe_mul(r, fe_sub(d1, d0))));
    }
}

// Fold round k and the partial sums of round k+1 in one pass (SURVEY.md §8d "fused fold + next-round sums"): a thread owns the
// quad (i, i+q, i+off, i+off+q), q = off/2, of both tables, folds it to the two elements (i, i+q) of the half-height
// tables (stored in place: no other thread touches this quad) and accumulates round k+1's products from the folded values
// while they are still in registers.  Per round the tables are read once and the half-height tables written once
// (48 B per input pair instead of 80), 6 multiplies per quad of which 2 stay unreduced in the 288-bit accumulators.
__global__ void __launch_bounds__(SC_THREADS) sumcheck_fold_sums_kernel(fe* __restrict__ m, fe* __restrict__ d, size_t off,
                                                                        const fe* __restrict__ r_dev, fe* __restrict__ partials) {
    __shared__ fe scratch[32];
    const fe r = fe_load(r_dev);
    const size_t q = off >> 1;
    fe_acc a1, a2;
    acc_zero(a1);
    acc_zero(a2);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // software pipeline: the eight loads of the next quad are issued before the six multiplies of the current one, so the
    // memory system always has a full set of requests per thread in flight (the kernel is HBM-bound: 192 B per 6 multiplies)
    fe c[8];
    if (i < q) {
        c[0] = fe_load(m + i); c[1] = fe_load(m + i + q); c[2] = fe_load(m + i + off); c[3] = fe_load(m + i + off + q);
        c[4] = fe_load(d + i); c[5] = fe_load(d + i + q); c[6] = fe_load(d + i + off); c[7] = fe_load(d + i + off + q);
    }
    for (; i < q; i += stride) {
        const size_t nx = i + stride;
        fe x[8];
        if (nx < q) {
            x[0] = fe_load(m + nx); x[1] = fe_load(m + nx + q); x[2] = fe_load(m + nx + off); x[3] = fe_load(m + nx + off + q);
            x[4] = fe_load(d + nx); x[5] = fe_load(d + nx + q); x[6] = fe_load(d + nx + off); x[7] = fe_load(d + nx + off + q);
        }
        fe m0 = fe_add(c[0], fe_mul(r, fe_sub(c[2], c[0]))), m1 = fe_add(c[1], fe_mul(r, fe_sub(c[3], c[1])));
        fe d0 = fe_add(c[4], fe_mul(r, fe_sub(c[6], c[4]))), d1 = fe_add(c[5], fe_mul(r, fe_sub(c[7], c[5])));
        fe_store(m + i, m0);
        fe_store(m + i + q, m1);
        fe_store(d + i, d0);
        fe_store(d + i + q, d1);
        acc_mul_add(a1, m1, d1);
        acc_mul_add(a2, fe_sub(fe_add(m1, m1), m0), fe_sub(fe_add(d1, d1), d0));
        if (nx < q) {
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = x[k];
        }
    }
    fe s1 = block_sum(acc_reduce(a1), scratch);
    fe s2 = block_sum(acc_reduce(a2), scratch);
    if (threadIdx.x == 0) {
        fe_store(partials + 2 * blockIdx.x, s1);
        fe_store(partials + 2 * blockIdx.x + 1, s2);
    }
}

// ---------------------------------------------------------------- width-w tables (System path, sumcheck.rs:21-38, 204-247)
// matrix is the trace, row-major [height][width]; the composition is a sparse polynomial over the row
//     comp(x) = sum_t coef[t] * prod_{k < len[t]} x[cols[off[t] + k]]
// (the C-ABI stand-in for the reference's closure argument, sumcheck.rs:176).  Terms live in shared memory; a thread
// interpolates one row pair, evaluates the composition and accumulates comp * delta unreduced.
__global__ void __launch_bounds__(SC_THREADS) wsumcheck_partial_kernel(const fe* __restrict__ m, const fe* __restrict__ d, size_t off, int width,
                                                                       fe r, fe sm1, int is_one, WTerms terms, fe* __restrict__ partials) {
    __shared__ fe scratch[32];
    __shared__ fe t_coef[W_MAX_TERMS];
    __shared__ uint32_t t_len[W_MAX_TERMS], t_off[W_MAX_TERMS], t_cols[W_MAX_COLS];
    for (int t = threadIdx.x; t < terms.n_terms; t += blockDim.x) { t_coef[t] = terms.coef[t]; t_len[t] = terms.len[t]; t_off[t] = terms.off[t]; }
    for (int c = threadIdx.x; c < terms.n_cols; c += blockDim.x) t_cols[c] = terms.cols[c];
    __syncthreads();
    fe_acc a;
    acc_zero(a);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < off; i += stride) {
        fe row[W_MAX_WIDTH];
        fe dd;
        if (is_one) {  // :208-218 (r * x with r = 1)
            dd = fe_load_nc(d + i + off);
            for (int j = 0; j < width; j++) row[j] = fe_load_nc(m + (i + off) * width + j);
        } else {       // :219-231
            dd = fe_add(fe_mul(sm1, fe_load_nc(d + i)), fe_mul(r, fe_load_nc(d + i + off)));
            for (int j = 0; j < width; j++)
                row[j] = fe_add(fe_mul(sm1, fe_load_nc(m + i * width + j)), fe_mul(r, fe_load_nc(m + (i + off) * width + j)));
        }
        fe comp = fe_zero();
        for (int t = 0; t < terms.n_terms; t++) {
            fe p = t_coef[t];
            for (uint32_t k = 0; k < t_len[t]; k++) p = fe_mul(p, row[t_cols[t_off[t] + k]]);
            comp = fe_add(comp, p);
        }
        acc_mul_add(a, comp, dd);
    }
    fe s = block_sum(acc_reduce(a), scratch);
    if (threadIdx.x == 0) fe_store(partials + blockIdx.x, s);
}
// All evaluation points r = 1 .. TD of one round in a single pass over the tables (the reference makes one pass per point,
// sumcheck.rs:185-187).  The interpolated row at integer r is x0 + r (x1 - x0), so consecutive points differ by the row
// difference: row_1 = x1, row_{r+1} = row_r + (x1 - x0) — additions only, no interpolation multiplies; same for delta.
// NARROW: width <= 4 — the row lives in registers and a column reference is a 4-way select instead of a local-memory lookup
template <int TD, bool NARROW = false>
__global__ void __launch_bounds__(SC_THREADS) wsumcheck_points_kernel(const fe* __restrict__ m, const fe* __restrict__ d, size_t off, int width,
                                                                      WTerms terms, fe* __restrict__ partials) {
    __shared__ fe scratch[32];
    __shared__ fe t_coef[W_MAX_TERMS];
    __shared__ uint32_t t_len[W_MAX_TERMS], t_off[W_MAX_TERMS], t_cols[W_MAX_COLS];
    for (int t = threadIdx.x; t < terms.n_terms; t += blockDim.x) { t_coef[t] = terms.coef[t]; t_len[t] = terms.len[t]; t_off[t] = terms.off[t]; }
    for (int c = threadIdx.x; c < terms.n_cols; c += blockDim.x) t_cols[c] = terms.cols[c];
    __syncthreads();
    fe_acc a[TD];
#pragma unroll
    for (int k = 0; k < TD; k++) acc_zero(a[k]);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int WM = NARROW ? 4 : W_MAX_WIDTH;
    for (; i < off; i += stride) {
        fe row[WM], diff[WM];
#pragma unroll
        for (int j = 0; j < WM; j++) {
            if (j < width) {
                fe x0 = fe_load_nc(m + i * width + j), x1 = fe_load_nc(m + (i + off) * width + j);
                row[j] = x1;
                diff[j] = fe_sub(x1, x0);
            }
        }
        fe d0 = fe_load_nc(d + i), dd = fe_load_nc(d + i + off);
        const fe ddiff = fe_sub(dd, d0);
#pragma unroll
        for (int k = 0; k < TD; k++) {  // point r = k + 1
            fe comp = fe_zero();
            for (int t = 0; t < terms.n_terms; t++) {
                fe p = t_coef[t];
                for (uint32_t c = 0; c < t_len[t]; c++) {
                    const uint32_t col = t_cols[t_off[t] + c];
                    if (NARROW) {
                        fe x = row[0];
#pragma unroll
                        for (int j = 1; j < 4; j++) x = col == (uint32_t)j ? row[j] : x;
                        p = fe_mul(p, x);
                    } else {
                        p = fe_mul(p, row[col]);
                    }
                }
                comp = fe_add(comp, p);
            }
            acc_mul_add(a[k], comp, dd);
            if (k + 1 < TD) {
#pragma unroll
                for (int j = 0; j < WM; j++)
                    if (j < width) row[j] = fe_add(row[j], diff[j]);
                dd = fe_add(dd, ddiff);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < TD; k++) {
        fe sk = block_sum(acc_reduce(a[k]), scratch);
        if (threadIdx.x == 0) fe_store(partials + (size_t)TD * blockIdx.x + k, sk);
    }
}
// fold (:234-247): rows i < off of the matrix and of delta, x <- x + r (x[i+off] - x)
__global__ void __launch_bounds__(256) wsumcheck_fold_kernel(fe* __restrict__ m, fe* __restrict__ d, size_t off, int width, fe r,
                                                             const fe* __restrict__ r_dev) {
    if (r_dev) r = fe_load(r_dev);
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t total = off * (size_t)(width + 1);
    for (; e < total; e += stride) {
        fe* base;
        size_t idx, hi;
        if (e < off * (size_t)width) { base = m; idx = e; hi = e + off * (size_t)width; }
        else { base = d; idx = e - off * (size_t)width; hi = idx + off; }
        fe x0 = fe_load(base + idx), x1 = fe_load(base + hi);
        fe_store(base + idx, fe_add(x0, fe_mul(r, fe_sub(x1, x0))));
    }
}

static inline fe fe_zero_host() { fe z; z.v[0] = z.v[1] = z.v[2] = z.v[3] = 0u; return z; }
static inline unsigned blocks_for(size_t n) {
    size_t b = (n + SC_THREADS - 1) / SC_THREADS;
    if (b > (size_t)SC_MAX_BLOCKS) b = SC_MAX_BLOCKS;
    if (b == 0) b = 1;
    return (unsigned)b;
}
static int fetch(const fe* dev, int count, hfe* out, cudaStream_t s) {
    uint8_t tmp[16 * 4];
    MLB_CUDA(cudaMemcpyAsync(tmp, dev, 16 * (size_t)count, cudaMemcpyDeviceToHost, s));
    MLB_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < count; i++) out[i] = hfe_load(tmp + 16 * i);
    return ML_OK;
}

int sumcheck_max_blocks() { return SC_MAX_BLOCKS; }
int sumcheck_sums_partials_launch(const fe* m, const fe* d, size_t height, fe* partials, int* n_blocks, cudaStream_t s) {
    const size_t off = height >> 1;
    const unsigned nb = blocks_for(off);
    ProfScope prof(PROF_SUMCHECK_SUMS, 32.0 * (double)height, s);
    sumcheck_sums_kernel<<<nb, SC_THREADS, 0, s>>>(m, d, off, partials);
    MLB_KERNEL_CHECK();
    *n_blocks = (int)nb;
    return ML_OK;
}
int sumcheck_sums_launch(Ctx* ctx, const fe* m, const fe* d, size_t height, hfe* s1, hfe* s2, cudaStream_t s) {
    (void)ctx;
    const size_t off = height >> 1;
    const unsigned nb = blocks_for(off);
    fe* partials;
    MLB_TRY(dev_alloc_async((void**)&partials, (size_t)(2 * nb + 2) * 16, s));
    ProfScope prof(PROF_SUMCHECK_SUMS, 32.0 * (double)height, s);  // read both tables once
    sumcheck_sums_kernel<<<nb, SC_THREADS, 0, s>>>(m, d, off, partials);
    MLB_KERNEL_CHECK();
    reduce_partials_kernel<<<1, 256, 0, s>>>(partials, (int)nb, 2, partials + 2 * nb);
    MLB_KERNEL_CHECK();
    hfe o[2];
    MLB_TRY(fetch(partials + 2 * nb, 2, o, s));
    MLB_TRY(dev_free_async(partials, s));
    *s1 = o[0];
    *s2 = o[1];
    return ML_OK;
}
int sumcheck_partial_sum_launch(Ctx* ctx, const fe* m, const fe* d, size_t height, hfe r, hfe* out, cudaStream_t s) {
    (void)ctx;
    const size_t off = height >> 1;
    const unsigned nb = blocks_for(off);
    fe* partials;
    MLB_TRY(dev_alloc_async((void**)&partials, (size_t)(nb + 1) * 16, s));
    sumcheck_partial_kernel<<<nb, SC_THREADS, 0, s>>>(m, d, off, to_dev_fe_h(r), to_dev_fe_h(hfe_sub(1, r)), r == 1 ? 1 : 0, partials);
    MLB_KERNEL_CHECK();
    reduce_partials_kernel<<<1, 256, 0, s>>>(partials, (int)nb, 1, partials + nb);
    MLB_KERNEL_CHECK();
    MLB_TRY(fetch(partials + nb, 1, out, s));
    MLB_TRY(dev_free_async(partials, s));
    return ML_OK;
}
// fold tables of `height` with the challenge at r_dev and leave round k+1's (e1, e2) partials; needs height >= 4
int sumcheck_fold_sums_launch(fe* m, fe* d, size_t height, const fe* r_dev, fe* partials, int* n_blocks, cudaStream_t s) {
    const size_t off = height >> 1;
    if (off < 2) { set_error("sumcheck_fold_sums: height must be at least 4"); return ML_ERR_SIZE; }
    // 128 registers per thread: two CTAs are resident per SM, so the grid is exactly one persistent wave (148 x 2)
    unsigned nb = blocks_for(off >> 1);
    if (nb > 148 * 2) nb = 148 * 2;
    ProfScope prof(PROF_SUMCHECK_FOLD, 48.0 * (double)height, s);  // read 2 tables (h), write 2 half tables; the sums ride along
    sumcheck_fold_sums_kernel<<<nb, SC_THREADS, 0, s>>>(m, d, off, r_dev, partials);
    MLB_KERNEL_CHECK();
    *n_blocks = (int)nb;
    return ML_OK;
}
int sumcheck_fold_launch(fe* m, fe* d, size_t height, hfe r, const fe* r_dev, cudaStream_t s) {
    const size_t off = height >> 1;
    if (off == 0) return ML_OK;
    ProfScope prof(PROF_SUMCHECK_FOLD, 48.0 * (double)height, s);  // read 2 tables (h), write 2 half tables
    sumcheck_fold_kernel<<<blocks_for(off), 256, 0, s>>>(m, d, off, to_dev_fe_h(r), r_dev);
    MLB_KERNEL_CHECK();
    return ML_OK;
}

}  // namespace mlb

namespace mlb {
int wsumcheck_limits(int what) { return what == 0 ? W_MAX_WIDTH : what == 1 ? W_MAX_TERMS : W_MAX_COLS; }
int wsumcheck_partial_sum_launch(const fe* m, const fe* d, size_t height, size_t width, hfe r, const fe* coef, const uint32_t* len,
                                 const uint32_t* off, const uint32_t* cols, size_t n_terms, size_t n_cols, hfe* out, cudaStream_t s) {
    const size_t half = height >> 1;
    const unsigned nb = blocks_for(half);
    fe* partials;
    MLB_TRY(dev_alloc_async((void**)&partials, (size_t)(nb + 1) * 16, s));
    WTerms t{coef, len, off, cols, (int)n_terms, (int)n_cols};
    wsumcheck_partial_kernel<<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, to_dev_fe_h(r), to_dev_fe_h(hfe_sub(1, r)), r == 1 ? 1 : 0, t, partials);
    MLB_KERNEL_CHECK();
    reduce_partials_kernel<<<1, 256, 0, s>>>(partials, (int)nb, 1, partials + nb);
    MLB_KERNEL_CHECK();
    MLB_TRY(fetch(partials + nb, 1, out, s));
    MLB_TRY(dev_free_async(partials, s));
    return ML_OK;
}
// evals_out[k] = partial_sum(r = k + 1), k < td, in one pass; returns ML_ERR_ARG when td is outside the fused range
int wsumcheck_points_launch(const fe* m, const fe* d, size_t height, size_t width, const fe* coef, const uint32_t* len, const uint32_t* off,
                            const uint32_t* cols, size_t n_terms, size_t n_cols, int td, hfe* evals_out, cudaStream_t s) {
    if (td < 1 || td > 4) return ML_ERR_ARG;
    const size_t half = height >> 1;
    const unsigned nb = blocks_for(half);
    fe* partials;
    MLB_TRY(dev_alloc_async((void**)&partials, ((size_t)nb + 1) * td * 16, s));
    WTerms t{coef, len, off, cols, (int)n_terms, (int)n_cols};
    switch (td) {
        case 1: wsumcheck_points_kernel<1><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
        case 2: wsumcheck_points_kernel<2><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
        case 3: wsumcheck_points_kernel<3><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
        default: wsumcheck_points_kernel<4><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
    }
    MLB_KERNEL_CHECK();
    reduce_partials_kernel<<<1, 256, 0, s>>>(partials, (int)nb, td, partials + (size_t)nb * td);
    MLB_KERNEL_CHECK();
    MLB_TRY(fetch(partials + (size_t)nb * td, td, evals_out, s));
    MLB_TRY(dev_free_async(partials, s));
    return ML_OK;
}
int wsumcheck_fold_launch(fe* m, fe* d, size_t height, size_t width, hfe r, cudaStream_t s) {
    const size_t half = height >> 1;
    if (half == 0) return ML_OK;
    size_t total = half * (width + 1), blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wsumcheck_fold_kernel<<<(unsigned)blocks, 256, 0, s>>>(m, d, half, (int)width, to_dev_fe_h(r), nullptr);
    MLB_KERNEL_CHECK();
    return ML_OK;
}
int wsumcheck_fold_dev_launch(fe* m, fe* d, size_t height, size_t width, const fe* r_dev, cudaStream_t s) {
    const size_t half = height >> 1;
    if (half == 0) return ML_OK;
    size_t total = half * (width + 1), blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wsumcheck_fold_kernel<<<(unsigned)blocks, 256, 0, s>>>(m, d, half, (int)width, fe_zero_host(), r_dev);
    MLB_KERNEL_CHECK();
    return ML_OK;
}
int wsumcheck_points_partials_launch(const fe* m, const fe* d, size_t height, size_t width, const WTerms& t, int td, fe* partials, int* n_blocks,
                                     cudaStream_t s) {
    if (td < 1 || td > W_MAX_TD) return ML_ERR_ARG;
    const size_t half = height >> 1;
    const unsigned nb = blocks_for(half);
    if (width <= 4) {
        switch (td) {
            case 1: wsumcheck_points_kernel<1, true><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            case 2: wsumcheck_points_kernel<2, true><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            case 3: wsumcheck_points_kernel<3, true><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            default: wsumcheck_points_kernel<4, true><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
        }
    } else {
        switch (td) {
            case 1: wsumcheck_points_kernel<1><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            case 2: wsumcheck_points_kernel<2><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            case 3: wsumcheck_points_kernel<3><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
            default: wsumcheck_points_kernel<4><<<nb, SC_THREADS, 0, s>>>(m, d, half, (int)width, t, partials); break;
        }
    }
    MLB_KERNEL_CHECK();
    *n_blocks = (int)nb;
    return ML_OK;
}
}  // namespace mlb
