// SHA-256 Merkle trees (src/merkle_tree/mod.rs:65-131) on sm_100a.
//
// Reference: leaf_i = SHA-256(bytes(data[i])) (:71, :178-182), node = SHA-256(left || right) (:184-189), every
// layer retained (:8-11).  FRI leaves are ReedSolomonPairs (code[i], code[i + n/2]) (src/fri/mod.rs:45-55).
//
// Hashing is integer-pipe bound (about 1.4k alu-pipe instructions per compression), so the kernels are built to
// keep every lane busy and every digest in registers:
//   - one thread owns 2^G consecutive leaves and walks its own subtree with a G-deep digest stack: all
//     2^G leaf hashes and 2^G - 1 node hashes run back to back with no shared memory, no barrier and no idle
//     lanes (a shared-memory reduction tree idles half the lanes per level);
//   - every layer is still written to HBM (the prover opens paths from them), 32 B per digest, once;
//   - upper layers repeat the same scheme on digests; the last <= 2048 nodes finish in one CTA.
// All layers live in one buffer: layer l starts at digest index 2L - (2L >> l) (merkle_layer_offset).
#include <cstdlib>
#include <cstring>
#include "field.cuh"
#include "internal.h"
#include "sha256.cuh"

namespace mlb {

// 16-byte read-only load that asks L2 to bring in the whole 128-byte line: a thread consumes its 8 leaves' 128 bytes
// over ~40 us, and without the hint every 32-byte sector request became its own 64-byte DRAM burst (ncu: DRAM reads
// 1.9x the algorithmic bytes, profiles/r1_ncu_top_kernels_v3.txt)
__device__ __forceinline__ uint4 ldg_line(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint8_t* layer_ptr(uint8_t* digests, size_t n_leaves, int layer, size_t idx) {
    return digests + 32 * ((2 * n_leaves - ((2 * n_leaves) >> layer)) + idx);
}

// Thread-private subtree walk.  A thread owns 2^G consecutive nodes of `base_layer` (leaves when LEAVES) and computes
// the G layers above them with a G-deep digest stack.  The walk is written as ONE loop whose body holds a single
// inlined copy of the compression function (37 KB of straight-line SASS): unrolling the 2^(G+1)-1 hashes instead
// makes a ~450 KB body that thrashes the instruction cache (ncu: stall_no_instruction dominant, profiles/r1_*).
// Even one extra copy costs more than it saves: a second, leaf-specialised compression (padding words as constants, 8 %
// fewer instructions per leaf hash) grew the body from 41 KB to 65 KB and made the 2^24-leaf launch 14 % SLOWER (5.91 -> 6.75 ms
// per commit for the group, measured in the second session).
// Control flow depends only on (j, lvl), identical for every thread, so there is no divergence; the stack is
// indexed dynamically and lives in local memory (32 bytes of traffic per 2400-instruction hash).
static const int MERKLE_THREADS = 128;
template <int LOG_G, bool LEAVES>
__device__ __forceinline__ void subtree_walk(const fe* __restrict__ code, uint8_t* __restrict__ digests, size_t n_leaves, int base_layer,
                                             size_t first) {
    // digest stack in shared memory, [level][word][thread]: conflict-free, dynamically indexed, and no local-memory
    // traffic (the per-thread local stack spilled through L1/L2 into DRAM: ncu showed 2x the algorithmic reads)
    __shared__ uint32_t stack_sm[LOG_G > 0 ? LOG_G : 1][8][MERKLE_THREADS];
    const int tx = threadIdx.x;
    uint32_t h[8];
    uint4 nx = make_uint4(0, 0, 0, 0), ny = nx;  // second half of the 32-byte sectors fetched for an even leaf
    int j = 0, lvl = -1;  // lvl < 0: next hash is leaf j; otherwise node(stack[lvl], h)
    size_t idx = first;
#pragma unroll 1
    while (j < (1 << LOG_G)) {
        uint32_t w[16], st[8];
        const bool leaf = lvl < 0;
        if (leaf) {
            idx = first + j;
            if (LEAVES) {
                uint4 x, y;
                if (LOG_G > 0 && (j & 1)) {
                    x = nx; y = ny;
                } else {
                    x = ldg_line(code + idx);
                    y = ldg_line(code + idx + n_leaves);
                    if (LOG_G > 0) {
                        nx = ldg_line(code + idx + 1);
                        ny = ldg_line(code + idx + 1 + n_leaves);
                    }
                }
                sha_words_from_le(x, w);
                sha_words_from_le(y, w + 4);
                w[8] = 0x80000000u; w[9] = 0; w[10] = 0; w[11] = 0; w[12] = 0; w[13] = 0; w[14] = 0; w[15] = 256u;
            } else {
                sha_load_digest(layer_ptr(digests, n_leaves, base_layer, idx), h);
            }
        } else {
            // siblings are written together, 64 contiguous bytes: single 32-byte digest stores turned into
            // read-modify-write DRAM traffic (ncu: DRAM reads 1.9x the algorithmic bytes before this change)
#pragma unroll
            for (int k = 0; k < 8; k++) { w[k] = stack_sm[lvl][k][tx]; w[8 + k] = h[k]; }
            if (LEAVES || lvl > 0) {
                uint8_t* dst = layer_ptr(digests, n_leaves, base_layer + lvl, idx - 1);
                sha_store_digest(dst, w);
                sha_store_digest(dst + 32, h);
            }
        }
        if (LEAVES || !leaf) {
            sha_iv(st);
            sha_compress(st, w);
            if (!leaf) sha_compress_pad512(st);
#pragma unroll
            for (int k = 0; k < 8; k++) h[k] = st[k];
            if (!leaf) idx >>= 1;
        }
        lvl = leaf ? 0 : lvl + 1;
        if (lvl < LOG_G && ((j >> lvl) & 1)) continue;  // left sibling is waiting on the stack: hash the parent next
        if (lvl < LOG_G) {
#pragma unroll
            for (int k = 0; k < 8; k++) stack_sm[lvl][k][tx] = h[k];
        } else {
            // top of this thread's subtree (adjacent threads write adjacent digests)
            if (LEAVES || LOG_G > 0) sha_store_digest(layer_ptr(digests, n_leaves, base_layer + LOG_G, idx), h);
        }
        j++;
        lvl = -1;
    }
}

// Leaves = RS pairs of a code of n_code = 2 * n_leaves elements.  Thread g hashes leaves [g*2^G, (g+1)*2^G).
template <int LOG_G>
__global__ void __launch_bounds__(128) merkle_rs_kernel(const fe* __restrict__ code, size_t n_leaves, uint8_t* __restrict__ digests) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((g << LOG_G) >= n_leaves) return;
    subtree_walk<LOG_G, true>(code, digests, n_leaves, 0, g << LOG_G);
}

// Thread g reads 2^G consecutive digests of `from_layer` and writes the G layers above them.
template <int LOG_G>
__global__ void __launch_bounds__(128) merkle_nodes_kernel(uint8_t* __restrict__ digests, size_t n_leaves, int from_layer) {
    const size_t count = n_leaves >> from_layer;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((g << LOG_G) >= count) return;
    subtree_walk<LOG_G, false>(nullptr, digests, n_leaves, from_layer, g << LOG_G);
}

// The last <= 4096 nodes of a tree: a chain of up to 12 dependent levels, too small for a grid and too much hashing for one
// SM (2048 + 1024 + ... node hashes at ~2400 instructions each kept a single CTA busy for ~65 us, 14 trees per commit).
// One thread-block cluster of 8 CTAs (8 SMs, 4096 threads) walks the levels with a cluster barrier between them: level data
// goes through global memory (L2), made visible across the cluster's CTAs by __threadfence() before barrier.cluster.
static const int TOP_CLUSTER = 8, TOP_THREADS = 512;
#ifndef MLB_TOP_MAX_COUNT
#define MLB_TOP_MAX_COUNT 4096
#endif
static const size_t TOP_MAX_COUNT = MLB_TOP_MAX_COUNT;  // nodes handed to the cluster kernel (32768 measured: nodes -0.27 ms, tops +0.45 ms per commit)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__global__ void __launch_bounds__(TOP_THREADS) merkle_top_kernel(uint8_t* digests, size_t n_leaves, int from_layer) {
    size_t count = n_leaves >> from_layer;
    int layer = from_layer;
    const size_t g = (size_t)cluster_cta_rank() * blockDim.x + threadIdx.x, stride = (size_t)TOP_CLUSTER * blockDim.x;
    while (count > 1) {
        const size_t next = count >> 1;
        for (size_t i = g; i < next; i += stride) {
            uint32_t l[8], r[8], o[8];
            sha_load_digest(layer_ptr(digests, n_leaves, layer, 2 * i), l);
            sha_load_digest(layer_ptr(digests, n_leaves, layer, 2 * i + 1), r);
            sha256_node64(l, r, o);
            sha_store_digest(layer_ptr(digests, n_leaves, layer + 1, i), o);
        }
        __threadfence();
        cluster_sync_all();
        count = next;
        layer++;
    }
}
static int merkle_top_launch(uint8_t* digests, size_t n_leaves, int layer, cudaStream_t s) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(TOP_CLUSTER, 1, 1);
    cfg.blockDim = dim3(TOP_THREADS, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = TOP_CLUSTER;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    MLB_CUDA(cudaLaunchKernelEx(&cfg, merkle_top_kernel, digests, n_leaves, layer));
    return ML_OK;
}

// Batched leaves (Merkle::batch_commit, :110-116): leaf_i = SHA-256(pair_0(i) || pair_1(i) || ... || pair_{B-1}(i)).
// Two 32-byte pairs fill one block, so B pairs take ceil((B+1)/2) compressions (33 for B = 64).
// PAIRS = false: src[j] is code j (n_code elements), pair = (code[i], code[i + n_leaves]);
// PAIRS = true : src[j] is an array of n_leaves ReedSolomonPairs (32 bytes each).
// STRIDED (with PAIRS): the pair arrays are equally spaced, src[j] = (const fe*)strided_base + j * stride_elems — the layout of a
// sharded prover's receive buffer ([polynomial][row][32 B]), which needs no pointer table.
template <bool PAIRS, bool STRIDED = false>
__global__ void __launch_bounds__(128) merkle_batched_kernel(const fe* const* __restrict__ src, int n_codes, size_t n_leaves,
                                                             uint8_t* __restrict__ digests, const fe* __restrict__ strided_base = nullptr,
                                                             size_t stride_elems = 0) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    uint32_t st[8];
    sha_iv(st);
    uint32_t w[16];
    const unsigned long long bits = (unsigned long long)n_codes * 256ull;
    for (int j = 0; j < n_codes; j += 2) {
        const fe* c0 = STRIDED ? strided_base + (size_t)j * stride_elems : src[j];
        uint4 x0 = PAIRS ? __ldg(reinterpret_cast<const uint4*>(c0 + 2 * i)) : __ldg(reinterpret_cast<const uint4*>(c0 + i));
        uint4 y0 = PAIRS ? __ldg(reinterpret_cast<const uint4*>(c0 + 2 * i + 1)) : __ldg(reinterpret_cast<const uint4*>(c0 + i + n_leaves));
        sha_words_from_le(x0, w);
        sha_words_from_le(y0, w + 4);
        if (j + 1 < n_codes) {
            const fe* c1 = STRIDED ? strided_base + (size_t)(j + 1) * stride_elems : src[j + 1];
            uint4 x1 = PAIRS ? __ldg(reinterpret_cast<const uint4*>(c1 + 2 * i)) : __ldg(reinterpret_cast<const uint4*>(c1 + i));
            uint4 y1 = PAIRS ? __ldg(reinterpret_cast<const uint4*>(c1 + 2 * i + 1)) : __ldg(reinterpret_cast<const uint4*>(c1 + i + n_leaves));
            sha_words_from_le(x1, w + 8);
            sha_words_from_le(y1, w + 12);
        } else {  // odd batch: padding starts in the second half of the last data block
            w[8] = 0x80000000u; w[9] = 0; w[10] = 0; w[11] = 0; w[12] = 0; w[13] = 0;
            w[14] = (uint32_t)(bits >> 32); w[15] = (uint32_t)bits;
        }
        sha_compress(st, w);
    }
    if ((n_codes & 1) == 0) {
        uint32_t p[16] = {0x80000000u, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (uint32_t)(bits >> 32), (uint32_t)bits};
        sha_compress(st, p);
    }
    sha_store_digest(layer_ptr(digests, n_leaves, 0, i), st);
}

// Generic byte items (Merkle<T>::commit / batch_commit for arbitrary AsRef<[u8]> items).
__global__ void merkle_bytes_kernel(const uint8_t* const* __restrict__ data, int n_batches, size_t item_bytes, size_t n_items,
                                    uint8_t* __restrict__ digests) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint32_t st[8];
    auto next = [&](size_t pos) -> uint8_t {
        size_t b = pos / item_bytes, off = pos - b * item_bytes;
        return data[b][i * item_bytes + off];
    };
    sha256_bytes(next, (size_t)n_batches * item_bytes, st);
    // digests may be only 32-byte aligned relative to a 256-byte aligned base: vector stores are fine
    sha_store_digest(layer_ptr(digests, n_items, 0, i), st);
}

// Subtree depth per thread.  A thread that walks 2^G inputs runs 2^(G+1) - 1 hashes back to back (≈1.9 us per compression when
// its warp has the SM sub-partition to itself): 8-input walks take 27 us (nodes) / 43 us (leaves) no matter how small the layer is.
// Shallower walks on small layers (more threads, shorter chains) were measured in round 2 (profiles/r2_serial_latency.txt): one
// serial 2^24 commit 10.76 -> 10.60 ms with MLB_WALK_MIN_THREADS = 2^15..2^18, at 97-112 launches instead of 76 — every level
// saved from a walk comes back as a launch gap — so the default stays 0 (always 8-input walks); the knobs remain for experiments:
// MLB_WALK_MIN_THREADS = threads a launch should keep before the walk gets shallower, MLB_TOP_MAX = nodes handed to the cluster top.
static size_t env_size(const char* name, size_t dflt) {
    const char* e = getenv(name);
    return e ? (size_t)strtoull(e, nullptr, 10) : dflt;
}
static int walk_depth(size_t count) {
    static const size_t min_threads = env_size("MLB_WALK_MIN_THREADS", 0);
    int g = 3;
    while (g > 1 && (count >> g) < min_threads) g--;
    return g;
}
static size_t top_max_count() {
    static const size_t v = env_size("MLB_TOP_MAX", TOP_MAX_COUNT);
    return v;
}
// Layers above `from_layer` (which must be complete).
static int upper_from(uint8_t* digests, size_t n_leaves, int from_layer, cudaStream_t s) {
    int total = (int)ilog2(n_leaves);
    int layer = from_layer;
    while (layer < total) {
        size_t count = n_leaves >> layer;
        if (count <= top_max_count()) {
            ProfScope prof(PROF_MERKLE_TOP, 64.0 * (double)count, s);
            MLB_TRY(merkle_top_launch(digests, n_leaves, layer, s));
            MLB_KERNEL_CHECK();
            return ML_OK;
        }
        int g = walk_depth(count);
        if (layer + g > total) g = total - layer;
        const size_t threads = count >> g;
        ProfScope prof(PROF_MERKLE_NODES, 60.0 * (double)count, s);  // read 32 B/digest, write the layers above (<= 28 B/digest)
        const unsigned grid = (unsigned)((threads + 127) / 128);
        if (g >= 3) merkle_nodes_kernel<3><<<grid, 128, 0, s>>>(digests, n_leaves, layer);
        else if (g == 2) merkle_nodes_kernel<2><<<grid, 128, 0, s>>>(digests, n_leaves, layer);
        else merkle_nodes_kernel<1><<<grid, 128, 0, s>>>(digests, n_leaves, layer);
        MLB_KERNEL_CHECK();
        layer += g;
    }
    return ML_OK;
}
int merkle_upper_launch(uint8_t* digests, size_t n_leaves, cudaStream_t s) { return upper_from(digests, n_leaves, 0, s); }

int merkle_rs_launch(const fe* code, size_t n_code, uint8_t* digests, cudaStream_t s) {
    const size_t L = n_code / 2;
    if (L == 0) return ML_ERR_NOT_POW2;
    if (L >= 8) {
        const int g = walk_depth(L);
        const size_t threads = L >> g;
        {   // read one 32-byte pair per leaf, write layers 0..g (32 * (1 + 1/2 + .. + 2^-g) bytes per leaf)
            ProfScope prof(PROF_MERKLE_LEAF, (32.0 + 64.0 - (64.0 / (double)(1 << g))) * (double)L, s);
            const unsigned grid = (unsigned)((threads + 127) / 128);
            if (g >= 3) merkle_rs_kernel<3><<<grid, 128, 0, s>>>(code, L, digests);
            else if (g == 2) merkle_rs_kernel<2><<<grid, 128, 0, s>>>(code, L, digests);
            else merkle_rs_kernel<1><<<grid, 128, 0, s>>>(code, L, digests);
            MLB_KERNEL_CHECK();
        }
        return upper_from(digests, L, g, s);
    }
    merkle_rs_kernel<0><<<1, 128, 0, s>>>(code, L, digests);
    MLB_KERNEL_CHECK();
    return upper_from(digests, L, 0, s);
}

int merkle_batched_rs_launch(const fe* const* codes, size_t n_codes, size_t n_code, uint8_t* digests, cudaStream_t s) {
    const size_t L = n_code / 2;
    if (L == 0 || n_codes == 0) return ML_ERR_NOT_POW2;
    {
        ProfScope prof(PROF_MERKLE_LEAF, (32.0 * (double)n_codes + 32.0) * (double)L, s);
        merkle_batched_kernel<false><<<(unsigned)((L + 127) / 128), 128, 0, s>>>(codes, (int)n_codes, L, digests);
        MLB_KERNEL_CHECK();
    }
    return upper_from(digests, L, 0, s);
}
int merkle_batched_pairs_launch(const uint8_t* const* pairs, size_t n_codes, size_t n_leaves, uint8_t* digests, cudaStream_t s) {
    if (n_leaves == 0 || n_codes == 0) return ML_ERR_NOT_POW2;
    {
        ProfScope prof(PROF_MERKLE_LEAF, (32.0 * (double)n_codes + 32.0) * (double)n_leaves, s);
        merkle_batched_kernel<true><<<(unsigned)((n_leaves + 127) / 128), 128, 0, s>>>((const fe* const*)pairs, (int)n_codes, n_leaves, digests);
        MLB_KERNEL_CHECK();
    }
    return upper_from(digests, n_leaves, 0, s);
}
// pair arrays at pairs_base + j * n_leaves * 32 bytes (one per code); layers above the leaves optional (a caller that shares the
// stream with other work may want the leaf pass alone)
int merkle_batched_pairs_strided_launch(const uint8_t* pairs_base, size_t n_codes, size_t n_leaves, uint8_t* digests, cudaStream_t s) {
    if (n_leaves == 0 || n_codes == 0) return ML_ERR_NOT_POW2;
    {
        ProfScope prof(PROF_MERKLE_LEAF, (32.0 * (double)n_codes + 32.0) * (double)n_leaves, s);
        merkle_batched_kernel<true, true><<<(unsigned)((n_leaves + 127) / 128), 128, 0, s>>>(nullptr, (int)n_codes, n_leaves, digests,
                                                                                              (const fe*)pairs_base, 2 * n_leaves);
        MLB_KERNEL_CHECK();
    }
    return upper_from(digests, n_leaves, 0, s);
}
int merkle_bytes_launch(const uint8_t* const* data, size_t n_batches, size_t item_bytes, size_t n_items, uint8_t* digests, cudaStream_t s) {
    merkle_bytes_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, s>>>(data, (int)n_batches, item_bytes, n_items, digests);
    MLB_KERNEL_CHECK();
    return upper_from(digests, n_items, 0, s);
}

}  // namespace mlb
