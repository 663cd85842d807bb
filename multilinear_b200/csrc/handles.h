// Opaque handle layouts behind include/multilinear_b200.h (host-side bookkeeping; all bulk data is in HBM).
#pragma once
#include <memory>
#include <vector>

#include "internal.h"
#include "sha256.cuh"

struct ml_transcript {
    mlb::HostSha256 sha;
};

// Merkle<T> (src/merkle_tree/mod.rs:8-11): every layer + the data stay resident on the device.
struct ml_merkle {
    enum Kind { BYTES = 0, RS_CODE = 1 } kind = BYTES;
    size_t n_leaves = 0;
    size_t item_bytes = 0;     // bytes per item per batch (32 for RS pairs)
    size_t n_batches = 0;      // 0: Merkle<T>; >0: Merkle<Vec<T>> built by batch_commit
    uint8_t* digests = nullptr;         // all layers back to back (merkle_layer_offset)
    std::vector<void*> data;            // BYTES: device copies [n_items][item_bytes]; RS_CODE: code pointers (n_code = 2*n_leaves elements)
    void** data_ptrs_dev = nullptr;     // device array of the pointers above (batched kernels, gathers)
    bool owns_data = true;
    bool owns_digests = true;
    uint8_t root[32];
    int device = 0;
    cudaStream_t stream = nullptr;  // stream-ordered allocations are returned to the pool on this stream
};

struct ml_fri {
    struct Layer {
        mlb::fe* code = nullptr;  // n elements, natural order; leaf i = (code[i], code[i + n/2])
        size_t n = 0;
        bool owns_code = true;
        ml_merkle* tree = nullptr;  // digests + root
    };
    std::vector<Layer> layers;
    bool has_last = false;
    mlb::hfe last = 0;
    int log_n0 = 0;  // log2 of the original domain (gen_pows.len())
    cudaStream_t stream = nullptr;
};

struct PathH {
    std::vector<uint8_t> value;    // 32 bytes (n_batches*32 for the batch layer)
    std::vector<uint8_t> digests;  // path_len * 32
    std::vector<uint8_t> dirs;     // path_len
};
struct QueryH {
    std::vector<PathH> paths;
};
struct ml_fri_proof {
    std::vector<uint8_t> commitments;  // n * 32
    std::vector<QueryH> queries;
    mlb::hfe last_elem = 0;
    uint8_t last_random[32];
};
struct ml_sumcheck {
    mlb::fe* matrix = nullptr;
    mlb::fe* delta = nullptr;
    bool owns_matrix = true;  // false: the matrix lives in a sharded prover's peer-visible arena
    size_t height = 0;
    cudaStream_t stream = nullptr;
};
// SumcheckTables with an arbitrary trace width (System path); composition = sparse polynomial over the row
struct ml_wsumcheck {
    mlb::fe* matrix = nullptr;  // [height][width] row-major
    mlb::fe* delta = nullptr;
    size_t width = 0, height = 0;
    mlb::fe* coef = nullptr;    // device copies of the composition terms
    uint32_t *len = nullptr, *off = nullptr, *cols = nullptr;
    size_t n_terms = 0, n_cols = 0;
    cudaStream_t stream = nullptr;
};
struct ml_pcs_proof {
    ml_fri_proof fri;
    std::vector<mlb::hfe> sumcheck;  // rounds * 2 (c1, c2)
    std::vector<mlb::hfe> inputs;
    mlb::hfe output = 0;
};
struct BQueryH {
    PathH batch_path;
    QueryH query;
};
struct ml_bfri_proof {
    uint8_t batch_commitment[32];
    std::vector<uint8_t> commitments;
    std::vector<BQueryH> queries;
    mlb::hfe last_elem = 0;
    uint8_t last_random[32];
};
struct ml_bpcs_proof {
    ml_bfri_proof fri;
    std::vector<mlb::hfe> sumcheck;
    std::vector<mlb::hfe> inputs, outputs;
};
