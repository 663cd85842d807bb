// Block-wide modular sums (exact integer arithmetic: any summation order gives the reference's result).
#pragma once
#include "field.cuh"

namespace mlb {

__device__ __forceinline__ fe fe_shfl_down(fe a, int delta) {
    fe r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
    return r;
}
__device__ __forceinline__ fe warp_sum(fe a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = fe_add(a, fe_shfl_down(a, d));
    return a;  // lane 0 holds the sum
}
// sum over the whole CTA (blockDim.x multiple of 32, <= 1024); result valid in thread 0. `scratch` >= 32 elements.
__device__ __forceinline__ fe block_sum(fe a, fe* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    __syncthreads();  // scratch may still be in use by a previous call
    if (lane == 0) scratch[warp] = a;
    __syncthreads();
    fe r = fe_zero();
    if (warp == 0) {
        r = lane < nwarps ? scratch[lane] : fe_zero();
        r = warp_sum(r);
    }
    return r;
}

static inline fe to_dev_fe_h(unsigned __int128 x) {
    fe r;
    r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); r.v[2] = (uint32_t)(x >> 64); r.v[3] = (uint32_t)(x >> 96);
    return r;
}

}  // namespace mlb
