// Training support (model.py:290-336: the k-means fits of the coarse quantizers and of the sub-quantizers).  Lloyd
// iterations entirely on the device: assignment = utils.predict_cluster over rows (direct-form squared L2 in NumPy's
// summation order, first minimum, float64), centroid update = per-cluster sums by float64 atomics.  Training has no parity
// contract (models are inputs of the hot path): the atomics make the last bits of a centroid run-dependent.
#pragma once
#include "common.cuh"

#define KM_WARPS 8

// one warp per row: nearest centroid (first minimum), squared distance added to *cost
// dynamic smem: x[KM_WARPS][d] doubles
__global__ void __launch_bounds__(KM_WARPS * 32)
k_km_assign(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ C, int k, int32_t* __restrict__ assign,
            double* __restrict__ cost) {
    extern __shared__ double sm_km[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * KM_WARPS + warp;
    if (i >= n) return;
    double* x = sm_km + (size_t)warp * d;
    for (int t = lane; t < d; t += 32) x[t] = X[i * d + t];
    __syncwarp();
    double best = 1e300;
    int bestc = 0x7fffffff;
    for (int c = lane; c < k; c += 32) {
        const double dist = sqdist_np<double>(x, C + (int64_t)c * d, d);
        if (dist < best) { best = dist; bestc = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bestc, o);
        if (ob < best || (ob == best && oc < bestc)) { best = ob; bestc = oc; }
    }
    if (lane == 0) { assign[i] = bestc; atomicAdd(cost, best); }
}

__global__ void k_km_accum(const double* __restrict__ X, int64_t n, int d, const int32_t* __restrict__ assign,
                           double* __restrict__ sums, unsigned long long* __restrict__ counts) {
    const int64_t total = n * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int t = (int)(e - i * d);
        const int a = assign[i];
        atomicAdd(&sums[(int64_t)a * d + t], X[e]);
        if (t == 0) atomicAdd(&counts[a], 1ull);
    }
}

// C[c] = mean of its rows, or row reseed[c] of X for an empty cluster
__global__ void k_km_update(const double* __restrict__ X, int d, int k, const double* __restrict__ sums,
                            const unsigned long long* __restrict__ counts, const int64_t* __restrict__ reseed, double* __restrict__ C) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * d; e += gridDim.x * blockDim.x) {
        const int c = e / d, t = e - c * d;
        const unsigned long long cnt = counts[c];
        C[e] = cnt ? sums[e] / (double)cnt : X[reseed[c] * d + t];
    }
}
