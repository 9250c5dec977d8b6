// Batch encode kernels: LOPQModel.predict over rows (model.py:543-602), LOPQModelPCA.apply_PCA
// (model.py:961-978).  All arithmetic is float64 (float32 only where NumPy promotion makes the
// reference compute in float32), distances in NumPy summation order, so codes are bit-exact up to
// the rounding of the BLAS dgemv the reference uses for the rotation.
#pragma once
#include "common.cuh"

// ---- apply_PCA: one block per vector ----------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(128) k_pca(ModelView mv, const XT* __restrict__ X, int64_t n, float* __restrict__ Y) {
    extern __shared__ double sm_pca[];   // [D0] centred input, then [D] output
    double* xc = sm_pca;
    double* y = sm_pca + mv.D0;
    __shared__ double red[4];
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    const XT* x = X + i * (int64_t)mv.D0;
    for (int d = threadIdx.x; d < mv.D0; d += blockDim.x) xc[d] = __dsub_rn((double)x[d], mv.pmu[d]);
    __syncthreads();
    double ss = 0.0;
    for (int e = threadIdx.x; e < mv.D; e += blockDim.x) {
        double acc = 0.0;
        for (int d = 0; d < mv.D0; ++d) acc = fma(xc[d], mv.P[(int64_t)d * mv.D + e], acc);
        y[e] = acc;
        ss = fma(acc, acc, ss);
    }
    if (mv.renorm) {
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        const double nrm = sqrt(tot);
        for (int e = threadIdx.x; e < mv.D; e += blockDim.x) Y[i * (int64_t)mv.D + e] = (float)(y[e] / nrm);
    } else {
        for (int e = threadIdx.x; e < mv.D; e += blockDim.x) Y[i * (int64_t)mv.D + e] = (float)y[e];
    }
}

// ---- coarse assignment + residual + local rotation: one warp per vector -----------------------
// predict_coarse (model.py:563-573 -> utils.py:33-53) and project (model.py:604-641).
// coarse_in != NULL: use the given coarse pair instead of the argmin (project / LUT probes).
#define ENC_WARPS 4
template <typename XT>
__global__ void __launch_bounds__(ENC_WARPS * 32)
k_coarse_project(ModelView mv, const XT* __restrict__ X, int64_t n, const int32_t* __restrict__ coarse_in,
                 int32_t* __restrict__ coarse_out, double* __restrict__ PX) {
    extern __shared__ double sm_r[];     // [ENC_WARPS][h]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * ENC_WARPS + warp;
    if (i >= n) return;                   // whole warp exits together
    const int h = mv.h, V = mv.V;
    double* r = sm_r + warp * h;
    const XT* x = X + i * (int64_t)mv.D;
    const bool f32 = (sizeof(XT) == 4) && mv.coarse_f32;
    for (int s = 0; s < 2; ++s) {
        int c;
        if (coarse_in) {
            c = coarse_in[i * 2 + s];
        } else {
            double best = 1e300;
            int bestv = 0x7fffffff;
            for (int v = lane; v < V; v += 32) {
                const double* C = mv.Cs + ((int64_t)s * V + v) * h;
                double d = f32 ? (double)sqdist_np<float>(x + s * h, C, h) : sqdist_np<double>(x + s * h, C, h);
                if (d < best) { best = d; bestv = v; }    // v ascending: first minimum wins
            }
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int ov = __shfl_xor_sync(0xffffffffu, bestv, o);
                if (ob < best || (ob == best && ov < bestv)) { best = ob; bestv = ov; }
            }
            c = bestv;
        }
        if (lane == 0 && coarse_out) coarse_out[i * 2 + s] = c;
        if (PX) {
            const double* C = mv.Cs + ((int64_t)s * V + c) * h;
            const double* mu = mv.mus + ((int64_t)s * V + c) * h;
            for (int d = lane; d < h; d += 32) r[d] = coarse_residual<XT>(x[s * h + d], C[d], mu[d], mv.coarse_f32);
            __syncwarp();
            const double* Rt = mv.Rt + ((int64_t)s * V + c) * h * (int64_t)h;
            for (int t = lane; t < h; t += 32) {
                double acc = 0.0;
                for (int d = 0; d < h; ++d) acc = fma(Rt[(int64_t)d * h + t], r[d], acc);
                PX[i * (int64_t)mv.D + s * h + t] = acc;
            }
            __syncwarp();
        }
    }
}

// ---- fine argmin: one thread per vector, sub-centroids broadcast from shared memory -----------
// predict_fine (model.py:593-602): per sub-vector argmin of ((fx - subC[j])**2).sum(1), first minimum.
// DS > 0: compile-time sub-vector length (registers); DS == 0: runtime ds <= 128 (local array).
#define FINE_THREADS 128
template <int DS>
__global__ void __launch_bounds__(FINE_THREADS)
k_fine_argmin(ModelView mv, const double* __restrict__ PX, int64_t n, uint8_t* __restrict__ fine, int kchunk) {
    extern __shared__ double sm_sub[];   // [kchunk][ds]
    const int ds = DS ? DS : mv.ds;
    const int64_t i = (int64_t)blockIdx.x * FINE_THREADS + threadIdx.x;
    const bool live = i < n;
    double pj[DS ? DS : 128];
    for (int j = 0; j < mv.M; ++j) {
        if (live) {
            const double* p = PX + i * (int64_t)mv.D + (int64_t)j * ds;
#pragma unroll
            for (int d = 0; d < ds; ++d) pj[d] = p[d];
        }
        double best = 1e300;
        int bestk = 0;
        for (int k0 = 0; k0 < mv.K; k0 += kchunk) {
            const int kc = min(kchunk, mv.K - k0);
            __syncthreads();
            const double* src = mv.subs + ((int64_t)j * mv.K + k0) * ds;
            for (int e = threadIdx.x; e < kc * ds; e += FINE_THREADS) sm_sub[e] = src[e];
            __syncthreads();
            if (live) {
                for (int kk = 0; kk < kc; ++kk) {
                    double d = sqdist_np<double>(pj, sm_sub + kk * ds, ds);
                    if (d < best) { best = d; bestk = k0 + kk; }
                }
            }
        }
        if (live) fine[i * (int64_t)mv.M + j] = (uint8_t)bestk;
    }
}

// ---- float64 LUT rows for explicit probes (get_subquantizer_distances, model.py:673-704) ------
// one block per vector, thread per sub-centroid; lut [n][M][K] float64
__global__ void __launch_bounds__(256)
k_lut64_probe(ModelView mv, const double* __restrict__ PX, int64_t n, double* __restrict__ lut) {
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    for (int k = threadIdx.x; k < mv.K; k += blockDim.x)
        for (int j = 0; j < mv.M; ++j) {
            const double* p = PX + i * (int64_t)mv.D + (int64_t)j * mv.ds;
            const double* c = mv.subs + ((int64_t)j * mv.K + k) * mv.ds;
            lut[(i * mv.M + j) * (int64_t)mv.K + k] = sqdist_np<double>(p, c, mv.ds);
        }
}
