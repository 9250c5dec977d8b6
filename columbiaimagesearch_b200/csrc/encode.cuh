// Batch encode kernels: LOPQModel.predict over rows (model.py:543-602), LOPQModelPCA.apply_PCA
// (model.py:961-978).  All arithmetic is float64 (float32 only where NumPy promotion makes the
// reference compute in float32), distances in NumPy summation order, so codes are bit-exact up to
// the rounding of the BLAS dgemv the reference uses for the rotation.
#pragma once
#include "common.cuh"

// ---- apply_PCA: one block per vector ----------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(128) k_pca(ModelView mv, const XT* __restrict__ X, int64_t n, float* __restrict__ Y,
                                             double* __restrict__ Y64 = nullptr) {
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
        for (int e = threadIdx.x; e < mv.D; e += blockDim.x) {
            if (Y) Y[i * (int64_t)mv.D + e] = (float)(y[e] / nrm);
            if (Y64) Y64[i * (int64_t)mv.D + e] = y[e] / nrm;
        }
    } else {
        for (int e = threadIdx.x; e < mv.D; e += blockDim.x) {
            if (Y) Y[i * (int64_t)mv.D + e] = (float)y[e];
            if (Y64) Y64[i * (int64_t)mv.D + e] = y[e];
        }
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

// ---- coarse assignment alone, 8 lanes per centroid ------------------------------------------------
// predict_coarse (model.py:563-573 -> utils.py:33-53) for h % 8 == 0, h <= 128: NumPy's pairwise sum of such a row is
// eight strided accumulators r_j = sum_i term(j + 8 i) combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).  Lane (v, j)
// builds r_j of centroid v and an xor-butterfly over the 8 lanes performs exactly those additions (IEEE addition is
// commutative), so the distance bits equal sqdist_np's -- with 8x the lanes busy of the one-lane-per-centroid form.
template <typename XT, typename T, typename CT, int HC>
__device__ __forceinline__ void coarse_argmin_split(int h_rt, int V, const XT* x, const CT* Cs, int ldc, int lane, int& bestv_out) {
    // x: the row's split (h values), Cs: the split's centroids [V][h] (float64 model values, or already converted to T)
    // HC: compile-time h (the 8-term strided sums unroll completely), 0: runtime h
    const int h = HC ? HC : h_rt;
    T best = (T)3.0e38;
    int bestv = 0x7fffffff;
    for (int t0 = 0; t0 < V * 8; t0 += 32) {
        const int task = t0 + lane, v = task >> 3, j = task & 7;
        T r = (T)0;
        if (v < V) {
            const CT* C = Cs + (int64_t)v * ldc + j;          // ldc: row stride of Cs (padded in shared memory: no bank conflicts
            const XT* xj = x + j;                             //      between the four centroid groups of a warp)
            if (HC) {
#pragma unroll
                for (int i = 0; i < (HC ? HC : 8); i += 8) {
                    const T t = Ex<T>::sub((T)xj[i], (T)C[i]);
                    const T sq = Ex<T>::mul(t, t);
                    r = (i == 0) ? sq : Ex<T>::add(r, sq);
                }
            } else {
                for (int i = 0; i + j < h; i += 8) {
                    const T t = Ex<T>::sub((T)xj[i], (T)C[i]);
                    const T sq = Ex<T>::mul(t, t);
                    r = (i == 0) ? sq : Ex<T>::add(r, sq);
                }
            }
        }
        r = Ex<T>::add(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = Ex<T>::add(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = Ex<T>::add(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (v < V && (r < best || (r == best && v < bestv))) { best = r; bestv = v; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ov = __shfl_xor_sync(0xffffffffu, bestv, o);
        if (ob < best || (ob == best && ov < bestv)) { best = ob; bestv = ov; }
    }
    bestv_out = bestv;
}

// Persistent blocks; a warp takes a row at a time.  The row is staged in shared memory by coalesced loads (the next row
// is already in flight in registers), and when they fit (2 V h values <= 48 KB: every search model) so are the centroids,
// converted once to the arithmetic type -- the inner loop then reads shared memory only.  c_smem = 0: centroids from
// global memory (k-means training with thousands of clusters).
#define COARSE_WARPS 8
// shared memory of a block: [c_smem: centroids as T [2][V][h] | centroids as float [2][V][h] | half norms [2V] | max norm [2]] | rows [COARSE_WARPS][D] XT
template <typename T> __host__ __device__ inline size_t coarse_c_bytes(int V, int h) {
    return (((size_t)2 * V * (h + 8) * (sizeof(T) + 4) + (size_t)(2 * V + 2) * 4 + 15) / 16) * 16;
}

// With the centroids in shared memory the assignment is first scored in float32, like the fine argmin: score_v =
// |c_v|^2 / 2 - x.c_v, one FFMA per dimension instead of a float64 subtract, multiply and add; the winner is accepted when
// the runner-up is more than 3 E away, E = (h + 16) 2^-24 (|x| + max|c|)^2 bounding the float32 evaluation error of a
// score (inputs rounded to float32, half norm, FMA chain, butterfly sum) plus, for float32 models, the rounding of the
// reference's own float32 distances.  Near ties (and everything when the centroids stay in global memory) take the
// exact path below, bit-compatible with NumPy's pairwise sums.
template <typename XT, typename T, int HC>
__device__ __forceinline__ void coarse_assign_rows(const ModelView& mv, const XT* __restrict__ X, int64_t n, int32_t* __restrict__ coarse_out,
                                                   unsigned char* smem_raw, int c_smem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = mv.D, h = HC ? HC : mv.h, V = mv.V;
    T* Cs_s = (T*)smem_raw;
    const int HP = h + 8;                                    // padded row stride of the staged centroids
    float* Cf_s = (float*)(Cs_s + (size_t)2 * V * HP);
    float* hn_s = Cf_s + (size_t)2 * V * HP;                 // [2V] half norms, then [2] max |c| per split
    const size_t cbytes = c_smem ? coarse_c_bytes<T>(V, h) : 0;
    XT* xs = (XT*)(smem_raw + cbytes) + (size_t)warp * D;
    if (c_smem) {
        for (int e = threadIdx.x; e < 2 * V * h; e += COARSE_WARPS * 32) {
            const int c = e / h, d = e - c * h;
            Cs_s[(size_t)c * HP + d] = (T)mv.Cs[e]; Cf_s[(size_t)c * HP + d] = (float)mv.Cs[e];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < 2 * V; c += COARSE_WARPS * 32) {
            float q = 0.0f;
            for (int i = 0; i < h; ++i) q = fmaf(Cf_s[(size_t)c * HP + i], Cf_s[(size_t)c * HP + i], q);
            hn_s[c] = 0.5f * q;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            float mx = 0.0f;
            for (int v = 0; v < V; ++v) mx = fmaxf(mx, hn_s[threadIdx.x * V + v]);
            hn_s[2 * V + threadIdx.x] = sqrtf(2.0f * mx) * 1.0001f;
        }
        __syncthreads();
    }
    constexpr int NX = 8;                                   // D <= 256: a lane holds up to 8 values of the next row
    XT nx[NX];
    int64_t i = (int64_t)blockIdx.x * COARSE_WARPS + warp;
    const int64_t step = (int64_t)gridDim.x * COARSE_WARPS;
    if (i < n) {
#pragma unroll
        for (int e = 0; e < NX; ++e) if (lane + 32 * e < D) nx[e] = X[i * (int64_t)D + lane + 32 * e];
    }
    const float U = 5.9604645e-08f;
    for (; i < n; i += step) {
        __syncwarp();
#pragma unroll
        for (int e = 0; e < NX; ++e) if (lane + 32 * e < D) xs[lane + 32 * e] = nx[e];
        __syncwarp();
        if (i + step < n) {
#pragma unroll
            for (int e = 0; e < NX; ++e) if (lane + 32 * e < D) nx[e] = X[(i + step) * (int64_t)D + lane + 32 * e];
        }
        for (int s = 0; s < 2; ++s) {
            int c = -1;
            if (c_smem) {
                // float32 stage: lane (v, j) sums the terms j, j+8, ... of centroid v; xor-butterfly over the 8 lanes
                float best = 3.0e38f, second = 3.0e38f, xx = 0.0f;
                int bv = 0;
                const XT* xj = xs + s * h + (lane & 7);
                for (int t0 = 0; t0 < V * 8; t0 += 32) {
                    const int v = (t0 + lane) >> 3;
                    float acc = 0.0f;
                    if (v < V) {
                        const float* cf = Cf_s + ((size_t)s * V + v) * HP + (lane & 7);
                        if (HC) {
#pragma unroll
                            for (int k = 0; k < (HC ? HC : 8); k += 8) { const float xv = (float)xj[k]; acc = fmaf(xv, cf[k], acc); if (t0 == 0) xx = fmaf(xv, xv, xx); }
                        } else {
                            for (int k = 0; k + (lane & 7) < h; k += 8) { const float xv = (float)xj[k]; acc = fmaf(xv, cf[k], acc); if (t0 == 0) xx = fmaf(xv, xv, xx); }
                        }
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    if (v < V) {
                        const float sc = hn_s[s * V + v] - acc;
                        second = fminf(second, fmaxf(sc, best));
                        if (sc < best) bv = v;
                        best = fminf(best, sc);
                    }
                }
                // |x|^2 of the split: lanes 0..7 hold the eight strided partial sums (V >= 1: group 0 always has a centroid)
                xx += __shfl_xor_sync(0xffffffffu, xx, 1);
                xx += __shfl_xor_sync(0xffffffffu, xx, 2);
                xx += __shfl_xor_sync(0xffffffffu, xx, 4);
                xx = __shfl_sync(0xffffffffu, xx, 0);
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {            // merge the four centroid groups of the warp
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
                    const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    second = fminf(fminf(second, os), fmaxf(best, ob));
                    if (ob < best) bv = ov;
                    best = fminf(best, ob);
                }
                const float sp = sqrtf(xx) + hn_s[2 * V + s];
                const float E = (float)(h + 16) * U * sp * sp * 1.01f + 1e-30f;
                if (second - best > 3.0f * E) c = bv;           // (warp-uniform: every lane holds the merged triple)
            }
            if (c < 0) {
                if (c_smem) coarse_argmin_split<XT, T, T, HC>(h, V, xs + s * h, Cs_s + (size_t)s * V * HP, HP, lane, c);
                else coarse_argmin_split<XT, T, double, HC>(h, V, xs + s * h, mv.Cs + (int64_t)s * V * h, h, lane, c);
            }
            if (lane == 0) coarse_out[i * 2 + s] = c;
        }
    }
}

template <typename XT>
__global__ void __launch_bounds__(COARSE_WARPS * 32)
k_coarse_assign(ModelView mv, const XT* __restrict__ X, int64_t n, int32_t* __restrict__ coarse_out, int c_smem) {
    extern __shared__ __align__(16) unsigned char sm_coarse[];
    const bool f32 = (sizeof(XT) == 4) && mv.coarse_f32;
    if (mv.h == 64) { if (f32) coarse_assign_rows<XT, float, 64>(mv, X, n, coarse_out, sm_coarse, c_smem); else coarse_assign_rows<XT, double, 64>(mv, X, n, coarse_out, sm_coarse, c_smem); }
    else if (f32) coarse_assign_rows<XT, float, 0>(mv, X, n, coarse_out, sm_coarse, c_smem);
    else coarse_assign_rows<XT, double, 0>(mv, X, n, coarse_out, sm_coarse, c_smem);
}

// one centroid from shared memory, in the widest loads its (compile-time) length allows; p is 16-byte aligned for DS % 4 == 0
template <int DS> __device__ __forceinline__ void lds_centroid(const float* p, float (&c)[DS]) {
    if (DS % 4 == 0) {
#pragma unroll
        for (int t = 0; t < DS / 4; ++t) { const float4 v = *((const float4*)p + t); c[4 * t] = v.x; c[4 * t + 1] = v.y; c[4 * t + 2] = v.z; c[4 * t + 3] = v.w; }
    } else if (DS % 2 == 0) {
#pragma unroll
        for (int t = 0; t < DS / 2; ++t) { const float2 v = *((const float2*)p + t); c[2 * t] = v.x; c[2 * t + 1] = v.y; }
    } else {
#pragma unroll
        for (int t = 0; t < DS; ++t) c[t] = p[t];
    }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// ---- fine argmin, float32 first stage with a float64 guard ------------------------------------------
// Same result as k_fine_argmin (bit-exact with the reference), on the float32 pipe at ONE FFMA per dimension: the argmin
// of |p - c_k|^2 is the argmin of the score  s_k = |c_k|^2 / 2 - p.c_k  (the |p|^2 term is common), evaluated as an FMA chain
// that starts from the precomputed half norm; a thread scores R rows against each centroid it reads from shared memory
// (one broadcast load of c_k serves R x DS FFMAs).  The float32 winner is accepted only if the runner-up is farther than
// three times a rigorous bound on the float32 evaluation error of a score,
//     |s32 - s| <= E = (DS + 4) * 2^-24 * (|p| + max_k |c_k|)^2
// (rounding of p and c to float32, of the half norm, and of the DS-step FMA chain; |p.c| <= |p||c|): then every other
// centroid is strictly farther in exact arithmetic, by a margin (>= E) far above what float64 rounding could flip.
// Otherwise -- a near tie or an exact tie, a few sub-vectors per 100 000 -- that sub-vector is redone in float64 in
// NumPy's order with the first-minimum rule (utils.py:33-53).
// Two table buffers at compile-time offsets (K <= 256): the table of sub-quantizer j+1 arrives by cp.async while j is
// scored, so a sub-quantizer costs one barrier and no load phase; the body is instantiated once per buffer so that every
// table address is a constant.
template <int DS, int R>
__device__ __forceinline__ void fine32_score(const ModelView& mv, const double* __restrict__ PX, int64_t n, uint8_t* __restrict__ fine,
                                             int j, int64_t i0, const float* sm_sub32, const float* c2h, unsigned int& guards) {
        const float U = 5.9604645e-08f;
        float pn[R][DS], p2[R], best[R], second[R];
        int bestk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t i = (i0 + r < n) ? i0 + r : i0;                 // (a dead second row repeats the first: not stored)
            const double* p64 = PX + i * (int64_t)mv.D + (int64_t)j * DS;
            p2[r] = 0.0f;
#pragma unroll
            for (int d = 0; d < DS; ++d) { const float v = (float)p64[d]; pn[r][d] = -v; p2[r] = fmaf(v, v, p2[r]); }
            best[r] = 3.0e38f; second[r] = 3.0e38f; bestk[r] = 0;
        }
        // centroids four at a time: the two smallest scores of the four merge into (best, second) with 3 min/max operations
        // per score instead of 5 (those run on the half-rate ALU pipe), and only the winning GROUP is tracked: its four
        // scores are recomputed (same FMA chains, same bits) after the loop to name the centroid
        const int K4 = mv.K & ~3;
#pragma unroll 1
        for (int k = 0; k < K4; k += 4) {
            float sc[R][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float c[DS];
#pragma unroll
                for (int t = 0; t < DS; ++t) c[t] = sm_sub32[(k + q) * DS + t];
                const float hn = c2h[k + q];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float v = hn;
#pragma unroll
                    for (int t = 0; t < DS; ++t) v = fmaf(pn[r][t], c[t], v);
                    sc[r][q] = v;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float a = fminf(sc[r][0], sc[r][1]), b = fmaxf(sc[r][0], sc[r][1]);
                const float c = fminf(sc[r][2], sc[r][3]), d = fmaxf(sc[r][2], sc[r][3]);
                const float lo1 = fminf(a, c);
                const float lo2 = fminf(fminf(fmaxf(a, c), b), d);               // second smallest of the four
                second[r] = fminf(fminf(second[r], fmaxf(best[r], lo1)), lo2);   // an exact tie leaves second == best
                if (lo1 < best[r]) bestk[r] = k;                                  // first group holding the minimum
                best[r] = fminf(best[r], lo1);
            }
        }
        for (int k = K4; k < mv.K; ++k) {                                         // (K not a multiple of 4)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float v = c2h[k];
#pragma unroll
                for (int t = 0; t < DS; ++t) v = fmaf(pn[r][t], sm_sub32[k * DS + t], v);
                second[r] = fminf(second[r], fmaxf(v, best[r]));
                if (v < best[r]) bestk[r] = k;
                best[r] = fminf(best[r], v);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {                                             // the winner inside its group
            const int k0 = bestk[r];
            if (k0 < K4) {
                int kk = k0 + 3;
#pragma unroll
                for (int q = 3; q >= 0; --q) {
                    float v = c2h[k0 + q];
#pragma unroll
                    for (int t = 0; t < DS; ++t) v = fmaf(pn[r][t], sm_sub32[(k0 + q) * DS + t], v);
                    if (v == best[r]) kk = k0 + q;                                // first of equal ones (a tie goes to the guard anyway)
                }
                bestk[r] = kk;
            }
        }
        const float cm = sqrtf(mv.c2max[j]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (i0 + r >= n) break;
            const float sp = sqrtf(p2[r]) + cm;
            const float E = (float)(DS + 4) * U * sp * sp * 1.01f + 1e-30f;
            int bk = bestk[r];
            if (!(second[r] - best[r] > 3.0f * E)) {       // cannot be proven in float32: exact evaluation of this sub-vector
                const double* p64 = PX + (i0 + r) * (int64_t)mv.D + (int64_t)j * DS;
                double pj[DS];
#pragma unroll
                for (int d = 0; d < DS; ++d) pj[d] = p64[d];
                double b64 = 1e300;
                for (int k = 0; k < mv.K; ++k) {
                    const double d64 = sqdist_np<double>(pj, mv.subs + ((int64_t)j * mv.K + k) * DS, DS);
                    if (d64 < b64) { b64 = d64; bk = k; }
                }
                ++guards;
            }
            fine[(i0 + r) * (int64_t)mv.M + j] = (uint8_t)bk;
        }
}

template <int DS, int R>
__global__ void __launch_bounds__(FINE_THREADS)
k_fine_argmin32(ModelView mv, const double* __restrict__ PX, int64_t n, uint8_t* __restrict__ fine, unsigned long long* __restrict__ nguard) {
    extern __shared__ __align__(16) float sm_fine[];     // 2 x ([256][DS] centroids + [256] half norms)
    const int64_t i0 = ((int64_t)blockIdx.x * FINE_THREADS + threadIdx.x) * R;
    const float* c2h_all = mv.subs32 + (int64_t)mv.M * mv.K * DS;   // half norms: stored after the centroids
    const bool wide = (mv.K % 4 == 0);
    auto fetch = [&](int j, int buf) {
        float* dst = sm_fine + buf * (256 * (DS + 1));
        const float* sc = mv.subs32 + (int64_t)j * mv.K * DS;
        const float* sh = c2h_all + (int64_t)j * mv.K;
        if (wide) {
            for (int e = threadIdx.x * 4; e < mv.K * DS; e += FINE_THREADS * 4) cp_async16(dst + e, sc + e);
            for (int e = threadIdx.x * 4; e < mv.K; e += FINE_THREADS * 4) cp_async16(dst + 256 * DS + e, sh + e);
        } else {
            for (int e = threadIdx.x; e < mv.K * DS; e += FINE_THREADS) cp_async4(dst + e, sc + e);
            for (int e = threadIdx.x; e < mv.K; e += FINE_THREADS) cp_async4(dst + 256 * DS + e, sh + e);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0, 0);
    unsigned int guards = 0;
    for (int j = 0; j < mv.M; j += 2) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                     // table j landed for everybody; everybody is through with table j-1
        if (j + 1 < mv.M) fetch(j + 1, 1);
        if (i0 < n) fine32_score<DS, R>(mv, PX, n, fine, j, i0, sm_fine, sm_fine + 256 * DS, guards);
        if (j + 1 >= mv.M) break;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (j + 2 < mv.M) fetch(j + 2, 0);
        if (i0 < n) fine32_score<DS, R>(mv, PX, n, fine, j + 1, i0, sm_fine + 256 * (DS + 1), sm_fine + 256 * (DS + 1) + 256 * DS, guards);
    }
    if (guards && nguard) atomicAdd(nguard, (unsigned long long)guards);
}

// All M tables resident (M * K * (DS + 1) floats fit one block's shared memory: D <= ~190 at K = 256): one persistent block
// per SM, the tables are loaded once, and the warps never meet again -- each takes 32 x R rows at a time through all M
// sub-quantizers (no barrier per sub-quantizer: in k_fine_argmin32 a fifth of the warp time is spent waiting for the
// slowest warp of the block at every table switch).
#define FINE_ALL_THREADS 640
template <int DS, int R>
__global__ void __launch_bounds__(FINE_ALL_THREADS, 1)
k_fine_argmin32_all(ModelView mv, const double* __restrict__ PX, int64_t n, uint8_t* __restrict__ fine, unsigned long long* __restrict__ nguard) {
    extern __shared__ __align__(16) float sm_fine[];     // [M][K][DS] centroids, then [M][K] half norms: the layout of mv.subs32
    const int nfl = mv.M * mv.K * (DS + 1);
    if (nfl % 4 == 0) for (int e = threadIdx.x * 4; e < nfl; e += FINE_ALL_THREADS * 4) cp_async16(sm_fine + e, mv.subs32 + e);
    else for (int e = threadIdx.x; e < nfl; e += FINE_ALL_THREADS) cp_async4(sm_fine + e, mv.subs32 + e);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float* c2h_all = sm_fine + mv.M * mv.K * DS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntile = (n + 32 * R - 1) / (32 * R);
    unsigned int guards = 0;
    for (int64_t tile = (int64_t)blockIdx.x * (FINE_ALL_THREADS / 32) + warp; tile < ntile; tile += (int64_t)gridDim.x * (FINE_ALL_THREADS / 32)) {
        const int64_t i0 = (tile * 32 + lane) * R;
        if (i0 >= n) continue;
        for (int j = 0; j < mv.M; ++j)
            fine32_score<DS, R>(mv, PX, n, fine, j, i0, sm_fine + j * mv.K * DS, c2h_all + j * mv.K, guards);
    }
    if (guards && nguard) atomicAdd(nguard, (unsigned long long)guards);
}

// ---- coarse assignment against MANY centroids (V in the thousands: the product's V = 2048 / 4096 models) ------------
// predict_coarse (model.py:563-573 -> utils.py:33-53) when the centroids do not fit the shared memory of k_coarse_assign.
// One row per thread (its split in registers), the split's centroids stream through shared memory in chunks of
// CBIG_CK (float32 copies + half norms, double-buffered by cp.async); scores |c|^2/2 - x.c, four centroids at a time, the
// same bookkeeping and the same kind of guard as k_fine_argmin32:  E = (h + 16) 2^-24 (|x| + max|c|)^2  (float32 inputs,
// half norm, FMA chain, and for float32 models the rounding of the reference's own float32 distances).  A row whose
// runner-up is not more than 3 E away goes to a list that k_coarse_redo settles with the exact NumPy-order arithmetic.
#define CBIG_THREADS 256
#define CBIG_CK 64
// shared memory: row tile [CBIG_THREADS][HC + 4] floats (float32 rows arrive by coalesced cp.async and each thread then reads
// its own row: stride HC + 4 keeps the 16-byte reads of a quarter warp on different banks) | 2 centroid chunk buffers
template <int HC> __host__ __device__ constexpr size_t cbig_smem_bytes() { return (size_t)CBIG_THREADS * (HC + 4) * 4 + (size_t)2 * CBIG_CK * (HC + 1) * 4; }

// R rows per thread: a centroid read from shared memory then serves R x HC FFMAs -- with one row per thread the kernel is
// bound by the shared-memory pipe (one 16-byte broadcast read per 4 FFMAs) as soon as V is large; R = 2 halves that
// traffic (one block of 8 warps per SM, 8 independent FMA chains per thread).  A block takes R x CBIG_THREADS rows.
template <typename XT, int HC, int R>
__global__ void __launch_bounds__(CBIG_THREADS, ((HC <= 64 && R == 1) ? 2 : 1))
k_coarse_big(ModelView mv, const XT* __restrict__ X, int64_t n, int32_t* __restrict__ coarse_out,
             unsigned long long* __restrict__ redo, unsigned int* __restrict__ nredo) {
    extern __shared__ __align__(16) float sm_cbig_all[];  // row tile | 2 x ([CBIG_CK][HC] centroids + [CBIG_CK] half norms)
    float* tile = sm_cbig_all;
    float* sm_cbig = sm_cbig_all + CBIG_THREADS * (HC + 4);
    const float U = 5.9604645e-08f;
    const int V = mv.V;
    const int64_t r0 = (int64_t)blockIdx.x * (CBIG_THREADS * R);
    for (int s = 0; s < 2; ++s) {
        const float* C32 = mv.Cs32 + (int64_t)s * V * HC;
        const float* H32 = mv.Chn32 + (int64_t)s * V;
        auto fetch = [&](int c0, int buf) {
            float* dst = sm_cbig + buf * (CBIG_CK * (HC + 1));
            const int cnt = min(CBIG_CK, V - c0);
            for (int e = threadIdx.x * 4; e < cnt * HC; e += CBIG_THREADS * 4) cp_async16(dst + e, C32 + (int64_t)c0 * HC + e);
            for (int e = threadIdx.x; e < cnt; e += CBIG_THREADS) cp_async4(dst + CBIG_CK * HC + e, H32 + c0 + e);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float pn[R][HC], xx[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int64_t i = r0 + rr * CBIG_THREADS + threadIdx.x;
            if (sizeof(XT) == 4) {
                // CBIG_THREADS rows of this split, coalesced: 16-byte pieces in row order (rows past the end repeat the last one)
                if (rr) __syncthreads();                  // everybody has taken the previous rows out of the tile
                for (int e = threadIdx.x; e < CBIG_THREADS * (HC / 4); e += CBIG_THREADS) {
                    const int r = e / (HC / 4), f = e - r * (HC / 4);
                    const int64_t row = min(r0 + rr * CBIG_THREADS + r, n - 1);
                    cp_async16(tile + r * (HC + 4) + f * 4, (const float*)X + row * (int64_t)mv.D + s * HC + f * 4);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
#pragma unroll
                for (int d = 0; d < HC; d += 4) {
                    const float4 v = *(const float4*)(tile + threadIdx.x * (HC + 4) + d);
                    pn[rr][d] = -v.x; pn[rr][d + 1] = -v.y; pn[rr][d + 2] = -v.z; pn[rr][d + 3] = -v.w;
                }
            } else {
                const XT* x = X + min(i, n - 1) * (int64_t)mv.D + s * HC;         // 16-byte aligned (checked by the caller)
#pragma unroll
                for (int d = 0; d < HC; d += 2) {
                    const double2 v = *(const double2*)((const double*)x + d);
                    pn[rr][d] = -(float)v.x; pn[rr][d + 1] = -(float)v.y;
                }
            }
            xx[rr] = 0.0f;
#pragma unroll
            for (int d = 0; d < HC; ++d) xx[rr] = fmaf(pn[rr][d], pn[rr][d], xx[rr]);
        }
        fetch(0, 0);
        float best[R], second[R];
        int bestg[R];                                     // first centroid of the group (of up to four) that holds the minimum
#pragma unroll
        for (int rr = 0; rr < R; ++rr) { best[rr] = 3.0e38f; second[rr] = 3.0e38f; bestg[rr] = 0; }
        int buf = 0;
        for (int c0 = 0; c0 < V; c0 += CBIG_CK, buf ^= 1) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                              // chunk landed for everybody; everybody is through with the other buffer
            if (c0 + CBIG_CK < V) fetch(c0 + CBIG_CK, buf ^ 1);
            const float* cs = sm_cbig + buf * (CBIG_CK * (HC + 1));
            const float* hn = cs + CBIG_CK * HC;
            const int cnt = min(CBIG_CK, V - c0), cnt4 = cnt & ~3;
#pragma unroll 1
            for (int k = 0; k < cnt4; k += 4) {
                float sc[R][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v[R];
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) v[rr] = hn[k + q];
#pragma unroll
                    for (int t = 0; t < HC; t += 4) {
                        const float4 c = *(const float4*)(cs + (k + q) * HC + t);
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) {
                            v[rr] = fmaf(pn[rr][t], c.x, v[rr]); v[rr] = fmaf(pn[rr][t + 1], c.y, v[rr]);
                            v[rr] = fmaf(pn[rr][t + 2], c.z, v[rr]); v[rr] = fmaf(pn[rr][t + 3], c.w, v[rr]);
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) sc[rr][q] = v[rr];
                }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    const float a = fminf(sc[rr][0], sc[rr][1]), b = fmaxf(sc[rr][0], sc[rr][1]);
                    const float c = fminf(sc[rr][2], sc[rr][3]), d = fmaxf(sc[rr][2], sc[rr][3]);
                    const float lo1 = fminf(a, c);
                    const float lo2 = fminf(fminf(fmaxf(a, c), b), d);
                    second[rr] = fminf(fminf(second[rr], fmaxf(best[rr], lo1)), lo2);
                    if (lo1 < best[rr]) bestg[rr] = c0 + k;
                    best[rr] = fminf(best[rr], lo1);
                }
            }
            for (int k = cnt4; k < cnt; ++k) {            // (V not a multiple of 4): groups of one
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    float v = hn[k];
#pragma unroll
                    for (int t = 0; t < HC; ++t) v = fmaf(pn[rr][t], cs[k * HC + t], v);
                    second[rr] = fminf(second[rr], fmaxf(v, best[rr]));
                    if (v < best[rr]) bestg[rr] = c0 + k;
                    best[rr] = fminf(best[rr], v);
                }
            }
        }
        __syncthreads();                                  // the buffers are refilled for the next split
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int64_t i = r0 + rr * CBIG_THREADS + threadIdx.x;
            if (i >= n) continue;
            // the winner inside its group: the same FMA chains again, from the float32 copies in global memory
            int bv = bestg[rr];
            const int gsz = ((bestg[rr] & (CBIG_CK - 1)) < (min(CBIG_CK, V - (bestg[rr] & ~(CBIG_CK - 1))) & ~3)) ? 4 : 1;
            for (int q = gsz - 1; q >= 0; --q) {
                float v = H32[bestg[rr] + q];
#pragma unroll
                for (int t = 0; t < HC; ++t) v = fmaf(pn[rr][t], C32[(int64_t)(bestg[rr] + q) * HC + t], v);
                if (v == best[rr]) bv = bestg[rr] + q;    // first of equal ones (a tie goes to the list anyway)
            }
            const float sp = sqrtf(xx[rr]) + mv.Cmax32[s];
            const float E = (float)(HC + 16) * U * sp * sp * 1.01f + 1e-30f;
            if (second[rr] - best[rr] > 3.0f * E) coarse_out[i * 2 + s] = bv;
            else redo[atomicAdd(nredo, 1u)] = ((unsigned long long)i << 1) | (unsigned long long)s;     // (capacity 2 n: cannot overflow)
        }
    }
}

// the listed (row, split) pairs in exact arithmetic: one warp each
template <typename XT>
__global__ void __launch_bounds__(256) k_coarse_redo(ModelView mv, const XT* __restrict__ X, int32_t* __restrict__ coarse_out,
                                                     const unsigned long long* __restrict__ redo, const unsigned int* __restrict__ nredo) {
    const unsigned int cnt = *nredo;
    const int lane = threadIdx.x & 31;
    const unsigned int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool f32 = (sizeof(XT) == 4) && mv.coarse_f32;
    for (unsigned int e = wid; e < cnt; e += nw) {
        const int64_t i = (int64_t)(redo[e] >> 1);
        const int s = (int)(redo[e] & 1);
        const XT* x = X + i * (int64_t)mv.D + s * mv.h;
        const double* C = mv.Cs + (int64_t)s * mv.V * mv.h;
        int c;
        if (f32) coarse_argmin_split<XT, float, double, 0>(mv.h, mv.V, x, C, mv.h, lane, c);
        else coarse_argmin_split<XT, double, double, 0>(mv.h, mv.V, x, C, mv.h, lane, c);
        if (lane == 0) coarse_out[i * 2 + s] = c;
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

// ---- batch rotation as a grouped float64 tensor-core GEMM ------------------------------------------
// project (model.py:604-641) for a whole batch: rows are bucketed by (split, coarse code); every bucket is the dense
// contraction  P[rows, h] = Resid[rows, h] . R[c]^T  and runs on the float64 tensor cores (mma.sync m8n8k4 f64,
// DMMA in SASS) with R[c] and a 64-row residual tile staged in shared memory -- instead of one mat-vec per row
// re-reading R[c] from L2 (k_coarse_project).  float64 throughout, so codes stay bit-comparable with the reference
// up to the summation order of the rotation (the reference's own order is BLAS-defined).
// Buckets: cnt / base / tile_base are indexed by b = s*V + c;  perm[s][.] lists the rows of split s bucket by bucket.

__global__ void k_enc_hist(const int32_t* __restrict__ coarse, int64_t n, int V, unsigned int* __restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) / 32) * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < n;
        for (int s = 0; s < 2; ++s) {
            const int c = live ? coarse[2 * i + s] : -1;
            const unsigned int peers = __match_any_sync(0xffffffffu, c);
            if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[s * V + c], (unsigned)__popc(peers));
        }
    }
}

// single block: per split, exclusive prefix of the bucket sizes (row offsets) and of their 64-row tile counts
__global__ void __launch_bounds__(1024) k_enc_offsets(int V, const unsigned int* __restrict__ cnt, unsigned int* __restrict__ base,
                              unsigned int* __restrict__ cursor, unsigned int* __restrict__ tile_base) {
    // one block: every thread owns a contiguous run of the 2 V buckets (split-major), a block scan joins the runs;
    // rows restart at the split boundary, tiles run through
    __shared__ unsigned int s_rows[1024], s_tiles[1024];
    const int nb = 2 * V, per = (nb + 1023) / 1024, lo = min(nb, (int)threadIdx.x * per), hi = min(nb, lo + per);
    unsigned int rows = 0, tiles = 0;
    for (int b = lo; b < hi; ++b) { rows += cnt[b]; tiles += (cnt[b] + 63u) / 64u; }
    s_rows[threadIdx.x] = rows; s_tiles[threadIdx.x] = tiles;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned int r = 0, t = 0;
        if ((int)threadIdx.x >= o) { r = s_rows[threadIdx.x - o]; t = s_tiles[threadIdx.x - o]; }
        __syncthreads();
        s_rows[threadIdx.x] += r; s_tiles[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned int rbase = s_rows[threadIdx.x] - rows, tbase = s_tiles[threadIdx.x] - tiles;     // exclusive prefixes (rows over both splits)
    // rows of split 0 in total = prefix at the first thread whose run starts at or after V
    __shared__ unsigned int s_split0;
    if (threadIdx.x == 0) s_split0 = 0;
    __syncthreads();
    if (lo < V && hi >= V) {                                  // the run that crosses (or ends at) the split boundary
        unsigned int r0 = rbase;
        for (int b = lo; b < V; ++b) r0 += cnt[b];
        s_split0 = r0;
    }
    __syncthreads();
    const unsigned int split0 = s_split0;
    for (int b = lo; b < hi; ++b) {
        base[b] = (b >= V) ? rbase - split0 : rbase;
        cursor[b] = 0u; tile_base[b] = tbase;
        rbase += cnt[b]; tbase += (cnt[b] + 63u) / 64u;
    }
    if (threadIdx.x == 1023) tile_base[nb] = s_tiles[1023];
}

__global__ void k_enc_scatter(const int32_t* __restrict__ coarse, int64_t n, int V, const unsigned int* __restrict__ base,
                              unsigned int* __restrict__ cursor, unsigned int* __restrict__ perm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) / 32) * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < n;
        const int lane = threadIdx.x & 31;
        for (int s = 0; s < 2; ++s) {
            const int c = live ? coarse[2 * i + s] : -1;
            const unsigned int peers = __match_any_sync(0xffffffffu, c);
            const int leader = __ffs(peers) - 1;
            unsigned int start = 0;
            if (live && lane == leader) start = atomicAdd(&cursor[s * V + c], (unsigned)__popc(peers));
            start = __shfl_sync(0xffffffffu, start, leader);
            if (live) perm[(size_t)s * n + base[s * V + c] + start + __popc(peers & ((1u << lane) - 1u))] = (unsigned int)i;
        }
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// h == 64.  A block (4 warps) walks `tpb` consecutive 64-row tiles (tiles are numbered bucket by bucket, so they mostly share
// one R[c]): R[c], C[c], mu[c] are loaded when the bucket changes, the raw rows of tile t+1 arrive by cp.async in the other
// buffer while tile t is contracted (row indices are fetched two tiles ahead), and the coarse residual is formed on the
// fly when the A fragments are read -- one barrier per tile, no load phase in front of the tensor-core loop.
// dynamic smem: Rt[64][68] doubles | C[64], mu[64] doubles | raw[2][64][68] XT | rows[2][64] int
#define ROT_H 64
#define ROT_LD 68
template <typename XT>
size_t rotate_smem_bytes() { return (size_t)ROT_H * ROT_LD * 8 + 128 * 8 + (size_t)2 * 64 * ROT_LD * sizeof(XT) + 2 * 64 * 4; }

template <typename XT>
__global__ void __launch_bounds__(128)
k_rotate_dmma(ModelView mv, const XT* __restrict__ X, int64_t n, const unsigned int* __restrict__ cnt,
              const unsigned int* __restrict__ base, const unsigned int* __restrict__ tile_base, const unsigned int* __restrict__ perm,
              double* __restrict__ PX, int tpb) {
    extern __shared__ __align__(16) double sm_rot[];
    double* Rs = sm_rot;                           // Rs[d][t] = Rt[d][t]
    double* Cm = Rs + ROT_H * ROT_LD;              // C[64] | mu[64]
    XT* Xs = (XT*)(Cm + 128);                      // raw rows, two tiles
    int* rows = (int*)(Xs + 2 * 64 * ROT_LD);
    const int V = mv.V, nb = 2 * V;
    const unsigned int total = tile_base[nb];
    const unsigned int t0 = blockIdx.x * (unsigned)tpb;
    if (t0 >= total) return;
    const unsigned int t1 = min(total, t0 + (unsigned)tpb);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ar = lane >> 2, ak = lane & 3;
    const int cr = tid >> 1, chalf = tid & 1;      // this thread copies half `chalf` of row `cr` of a tile
    constexpr int HALF = ROT_H / 2, NCP = HALF * (int)sizeof(XT) / 16;

    auto bucket_from = [&](int b, unsigned int t) { while (tile_base[b + 1] <= t) ++b; return b; };   // (empty buckets have no tiles)
    auto row_index = [&](unsigned int t, int b) -> int {                 // source row of this thread's copy row in tile t
        const unsigned int row0 = (t - tile_base[b]) * 64u;
        return (row0 + (unsigned)cr < cnt[b]) ? (int)perm[(size_t)(b / V) * n + base[b] + row0 + cr] : -1;
    };
    auto issue = [&](int idx, int b, int buf) {
        if (idx >= 0) {
            const XT* src = X + (int64_t)idx * mv.D + (b / V) * ROT_H + chalf * HALF;
            XT* dst = Xs + (size_t)(buf * 64 + cr) * ROT_LD + chalf * HALF;
#pragma unroll
            for (int c = 0; c < NCP; ++c) cp_async16((char*)dst + 16 * c, (const char*)src + 16 * c);
        }
        if (chalf == 0) rows[buf * 64 + cr] = idx;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int bcur;
    {
        int lo = 0, hi = nb;                       // tile_base[lo] <= t0 < tile_base[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_base[mid] <= t0) lo = mid; else hi = mid; }
        bcur = bucket_from(lo, t0);
    }
    issue(row_index(t0, bcur), bcur, 0);
    int bnext = bcur, idx1 = -1;
    if (t0 + 1 < t1) { bnext = bucket_from(bcur, t0 + 1); idx1 = row_index(t0 + 1, bnext); }
    int bload = -1;
    for (unsigned int t = t0; t < t1; ++t) {
        const int buf = (int)((t - t0) & 1u);
        if (bcur != bload) {                       // another bucket: its rotation, centroid and mean
            if (bload >= 0) __syncthreads();       // (the previous tile still reads the old ones)
            const double* Rt = mv.Rt + (int64_t)bcur * ROT_H * ROT_H;
            for (int e = tid; e < ROT_H * ROT_H; e += 128) Rs[(e >> 6) * ROT_LD + (e & 63)] = Rt[e];
            if (tid < 64) Cm[tid] = mv.Cs[(int64_t)bcur * ROT_H + tid];
            else Cm[tid] = mv.mus[(int64_t)bcur * ROT_H + tid - 64];
            bload = bcur;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                           // tile t (and R) in place; everybody is through with tile t-1
        if (t + 1 < t1) issue(idx1, bnext, buf ^ 1);
        int b2 = bnext, idx2 = -1;
        if (t + 2 < t1) { b2 = bucket_from(bnext, t + 2); idx2 = row_index(t + 2, b2); }   // (latency overlaps the contraction)
        const int s = bcur / V;
        double acc[2][8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const XT* xt = Xs + (size_t)(buf * 64 + 16 * warp + ar) * ROT_LD + ak;
#pragma unroll 4
        for (int k0 = 0; k0 < ROT_H; k0 += 4) {
            double a[2], bb[8];
            const double cc = Cm[k0 + ak], mm = Cm[64 + k0 + ak];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) a[mt] = coarse_residual<XT>(xt[(size_t)8 * mt * ROT_LD + k0], cc, mm, mv.coarse_f32);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) bb[nt] = Rs[(k0 + ak) * ROT_LD + 8 * nt + ar];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], bb[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int i = rows[buf * 64 + 16 * warp + 8 * mt + ar];
            if (i < 0) continue;
            double* o = PX + (int64_t)i * mv.D + s * ROT_H + 2 * ak;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) *(double2*)(o + 8 * nt) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
        bcur = bnext; bnext = b2; idx1 = idx2;
    }
}

// ---- the same grouped GEMM for any h that is a multiple of 64 (2048-d models: h = 1024) -------------------------------
// grid = (64-row tiles, h / 64 column blocks).  A block produces a 64 x 64 tile of P = Resid . R[c]^T, walking the
// contraction index in chunks of 64: the R chunk [64 d][64 t] and the residual chunk [64 rows][64 d] (recomputed from
// the raw rows: x - C - mu) are staged in shared memory, 16 DMMA k-steps per chunk.  Every R[c] (8 MB at h = 1024) is
// thus read once per 64 rows instead of once per row.
// MODE 0 (batch encode): tile rows are database rows i = perm[s][.], input X[i], output PX[i][s*h + t].
// MODE 1 (distance tables of a query batch): tile rows are LUT slots = perm[s][.], input Xq[desc[slot].q], output P64[slot][t];
//         `n` is then the slot capacity (row stride of perm).
template <typename XT, int MODE>
__global__ void __launch_bounds__(128)
k_rotate_dmma_g0(ModelView mv, const XT* __restrict__ X, int64_t n, const unsigned int* __restrict__ cnt,
                const unsigned int* __restrict__ base, const unsigned int* __restrict__ tile_base, const unsigned int* __restrict__ perm,
                const int32_t* __restrict__ desc, double* __restrict__ OUT) {
    extern __shared__ double sm_rot[];
    double* Rs = sm_rot;                       // Rs[d][t]
    double* Es = sm_rot + 64 * ROT_LD;         // Es[row][d]
    int* rows = (int*)(Es + 64 * ROT_LD);
    const int V = mv.V, nb = 2 * V, h = mv.h;
    const unsigned int tile = blockIdx.x;
    if (tile >= tile_base[nb]) return;
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_base[mid] <= tile) lo = mid; else hi = mid; }
    const int b = lo, s = b / V;
    const unsigned int row0 = (tile - tile_base[b]) * 64u;
    const int nrow = (int)min(64u, cnt[b] - row0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.y * 64;
    const double* Rt = mv.Rt + (int64_t)b * h * (int64_t)h;
    const double* C = mv.Cs + (int64_t)b * h;
    const double* mu = mv.mus + (int64_t)b * h;
    if (tid < 64) rows[tid] = tid < nrow ? (int)perm[(size_t)s * n + base[b] + row0 + tid] : -1;
    double acc[2][8][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const int ar = lane >> 2, ak = lane & 3;
    for (int d0 = 0; d0 < h; d0 += 64) {
        __syncthreads();                       // rows[] published / previous chunk consumed
        for (int e = tid; e < 64 * 64; e += 128) Rs[(e >> 6) * ROT_LD + (e & 63)] = Rt[(int64_t)(d0 + (e >> 6)) * h + t0 + (e & 63)];
        for (int e = tid; e < 64 * 64; e += 128) {
            const int r = e >> 6, d = d0 + (e & 63);
            const int i = rows[r];
            double v = 0.0;
            if (i >= 0) {
                const XT* x = MODE == 0 ? X + (int64_t)i * mv.D + s * h : X + (int64_t)desc[3 * i] * mv.D + s * h;
                v = coarse_residual<XT>(x[d], C[d], mu[d], mv.coarse_f32);
            }
            Es[r * ROT_LD + (e & 63)] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int k0 = 0; k0 < 64; k0 += 4) {
            double a[2], bb[8];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) a[mt] = Es[(16 * warp + 8 * mt + ar) * ROT_LD + k0 + ak];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) bb[nt] = Rs[(k0 + ak) * ROT_LD + 8 * nt + ar];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], bb[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int i = rows[16 * warp + 8 * mt + ar];
        if (i < 0) continue;
        double* o = (MODE == 0 ? OUT + (int64_t)i * mv.D + s * h : OUT + (int64_t)i * h) + t0 + 2 * ak;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *(double2*)(o + 8 * nt) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// The same, pipelined (k_rotate_dmma_g0 stays as the path for inputs that are not 16-byte aligned): the R chunk, the raw
// row chunk and the C / mu chunk of contraction step c+1 arrive by cp.async in the other buffers while step c runs on the
// tensor cores; the residual is formed when the A fragments are read.  One barrier per 64-wide contraction step.
// dynamic smem: R[2][64][68] doubles | Cmu[2][128] doubles | raw[2][64][68] XT | rows[64] int
template <typename XT>
size_t rotate_g_smem_bytes() { return (size_t)2 * 64 * ROT_LD * 8 + 2 * 128 * 8 + (size_t)2 * 64 * ROT_LD * sizeof(XT) + 64 * 4; }

template <typename XT, int MODE>
__global__ void __launch_bounds__(128)
k_rotate_dmma_g(ModelView mv, const XT* __restrict__ X, int64_t n, const unsigned int* __restrict__ cnt,
                const unsigned int* __restrict__ base, const unsigned int* __restrict__ tile_base, const unsigned int* __restrict__ perm,
                const int32_t* __restrict__ desc, double* __restrict__ OUT) {
    extern __shared__ __align__(16) double sm_rot[];
    double* Rs = sm_rot;                           // [2][64][68]: Rs[buf][d][t]
    double* Cm = Rs + 2 * 64 * ROT_LD;             // [2][C 64 | mu 64]
    XT* Xs = (XT*)(Cm + 2 * 128);                  // [2][64][68] raw rows
    int* rows = (int*)(Xs + 2 * 64 * ROT_LD);
    const int V = mv.V, nb = 2 * V, h = mv.h;
    const unsigned int tile = blockIdx.x;
    if (tile >= tile_base[nb]) return;
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_base[mid] <= tile) lo = mid; else hi = mid; }
    const int b = lo, s = b / V;
    const unsigned int row0 = (tile - tile_base[b]) * 64u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t0 = blockIdx.y * 64;
    const double* Rt = mv.Rt + (int64_t)b * h * (int64_t)h + t0;
    const double* C = mv.Cs + (int64_t)b * h;
    const double* mu = mv.mus + (int64_t)b * h;
    const int cr = tid >> 1, chalf = tid & 1;      // this thread copies half `chalf` of row `cr` of every chunk
    constexpr int NCX = 32 * (int)sizeof(XT) / 16;
    int myrow = -1;
    if (row0 + (unsigned)cr < cnt[b]) myrow = (int)perm[(size_t)s * n + base[b] + row0 + cr];
    if (chalf == 0) rows[cr] = myrow;
    const XT* xrow = nullptr;
    if (myrow >= 0) xrow = (MODE == 0 ? X + (int64_t)myrow * mv.D : X + (int64_t)desc[3 * myrow] * mv.D) + s * h + chalf * 32;
    auto issue = [&](int d0, int buf) {
        // R chunk: rows d0 .. d0+63 of Rt, 64 doubles each (512 B = 32 x 16 B): thread -> (row cr, half chalf): 16 copies
        {
            const char* src = (const char*)(Rt + (int64_t)(d0 + cr) * h + chalf * 32);
            char* dst = (char*)(Rs + (size_t)(buf * 64 + cr) * ROT_LD + chalf * 32);
#pragma unroll
            for (int c = 0; c < 16; ++c) cp_async16(dst + 16 * c, src + 16 * c);
        }
        if (xrow) {
            const char* src = (const char*)(xrow + d0);
            char* dst = (char*)(Xs + (size_t)(buf * 64 + cr) * ROT_LD + chalf * 32);
#pragma unroll
            for (int c = 0; c < NCX; ++c) cp_async16(dst + 16 * c, src + 16 * c);
        }
        if (tid < 32) cp_async16(Cm + buf * 128 + 2 * tid, C + d0 + 2 * tid);
        else if (tid < 64) cp_async16(Cm + buf * 128 + 64 + 2 * (tid - 32), mu + d0 + 2 * (tid - 32));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[2][8][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const int ar = lane >> 2, ak = lane & 3;
    issue(0, 0);
    int buf = 0;
    for (int d0 = 0; d0 < h; d0 += 64, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                           // chunk d0 landed for everybody; everybody is through with chunk d0-64
        if (d0 + 64 < h) issue(d0 + 64, buf ^ 1);
        const XT* xt = Xs + (size_t)(buf * 64 + 16 * warp + ar) * ROT_LD + ak;
        const double* rs = Rs + (size_t)buf * 64 * ROT_LD;
        const double* cm = Cm + buf * 128;
#pragma unroll 4
        for (int k0 = 0; k0 < 64; k0 += 4) {
            double a[2], bb[8];
            const double cc = cm[k0 + ak], mm = cm[64 + k0 + ak];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) a[mt] = coarse_residual<XT>(xt[(size_t)8 * mt * ROT_LD + k0], cc, mm, mv.coarse_f32);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) bb[nt] = rs[(k0 + ak) * ROT_LD + 8 * nt + ar];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], a[mt], bb[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int i = rows[16 * warp + 8 * mt + ar];
        if (i < 0) continue;
        double* o = (MODE == 0 ? OUT + (int64_t)i * mv.D + s * h : OUT + (int64_t)i * h) + t0 + 2 * ak;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *(double2*)(o + 8 * nt) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// bucketing of LUT slots by (split, coarse code) for MODE 1: desc [slot][3] = (q, split, c); nslot read from the device
__global__ void k_slot_hist(const int32_t* __restrict__ desc, const unsigned int* __restrict__ nslot_p, int V, unsigned int* __restrict__ cnt) {
    const unsigned int nslot = *nslot_p;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x)
        atomicAdd(&cnt[desc[3 * i + 1] * V + desc[3 * i + 2]], 1u);
}
__global__ void k_slot_scatter(const int32_t* __restrict__ desc, const unsigned int* __restrict__ nslot_p, int V, size_t cap,
                               const unsigned int* __restrict__ base, unsigned int* __restrict__ cursor, unsigned int* __restrict__ perm) {
    const unsigned int nslot = *nslot_p;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslot; i += gridDim.x * blockDim.x) {
        const int s = desc[3 * i + 1], b = s * V + desc[3 * i + 2];
        perm[(size_t)s * cap + base[b] + atomicAdd(&cursor[b], 1u)] = i;
    }
}
