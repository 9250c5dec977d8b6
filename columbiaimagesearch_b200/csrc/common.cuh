// Shared device helpers of libb200lopq: exact NumPy-order arithmetic, model view, error macros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#define B2L_MAX_V 64          // full V*V visit tables per query (large-V traversal is a later row)
#define B2L_MAX_K 256         // subquantizer clusters; fine codes are bytes
#define B2L_LUT_ROWS 256      // LUT rows staged in shared memory (indexed by a code byte)

// Device-side view of the model (all pointers device memory).
struct ModelView {
    int D, V, M, K, h, m, ds;  // h = D/2 (coarse split), m = M/2 (fine splits per coarse split), ds = D/M
    int MP;                    // padded code row stride in bytes on the device (power of two >= 4, or M rounded up)
    int G;                     // lane groups per warp in the scan = 32 / MP (0 when the fast scan is unavailable)
    int SW;                    // code-row swizzle mask (MP-1 with the fast scan, else 0): stored byte s of in-cell row i
                               // is code byte (i & SW) ^ s  (see scan.cuh)
    int coarse_f32;            // coarse centroids were float32 in the model object
    const double* Cs;          // [2][V][h]
    const double* mus;         // [2][V][h]
    const double* Rt;          // [2][V][h(d)][h(t)] : Rt[s][c][d][t] = R[s][c][t][d]  (coalesced mat-vec)
    const double* subs;        // [M][K][ds]
    const float* subs32;       // [M][K][ds] the same, rounded to float32 (first stage of the fine argmin)
    const float* subs32T;      // [M][ds][K] float32, centroid index fastest (coalesced one-thread-per-centroid reads)
    const float* c2max;        // [M] upper bound on max_k |subs[j][k]|^2
    const float* Cs32;         // [2][V][h] coarse centroids rounded to float32, [2][V] their half norms, [2] max |c| per split
    const float* Chn32;        //   (first stage of the coarse assignment against many centroids, k_coarse_big)
    const float* Cmax32;
    // PCA (LOPQModelPCA)
    int D0, renorm;
    const double* P;           // [D0][D]
    const double* pmu;         // [D0]
};

// ---- exact (never contracted) arithmetic in either precision ------------------------------------
template <typename T> struct Ex;
template <> struct Ex<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};
template <> struct Ex<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};

// NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum) of
// term(i), i in [lo, lo+n): the order `ndarray.sum(axis=-1)` uses on a contiguous axis, which is what
// utils.py:47 / search.py:39 / model.py:702 evaluate.  Verified against NumPy 2.3 for n = 1..4096.
template <typename T, typename F>
__device__ __forceinline__ T pairwise_leaf(const F& term, int lo, int n) {      // n <= 128
    if (n < 8) {
        T r = (T)0;
        for (int i = 0; i < n; ++i) r = Ex<T>::add(r, term(lo + i));
        return r;
    }
    T r0 = term(lo + 0), r1 = term(lo + 1), r2 = term(lo + 2), r3 = term(lo + 3);
    T r4 = term(lo + 4), r5 = term(lo + 5), r6 = term(lo + 6), r7 = term(lo + 7);
    int i = 8;
    const int n8 = n - (n % 8);
    for (; i < n8; i += 8) {
        r0 = Ex<T>::add(r0, term(lo + i + 0)); r1 = Ex<T>::add(r1, term(lo + i + 1));
        r2 = Ex<T>::add(r2, term(lo + i + 2)); r3 = Ex<T>::add(r3, term(lo + i + 3));
        r4 = Ex<T>::add(r4, term(lo + i + 4)); r5 = Ex<T>::add(r5, term(lo + i + 5));
        r6 = Ex<T>::add(r6, term(lo + i + 6)); r7 = Ex<T>::add(r7, term(lo + i + 7));
    }
    T res = Ex<T>::add(Ex<T>::add(Ex<T>::add(r0, r1), Ex<T>::add(r2, r3)),
                       Ex<T>::add(Ex<T>::add(r4, r5), Ex<T>::add(r6, r7)));
    for (; i < n; ++i) res = Ex<T>::add(res, term(lo + i));
    return res;
}

// n > 128: NumPy halves the range (first half rounded down to a multiple of 8) and adds the two partial sums.
// Evaluated with an explicit frame stack (no device recursion: its stack need cannot be sized statically).
template <typename T, typename F>
__device__ T pairwise_sum(const F& term, int lo, int n) {
    if (n <= 128) return pairwise_leaf<T>(term, lo, n);
    int f_lo[24], f_n[24], f_state[24];
    T f_left[24];
    int sp = 0;
    f_lo[0] = lo; f_n[0] = n; f_state[0] = 0; f_left[0] = (T)0;
    T ret = (T)0;
    while (sp >= 0) {
        if (f_n[sp] <= 128) { ret = pairwise_leaf<T>(term, f_lo[sp], f_n[sp]); --sp; continue; }
        int n2 = f_n[sp] / 2;
        n2 -= n2 % 8;
        if (f_state[sp] == 0) {
            f_state[sp] = 1;
            f_lo[sp + 1] = f_lo[sp]; f_n[sp + 1] = n2; f_state[sp + 1] = 0;
            ++sp;
        } else if (f_state[sp] == 1) {
            f_left[sp] = ret; f_state[sp] = 2;
            f_lo[sp + 1] = f_lo[sp] + n2; f_n[sp + 1] = f_n[sp] - n2; f_state[sp + 1] = 0;
            ++sp;
        } else {
            ret = Ex<T>::add(f_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

// squared L2 distance ((x - c)**2).sum() in NumPy order, x and c already in the compute type T
template <typename T, typename XP, typename CP>
__device__ __forceinline__ T sqdist_np(const XP* x, const CP* c, int n) {
    auto term = [&](int i) -> T {
        T t = Ex<T>::sub((T)x[i], (T)c[i]);
        return Ex<T>::mul(t, t);
    };
    return pairwise_sum<T>(term, 0, n);
}

// residual of the coarse stage, model.py:635-637:  r = cx - C[c]  (float32 when both are float32),
// then r - mu[c] in float64.
template <typename XT>
__device__ __forceinline__ double coarse_residual(XT x, double C, double mu, int coarse_f32) {
    double r;
    if (sizeof(XT) == 4 && coarse_f32) r = (double)__fsub_rn((float)x, (float)C);
    else r = __dsub_rn((double)x, C);
    return __dsub_rn(r, mu);
}

// code byte j of the stored (swizzled) row of in-cell index `incell`
__device__ __forceinline__ uint8_t code_byte(const uint8_t* row, int64_t incell, int j, int SW) {
    return row[((int)incell & SW) ^ j];
}

__device__ __forceinline__ int next_pow2_dev(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// in-shared-memory bitonic sort of n (power of two) 64-bit keys, ascending, by the whole block
__device__ inline void bitonic_sort_u64(unsigned long long* buf, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = buf[i], b = buf[ixj];
                    bool asc = ((i & k) == 0);
                    if ((a > b) == asc) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

#define B2L_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
