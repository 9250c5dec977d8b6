// Selection kernels: merge of the per-segment partial lists, float64 re-rank of the survivors in the
// reference's summation order (search.py:173: left-to-right python sum over the M LUT entries, each
// entry ((fx - subC[j][k])**2).sum() in NumPy order, model.py:702), certification, and the final
// merge over ranks (consumes the all-gathered record buffers directly).
#pragma once
#include "common.cuh"
#include "plan.cuh"

// Record buffer of one rank for nq queries x k results (device memory, 8-byte arrays first).
struct RecView {
    double* d64;                 // [nq][k] exact float64 distance
    unsigned long long* pos;     // [nq][k] retrieval position (global over ranks)
    int64_t* rowid;              // [nq][k]
    double* lb;                  // [nq] lower bound on the exact distance of any candidate NOT in the list (+inf: none)
    int64_t* ncand;              // [nq] retrieved codes (global)
    int32_t* cell;               // [nq][k] c0*V + c1
    int32_t* count;              // [nq] valid entries
    int32_t* visited;            // [nq]
    uint8_t* fine;               // [nq][k][M]
};
__host__ __device__ inline size_t rec_bytes(int nq, int k, int M) {
    size_t b = (size_t)nq * k * 8 * 3 + (size_t)nq * 8 * 2 + (size_t)nq * k * 4 + (size_t)nq * 4 * 2 + (size_t)nq * k * M;
    return (b + 255) & ~(size_t)255;
}
__host__ __device__ inline RecView rec_view(void* base, int nq, int k, int M) {
    RecView r;
    unsigned char* p = (unsigned char*)base;
    r.d64 = (double*)p; p += (size_t)nq * k * 8;
    r.pos = (unsigned long long*)p; p += (size_t)nq * k * 8;
    r.rowid = (int64_t*)p; p += (size_t)nq * k * 8;
    r.lb = (double*)p; p += (size_t)nq * 8;
    r.ncand = (int64_t*)p; p += (size_t)nq * 8;
    r.cell = (int32_t*)p; p += (size_t)nq * k * 4;
    r.count = (int32_t*)p; p += (size_t)nq * 4;
    r.visited = (int32_t*)p; p += (size_t)nq * 4;
    r.fine = (uint8_t*)p;
    return r;
}

struct IndexView {
    const uint8_t* codes;        // [rows][MP]
    const int64_t* rowids;       // [rows]
    const int64_t* cell_start;   // [ncell]
    const int64_t* lsize;        // [ncell]
};

// exact ADC distance of one code: float64, reference summation order
__device__ inline double exact_adc(const ModelView& mv, const uint8_t* code, const double* p0, const double* p1) {
    double acc = 0.0;
    for (int j = 0; j < mv.M; ++j) {
        const int s = j / mv.m;
        const double* p = (s ? p1 : p0) + (j - s * mv.m) * mv.ds;
        const double* c = mv.subs + ((int64_t)j * mv.K + code[j]) * mv.ds;
        const double e = sqdist_np<double>(p, c, mv.ds);
        acc = (j == 0) ? e : __dadd_rn(acc, e);
    }
    return acc;
}

// bitonic sort of n (power of two) entries by (dkey, pkey) ascending, payload idx
__device__ inline void bitonic_sort_dp(unsigned long long* dk, unsigned int* pk, int* idx, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = dk[i], b = dk[ixj];
                    const unsigned int pa = pk[i], pb = pk[ixj];
                    const bool gt = (a > b) || (a == b && pa > pb);
                    const bool asc = ((i & k) == 0);
                    if (gt == asc) {
                        dk[i] = b; dk[ixj] = a; pk[i] = pb; pk[ixj] = pa;
                        const int t = idx[i]; idx[i] = idx[ixj]; idx[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

#define SEL_SB 2048     // merge buffer entries
#define SEL_THREADS 256
// one block per query.  dynamic smem: keys[SEL_SB] u64 | dk[KP] u64 | pk[KP] u32 | idx[KP] int | rows[KP] i64 | vis[KP] int
__global__ void __launch_bounds__(SEL_THREADS)
k_merge_rerank(ModelView mv, IndexView ix, PlanView pv, const unsigned long long* __restrict__ partial,
               const double* __restrict__ P64, int KP, int k, double eps_rel, void* recbuf) {
    extern __shared__ __align__(16) unsigned char sm_sel[];
    unsigned long long* keys = (unsigned long long*)sm_sel;
    unsigned long long* dk = keys + SEL_SB;
    int64_t* rows = (int64_t*)(dk + KP);
    unsigned int* pk = (unsigned int*)(rows + KP);
    int* idx = (int*)(pk + KP);
    int* visv = idx + KP;
    const int q = blockIdx.x, tid = threadIdx.x;
    const int npart = pv.npart[q];
    const unsigned long long* src = partial + (size_t)pv.pbase[q] * KP;
    const int64_t total = (int64_t)npart * KP;
    RecView rv = rec_view(recbuf, pv.nq, k, mv.M);

    for (int i = tid; i < KP; i += SEL_THREADS) keys[i] = B2L_KEY_EMPTY;
    const int chunk = SEL_SB - KP;
    for (int64_t off = 0; off < total; off += chunk) {
        for (int i = tid; i < chunk; i += SEL_THREADS) {
            const int64_t e = off + i;
            keys[KP + i] = (e < total) ? src[e] : B2L_KEY_EMPTY;
        }
        __syncthreads();
        bitonic_sort_u64(keys, SEL_SB);
    }
    __syncthreads();
    // exact float64 distances of the k' survivors
    const int nv = pv.nvis[q];
    const int64_t o = (int64_t)q * pv.maxvis;
    for (int i = tid; i < KP; i += SEL_THREADS) {
        const unsigned long long key = keys[i];
        dk[i] = 0x7FF0000000000000ull;       // +inf
        pk[i] = 0xFFFFFFFFu;
        idx[i] = i;
        rows[i] = -1;
        visv[i] = -1;
        if (key != B2L_KEY_EMPTY) {
            const unsigned int pos = (unsigned int)(key & 0xFFFFFFFFull);
            int v = -1;
            for (int t = 0; t < nv; ++t) {
                if (pv.vis_pbase[o + t] >= 0) {
                    const int64_t b = pv.vis_base[o + t];
                    if ((int64_t)pos >= b && (int64_t)pos < b + ix.lsize[pv.vis_cell[o + t]]) { v = t; break; }
                }
            }
            if (v < 0) continue;                           // cannot happen: positions come from scanned cells
            const int cell = pv.vis_cell[o + v];
            const int64_t row = ix.cell_start[cell] + ((int64_t)pos - pv.vis_base[o + v]);
            const double d = exact_adc(mv, ix.codes + row * mv.MP, P64 + (int64_t)pv.vis_lut0[o + v] * mv.h,
                                       P64 + (int64_t)pv.vis_lut1[o + v] * mv.h);
            dk[i] = (unsigned long long)__double_as_longlong(d);
            pk[i] = pos;
            rows[i] = row;
            visv[i] = v;
        }
    }
    __syncthreads();
    int ncoll = 0;                                           // keys are sorted: EMPTY entries are last
    {
        int lo = 0, hi = KP;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] != B2L_KEY_EMPTY) lo = mid + 1; else hi = mid; }
        ncoll = lo;
    }
    bitonic_sort_dp(dk, pk, idx, KP);
    const int nout = min(k, ncoll);
    for (int i = tid; i < nout; i += SEL_THREADS) {
        const int sidx = idx[i];
        const int64_t row = rows[sidx];
        const int64_t e = (int64_t)q * k + i;
        rv.d64[e] = __longlong_as_double((long long)dk[i]);
        rv.pos[e] = pk[i];
        rv.rowid[e] = ix.rowids[row];
        rv.cell[e] = pv.vis_cell[o + visv[sidx]];
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = ix.codes[row * mv.MP + j];
    }
    if (tid == 0) {
        double lb = __longlong_as_double(0x7FF0000000000000ll);
        if (pv.ncand_local[q] > (int64_t)ncoll) {            // some local candidates were not collected
            const float amax = __uint_as_float((unsigned int)(keys[KP - 1] >> 32));
            lb = (double)amax * (1.0 - eps_rel) - 1e-300;
        }
        rv.lb[q] = lb;
        rv.count[q] = nout;
        rv.visited[q] = nv;
        rv.ncand[q] = pv.ncand[q];
    }
}

// Final merge over ranks: one block per query; consumes the gathered record buffers
// (rank-major, each rec_bytes(nq,k,M) long).  dynamic smem: dk[n] u64 | pk[n] u32 | idx[n] int, n = pow2 >= nranks*k
__global__ void __launch_bounds__(128)
k_final(int V, int M, const void* __restrict__ recs_all, int nranks, int nq, int k, int n,
        int64_t* __restrict__ rowid, double* __restrict__ dist, int32_t* __restrict__ coarse, uint8_t* __restrict__ fine,
        int32_t* __restrict__ count, int32_t* __restrict__ visited, uint8_t* __restrict__ certified) {
    extern __shared__ __align__(16) unsigned char sm_fin[];
    unsigned long long* dk = (unsigned long long*)sm_fin;
    unsigned int* pk = (unsigned int*)(dk + n);
    int* idx = (int*)(pk + n);
    const int q = blockIdx.x, tid = threadIdx.x;
    const size_t rb = rec_bytes(nq, k, M);
    double lbmin = __longlong_as_double(0x7FF0000000000000ll);
    int total = 0;
    for (int r = 0; r < nranks; ++r) {
        RecView rv = rec_view((unsigned char*)recs_all + rb * r, nq, k, M);
        const int c = rv.count[q];
        for (int i = tid; i < k; i += blockDim.x) {
            const int e = r * k + i;
            if (i < c) {
                dk[e] = (unsigned long long)__double_as_longlong(rv.d64[(int64_t)q * k + i]);
                pk[e] = (unsigned int)rv.pos[(int64_t)q * k + i];
            } else { dk[e] = 0x7FF0000000000000ull; pk[e] = 0xFFFFFFFFu; }
            idx[e] = e;
        }
        total += c;
        lbmin = fmin(lbmin, rv.lb[q]);
    }
    for (int e = nranks * k + tid; e < n; e += blockDim.x) { dk[e] = 0x7FF0000000000000ull; pk[e] = 0xFFFFFFFFu; idx[e] = e; }
    __syncthreads();
    bitonic_sort_dp(dk, pk, idx, n);
    const int nout = min(k, total);
    for (int i = tid; i < nout; i += blockDim.x) {
        const int e = idx[i], r = e / k, j = e % k;
        RecView rv = rec_view((unsigned char*)recs_all + rb * r, nq, k, M);
        const int64_t s = (int64_t)q * k + j, d = (int64_t)q * k + i;
        if (rowid) rowid[d] = rv.rowid[s];
        if (dist) dist[d] = rv.d64[s];
        if (coarse) { coarse[d * 2] = rv.cell[s] / V; coarse[d * 2 + 1] = rv.cell[s] % V; }
        if (fine) for (int t = 0; t < M; ++t) fine[d * M + t] = rv.fine[s * M + t];
    }
    if (tid == 0) {
        RecView r0 = rec_view((unsigned char*)recs_all, nq, k, M);
        count[q] = nout;
        if (visited) visited[q] = r0.visited[q];
        // certified: no uncollected candidate of any rank can precede the k-th result
        bool ok = true;
        if (nout == k && k > 0) ok = __longlong_as_double((long long)dk[k - 1]) < lbmin;
        else ok = !(lbmin < __longlong_as_double(0x7FF0000000000000ll));
        if (certified) certified[q] = ok ? 1 : 0;
    }
}
