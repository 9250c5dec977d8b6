// Selection kernels: selection among the scan's candidates, float64 re-rank of the survivors in the
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

// Where k_select puts the records of query q: block q / nq_home (the query's HOME rank in the multi-GPU exchange, whose
// record mailbox base[] points into -- a peer-mapped window -- at this rank's block), entry q % nq_home.  Single GPU and the
// host-driven (all-gather) protocol: nq_home = nq, base[0] = the record buffer.
struct RecRoute {
    unsigned char* base[8];
    int nq_home;
};

struct IndexView {
    const uint8_t* codes;        // [rows][MP]
    const int64_t* rowids;       // [rows]
    const int64_t* cell_start;   // [ncell]
    const int64_t* lsize;        // [ncell]
};

// exact ADC distance of one code (stored row `code`, in-cell index `incell`): float64, reference summation order
__device__ inline double exact_adc(const ModelView& mv, const uint8_t* code, int64_t incell, const double* p0, const double* p1) {
    double acc = 0.0;
    for (int j = 0; j < mv.M; ++j) {
        const int s = j / mv.m;
        const double* p = (s ? p1 : p0) + (j - s * mv.m) * mv.ds;
        const double* c = mv.subs + ((int64_t)j * mv.K + code_byte(code, incell, j, mv.SW)) * mv.ds;
        const double e = sqdist_np<double>(p, c, mv.ds);
        acc = (j == 0) ? e : __dadd_rn(acc, e);
    }
    return acc;
}

// One squared sub-distance by 8 consecutive lanes (n % 8 == 0, n <= 128; every lane of the warp must call it).  NumPy's
// pairwise leaf is eight strided accumulators r_j = sum_i term(j + 8 i) combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7));
// lane j builds r_j and an xor butterfly performs exactly those additions (IEEE addition is commutative), so the bits equal
// sqdist_np's -- with 8 loads in sequence per term instead of n, and 64-byte coalesced reads.
__device__ __forceinline__ double sqdist_np_lanes8(const double* p, const double* c, int n, int j, bool live) {
    double r = 0.0;
    if (live) {
        for (int i = j; i < n; i += 8) {
            const double t = __dsub_rn(p[i], c[i]);
            const double sq = __dmul_rn(t, t);
            r = (i == j) ? sq : __dadd_rn(r, sq);
        }
    }
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    return r;
}

// part[t] = exact sub-distance of term t = (candidate i0 + t / M, sub-quantizer t % M), t < nterms, by the whole block
__device__ __forceinline__ void exact_terms(const ModelView& mv, const IndexView& ix, const PlanView& pv, const double* __restrict__ P64,
                                            int64_t o, const int64_t* rows, const int* visv, const unsigned int* pk, int i0, int nterms,
                                            double* part) {
    const int M = mv.M, tid = threadIdx.x, nthr = blockDim.x;
    auto pointers = [&](int t, const double*& p, const double*& c) -> bool {
        const int i = i0 + t / M, j = t % M;
        if (rows[i] < 0) return false;
        const int v = visv[i];
        const int64_t incell = (int64_t)pk[i] - pv.vis_base[o + v];
        const int s = j / mv.m;
        p = P64 + (int64_t)(s ? pv.vis_lut1[o + v] : pv.vis_lut0[o + v]) * mv.h + (j - s * mv.m) * mv.ds;
        c = mv.subs + ((int64_t)j * mv.K + code_byte(ix.codes + rows[i] * mv.MP, incell, j, mv.SW)) * mv.ds;
        return true;
    };
    if ((mv.ds & 7) == 0 && mv.ds >= 16 && mv.ds <= 128) {                   // (at ds = 8 a term is 8 loads: one thread does it faster)
        const int g8 = tid >> 3, j8 = tid & 7, ng = nthr >> 3;               // 8 lanes per term
        for (int t0 = 0; t0 < nterms; t0 += ng) {                            // (warp-uniform trip count: the shuffles need every lane)
            const int t = t0 + g8;
            const double* p = nullptr; const double* c = nullptr;
            const bool live = t < nterms && pointers(t, p, c);
            const double d = sqdist_np_lanes8(p, c, mv.ds, j8, live);
            if (live && j8 == 0) part[t] = d;
        }
    } else {
        for (int t = tid; t < nterms; t += nthr) {
            const double* p; const double* c;
            if (pointers(t, p, c)) part[t] = sqdist_np<double>(p, c, mv.ds);
        }
    }
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

#define SEL_HB 1024       // histogram buckets of the selection
#define SEL_LIST 1024     // capacity of the boundary list that is actually sorted (>= KP)
#define SEL_PART 2048     // float64 partial sums staged per chunk of candidates
#define SEL_THREADS 128
__host__ __device__ inline size_t select_smem_bytes(int KP) {
    return (size_t)SEL_LIST * 8 + (size_t)KP * 28 + (size_t)SEL_PART * 8 + (size_t)SEL_HB * 4 + 64;
}

// histogram bucket of a key (monotone non-decreasing in the key).  Float32 scan: keys are float bits; packed scan:
// keys are integer code sums, `lo_bits` / `scale` then hold the integer minimum and SEL_HB / (range + 1).
__device__ __forceinline__ int sel_bucket(unsigned int dbits, unsigned int lo_bits, float scale, int packed) {
    const float x = packed ? (float)(dbits - lo_bits) : (__uint_as_float(dbits) - __uint_as_float(lo_bits));
    return min(SEL_HB - 1, (int)(x * scale));
}

// one block per query.  The scan appended every candidate whose float32 distance is <= the query's final bound
// gthr[q] (and at least KP of them are).  Select the KP smallest without sorting them all: a 1024-bucket histogram
// over [min, max] of the passing distances locates the bucket holding the KP-th smallest; only candidates up to that
// bucket are sorted (bitonic, by (dist32, retrieval position)).  The KP best are re-evaluated in float64 in the
// reference's summation order, ordered by (dist64, retrieval position), and the first k emitted with the
// certification bound.
// dynamic smem: list[SEL_LIST] u64 | dk[KP] u64 | rows[KP] i64 | part[SEL_PART] f64 | pk[KP] u32 | idx[KP] int | vis[KP] int | hist[SEL_HB] u32
__global__ void __launch_bounds__(SEL_THREADS, 8)
k_select(ModelView mv, IndexView ix, PlanView pv, const unsigned long long* __restrict__ cand,
         const unsigned int* __restrict__ cand_cnt, const unsigned int* __restrict__ gthr, int cand_cap,
         const double* __restrict__ P64, int KP, int k, double eps_rel, RecRoute route,
         int packed, const double* __restrict__ qB, const double* __restrict__ qDelta, const double* __restrict__ qSlack,
         uint8_t* __restrict__ need2, const unsigned int* __restrict__ qmargin) {
    extern __shared__ __align__(16) unsigned char sm_sel[];
    unsigned long long* keys = (unsigned long long*)sm_sel;
    unsigned long long* dk = keys + SEL_LIST;
    int64_t* rows = (int64_t*)(dk + KP);
    double* part = (double*)(rows + KP);
    unsigned int* pk = (unsigned int*)(part + SEL_PART);
    int* idx = (int*)(pk + KP);
    int* visv = idx + KP;
    unsigned int* hist = (unsigned int*)(visv + KP);
    __shared__ unsigned int s_red[3][SEL_THREADS / 32];
    __shared__ unsigned int s_scan[SEL_THREADS / 32];
    __shared__ int s_n, s_bstar;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int qh = q / route.nq_home, ql = q - qh * route.nq_home;       // home block, index inside it
    RecView rv = rec_view(route.base[qh], route.nq_home, k, mv.M);
    if (pv.ncand_local[q] == 0) {                     // nothing of this query is stored on this rank (uniform exit)
        if (tid == 0) {
            rv.lb[ql] = __longlong_as_double(0x7FF0000000000000ll);
            rv.count[ql] = 0; rv.visited[ql] = pv.nvis[q]; rv.ncand[ql] = pv.ncand[q];
            if (need2) need2[q] = 0;
        }
        return;
    }
    const unsigned int appended = cand_cnt[q];
    const int n = (int)min(appended, (unsigned int)cand_cap);
    const unsigned int bound = gthr[q] + ((packed && qmargin) ? qmargin[q] : 0u);     // what the scan appended up to
    const unsigned long long* src = cand + (size_t)q * cand_cap;

    // ---- pass 1: range and count of the passing candidates
    unsigned int lo = 0xFFFFFFFFu, hi = 0u, cnt = 0u;
    for (int i = tid; i < n; i += SEL_THREADS) {
        const unsigned int d = (unsigned int)(src[i] >> 32);
        if (d <= bound) { lo = min(lo, d); hi = max(hi, d); ++cnt; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { s_red[0][wid] = lo; s_red[1][wid] = hi; s_red[2][wid] = cnt; }
    for (int i = tid; i < SEL_HB; i += SEL_THREADS) hist[i] = 0u;
    if (tid == 0) s_n = 0;
    __syncthreads();
    lo = 0xFFFFFFFFu; hi = 0u; cnt = 0u;
    for (int w = 0; w < SEL_THREADS / 32; ++w) { lo = min(lo, s_red[0][w]); hi = max(hi, s_red[1][w]); cnt += s_red[2][w]; }
    const int npass = (int)cnt;
    const float flo = __uint_as_float(lo), fhi = __uint_as_float(hi);
    float scale = 0.0f;
    if (npass > 0 && hi > lo) scale = packed ? (float)(SEL_HB - 1) / (float)(hi - lo) : (float)(SEL_HB - 1) / (fhi - flo);
    const int want = min(KP, npass);

    // ---- pass 2: histogram, then the bucket b* where the running count reaches `want`
    for (int i = tid; i < n; i += SEL_THREADS) {
        const unsigned int d = (unsigned int)(src[i] >> 32);
        if (d <= bound) atomicAdd(&hist[sel_bucket(d, lo, scale, packed)], 1u);
    }
    __syncthreads();
    {
        constexpr int PER = SEL_HB / SEL_THREADS;
        unsigned int loc[PER], sum = 0;
#pragma unroll
        for (int e = 0; e < PER; ++e) { loc[e] = hist[tid * PER + e]; sum += loc[e]; }
        unsigned int inc = sum;
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        unsigned int base = 0;
        for (int w = 0; w < wid; ++w) base += s_scan[w];
        unsigned int run = base + inc - sum;                       // exclusive prefix of this thread's first bucket
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            if (run < (unsigned)want && run + loc[e] >= (unsigned)want) s_bstar = tid * PER + e;
            run += loc[e];
        }
        if (want == 0 && tid == 0) s_bstar = -1;
    }
    __syncthreads();
    const int bstar = s_bstar;

    // ---- pass 3: gather the candidates up to bucket b*, sort them
    for (int i = tid; i < n; i += SEL_THREADS) {
        const unsigned long long key = src[i];
        const unsigned int d = (unsigned int)(key >> 32);
        if (d <= bound && sel_bucket(d, lo, scale, packed) <= bstar) {
            const int j = atomicAdd(&s_n, 1);
            if (j < SEL_LIST) keys[j] = key;
        }
    }
    __syncthreads();
    const int nl = s_n;
    const bool lost = appended > (unsigned int)cand_cap || nl > SEL_LIST;      // some candidate could not be kept
    const int nk = min(nl, SEL_LIST);
    const int np2 = max(KP, next_pow2_dev(nk));
    for (int i = nk + tid; i < np2; i += SEL_THREADS) keys[i] = B2L_KEY_EMPTY;
    __syncthreads();
    bitonic_sort_u64(keys, np2);

    // ---- exact float64 distances of the KP best: locate the rows, then one (candidate, sub-quantizer) term per thread
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
            pk[i] = pos;
            rows[i] = ix.cell_start[pv.vis_cell[o + v]] + ((int64_t)pos - pv.vis_base[o + v]);
            visv[i] = v;
        }
    }
    __syncthreads();
    const int M = mv.M, CH = SEL_PART / M;
    for (int i0 = 0; i0 < KP; i0 += CH) {
        const int nc = min(CH, KP - i0);
        exact_terms(mv, ix, pv, P64, o, rows, visv, pk, i0, nc * M, part);
        __syncthreads();
        for (int i = i0 + tid; i < i0 + nc; i += SEL_THREADS) {
            if (rows[i] >= 0) {
                const double* e = part + (i - i0) * M;
                double acc = e[0];
                for (int j = 1; j < M; ++j) acc = __dadd_rn(acc, e[j]);      // left-to-right, search.py:173
                dk[i] = (unsigned long long)__double_as_longlong(acc);
            }
        }
        __syncthreads();
    }
    const int ncoll = min(nk, KP);
    bitonic_sort_dp(dk, pk, idx, KP);
    const int nout = min(k, ncoll);
    for (int i = tid; i < nout; i += SEL_THREADS) {
        const int sidx = idx[i];
        const int64_t row = rows[sidx];
        const int v = visv[sidx];
        const int64_t e = (int64_t)ql * k + i;
        rv.d64[e] = __longlong_as_double((long long)dk[i]);
        rv.pos[e] = pk[i];
        rv.rowid[e] = ix.rowids[row];
        rv.cell[e] = pv.vis_cell[o + v];
        const int64_t incell = (int64_t)pk[i] - pv.vis_base[o + v];
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(ix.codes + row * mv.MP, incell, j, mv.SW);
    }
    if (tid == 0) {
        double lb = __longlong_as_double(0x7FF0000000000000ll);
        if (lost) lb = -1.0;                                  // never certified: exact re-rank
        else if (pv.ncand_local[q] > (int64_t)ncoll) {        // some local candidates are not among the KP collected
            const unsigned int kb = (unsigned int)(keys[KP - 1] >> 32);
            if (packed) lb = (qB[q] + qDelta[q] * ((double)kb - 1.0) - qSlack[q]) * (1.0 - 1e-6) - 1e-300;      // see plan.cuh (k_lut_quant)
            else lb = (double)__uint_as_float(kb) * (1.0 - eps_rel) - 1e-300;
        }
        rv.lb[ql] = lb;
        rv.count[ql] = nout;
        rv.visited[ql] = nv;
        rv.ncand[ql] = pv.ncand[q];
        // Locally not certifiable from the KP best (near-ties between the k-th and the KP-th candidate: concentrated
        // distances) but nothing was lost: k_select2 re-ranks EVERY appended candidate of the query exactly, which moves the
        // bound of "everything not looked at" from the KP-th best out to the scan's final bound.
        if (need2) {
            const bool okl = (nout == k) ? (__longlong_as_double((long long)dk[k - 1]) < lb)
                                         : !(lb < __longlong_as_double(0x7FF0000000000000ll));
            need2[q] = (!okl && !lost) ? 1 : 0;
        }
    }
}

// Second chance for the queries k_select flagged (a handful per batch, or none: every other block leaves at once).
// All candidates the scan appended with key <= the final bound (at most SEL2_LIST, else the query stays as it is and goes
// down the fallback chain) are evaluated in float64 and sorted by (dist64, retrieval position); every code that was NOT
// appended has a key above the bound, so the certification bound is that of the bound itself.
// dynamic smem: dk[SEL2_LIST] u64 | rows[SEL2_LIST] i64 | part[SEL_PART] f64 | pk[SEL2_LIST] u32 | idx[SEL2_LIST] int | vis[SEL2_LIST] int
#define SEL2_LIST 2048
__host__ __device__ inline size_t select2_smem_bytes() { return (size_t)SEL2_LIST * 28 + (size_t)SEL_PART * 8 + 64; }

__global__ void __launch_bounds__(256)
k_select2(ModelView mv, IndexView ix, PlanView pv, const unsigned long long* __restrict__ cand,
          const unsigned int* __restrict__ cand_cnt, const unsigned int* __restrict__ gthr, int cand_cap,
          const double* __restrict__ P64, int k, double eps_rel, RecRoute route,
          int packed, const double* __restrict__ qB, const double* __restrict__ qDelta, const double* __restrict__ qSlack,
          const uint8_t* __restrict__ need2, const unsigned int* __restrict__ qmargin) {
    const int q = blockIdx.x, tid = threadIdx.x;
    if (!need2[q]) return;
    extern __shared__ __align__(16) unsigned char sm_sel2[];
    unsigned long long* dk = (unsigned long long*)sm_sel2;
    int64_t* rows = (int64_t*)(dk + SEL2_LIST);
    double* part = (double*)(rows + SEL2_LIST);
    unsigned int* pk = (unsigned int*)(part + SEL_PART);
    int* idx = (int*)(pk + SEL2_LIST);
    int* visv = idx + SEL2_LIST;
    __shared__ int s_n;
    const int qh = q / route.nq_home, ql = q - qh * route.nq_home;
    RecView rv = rec_view(route.base[qh], route.nq_home, k, mv.M);
    const unsigned int appended = cand_cnt[q];
    if (appended > (unsigned int)cand_cap) return;                    // (k_select already marked it lost)
    const int n = (int)appended;
    const unsigned int bound = gthr[q] + ((packed && qmargin) ? qmargin[q] : 0u);
    const unsigned long long* src = cand + (size_t)q * cand_cap;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long key = src[i];
        if ((unsigned int)(key >> 32) <= bound) {
            const int j = atomicAdd(&s_n, 1);
            if (j < SEL2_LIST) { pk[j] = (unsigned int)(key & 0xFFFFFFFFull); }
        }
    }
    __syncthreads();
    const int nl = s_n;
    if (nl > SEL2_LIST || nl < 1) return;                             // too many to take here: the fallback chain handles it
    int np2 = 1;
    while (np2 < nl) np2 <<= 1;
    const int nv = pv.nvis[q];
    const int64_t o = (int64_t)q * pv.maxvis;
    for (int i = tid; i < np2; i += blockDim.x) {
        dk[i] = 0x7FF0000000000000ull;
        idx[i] = i;
        rows[i] = -1;
        visv[i] = -1;
        if (i >= nl) { pk[i] = 0xFFFFFFFFu; continue; }
        const unsigned int pos = pk[i];
        int v = -1;
        for (int t = 0; t < nv; ++t) {
            if (pv.vis_pbase[o + t] >= 0) {
                const int64_t b = pv.vis_base[o + t];
                if ((int64_t)pos >= b && (int64_t)pos < b + ix.lsize[pv.vis_cell[o + t]]) { v = t; break; }
            }
        }
        if (v < 0) { pk[i] = 0xFFFFFFFFu; continue; }
        rows[i] = ix.cell_start[pv.vis_cell[o + v]] + ((int64_t)pos - pv.vis_base[o + v]);
        visv[i] = v;
    }
    __syncthreads();
    const int M = mv.M, CH = SEL_PART / M;
    for (int i0 = 0; i0 < nl; i0 += CH) {
        const int nc = min(CH, nl - i0);
        exact_terms(mv, ix, pv, P64, o, rows, visv, pk, i0, nc * M, part);
        __syncthreads();
        for (int i = i0 + tid; i < i0 + nc; i += blockDim.x) {
            if (rows[i] >= 0) {
                const double* e = part + (i - i0) * M;
                double acc = e[0];
                for (int j = 1; j < M; ++j) acc = __dadd_rn(acc, e[j]);
                dk[i] = (unsigned long long)__double_as_longlong(acc);
            }
        }
        __syncthreads();
    }
    bitonic_sort_dp(dk, pk, idx, np2);
    const int nout = min(k, nl);
    for (int i = tid; i < nout; i += blockDim.x) {
        const int sidx = idx[i];
        const int64_t row = rows[sidx];
        const int v = visv[sidx];
        const int64_t e = (int64_t)ql * k + i;
        rv.d64[e] = __longlong_as_double((long long)dk[i]);
        rv.pos[e] = pk[i];
        rv.rowid[e] = ix.rowids[row];
        rv.cell[e] = pv.vis_cell[o + v];
        const int64_t incell = (int64_t)pk[i] - pv.vis_base[o + v];
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(ix.codes + row * mv.MP, incell, j, mv.SW);
    }
    if (tid == 0) {
        double lb = __longlong_as_double(0x7FF0000000000000ll);
        if (pv.ncand_local[q] > (int64_t)nl) {                        // codes that were never appended: key >= bound + 1 (packed) / > bound
            if (packed) lb = (qB[q] + qDelta[q] * (double)bound - qSlack[q]) * (1.0 - 1e-6) - 1e-300;
            else lb = (double)__uint_as_float(bound) * (1.0 - eps_rel) - 1e-300;
        }
        rv.lb[ql] = lb;
        rv.count[ql] = nout;
    }
}

// Final merge over ranks: one block per query; consumes the gathered record buffers
// (rank-major, each rec_bytes(nq,k,M) long).  dynamic smem: dk[n] u64 | pk[n] u32 | idx[n] int, n = pow2 >= nranks*k
__global__ void __launch_bounds__(128)
k_final(int V, int M, const void* __restrict__ recs_all, int nranks, int nq, int k, int n,
        int64_t* __restrict__ rowid, double* __restrict__ dist, int32_t* __restrict__ coarse, uint8_t* __restrict__ fine,
        int32_t* __restrict__ count, int32_t* __restrict__ visited, uint8_t* __restrict__ certified,
        unsigned int* __restrict__ n_uncertified = nullptr, int force_unc = 0) {
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
    for (int i = nout + tid; i < k; i += blockDim.x) {           // rows beyond the count come back zero-filled
        const int64_t d = (int64_t)q * k + i;
        if (rowid) rowid[d] = 0;
        if (dist) dist[d] = 0.0;
        if (coarse) { coarse[d * 2] = 0; coarse[d * 2 + 1] = 0; }
        if (fine) for (int t = 0; t < M; ++t) fine[d * M + t] = 0;
    }
    if (tid == 0) {
        RecView r0 = rec_view((unsigned char*)recs_all, nq, k, M);
        count[q] = nout;
        if (visited) visited[q] = r0.visited[q];
        // certified: no uncollected candidate of any rank can precede the k-th result
        bool ok = true;
        if (nout == k && k > 0) ok = __longlong_as_double((long long)dk[k - 1]) < lbmin;
        else ok = !(lbmin < __longlong_as_double(0x7FF0000000000000ll));
        if (force_unc && q % 5 == 0) ok = false;                  // test knob (b2l_debug_force_redo bit 2): exercise the fallback chain
        if (certified) certified[q] = ok ? 1 : 0;
        if (!ok && n_uncertified) atomicAdd(n_uncertified, 1u);
    }
}
