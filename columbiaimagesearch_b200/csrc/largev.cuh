// Large-V multi-index search (the product's shipped configurations: V = 2048 / 4096, M = 8, PCA to 128 / 256-d,
// conf/conf_search_*_release.json:12-16).  V*V = 4M .. 16M cells, almost all of them empty, so nothing here is dense in V*V:
//
//   * sparse cell directory: rows sorted by cell id; run-length encoding gives the non-empty cells `ucell[nu]` with their
//     row ranges `ustart[nu+1]` (no padding between cells), and an open-addressing hash cell -> run index;
//   * k_walk (one block per query): search.multisequence (search.py:13-82) + get_result_quota (search.py:110-135) without
//     a heap.  The reference pops cells in the order (d0[i0] + d1[i1], (i0, i1)), i* = ranks in the per-split argsort
//     (heap ties are broken by the index tuple), so the traversal IS the sorted order of that key.  The block walks it in
//     distance slabs: an exponential + bisection search on the slab's upper edge T picks the next <= WALK_CAP cells (every
//     row i0 contributes the run of i1 with key <= T, found by binary search in the sorted d1), the slab is sorted by
//     (distance, i0, i1), the sizes of its cells come from the hash directory, and a prefix sum finds where the quota is
//     reached.  `visited` counts empty cells too, as the reference does;
//   * every distinct (split, coarse code) among a query's non-empty visited cells gets one projection slot; the
//     projections are one grouped float64 tensor-core GEMM (k_rotate_dmma_g, MODE 1);
//   * k_cand_dist: the ADC distance of EVERY retrieved code directly in float64, in the reference's summation order
//     (sub-distance in NumPy pairwise order, left-to-right sum over the M sub-quantizers, search.py:173) -- with tiny
//     cells and hundreds of distinct coarse codes per query a 256-entry table per coarse code would serve a handful of
//     codes, so the few entries needed are evaluated on the fly; no quantised tables, nothing to certify;
//   * a stable segmented radix sort by the float64 distance keeps ties in retrieval order (search.py:210), and
//     k_emit_sorted writes the first k of every query as records.
#pragma once
#include "common.cuh"
#include "select.cuh"

#define WALK_THREADS 256
#define WALK_CAP 2048             // cells per distance slab (power of two)
#define B2L_MAX_V_SPARSE 4096

struct SparseDir {
    const unsigned int* hkeys;    // [hmask + 1] cell id or 0xFFFFFFFF
    const unsigned int* hvals;    // [hmask + 1] run index
    unsigned int hmask;
    const unsigned int* ustart;   // [nu + 1] first row of the run
    const unsigned int* ucell;    // [nu]
    unsigned int nu;
};

__device__ __forceinline__ unsigned int hash_cell(unsigned int c) { return (c * 2654435761u) ^ (c >> 15); }

__global__ void k_dir_build(const unsigned int* __restrict__ ucell, unsigned int nu, unsigned int* __restrict__ hkeys,
                            unsigned int* __restrict__ hvals, unsigned int hmask) {
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < nu; r += gridDim.x * blockDim.x) {
        const unsigned int c = ucell[r];
        unsigned int p = hash_cell(c) & hmask;
        while (true) {
            const unsigned int old = atomicCAS(&hkeys[p], 0xFFFFFFFFu, c);
            if (old == 0xFFFFFFFFu || old == c) { hvals[p] = r; break; }
            p = (p + 1) & hmask;
        }
    }
}

__device__ __forceinline__ int dir_lookup(const SparseDir& d, unsigned int c) {
    unsigned int p = hash_cell(c) & d.hmask;
    while (true) {
        const unsigned int k = d.hkeys[p];
        if (k == c) return (int)d.hvals[p];
        if (k == 0xFFFFFFFFu) return -1;
        p = (p + 1) & d.hmask;
    }
}

// exclusive prefix of the run lengths -> ustart[nu + 1] (single block; nu up to a few million: strided serial chunks)
__global__ void __launch_bounds__(1024) k_run_starts(const unsigned int* __restrict__ counts, unsigned int nu, unsigned int* __restrict__ ustart) {
    __shared__ unsigned long long part[1024];
    const unsigned int per = (nu + 1023) / 1024, a = threadIdx.x * per, b = min(nu, a + per);
    unsigned long long s = 0;
    for (unsigned int i = a; i < b; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned long long v = 0;
        if ((int)threadIdx.x >= o) v = part[threadIdx.x - o];
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - s;
    for (unsigned int i = a; i < b; ++i) { ustart[i] = (unsigned int)run; run += counts[i]; }
    if (threadIdx.x == 1023) ustart[nu] = (unsigned int)part[1023];
}

// rows of the sorted order -> the cell-contiguous layout (one thread per run; runs are short at large V)
__global__ void k_scatter_runs(const unsigned int* __restrict__ sorted_src, const unsigned int* __restrict__ ustart, unsigned int nu,
                               const uint8_t* __restrict__ fine_in, const int64_t* __restrict__ rowid_in, int M, int MP, int SW,
                               uint8_t* __restrict__ codes, int64_t* __restrict__ rowids) {
    for (unsigned int r = blockIdx.x * blockDim.x + threadIdx.x; r < nu; r += gridDim.x * blockDim.x) {
        const unsigned int a = ustart[r], b = ustart[r + 1];
        for (unsigned int i = a; i < b; ++i) {
            const int64_t src = sorted_src[i];
            const int sw = (int)(i - a) & SW;
            for (int s = 0; s < MP; ++s) {
                const int j = s ^ sw;
                codes[(int64_t)i * MP + s] = (j < M) ? fine_in[src * M + j] : (uint8_t)0;
            }
            rowids[i] = rowid_in[src];
        }
    }
}

struct WalkCounters {
    unsigned int n_lut;               // projection slots (query, split, coarse code)
    unsigned int err;                 // 1: a distance slab could not be bounded (mass tie of coarse distances); 2: segment list overflow
    unsigned long long cand_total;    // retrieved codes, all queries
    unsigned int presel_fallback;     // queries of the current group whose float32 preselection kept too many candidates
    unsigned int pad_;
};

struct WalkView {                     // per-batch arrays of the large-V plan
    int nq, segcap;
    int32_t* nvis;                    // [nq] cells visited (incl. empty)
    int32_t* nseg;                    // [nq] non-empty visited cells
    unsigned int* ncand;              // [nq] retrieved codes
    uint4* seg;                       // [nq][segcap] (first row, size, retrieval position of the first code, cell id)
    int32_t* slot0;                   // [nq][V] projection slot of (split 0, c) (defined for the codes the query uses)
    int32_t* slot1;                   // [nq][V]
    int32_t* lut_desc;                // [cap][3] (q, split, c)
    WalkCounters* cnt;
    // the traversal itself (b2l_cell_order at large V): the first viscap visited cells and their distances; NULL otherwise
    int32_t* vis_cells;               // [nq][viscap] c0 * V + c1
    double* vis_dists;                // [nq][viscap]
    int viscap;
    long long max_visit;              // stop after this many cells (>= 1), whatever the quota
};

// 64-bit key + 32-bit tag, ascending by (key, tag); n power of two
__device__ inline void bitonic_sort_kt(unsigned long long* dk, unsigned int* tg, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = dk[i], b = dk[ixj];
                    const unsigned int ta = tg[i], tb = tg[ixj];
                    const bool gt = (a > b) || (a == b && ta > tb);
                    if (gt == ((i & k) == 0)) { dk[i] = b; dk[ixj] = a; tg[i] = tb; tg[ixj] = ta; }
                }
            }
            __syncthreads();
        }
    }
}

// block-wide inclusive scan of v[0..n) (n <= WALK_CAP = 8 * WALK_THREADS), in place; returns the total
__device__ inline unsigned int block_scan_inplace(unsigned int* v, int n, unsigned int* s_red) {
    constexpr int PER = WALK_CAP / WALK_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned int loc[PER], sum = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) { const int i = tid * PER + e; loc[e] = i < n ? v[i] : 0u; sum += loc[e]; }
    unsigned int inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) s_red[wid] = inc;
    __syncthreads();
    unsigned int base = 0, total = 0;
    for (int w = 0; w < WALK_THREADS / 32; ++w) { if (w < wid) base += s_red[w]; total += s_red[w]; }
    unsigned int run = base + inc - sum;
#pragma unroll
    for (int e = 0; e < PER; ++e) { const int i = tid * PER + e; run += loc[e]; if (i < n) v[i] = run; }
    __syncthreads();
    return total;
}

// dynamic smem: ds[2][V] f64 | scratch (argsort: keys[np2] u64 + tags[np2] u32; slabs: keys[CAP] u64, tags, sizes, flags [CAP] u32)
//               | ord[2][V] u16 | pc[V] u16 | used[2][V/32+1] u32
__host__ __device__ inline size_t walk_scratch_bytes(int V) {
    int np2 = 1;
    while (np2 < V) np2 <<= 1;
    const size_t a = (size_t)np2 * 12, b = (size_t)WALK_CAP * 20;
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t walk_smem_bytes(int V) {
    return (size_t)2 * V * 8 + walk_scratch_bytes(V) + (size_t)2 * V * 2 + (size_t)(V + 8) * 2 + (size_t)2 * (V / 32 + 1) * 4 + 64;
}

template <typename XT>
__global__ void __launch_bounds__(WALK_THREADS)
k_walk(ModelView mv, const XT* __restrict__ Xq, int64_t quota, SparseDir dir, WalkView wv) {
    extern __shared__ __align__(16) unsigned char sm_w[];
    const int V = mv.V, h = mv.h;
    int np2 = 1;
    while (np2 < V) np2 <<= 1;
    double* ds = (double*)sm_w;                                       // [2][V] coarse distances, sorted ascending
    unsigned char* scratch = (unsigned char*)(ds + 2 * V);
    unsigned long long* ak = (unsigned long long*)scratch;            // argsort keys [np2]
    unsigned int* at = (unsigned int*)(ak + np2);                     // argsort tags [np2]
    unsigned long long* sk = (unsigned long long*)scratch;            // slab keys [CAP] (distance bits, later the run index)
    unsigned int* st = (unsigned int*)(sk + WALK_CAP);                // slab tags [CAP] (i0 << 12 | i1, later the cell id)
    unsigned int* ssz = st + WALK_CAP;                                // slab sizes -> inclusive prefix
    unsigned int* sfl = ssz + WALK_CAP;                               // slab non-empty flags -> inclusive prefix
    unsigned short* ord = (unsigned short*)(scratch + walk_scratch_bytes(V));   // [2][V] argsort
    unsigned short* pc = ord + 2 * V;                                 // [V] popped prefix of row i0
    unsigned int* used = (unsigned int*)(pc + V + 8 - (V & 7 ? 0 : 0));
    used = (unsigned int*)(((uintptr_t)used + 3) & ~(uintptr_t)3);    // [2][V/32+1] coarse codes with a non-empty visited cell
    __shared__ unsigned int s_red[WALK_THREADS / 32];
    __shared__ unsigned int s_cnt, s_cut, s_any;
    __shared__ unsigned long long s_tminbits;
    const int q = blockIdx.x, tid = threadIdx.x;
    const XT* x = Xq + (int64_t)q * mv.D;
    const bool f32 = (sizeof(XT) == 4) && mv.coarse_f32;
    const int UW = V / 32 + 1;

    // ---- coarse distances (search.py:37-39) and their argsort (stable: ties by cluster index)
    for (int s = 0; s < 2; ++s) {
        for (int i = tid; i < np2; i += WALK_THREADS) {
            unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
            if (i < V) {
                const double* C = mv.Cs + ((int64_t)s * V + i) * h;
                const double d = f32 ? (double)sqdist_np<float>(x + s * h, C, h) : sqdist_np<double>(x + s * h, C, h);
                key = (unsigned long long)__double_as_longlong(d);    // distances are >= 0: the bit pattern orders like the value
            }
            ak[i] = key;
            at[i] = (unsigned int)i;
        }
        __syncthreads();
        bitonic_sort_kt(ak, at, np2);
        for (int i = tid; i < V; i += WALK_THREADS) {
            ds[s * V + i] = __longlong_as_double((long long)ak[i]);
            ord[s * V + i] = (unsigned short)at[i];
        }
        __syncthreads();
    }
    for (int i = tid; i < V; i += WALK_THREADS) pc[i] = 0;
    for (int i = tid; i < 2 * UW; i += WALK_THREADS) used[i] = 0u;
    __syncthreads();

    const double* d0 = ds;
    const double* d1 = ds + V;
    auto celld = [&](int i0, int i1) -> double {                      // search.py:52-57: 0 + d0 + d1 in the compute type
        return f32 ? (double)__fadd_rn((float)d0[i0], (float)d1[i1]) : __dadd_rn(d0[i0], d1[i1]);
    };
    auto upper = [&](int i0, double T) -> int {                       // #i1 with key(i0, i1) <= T, at least pc[i0]
        int lo = pc[i0], hi = V;                                      // key(i0, i1) <= T for i1 < lo (popped), search in [lo, V]
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (celld(i0, mid) <= T) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    auto count_le = [&](double T) -> unsigned int {                   // cells not yet popped with key <= T (block-wide)
        if (tid == 0) s_cnt = 0u;
        __syncthreads();
        unsigned int c = 0;
        for (int i0 = tid; i0 < V; i0 += WALK_THREADS) {
            if (celld(i0, 0) > T) break;                              // rows are sorted by d0: no later row of this thread qualifies
            c += (unsigned int)(upper(i0, T) - pc[i0]);
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
        __syncthreads();
        const unsigned int r = s_cnt;
        __syncthreads();
        return r;
    };

    const double Tmax = celld(V - 1, V - 1);
    unsigned long long got = 0;                                       // retrieved codes so far (same value in every thread)
    unsigned int nvis = 0, nseg = 0;
    double step0 = 0.0;
    bool done = false;
    uint4* segq = wv.seg + (size_t)q * wv.segcap;
    while (!done) {
        // ---- the smallest key not yet popped
        if (tid == 0) { s_tminbits = 0xFFFFFFFFFFFFFFFFull; }
        __syncthreads();
        {
            unsigned long long mn = 0xFFFFFFFFFFFFFFFFull;
            for (int i0 = tid; i0 < V; i0 += WALK_THREADS) {
                const int p = pc[i0];
                if (p < V) mn = min(mn, (unsigned long long)__double_as_longlong(celld(i0, p)));
                if (p == 0) break;                                    // rows after the first untouched one only hold larger keys
            }
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if ((tid & 31) == 0) atomicMin(&s_tminbits, mn);
        }
        __syncthreads();
        const unsigned long long tmb = s_tminbits;
        __syncthreads();
        if (tmb == 0xFFFFFFFFFFFFFFFFull) break;                      // every cell has been popped
        const double Tmin = __longlong_as_double((long long)tmb);
        // ---- upper edge of the slab: count(T) in [WALK_CAP/2, WALK_CAP] where possible
        double lo = Tmin;
        unsigned int clo = count_le(lo);
        if (clo > WALK_CAP) { if (tid == 0) atomicExch(&wv.cnt->err, 1u); break; }
        if (clo < WALK_CAP / 2 && lo < Tmax) {
            double step = step0 > 0.0 ? step0 : fmax(Tmin * 1e-3, 1e-12);
            double hi = fmin(Tmax, lo + step);
            bool bracket = false;
            for (int it = 0; it < 200; ++it) {
                if (!(hi > lo)) { hi = Tmax; }
                const unsigned int c = count_le(hi);
                if (c > WALK_CAP) { bracket = true; break; }
                lo = hi; clo = c;
                if (c >= WALK_CAP / 2 || hi >= Tmax) break;
                step *= 2.0;
                hi = fmin(Tmax, lo + step);
            }
            if (bracket) {
                for (int it = 0; it < 80; ++it) {
                    const double mid = lo + (hi - lo) * 0.5;
                    if (!(mid > lo) || !(mid < hi)) break;
                    const unsigned int c = count_le(mid);
                    if (c > WALK_CAP) hi = mid;
                    else { lo = mid; clo = c; if (c >= WALK_CAP / 2) break; }
                }
            }
        }
        const double T = lo;
        const int n = (int)clo;
        step0 = fmax(T - Tmin, step0 * 0.5);
        // ---- gather the slab: per-thread counts -> exclusive offsets -> (key, tag) pairs
        {
            unsigned int c = 0;
            for (int i0 = tid; i0 < V; i0 += WALK_THREADS) {
                if (celld(i0, 0) > T) break;
                c += (unsigned int)(upper(i0, T) - pc[i0]);
            }
            // exclusive scan over the threads (8 warps)
            const int lane = tid & 31, wid = tid >> 5;
            unsigned int inc = c;
            for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) s_red[wid] = inc;
            __syncthreads();
            unsigned int off = inc - c;
            for (int w = 0; w < wid; ++w) off += s_red[w];
            __syncthreads();
            for (int i0 = tid; i0 < V; i0 += WALK_THREADS) {
                if (celld(i0, 0) > T) break;
                const int a = pc[i0], b = upper(i0, T);
                for (int i1 = a; i1 < b; ++i1) {
                    sk[off] = (unsigned long long)__double_as_longlong(celld(i0, i1));
                    st[off] = ((unsigned int)i0 << 12) | (unsigned int)i1;
                    ++off;
                }
                pc[i0] = (unsigned short)b;                           // (read by nobody else until the next barrier)
            }
        }
        int npad = 1;
        while (npad < n) npad <<= 1;
        for (int i = n + tid; i < npad; i += WALK_THREADS) { sk[i] = 0xFFFFFFFFFFFFFFFFull; st[i] = 0xFFFFFFFFu; }
        __syncthreads();
        bitonic_sort_kt(sk, st, npad);
        // ---- cells of the slab in visit order: directory look-up
        for (int i = tid; i < n; i += WALK_THREADS) {
            const unsigned int tg = st[i];
            const int c0 = ord[tg >> 12], c1 = ord[V + (tg & 4095u)];
            const unsigned int cell = (unsigned int)c0 * (unsigned int)V + (unsigned int)c1;
            const int r = dir_lookup(dir, cell);
            const unsigned int sz = r >= 0 ? dir.ustart[r + 1] - dir.ustart[r] : 0u;
            if (wv.vis_cells && (long long)nvis + i < (long long)wv.viscap) {
                wv.vis_cells[(size_t)q * wv.viscap + nvis + i] = (int32_t)cell;
                wv.vis_dists[(size_t)q * wv.viscap + nvis + i] = __longlong_as_double((long long)sk[i]);
            }
            st[i] = cell;
            sk[i] = (unsigned long long)(unsigned int)r;
            ssz[i] = sz;
            sfl[i] = sz ? 1u : 0u;
        }
        if (tid == 0) { s_cut = 0xFFFFFFFFu; }
        __syncthreads();
        block_scan_inplace(ssz, n, s_red);
        block_scan_inplace(sfl, n, s_red);
        // ---- quota cut (search.py:128-133): the first cell at which the retrieved count reaches the quota is the last one
        for (int i = tid; i < n; i += WALK_THREADS) {
            if ((long long)(got + ssz[i]) >= quota) { atomicMin(&s_cut, (unsigned int)i); break; }
        }
        if (tid == 0 && (long long)nvis + n >= wv.max_visit) atomicMin(&s_cut, (unsigned int)(wv.max_visit - 1 - (long long)nvis));
        __syncthreads();
        const unsigned int cut = s_cut;
        const int ncut = cut == 0xFFFFFFFFu ? n : (int)cut + 1;
        for (int i = tid; i < ncut; i += WALK_THREADS) {
            const unsigned int incl = ssz[i], prev = i ? ssz[i - 1] : 0u;
            const unsigned int sz = incl - prev;
            if (sz) {
                const unsigned int p = nseg + sfl[i] - 1u;
                const unsigned int cell = st[i], r = (unsigned int)sk[i];
                if (p < (unsigned int)wv.segcap) segq[p] = make_uint4(dir.ustart[r], sz, (unsigned int)(got + prev), cell);
                else atomicExch(&wv.cnt->err, 2u);
                const unsigned int c0 = cell / (unsigned int)V, c1 = cell - c0 * (unsigned int)V;
                atomicOr(&used[c0 >> 5], 1u << (c0 & 31));
                atomicOr(&used[UW + (c1 >> 5)], 1u << (c1 & 31));
            }
        }
        __syncthreads();
        got += ssz[ncut - 1];
        nseg += sfl[ncut - 1];
        nvis += (unsigned int)ncut;
        if (cut != 0xFFFFFFFFu) done = true;
        __syncthreads();
    }
    // ---- projection slots of the coarse codes this query needs
    __syncthreads();
    if (tid == 0) {
        unsigned int ns = 0;
        for (int w = 0; w < 2 * UW; ++w) ns += __popc(used[w]);
        const unsigned int base = ns ? atomicAdd(&wv.cnt->n_lut, ns) : 0u;
        s_cnt = base;
        wv.nvis[q] = (int32_t)nvis;
        wv.nseg[q] = (int32_t)nseg;
        wv.ncand[q] = (unsigned int)got;
        if (got) atomicAdd(&wv.cnt->cand_total, got);
    }
    __syncthreads();
    if (tid < 2) {                                                    // one thread per split walks its bitmap
        unsigned int slot = s_cnt;
        if (tid == 1) for (int w = 0; w < UW; ++w) slot += __popc(used[w]);
        int32_t* sl = (tid ? wv.slot1 : wv.slot0) + (size_t)q * V;
        for (int w = 0; w < UW; ++w) {
            unsigned int bits = used[tid * UW + w];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const int c = w * 32 + b;
                sl[c] = (int32_t)slot;
                int32_t* de = wv.lut_desc + 3 * (size_t)slot;
                de[0] = q; de[1] = tid; de[2] = c;
                ++slot;
            }
        }
    }
}

// group of queries [qa, qa + ng): flattened candidates; qoff[ng + 1] = prefix of their candidate counts.
// keys[i] = float64 ADC distance bits, vals[i] = retrieval position (candidates are enumerated in retrieval order, so a stable
// sort by key gives the order of the reference's sorted(), search.py:210)
__global__ void __launch_bounds__(256)
k_cand_dist(ModelView mv, const uint8_t* __restrict__ codes, WalkView wv, int qa, int ng, const unsigned long long* __restrict__ qoff,
            const double* __restrict__ P64, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals) {
    const unsigned long long total = qoff[ng];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = ng;                                          // query: last g with qoff[g] <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (qoff[mid] <= i) lo = mid; else hi = mid; }
        const int q = qa + lo;
        const unsigned int pos = (unsigned int)(i - qoff[lo]);
        const uint4* segq = wv.seg + (size_t)q * wv.segcap;
        int a = 0, b = wv.nseg[q];                                    // segment: last with base position <= pos
        while (b - a > 1) { const int mid = (a + b) >> 1; if (segq[mid].z <= pos) a = mid; else b = mid; }
        const uint4 sg = segq[a];
        const unsigned int incell = pos - sg.z;
        const unsigned int c0 = sg.w / (unsigned int)mv.V, c1 = sg.w - c0 * (unsigned int)mv.V;
        const double* p0 = P64 + (int64_t)wv.slot0[(size_t)q * mv.V + c0] * mv.h;
        const double* p1 = P64 + (int64_t)wv.slot1[(size_t)q * mv.V + c1] * mv.h;
        const double d = exact_adc(mv, codes + (int64_t)(sg.x + incell) * mv.MP, (int64_t)incell, p0, p1);
        keys[i] = (unsigned long long)__double_as_longlong(d);
        vals[i] = pos;
    }
}

// first k of every query of the group (sorted keys / positions) -> records; one block per query
__global__ void __launch_bounds__(128)
k_emit_sorted(ModelView mv, const uint8_t* __restrict__ codes, const int64_t* __restrict__ rowids, WalkView wv, int qa,
              const unsigned long long* __restrict__ qoff, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
              int k, void* recbuf, int nq_rec, int qrec0) {
    const int g = blockIdx.x, q = qa + g;            // q: index in the plan arrays of this sub-batch
    const int qr = qrec0 + q;                        // index of the query in the record buffer (whole batch)
    RecView rv = rec_view(recbuf, nq_rec, k, mv.M);
    const unsigned long long o = qoff[g];
    const unsigned int n = wv.ncand[q];
    const int nout = (int)min((unsigned int)k, n);
    const uint4* segq = wv.seg + (size_t)q * wv.segcap;
    const int ns = wv.nseg[q];
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        const unsigned int pos = vals[o + i];
        int a = 0, b = ns;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (segq[mid].z <= pos) a = mid; else b = mid; }
        const uint4 sg = segq[a];
        const unsigned int incell = pos - sg.z;
        const int64_t row = (int64_t)sg.x + incell;
        const int64_t e = (int64_t)qr * k + i;
        rv.d64[e] = __longlong_as_double((long long)keys[o + i]);
        rv.pos[e] = pos;
        rv.rowid[e] = rowids[row];
        rv.cell[e] = (int32_t)sg.w;
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(codes + row * mv.MP, (int64_t)incell, j, mv.SW);
    }
    if (threadIdx.x == 0) {
        rv.lb[qr] = __longlong_as_double(0x7FF0000000000000ll);      // every retrieved code was ranked exactly
        rv.count[qr] = nout;
        rv.visited[qr] = wv.nvis[q];
        rv.ncand[qr] = (int64_t)n;
    }
}


// The same result without sorting everything: the product asks for the first 100 of ~10 000 retrieved codes per query.
// One block per query: its keys in shared memory, an 8-pass radix SELECT finds the k-th smallest key T exactly, the block
// keeps every key < T and, of the keys equal to T, the ones with the smallest retrieval positions (ordered compaction: the
// stable order of the reference's sorted(), search.py:210), sorts those k pairs (bitonic, by (key, position)) and writes
// the records.  Replaces 11 segmented radix-sort passes + k_emit_sorted for k <= SELK_MAXK and <= SELK_MAXN candidates.
#define SELK_THREADS 256
#define SELK_MAXK 1024
#define SELK_MAXN 12288
inline size_t selk_smem_bytes(unsigned int nmax) { return (size_t)nmax * 8 + (size_t)SELK_MAXK * 12 + 64; }

__global__ void __launch_bounds__(SELK_THREADS)
k_select_emit(ModelView mv, const uint8_t* __restrict__ codes, const int64_t* __restrict__ rowids, WalkView wv, int qa,
              const unsigned long long* __restrict__ qoff, const unsigned long long* __restrict__ keys, unsigned int nmax,
              int k, void* recbuf, int nq_rec, int qrec0) {
    extern __shared__ __align__(16) unsigned long long sm_selk[];      // keys [nmax] | wkey [SELK_MAXK] | widx [SELK_MAXK]
    __shared__ unsigned int hist[256], wsum[SELK_THREADS / 32];
    __shared__ unsigned int s_cnt, s_digit, s_rem;
    unsigned long long* skey = sm_selk;
    unsigned long long* wkey = sm_selk + nmax;
    unsigned int* widx = (unsigned int*)(wkey + SELK_MAXK);
    const int g = blockIdx.x, q = qa + g, tid = threadIdx.x;
    const int qr = qrec0 + q;
    RecView rv = rec_view(recbuf, nq_rec, k, mv.M);
    const unsigned long long o = qoff[g];
    const unsigned int n = wv.ncand[q];
    const unsigned int kk = min((unsigned int)k, n);
    if (tid == 0) {
        rv.lb[qr] = __longlong_as_double(0x7FF0000000000000ll);      // every retrieved code was ranked exactly
        rv.count[qr] = (int32_t)kk;
        rv.visited[qr] = wv.nvis[q];
        rv.ncand[qr] = (int64_t)n;
    }
    if (kk == 0) return;
    for (unsigned int i = tid; i < n; i += SELK_THREADS) skey[i] = keys[o + i];
    // ---- radix select: T = the kk-th smallest key, rem = how many keys equal to T belong to the first kk
    unsigned long long prefix = 0ull, mask = 0ull;
    unsigned int rem = kk;
    for (int shift = 56; shift >= 0; shift -= 8) {
        hist[tid] = 0u;
        __syncthreads();
        for (unsigned int i = tid; i < n; i += SELK_THREADS) {
            const unsigned long long key = skey[i];
            if ((key & mask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {                                              // first digit whose cumulative count reaches rem
            unsigned int c[8], tot = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { c[e] = hist[tid * 8 + e]; tot += c[e]; }
            unsigned int incl = tot;
            for (int d = 1; d < 32; d <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
            unsigned int before = incl - tot;
            const bool mine = before < rem && rem <= incl;          // exactly one lane
            if (mine) {
                int e = 0;
                while (before + c[e] < rem) { before += c[e]; ++e; }
                s_digit = (unsigned int)(tid * 8 + e); s_rem = rem - before;
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)s_digit << shift;
        mask |= 0xFFull << shift;
        rem = s_rem;
        __syncthreads();
    }
    const unsigned long long T = prefix;
    // ---- keep: key < T (all of them), key == T (the first `rem` in retrieval order)
    if (tid == 0) s_cnt = 0u;
    __syncthreads();
    unsigned int ties_seen = 0;
    for (unsigned int b0 = 0; b0 < n; b0 += SELK_THREADS) {
        const unsigned int i = b0 + tid;
        const unsigned long long key = (i < n) ? skey[i] : ~0ull;
        const bool lt = (i < n) && key < T, tie = (i < n) && key == T;
        const unsigned int bal = __ballot_sync(0xffffffffu, tie);
        if ((tid & 31) == 0) wsum[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned int ord = ties_seen, tot = 0;
#pragma unroll
        for (int w = 0; w < SELK_THREADS / 32; ++w) { if (w < (tid >> 5)) ord += wsum[w]; tot += wsum[w]; }
        ord += __popc(bal & ((1u << (tid & 31)) - 1u));
        if (lt || (tie && ord < rem)) {
            const unsigned int slot = atomicAdd(&s_cnt, 1u);
            wkey[slot] = key; widx[slot] = i;
        }
        ties_seen += tot;
        __syncthreads();
    }
    // ---- sort the kk pairs by (key, position)
    unsigned int P = 1;
    while (P < kk) P <<= 1;
    for (unsigned int i = kk + tid; i < P; i += SELK_THREADS) { wkey[i] = ~0ull; widx[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (unsigned int kb = 2; kb <= P; kb <<= 1) {
        for (unsigned int j = kb >> 1; j > 0; j >>= 1) {
            for (unsigned int i = tid; i < P; i += SELK_THREADS) {
                const unsigned int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long ka = wkey[i], kc = wkey[ixj];
                    const unsigned int ia = widx[i], ic = widx[ixj];
                    const bool gt = ka > kc || (ka == kc && ia > ic);
                    if (gt == ((i & kb) == 0)) { wkey[i] = kc; wkey[ixj] = ka; widx[i] = ic; widx[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    // ---- records (as k_emit_sorted)
    const uint4* segq = wv.seg + (size_t)q * wv.segcap;
    const int ns = wv.nseg[q];
    for (unsigned int i = tid; i < kk; i += SELK_THREADS) {
        const unsigned int pos = widx[i];
        int a = 0, b = ns;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (segq[mid].z <= pos) a = mid; else b = mid; }
        const uint4 sg = segq[a];
        const unsigned int incell = pos - sg.z;
        const int64_t row = (int64_t)sg.x + incell;
        const int64_t e = (int64_t)qr * k + i;
        rv.d64[e] = __longlong_as_double((long long)wkey[i]);
        rv.pos[e] = pos;
        rv.rowid[e] = rowids[row];
        rv.cell[e] = (int32_t)sg.w;
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(codes + row * mv.MP, (int64_t)incell, j, mv.SW);
    }
}

// ---- float32 preselection of the retrieved codes -------------------------------------------------------------------------
// The exact float64 ADC of EVERY retrieved code (k_cand_dist) is bound by the float64 pipe (3 non-fused operations per
// dimension and candidate); the reference only reports the first k.  One block per query:
//   A. every candidate's distance in float32 (FFMA, float32 copies of the projections and of the codebook), kept in shared
//      memory, with a rigorous error bound per query  |d32 - d| <= E = (D + 4) 2^-24 * 1.01 * 2 (max |p|^2 + sum_j max_k |c_jk|^2)
//      (inputs rounded to float32: 2u on every difference, 4u (|p_i| + |c_i|)^2 on its square; D-term FMA accumulation:
//      D u sum t_i^2; sum_i (|p_i| + |c_i|)^2 <= 2 (|p|^2 + |c|^2));
//   B. d32_k = the k-th smallest float32 distance (radix select on the float bits): at least k candidates have d <= d32_k + E;
//   C. survivors = candidates with d32 <= d32_k + 2 E: every member of the exact first k (ties included) is one of them;
//   D. the exact float64 ADC of the survivors only (k plus a handful), bitonic sort by (distance, retrieval position) -- the
//      order of the reference's stable sorted() (search.py:210) -- and the records.
// Same records as the full evaluation, bit for bit; a query that keeps more than PRS_SCAP survivors (massive near ties)
// raises presel_fallback and the host sends the group through the full evaluation.
#define PRS_THREADS 256
#define PRS_MAXN 12288
#define PRS_SCAP 2048
inline size_t presel_smem_bytes(unsigned int nmax) { return (size_t)nmax * 4 + (size_t)PRS_SCAP * 12 + 64; }

// float32 copies of the projection slots and their squared norms (rounded up)
__global__ void __launch_bounds__(256) k_slot_prep(const double* __restrict__ P64, const unsigned int* __restrict__ nslot_p, int h,
                                                  float* __restrict__ P32, float* __restrict__ n2) {
    const unsigned int nslot = *nslot_p;
    const int lane = threadIdx.x & 31;
    const unsigned int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned int sl = wid; sl < nslot; sl += nw) {
        double acc = 0.0;
        for (int t = lane; t < h; t += 32) {
            const double v = P64[(size_t)sl * h + t];
            P32[(size_t)sl * h + t] = (float)v;
            acc += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) n2[sl] = __double2float_ru(acc * (1.0 + 1e-12));
    }
}

__global__ void __launch_bounds__(PRS_THREADS, 3)
k_presel_emit(ModelView mv, const uint8_t* __restrict__ codes, const int64_t* __restrict__ rowids, WalkView wv, int qa,
              const double* __restrict__ P64, const float* __restrict__ P32, const float* __restrict__ n2, float c2tot,
              unsigned int nmax, int k, void* recbuf, int nq_rec, int qrec0) {
    extern __shared__ __align__(16) unsigned char sm_prs[];           // d32 [nmax] | skey [SCAP] | spos [SCAP]
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_cnt, s_digit, s_rem;
    __shared__ float s_smax[PRS_THREADS / 32];
    float* d32s = (float*)sm_prs;
    unsigned long long* skey = (unsigned long long*)(sm_prs + (((size_t)nmax * 4 + 15) & ~(size_t)15));
    unsigned int* spos = (unsigned int*)(skey + PRS_SCAP);
    const int g = blockIdx.x, q = qa + g, tid = threadIdx.x;
    const int qr = qrec0 + q;
    RecView rv = rec_view(recbuf, nq_rec, k, mv.M);
    const unsigned int n = wv.ncand[q];
    const unsigned int kk = min((unsigned int)k, n);
    if (tid == 0) {
        rv.lb[qr] = __longlong_as_double(0x7FF0000000000000ll);      // every possible member of the first k was ranked exactly
        rv.count[qr] = (int32_t)kk;
        rv.visited[qr] = wv.nvis[q];
        rv.ncand[qr] = (int64_t)n;
    }
    if (kk == 0) return;
    const uint4* segq = wv.seg + (size_t)q * wv.segcap;
    const int ns = wv.nseg[q];
    const int32_t* sl0 = wv.slot0 + (size_t)q * mv.V;
    const int32_t* sl1 = wv.slot1 + (size_t)q * mv.V;
    const float U24 = 5.9604645e-08f;
    auto segment_of = [&](unsigned int pos) {
        int a = 0, b = ns;
        while (b - a > 1) { const int mid = (a + b) >> 1; if (segq[mid].z <= pos) a = mid; else b = mid; }
        return a;
    };
    // ---- A: float32 distances; S = the largest |p|^2 of a candidate of this query (one error bound per query)
    float smax = 0.0f;
    for (unsigned int pos = tid; pos < n; pos += PRS_THREADS) {
        const uint4 sg = segq[segment_of(pos)];
        const unsigned int incell = pos - sg.z;
        const unsigned int c0 = sg.w / (unsigned int)mv.V, c1 = sg.w - c0 * (unsigned int)mv.V;
        const int s0 = sl0[c0], s1 = sl1[c1];
        const uint8_t* code = codes + (int64_t)(sg.x + incell) * mv.MP;
        float d32 = 0.0f;
        for (int j = 0; j < mv.M; ++j) {
            const int s = j / mv.m;
            const float* p = P32 + (size_t)(s ? s1 : s0) * mv.h + (j - s * mv.m) * mv.ds;
            const float* c = mv.subs32 + ((size_t)j * mv.K + code_byte(code, (int64_t)incell, j, mv.SW)) * mv.ds;
            if ((mv.ds & 3) == 0) {                                   // 16-byte reads (rows, sub-vectors and centroids are 16-byte aligned)
                for (int t = 0; t < mv.ds; t += 4) {
                    const float4 pv = *(const float4*)(p + t), cv = *(const float4*)(c + t);
                    float df = pv.x - cv.x; d32 = fmaf(df, df, d32);
                    df = pv.y - cv.y; d32 = fmaf(df, df, d32);
                    df = pv.z - cv.z; d32 = fmaf(df, df, d32);
                    df = pv.w - cv.w; d32 = fmaf(df, df, d32);
                }
            } else {
                for (int t = 0; t < mv.ds; ++t) { const float df = p[t] - c[t]; d32 = fmaf(df, df, d32); }
            }
        }
        d32s[pos] = d32;
        smax = fmaxf(smax, n2[s0] + n2[s1]);
    }
    for (int o = 16; o > 0; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    if ((tid & 31) == 0) s_smax[tid >> 5] = smax;
    __syncthreads();
    smax = s_smax[0];
#pragma unroll
    for (int w = 1; w < PRS_THREADS / 32; ++w) smax = fmaxf(smax, s_smax[w]);
    const float E = (float)(mv.D + 4) * U24 * 1.01f * 2.0f * (smax * 1.0001f + c2tot) * 1.001f + 1e-37f;
    // ---- B: the kk-th smallest float32 distance (non-negative floats: the bit patterns order like the values)
    unsigned int prefix = 0u, mask = 0u, rem = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0u;
        __syncthreads();
        for (unsigned int i = tid; i < n; i += PRS_THREADS) {
            const unsigned int key = __float_as_uint(d32s[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            unsigned int c[8], tot = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { c[e] = hist[tid * 8 + e]; tot += c[e]; }
            unsigned int incl = tot;
            for (int d = 1; d < 32; d <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
            unsigned int before = incl - tot;
            if (before < rem && rem <= incl) {
                int e = 0;
                while (before + c[e] < rem) { before += c[e]; ++e; }
                s_digit = (unsigned int)(tid * 8 + e); s_rem = rem - before;
            }
        }
        __syncthreads();
        prefix |= s_digit << shift;
        mask |= 0xFFu << shift;
        rem = s_rem;
        __syncthreads();
    }
    // at least kk candidates have d <= d32_k + E; a candidate of the exact first kk has d32 - E <= d <= d32_k + E
    const float Ustar = (__uint_as_float(prefix) + 2.0f * E) * (1.0f + 8.0f * U24);
    // ---- C: survivors
    if (tid == 0) s_cnt = 0u;
    __syncthreads();
    for (unsigned int i = tid; i < n; i += PRS_THREADS) {
        if (d32s[i] <= Ustar) {
            const unsigned int slot = atomicAdd(&s_cnt, 1u);
            if (slot < PRS_SCAP) spos[slot] = i;
        }
    }
    __syncthreads();
    const unsigned int nsurv = s_cnt;
    if (nsurv > PRS_SCAP) {                                           // too many near ties: the host reruns the group in full
        if (tid == 0) atomicAdd(&wv.cnt->presel_fallback, 1u);
        return;
    }
    // ---- D: exact distances of the survivors, sorted by (distance, retrieval position)
    for (unsigned int i = tid; i < nsurv; i += PRS_THREADS) {
        const unsigned int pos = spos[i];
        const uint4 sg = segq[segment_of(pos)];
        const unsigned int incell = pos - sg.z;
        const unsigned int c0 = sg.w / (unsigned int)mv.V, c1 = sg.w - c0 * (unsigned int)mv.V;
        const double d = exact_adc(mv, codes + (int64_t)(sg.x + incell) * mv.MP, (int64_t)incell,
                                   P64 + (int64_t)sl0[c0] * mv.h, P64 + (int64_t)sl1[c1] * mv.h);
        skey[i] = (unsigned long long)__double_as_longlong(d);
    }
    unsigned int P = 1;
    while (P < nsurv) P <<= 1;
    for (unsigned int i = nsurv + tid; i < P; i += PRS_THREADS) { skey[i] = ~0ull; spos[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (unsigned int kb = 2; kb <= P; kb <<= 1) {
        for (unsigned int j = kb >> 1; j > 0; j >>= 1) {
            for (unsigned int i = tid; i < P; i += PRS_THREADS) {
                const unsigned int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long ka = skey[i], kc = skey[ixj];
                    const unsigned int ia = spos[i], ic = spos[ixj];
                    const bool gt = ka > kc || (ka == kc && ia > ic);
                    if (gt == ((i & kb) == 0)) { skey[i] = kc; skey[ixj] = ka; spos[i] = ic; spos[ixj] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned int i = tid; i < kk; i += PRS_THREADS) {
        const unsigned int pos = spos[i];
        const uint4 sg = segq[segment_of(pos)];
        const unsigned int incell = pos - sg.z;
        const int64_t row = (int64_t)sg.x + incell;
        const int64_t e = (int64_t)qr * k + i;
        rv.d64[e] = __longlong_as_double((long long)skey[i]);
        rv.pos[e] = pos;
        rv.rowid[e] = rowids[row];
        rv.cell[e] = (int32_t)sg.w;
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(codes + row * mv.MP, (int64_t)incell, j, mv.SW);
    }
}
