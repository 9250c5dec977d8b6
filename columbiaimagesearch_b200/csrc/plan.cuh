// Query planning kernels: multi-sequence cell order + quota cut (search.py:13-82, 110-135),
// work-list construction for the scan, and the distance-table (LUT) build (model.py:673-704).
#pragma once
#include "common.cuh"

struct PlanCounters {               // zeroed before every batch
    unsigned int n_lut;             // LUT slots (query, split, coarse code)
    unsigned int n_partial;         // partial top-k lists (query, visited local cell, segment)
    unsigned int n_items;           // scan work items (cell, segment, query group)
    unsigned int n_pairs;           // (query, visited local non-empty cell) pairs
    unsigned long long cand_local;  // sum over queries of retrieved codes stored on this rank
    unsigned int next_item;         // dynamic work counter of the persistent scan kernel
    unsigned int max_parts;         // max partial lists of one query
};

struct PlanView {                   // per-batch device arrays, maxvis = V*V
    int nq, maxvis, segc;
    int32_t* nvis;                  // [nq] cells visited (incl. empty / non-local)
    int64_t* ncand;                 // [nq] retrieved codes, global
    int64_t* ncand_local;           // [nq] retrieved codes on this rank
    int32_t* npart;                 // [nq] partial lists of the query
    int32_t* pbase;                 // [nq] first partial list
    int32_t* vis_cell;              // [nq][maxvis] cell id = c0*V + c1
    int64_t* vis_base;              // [nq][maxvis] retrieval position of the cell's first code
    int32_t* vis_lut0;              // [nq][maxvis] LUT slot of (split 0, c0), -1 if nothing to scan
    int32_t* vis_lut1;              // [nq][maxvis]
    int32_t* vis_pbase;             // [nq][maxvis] partial-list offset inside the query, -1 if none
    double*  vis_dist;              // [nq][maxvis] cell distance d0+d1 (NULL unless the cell order itself is requested)
    int32_t* lut_desc;              // [cap][3]  (q, split, c)
    unsigned int* cell_qcount;      // [V*V] queries visiting the (local, non-empty) cell
    unsigned int* cell_fill;        // [V*V]
    unsigned int* cellq_off;        // [V*V+1]
    unsigned int* item_base;        // [nsegmax*V*V+1], f = seg*ncell + cell
    unsigned int* item_f;           // [item_cap] f of every work item (items beyond the capacity are located by search)
    unsigned int item_cap;
    int2* cellq;                    // [n_pairs] (q, visit index), grouped by cell
    PlanCounters* cnt;
    const int32_t* cell_segc;       // [V*V] per-cell segment length (low-batch scan: items of equal size over all cells); NULL: segc
};

// per-query scratch the later kernels expect initialised; done by the query's own block of k_coarse_order instead of
// one memset per array (all pointers may be NULL)
struct InitView {
    unsigned int* gthr;     // [nq]      <- "no bound yet"
    unsigned int* gtab;     // [nq][E]   <- "nothing seen"
    unsigned int* qmin;     // [nq][M]   <- +inf pattern
    unsigned int* qmax;     // [nq]      <- 0
    int E, M;
};

// ---- cell order + quota: one warp per query ---------------------------------------------------
// shared memory per block: d[2V] double, hd[V+2] double, ord[2V], pc[V], hi0[V+2], hi1[V+2], slotmap[2V] int
template <typename XT>
__global__ void __launch_bounds__(32)
k_coarse_order(ModelView mv, const XT* __restrict__ Xq, int64_t quota,
               const int64_t* __restrict__ gsize, const int64_t* __restrict__ lsize, PlanView pv, InitView iv) {
    extern __shared__ double sm_co[];
    const int V = mv.V, h = mv.h;
    double* d = sm_co;                       // [2][V]
    double* hd = d + 2 * V;                  // heap distances [V+2]
    int* ord = (int*)(hd + V + 2);           // [2][V]
    int* pc = ord + 2 * V;                   // [V] popped prefix length per row
    int* hi0 = pc + V;                       // [V+2]
    int* hi1 = hi0 + V + 2;                  // [V+2]
    int* slotmap = hi1 + V + 2;              // [2][V]
    const int q = blockIdx.x, lane = threadIdx.x;
    const XT* x = Xq + (int64_t)q * mv.D;
    const bool f32 = (sizeof(XT) == 4) && mv.coarse_f32;
    if (iv.gthr && lane == 0) iv.gthr[q] = 0x7f7f7f7fu;
    if (iv.qmax && lane == 0) iv.qmax[q] = 0u;
    if (iv.gtab) for (int e = lane; e < iv.E; e += 32) iv.gtab[(size_t)q * iv.E + e] = 0x7f7f7f7fu;
    if (iv.qmin) for (int e = lane; e < iv.M; e += 32) iv.qmin[(size_t)q * iv.M + e] = 0xFFFFFFFFu;

    for (int idx = lane; idx < 2 * V; idx += 32) {
        const int s = idx / V;
        const double* C = mv.Cs + (int64_t)idx * h;          // (s*V + v)*h
        d[idx] = f32 ? (double)sqdist_np<float>(x + s * h, C, h) : sqdist_np<double>(x + s * h, C, h);
        slotmap[idx] = -1;
    }
    for (int v = lane; v < V; v += 32) pc[v] = 0;
    __syncwarp();
    for (int idx = lane; idx < 2 * V; idx += 32) {            // stable argsort by rank counting
        const int s = idx / V, v = idx % V;
        const double dv = d[idx];
        int rank = 0;
        for (int u = 0; u < V; ++u) {
            const double du = d[s * V + u];
            rank += (du < dv) || (du == dv && u < v);
        }
        ord[s * V + rank] = v;
    }
    __syncwarp();
    if (lane != 0) return;

    int32_t* vcell = pv.vis_cell + (int64_t)q * pv.maxvis;
    int64_t* vbase = pv.vis_base + (int64_t)q * pv.maxvis;
    int32_t* vl0 = pv.vis_lut0 + (int64_t)q * pv.maxvis;
    int32_t* vl1 = pv.vis_lut1 + (int64_t)q * pv.maxvis;
    int32_t* vpb = pv.vis_pbase + (int64_t)q * pv.maxvis;

    auto celld = [&](int i0, int i1) -> double {             // search.py:52-57: 0 + d0 + d1 in the compute type
        const double a = d[ord[i0]], b = d[V + ord[V + i1]];
        return f32 ? (double)__fadd_rn((float)a, (float)b) : __dadd_rn(a, b);
    };
    int hn = 0;
    hd[0] = celld(0, 0); hi0[0] = 0; hi1[0] = 0; hn = 1;
    int nv = 0, nslots = 0, npart = 0;
    int64_t got = 0, got_local = 0;
    while (hn > 0) {
        int b = 0;                                            // pop the minimum (dist, (i0, i1))
        for (int e = 1; e < hn; ++e) {
            if (hd[e] < hd[b] || (hd[e] == hd[b] && (hi0[e] < hi0[b] || (hi0[e] == hi0[b] && hi1[e] < hi1[b])))) b = e;
        }
        const int i0 = hi0[b], i1 = hi1[b];
        const double dpop = hd[b];
        --hn;
        hd[b] = hd[hn]; hi0[b] = hi0[hn]; hi1[b] = hi1[hn];
        pc[i0] = i1 + 1;
        const int c0 = ord[i0], c1 = ord[V + i1];
        const int cell = c0 * V + c1;
        const int64_t gs = gsize[cell], ls = lsize[cell];
        vcell[nv] = cell;
        vbase[nv] = got;
        if (pv.vis_dist) pv.vis_dist[(int64_t)q * pv.maxvis + nv] = dpop;
        if (ls > 0) {
            if (slotmap[c0] < 0) slotmap[c0] = nslots++;
            if (slotmap[V + c1] < 0) slotmap[V + c1] = nslots++;
            vl0[nv] = slotmap[c0];
            vl1[nv] = slotmap[V + c1];
            vpb[nv] = npart;
            npart += (int)((ls + pv.segc - 1) / pv.segc);
            got_local += ls;
        } else {
            vl0[nv] = -1; vl1[nv] = -1; vpb[nv] = -1;
        }
        got += gs;
        ++nv;
        if (got >= quota) break;
        // search.py:70-80 push rules; pc[] encodes the `traversed` set (a staircase)
        if ((i1 == 0 || pc[i0 + 1 < V ? i0 + 1 : i0] >= i1) && i0 + 1 < V) {
            hd[hn] = celld(i0 + 1, i1); hi0[hn] = i0 + 1; hi1[hn] = i1; ++hn;
        }
        if ((i0 == 0 || pc[i0 - 1] >= i1 + 2) && i1 + 1 < V) {
            hd[hn] = celld(i0, i1 + 1); hi0[hn] = i0; hi1[hn] = i1 + 1; ++hn;
        }
    }
    pv.nvis[q] = nv;
    pv.ncand[q] = got;
    pv.ncand_local[q] = got_local;
    pv.npart[q] = npart;
    const unsigned int lbase = nslots ? atomicAdd(&pv.cnt->n_lut, (unsigned)nslots) : 0u;
    pv.pbase[q] = npart ? (int)atomicAdd(&pv.cnt->n_partial, (unsigned)npart) : 0;
    atomicMax(&pv.cnt->max_parts, (unsigned)npart);
    if (got_local) atomicAdd(&pv.cnt->cand_local, (unsigned long long)got_local);
    for (int idx = 0; idx < 2 * V; ++idx) {
        if (slotmap[idx] >= 0) {
            int32_t* de = pv.lut_desc + 3 * (int64_t)(lbase + slotmap[idx]);
            de[0] = q; de[1] = idx / V; de[2] = idx % V;
        }
    }
    int npairs = 0;
    for (int v = 0; v < nv; ++v) {
        if (vpb[v] >= 0) {
            vl0[v] += (int)lbase; vl1[v] += (int)lbase;
            atomicAdd(&pv.cell_qcount[vcell[v]], 1u);
            ++npairs;
        }
    }
    if (npairs) atomicAdd(&pv.cnt->n_pairs, (unsigned)npairs);
}

// ---- offsets of the work list: single block ------------------------------------------------------
// Work items are ordered SEGMENT-major: all (cell, query group) items of segment 0, then segment 1, ...
// so the segments of one query are spread over time and later ones start with the pruning bound the
// earlier ones published (ScanArgs::gthr); inside a segment level consecutive items share a code segment
// (L2 reuse).  item_base is indexed by f = seg * ncell + cell.
__global__ void __launch_bounds__(1024)
k_plan(int ncell, int nsegmax, int G, int segc, const int64_t* __restrict__ lsize, PlanView pv) {
    __shared__ unsigned int sq[1024], si[1024];
    // pass 1: per-cell query-list offsets
    {
        const int per = (ncell + 1023) / 1024;
        const int c0 = threadIdx.x * per;
        unsigned int aq = 0;
        for (int c = c0; c < min(ncell, c0 + per); ++c) aq += pv.cell_qcount[c];
        sq[threadIdx.x] = aq;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {               // Hillis-Steele inclusive scan
            unsigned int vq = 0;
            if ((int)threadIdx.x >= o) vq = sq[threadIdx.x - o];
            __syncthreads();
            sq[threadIdx.x] += vq;
            __syncthreads();
        }
        unsigned int bq = sq[threadIdx.x] - aq;            // exclusive
        for (int c = c0; c < min(ncell, c0 + per); ++c) {
            pv.cellq_off[c] = bq;
            pv.cell_fill[c] = 0;
            bq += pv.cell_qcount[c];
        }
        if (threadIdx.x == 1023) pv.cellq_off[ncell] = sq[1023];
        __syncthreads();
    }
    // pass 2: item offsets over (segment, cell)
    const int F = nsegmax * ncell;
    const int per = (F + 1023) / 1024;
    const int f0 = threadIdx.x * per;
    auto items_of = [&](int f) -> unsigned int {
        const int seg = f / ncell, c = f - seg * ncell;
        const int sc = pv.cell_segc ? pv.cell_segc[c] : segc;
        const unsigned int nseg = (unsigned)((lsize[c] + sc - 1) / sc);
        return ((unsigned)seg < nseg) ? (pv.cell_qcount[c] + G - 1) / G : 0u;
    };
    unsigned int ai = 0;
    for (int f = f0; f < min(F, f0 + per); ++f) ai += items_of(f);
    si[threadIdx.x] = ai;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned int vi = 0;
        if ((int)threadIdx.x >= o) vi = si[threadIdx.x - o];
        __syncthreads();
        si[threadIdx.x] += vi;
        __syncthreads();
    }
    unsigned int bi = si[threadIdx.x] - ai;
    for (int f = f0; f < min(F, f0 + per); ++f) { pv.item_base[f] = bi; bi += items_of(f); }
    if (threadIdx.x == 1023) { pv.item_base[F] = si[1023]; pv.cnt->n_items = si[1023]; }
}

// ---- item -> (segment, cell) table, so a scan block finds its item with one load instead of a binary search ----
__global__ void k_item_table(int F, PlanView pv) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) {
        const unsigned int a = pv.item_base[f], b = min(pv.item_base[f + 1], pv.item_cap);
        for (unsigned int i = a; i < b; ++i) pv.item_f[i] = (unsigned int)f;
    }
}

// ---- group the (query, visit) pairs by cell -----------------------------------------------------
__global__ void k_fill(PlanView pv) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= pv.nq) return;
    const int nv = pv.nvis[q];
    for (int v = 0; v < nv; ++v) {
        if (pv.vis_pbase[(int64_t)q * pv.maxvis + v] >= 0) {
            const int cell = pv.vis_cell[(int64_t)q * pv.maxvis + v];
            const unsigned int idx = atomicAdd(&pv.cell_fill[cell], 1u);
            pv.cellq[pv.cellq_off[cell] + idx] = make_int2(q, v);
        }
    }
}

// ---- LUT build: one block per (query, split, coarse code) ---------------------------------------
// project (model.py:604-641) then ((fx - subC[j])**2).sum(1) for the m sub-vectors of the split
// (model.py:696-704).  Writes the float64 projection (for the exact re-rank) and the float32 table
// rounded from float64, k-major: lut32[slot][k][j], j < m  (so one scan-LUT row is contiguous).
// DS > 0: compile-time sub-vector length (unrolled, registers); DS == 0: runtime ds.
// dynamic smem: r[h] | p[h] | psum[LUT_THREADS] doubles
#define LUT_THREADS 256
template <typename XT, int DS>
__global__ void __launch_bounds__(LUT_THREADS)
k_lut(ModelView mv, const XT* __restrict__ Xq, const int32_t* __restrict__ lut_desc, const PlanCounters* __restrict__ cnt,
      double* __restrict__ P64, float* __restrict__ lut32, double* __restrict__ lut64, int have_p = 0) {
    extern __shared__ double sm_lut[];
    const int h = mv.h, m = mv.m, V = mv.V;
    const int ds = DS ? DS : mv.ds;
    double* r = sm_lut;
    double* p = sm_lut + h;
    double* psum = p + h;
    const int tid = threadIdx.x;
    const int nslot = (int)cnt->n_lut;             // grid-stride over the slots the plan produced (no host round trip)
  for (int slot = blockIdx.x; slot < nslot; slot += gridDim.x) {
    __syncthreads();
    const int q = lut_desc[3 * slot], s = lut_desc[3 * slot + 1], c = lut_desc[3 * slot + 2];
    const XT* x = Xq + (int64_t)q * mv.D + s * h;
    const double* C = mv.Cs + ((int64_t)s * V + c) * h;
    const double* mu = mv.mus + ((int64_t)s * V + c) * h;
    if (have_p == 1) {     // the projection was produced by the grouped GEMM (k_rotate_dmma_g, large h)
        for (int d = tid; d < h; d += LUT_THREADS) p[d] = P64[(int64_t)slot * h + d];
    } else
    for (int d = tid; d < h; d += LUT_THREADS) r[d] = coarse_residual<XT>(x[d], C[d], mu[d], mv.coarse_f32);
    __syncthreads();
    // rotation p[t] = sum_d Rt[d][t] r[d]: `parts` threads share one output (contiguous d ranges, summed in order)
    const double* Rt = mv.Rt + ((int64_t)s * V + c) * h * (int64_t)h;
    int parts = 1;
    while (parts * 2 * h <= LUT_THREADS && parts * 2 <= h) parts *= 2;       // h = 64 -> 4 parts
    const int tw = LUT_THREADS / parts;                                       // outputs handled per sweep
    const int part = tid / tw, tl = tid - part * tw;
    const int dlen = h / parts, d0 = part * dlen;
    for (int t0 = 0; t0 < (have_p == 1 ? 0 : h); t0 += tw) {
        const int t = t0 + tl;
        double acc = 0.0;
        if (t < h) {
#pragma unroll 8
            for (int d = d0; d < d0 + dlen; ++d) acc = fma(Rt[(int64_t)d * h + t], r[d], acc);
        }
        if (parts == 1) { if (t < h) { p[t] = acc; P64[(int64_t)slot * h + t] = acc; } }
        else {
            psum[tid] = acc;
            __syncthreads();
            if (part == 0 && t < h) {
                double a = psum[tl];
                for (int pp = 1; pp < parts; ++pp) a += psum[pp * tw + tl];
                p[t] = a; P64[(int64_t)slot * h + t] = a;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (have_p == 2) continue;             // projection only (large-V search evaluates table entries on the fly)
    for (int k = tid; k < B2L_LUT_ROWS; k += LUT_THREADS) {
        const bool live = k < mv.K;
        float* o32 = lut32 ? lut32 + ((int64_t)slot * B2L_LUT_ROWS + k) * m : nullptr;
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            double e = 0.0;
            if (live) {
                const double* cj = mv.subs + (((int64_t)s * m + j) * mv.K + k) * ds;
                if (DS) {
                    double cv[DS ? DS : 1];
#pragma unroll
                    for (int d = 0; d < DS; ++d) cv[d] = cj[d];
                    e = sqdist_np<double>(p + j * DS, cv, DS);
                } else e = sqdist_np<double>(p + j * ds, cj, ds);
            }
            if (o32) o32[j] = (float)e;
            if (lut64 && live) lut64[((int64_t)slot * m + j) * mv.K + k] = e;
        }
    }
  }
}

// ---- LUT build, register-resident sub-centroids ----------------------------------------------------
// Same arithmetic as k_lut for the shapes where a thread can keep its share of the sub-quantizer codebook in
// registers: block = 512 threads bound to ONE coarse split (blockIdx.x & 1); thread (k = tid % 256, jh = tid / 256)
// holds centroid k of sub-quantizers jh*MH .. jh*MH+MH-1 of that split (MH*DS doubles) for the whole kernel and walks
// over the slots of its split, so the codebook is read once per block instead of once per slot.
// Also reduces, per query, the minimum of every table column and the overall maximum (qmin / qmax, for the 16-bit
// tables of the packed scan; NULL: skipped).  Needs m == 2*MH, ds == DS, K <= 256.  dynamic smem: r[h] | p[h] | psum[512] doubles
#define LUTR_THREADS 512
template <typename XT, int DS, int MH>
__global__ void __launch_bounds__(LUTR_THREADS)
k_lut_reg(ModelView mv, const XT* __restrict__ Xq, const int32_t* __restrict__ lut_desc, const PlanCounters* __restrict__ cnt,
          double* __restrict__ P64, float* __restrict__ lut32, unsigned int* __restrict__ qmin, unsigned int* __restrict__ qmax) {
    extern __shared__ double sm_lutr[];
    __shared__ unsigned int s_cmin[2 * MH], s_cmax;
    const int h = mv.h, m = mv.m, V = mv.V;
    double* r = sm_lutr;
    double* p = sm_lutr + h;
    double* psum = p + h;
    const int tid = threadIdx.x, k = tid & 255, jh = tid >> 8, lane = tid & 31;
    const int s = blockIdx.x & 1, nb = gridDim.x >> 1, b = blockIdx.x >> 1;
    const int nslot = (int)cnt->n_lut;
    const bool live = k < mv.K;
    double cv[MH][DS];
#pragma unroll
    for (int jj = 0; jj < MH; ++jj) {
        const double* cj = mv.subs + (((int64_t)s * m + jh * MH + jj) * mv.K + (live ? k : 0)) * DS;
#pragma unroll
        for (int d = 0; d < DS; ++d) cv[jj][d] = cj[d];
    }
    int parts = 1;                            // same split of the d range as k_lut (same float64 summation order)
    while (parts * 2 * h <= LUT_THREADS && parts * 2 <= h) parts *= 2;        // h = 64 -> 4 parts
    const int tw = LUTR_THREADS / parts;
    const int part = tid / tw, tl = tid - part * tw;
    const int dlen = h / parts, d0 = part * dlen;
    for (int slot = b; slot < nslot; slot += nb) {
        if (lut_desc[3 * slot + 1] != s) continue;                            // block-uniform
        __syncthreads();
        const int q = lut_desc[3 * slot], c = lut_desc[3 * slot + 2];
        const XT* x = Xq + (int64_t)q * mv.D + s * h;
        const double* C = mv.Cs + ((int64_t)s * V + c) * h;
        const double* mu = mv.mus + ((int64_t)s * V + c) * h;
        for (int d = tid; d < h; d += LUTR_THREADS) r[d] = coarse_residual<XT>(x[d], C[d], mu[d], mv.coarse_f32);
        if (tid < 2 * MH) s_cmin[tid] = 0xFFFFFFFFu;
        if (tid == 0) s_cmax = 0u;
        __syncthreads();
        const double* Rt = mv.Rt + ((int64_t)s * V + c) * h * (int64_t)h;
        for (int t0 = 0; t0 < h; t0 += tw) {
            const int t = t0 + tl;
            double acc = 0.0;
            if (t < h) {
#pragma unroll 8
                for (int d = d0; d < d0 + dlen; ++d) acc = fma(Rt[(int64_t)d * h + t], r[d], acc);
            }
            if (parts == 1) { if (t < h) { p[t] = acc; P64[(int64_t)slot * h + t] = acc; } }
            else {
                psum[tid] = acc;
                __syncthreads();
                if (part == 0 && t < h) {
                    double a = psum[tl];
                    for (int pp = 1; pp < parts; ++pp) a += psum[pp * tw + tl];
                    p[t] = a; P64[(int64_t)slot * h + t] = a;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        float e32[MH];
#pragma unroll
        for (int jj = 0; jj < MH; ++jj) {
            const double e = live ? sqdist_np<double>(p + (jh * MH + jj) * DS, cv[jj], DS) : 0.0;
            e32[jj] = (float)e;
        }
        float* o32 = lut32 + ((int64_t)slot * B2L_LUT_ROWS + k) * m + jh * MH;
        if (MH == 4) *(float4*)o32 = make_float4(e32[0], e32[1], e32[2], e32[3]);
        else {
#pragma unroll
            for (int jj = 0; jj < MH; ++jj) o32[jj] = e32[jj];
        }
        if (qmin) {              // per query: minimum of every table column and the overall maximum (QuantView, 16-bit tables)
            unsigned int mx = 0u;
#pragma unroll
            for (int jj = 0; jj < MH; ++jj) {
                const unsigned int v = live ? __float_as_uint(e32[jj]) : 0xFFFFFFFFu;
                const unsigned int wm = __reduce_min_sync(0xffffffffu, v);
                if (lane == 0) atomicMin(&s_cmin[jh * MH + jj], wm);
                mx = max(mx, live ? v : 0u);
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) atomicMax(&s_cmax, mx);
            __syncthreads();
            if (tid < 2 * MH) atomicMin(&qmin[(size_t)q * mv.M + s * m + tid], s_cmin[tid]);
            if (tid == 0) atomicMax(&qmax[q], s_cmax);
        }
    }
}

// ---- LUT entries in float32 (packed scan only) --------------------------------------------------------
// The packed scan only needs table entries to within a stated error: its candidates are re-evaluated in float64 from
// the float64 projection P64 and the exact sub-quantizer codebook (k_select).  So the 2048 entries of a table are
// evaluated in float32 (FSUB + FFMA on the float32 codebook, register-resident: 8 sub-quantizers x 8 floats per
// thread); the projection itself stays float64.  The evaluation error of an entry,
//     |e32 - e| <= Emax = 8 * 2^-24 * (sqrt(emax * S2) + ds * emax) + 1e-13 * S2,   S2 >= |p|^2 + |c|^2,
// is folded into the certification bound by k_lut_quant (QuantView::slack = M * Emax).
// Block = 256 threads bound to one coarse split; needs m == MJ, ds == DS, K <= 256.
// dynamic smem: r[h] | p[h] | psum[256] doubles | p32[h] floats
template <typename XT, int DS, int MJ>
__global__ void __launch_bounds__(256, 2)
k_lut_f32(ModelView mv, const XT* __restrict__ Xq, const int32_t* __restrict__ lut_desc, const PlanCounters* __restrict__ cnt,
          double* __restrict__ P64, float* __restrict__ lut32, unsigned int* __restrict__ qmin, unsigned int* __restrict__ qmax) {
    extern __shared__ double sm_l32[];
    __shared__ unsigned int s_cmin[MJ], s_cmax;
    const int h = mv.h, V = mv.V;
    double* r = sm_l32;
    double* p = sm_l32 + h;
    double* psum = p + h;
    float* p32 = (float*)(psum + 256);
    const int tid = threadIdx.x, k = tid, lane = tid & 31;
    const int s = blockIdx.x & 1, nb = gridDim.x >> 1, b = blockIdx.x >> 1;
    const int nslot = (int)cnt->n_lut;
    const bool live = k < mv.K;
    float cv[MJ][DS];
#pragma unroll
    for (int jj = 0; jj < MJ; ++jj) {
        const float* cj = mv.subs32 + (((int64_t)s * MJ + jj) * mv.K + (live ? k : 0)) * DS;
#pragma unroll
        for (int d = 0; d < DS; ++d) cv[jj][d] = cj[d];
    }
    int parts = 1;                            // same split of the d range as k_lut (same float64 projection bits)
    while (parts * 2 * h <= LUT_THREADS && parts * 2 <= h) parts *= 2;
    const int tw = 256 / parts;
    const int part = tid / tw, tl = tid - part * tw;
    const int dlen = h / parts, d0 = part * dlen;
    for (int slot = b; slot < nslot; slot += nb) {
        if (lut_desc[3 * slot + 1] != s) continue;                            // block-uniform
        __syncthreads();
        const int q = lut_desc[3 * slot], c = lut_desc[3 * slot + 2];
        const XT* x = Xq + (int64_t)q * mv.D + s * h;
        const double* C = mv.Cs + ((int64_t)s * V + c) * h;
        const double* mu = mv.mus + ((int64_t)s * V + c) * h;
        for (int d = tid; d < h; d += 256) r[d] = coarse_residual<XT>(x[d], C[d], mu[d], mv.coarse_f32);
        if (tid < MJ) s_cmin[tid] = 0xFFFFFFFFu;
        if (tid == 0) s_cmax = 0u;
        __syncthreads();
        const double* Rt = mv.Rt + ((int64_t)s * V + c) * h * (int64_t)h;
        for (int t0 = 0; t0 < h; t0 += tw) {
            const int t = t0 + tl;
            double acc = 0.0;
            if (t < h) {
#pragma unroll 8
                for (int d = d0; d < d0 + dlen; ++d) acc = fma(Rt[(int64_t)d * h + t], r[d], acc);
            }
            if (parts == 1) { if (t < h) { p[t] = acc; p32[t] = (float)acc; P64[(int64_t)slot * h + t] = acc; } }
            else {
                psum[tid] = acc;
                __syncthreads();
                if (part == 0 && t < h) {
                    double a = psum[tl];
                    for (int pp = 1; pp < parts; ++pp) a += psum[pp * tw + tl];
                    p[t] = a; p32[t] = (float)a; P64[(int64_t)slot * h + t] = a;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        float e32[MJ];
#pragma unroll
        for (int jj = 0; jj < MJ; ++jj) {
            float d = 0.0f;
#pragma unroll
            for (int t = 0; t < DS; ++t) { const float df = p32[jj * DS + t] - cv[jj][t]; d = fmaf(df, df, d); }
            e32[jj] = live ? d : 0.0f;
        }
        float* o32 = lut32 + ((int64_t)slot * B2L_LUT_ROWS + k) * MJ;
#pragma unroll
        for (int jj = 0; jj < MJ; jj += 4) *(float4*)(o32 + jj) = make_float4(e32[jj], e32[jj + 1], e32[jj + 2], e32[jj + 3]);
        unsigned int mx = 0u;
#pragma unroll
        for (int jj = 0; jj < MJ; ++jj) {
            const unsigned int v = live ? __float_as_uint(e32[jj]) : 0xFFFFFFFFu;
            const unsigned int wm = __reduce_min_sync(0xffffffffu, v);
            if (lane == 0) atomicMin(&s_cmin[jj], wm);
            mx = max(mx, live ? v : 0u);
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) atomicMax(&s_cmax, mx);
        __syncthreads();
        if (tid < MJ) atomicMin(&qmin[(size_t)q * mv.M + s * MJ + tid], s_cmin[tid]);
        if (tid == 0) atomicMax(&qmax[q], s_cmax);
    }
}

// ---- LUT entries in float32 for any sub-vector length (packed scan, large models) ----------------------------------
// Same contract as k_lut_f32 (entries within Emax of the float64 value, folded into the certification bound by
// k_lut_quant), for shapes whose codebook does not fit a thread's registers (2048-d: ds = 64).  The projection comes
// from the grouped GEMM.  Block = 256 threads = centroids; it takes SB slots of one coarse split at a time (perm lists
// the slots split by split), so a codebook element read once (subs32T: centroid index fastest, coalesced) serves SB
// tables.  dynamic smem: p32[SB][h] floats
template <int SB>
__global__ void __launch_bounds__(256)
k_lut_entries_f32(ModelView mv, const double* __restrict__ P64, const unsigned int* __restrict__ perm, const unsigned int* __restrict__ base,
                  const unsigned int* __restrict__ bcnt, size_t cap, float* __restrict__ lut32) {
    extern __shared__ float sm_le[];
    __shared__ unsigned int s_slot[SB];
    const int h = mv.h, m = mv.m, ds = mv.ds, V = mv.V, K = mv.K;
    const int k = threadIdx.x;
    const bool live = k < K;
    const unsigned int nr0 = base[V - 1] + bcnt[V - 1], nr1 = base[2 * V - 1] + bcnt[2 * V - 1];
    const unsigned int g0 = (nr0 + SB - 1) / SB, g1 = (nr1 + SB - 1) / SB;
    for (unsigned int g = blockIdx.x; g < g0 + g1; g += gridDim.x) {
        const int s = g >= g0;
        const unsigned int r0 = (g - (s ? g0 : 0u)) * SB, nr = s ? nr1 : nr0;
        const int ns = (int)min((unsigned int)SB, nr - r0);
        __syncthreads();
        if (k < SB) s_slot[k] = k < ns ? perm[(size_t)s * cap + r0 + k] : 0u;
        __syncthreads();
        for (int e = k; e < SB * h; e += 256) {
            const int i = e / h;
            sm_le[e] = i < ns ? (float)P64[(int64_t)s_slot[i] * h + (e - i * h)] : 0.0f;
        }
        __syncthreads();
        for (int j0 = 0; j0 < m; j0 += 4) {
            float out[SB][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = j0 + jj;
                float acc[SB];
#pragma unroll
                for (int i = 0; i < SB; ++i) acc[i] = 0.0f;
                if (j < m) {
                    const float* c = mv.subs32T + ((int64_t)(s * m + j) * ds) * K + (live ? k : 0);
                    const float* pj = sm_le + j * ds;
#pragma unroll 4
                    for (int t = 0; t < ds; ++t) {
                        const float cv = c[(int64_t)t * K];
#pragma unroll
                        for (int i = 0; i < SB; ++i) { const float df = pj[i * h + t] - cv; acc[i] = fmaf(df, df, acc[i]); }
                    }
                }
#pragma unroll
                for (int i = 0; i < SB; ++i) out[i][jj] = live ? acc[i] : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < SB; ++i) {
                if (i >= ns) break;
                float* o = lut32 + ((int64_t)s_slot[i] * B2L_LUT_ROWS + k) * m + j0;
                if (j0 + 4 <= m && (m & 3) == 0) *(float4*)o = make_float4(out[i][0], out[i][1], out[i][2], out[i][3]);
                else for (int jj = 0; jj < 4 && j0 + jj < m; ++jj) o[jj] = out[i][jj];
            }
        }
    }
}

// ---- 16-bit quantised tables for the packed scan ---------------------------------------------------
// Per query q: bias b[q][j] = min over the query's tables of sub-quantizer j (both splits: j < M) and over the 256
// centroids; step Delta[q] = (max entry - min bias) / QMAX * (1 + 2^-20); code = floor((e - b[q][j]) / Delta[q]),
// clamped to [0, QMAX], QMAX = 65535 / M so that a sum of M codes fits 16 bits.  Then for every candidate
//     float32 ADC sum  >=  B[q] + Delta[q] * (S - 0.1),   S = sum of its M codes, B[q] = sum_j b[q][j]
// (floor never rounds up by more than the float32 error of the scaled value, < 0.002 per term).
struct QuantView {
    unsigned int* qmin;     // [nq][M] float bits (entries are >= 0, so unsigned order == float order)
    unsigned int* qmax;     // [nq]    float bits
    double* B;              // [nq]
    double* delta;          // [nq]
    double* slack;          // [nq] M * (bound on the evaluation error of a table entry); 0 for the float64-built tables
    unsigned int* margin;   // [nq] the scan appends candidates up to (bound + margin) code units: wide enough that re-ranking
                            //      every appended candidate exactly always certifies the first k (see k_select2)
    int qmax_code;          // QMAX
    int f32_entries;        // the tables were evaluated in float32 (k_lut_f32)
    int ds;                 // sub-vector length
    float c2m;              // max_j max_k |subs[j][k]|^2 (upper bound)
};

// one block per LUT slot: ranges of its m columns -> per-query minima / maximum
__global__ void __launch_bounds__(256)
k_lut_range(int m, const int32_t* __restrict__ lut_desc, const PlanCounters* __restrict__ cnt, const float* __restrict__ lut32,
            int K, QuantView qv, int M) {
    __shared__ unsigned int s_min[32], s_max;
    const int nslot = (int)cnt->n_lut;
    for (int slot = blockIdx.x; slot < nslot; slot += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 32) s_min[threadIdx.x] = 0x7f800000u;
        if (threadIdx.x == 0) s_max = 0u;
        __syncthreads();
        const int q = lut_desc[3 * slot], s = lut_desc[3 * slot + 1];
        const float* t = lut32 + (size_t)slot * B2L_LUT_ROWS * m;
        unsigned int mx = 0u;
        for (int e = threadIdx.x; e < K * m; e += blockDim.x) {
            const unsigned int v = __float_as_uint(t[e]);
            atomicMin(&s_min[e % m], v);
            mx = max(mx, v);
        }
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
        __syncthreads();
        if (threadIdx.x < m) atomicMin(&qv.qmin[(size_t)q * M + s * m + threadIdx.x], s_min[threadIdx.x]);
        if (threadIdx.x == 0) atomicMax(&qv.qmax[q], s_max);
    }
}

// one block per LUT slot: scale and bias of the slot's query (recomputed by every block of the query, identical values),
// then lut16[slot][k][j] = clamp(floor((e - b) * inv), 0, QMAX)
__global__ void __launch_bounds__(256)
k_lut_quant(int m, const int32_t* __restrict__ lut_desc, const PlanCounters* __restrict__ cnt, const float* __restrict__ lut32,
            QuantView qv, int M, unsigned short* __restrict__ lut16) {
    __shared__ float s_b[64], s_inv;
    const int nslot = (int)cnt->n_lut;
    for (int slot = blockIdx.x; slot < nslot; slot += gridDim.x) {
        const int q = lut_desc[3 * slot], s = lut_desc[3 * slot + 1];
        __syncthreads();
        if (threadIdx.x == 0) {
            float bmin = __int_as_float(0x7f800000);
            double B = 0.0;
            bool any = false;
            for (int j = 0; j < M; ++j) {
                const unsigned int b = qv.qmin[(size_t)q * M + j];
                if (b < 0x7f800000u) { const float f = __uint_as_float(b); bmin = fminf(bmin, f); B += (double)f; any = true; }
            }                                                           // (a sub-quantizer without a table on this rank: bias 0)
            const float range = any ? fmaxf(__uint_as_float(qv.qmax[q]) - bmin, 1e-30f) : 1.0f;
            const double delta = (double)range / (double)qv.qmax_code * (1.0 + 9.5367431640625e-07);
            qv.delta[q] = delta;
            qv.B[q] = B;
            double slack = 0.0;
            if (qv.f32_entries) {                // evaluation error of the float32 entries (see k_lut_f32)
                const float emax = __uint_as_float(qv.qmax[q]);
                const float sp = sqrtf(emax) + sqrtf(qv.c2m);
                const float S2 = (sp * sp + qv.c2m) * 1.001f;
                slack = (double)M * ((double)(8.0f * 5.9604645e-08f) * ((double)sqrtf(emax * S2) + (double)qv.ds * (double)emax) * 1.01 + 1e-13 * (double)S2);
            }
            qv.slack[q] = slack;
            // A candidate with code sum S has a float32 table sum in [B + Delta*S, B + Delta*(S + M)) (each floor loses < 1 unit),
            // known to +-slack.  With every candidate up to bound + margin re-ranked exactly, the KP >= k candidates of smallest
            // S (all <= bound) have exact distances < B + Delta*(bound + M) + slack, and everything not appended is
            // > B + Delta*(bound + margin) - slack: margin > M + 2*slack/Delta makes the first k certifiable by construction.
            qv.margin[q] = (unsigned int)min(8192.0, (double)M + ceil(2.0 * slack / delta) + 2.0);
            s_inv = (float)(1.0 / delta);
        }
        if (threadIdx.x < m) {
            const unsigned int b = qv.qmin[(size_t)q * M + s * m + threadIdx.x];
            s_b[threadIdx.x] = b < 0x7f800000u ? __uint_as_float(b) : 0.0f;
        }
        __syncthreads();
        const float inv = s_inv;
        const float* t = lut32 + (size_t)slot * B2L_LUT_ROWS * m;
        unsigned short* o = lut16 + (size_t)slot * B2L_LUT_ROWS * m;
        if ((m & 3) == 0) {                         // four entries per step: one 16-byte load, one 8-byte store
            for (int e4 = threadIdx.x; e4 < B2L_LUT_ROWS * m / 4; e4 += blockDim.x) {
                const float4 v = *(const float4*)(t + 4 * e4);
                const int j = (4 * e4) % m;
                const int q0 = min(qv.qmax_code, max(0, (int)floorf(__fmul_rn(__fsub_rn(v.x, s_b[j]), inv))));
                const int q1 = min(qv.qmax_code, max(0, (int)floorf(__fmul_rn(__fsub_rn(v.y, s_b[j + 1]), inv))));
                const int q2 = min(qv.qmax_code, max(0, (int)floorf(__fmul_rn(__fsub_rn(v.z, s_b[j + 2]), inv))));
                const int q3 = min(qv.qmax_code, max(0, (int)floorf(__fmul_rn(__fsub_rn(v.w, s_b[j + 3]), inv))));
                *(uint2*)(o + 4 * e4) = make_uint2((unsigned)q0 | ((unsigned)q1 << 16), (unsigned)q2 | ((unsigned)q3 << 16));
            }
        } else {
            for (int e = threadIdx.x; e < B2L_LUT_ROWS * m; e += blockDim.x) {
                const float x = __fmul_rn(__fsub_rn(t[e], s_b[e % m]), inv);
                o[e] = (unsigned short)min(qv.qmax_code, max(0, (int)floorf(x)));
            }
        }
    }
}
