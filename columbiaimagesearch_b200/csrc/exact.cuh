// Exact (float64, full sort) path: ranks every retrieved code of one query exactly as
// search.py:137-177 + the stable sort of :210 do.  Used when k is too large for the in-kernel
// selection, when the code width has no fast-scan instantiation, and for queries whose float32
// selection could not be certified.
#pragma once
#include "common.cuh"
#include "plan.cuh"
#include "select.cuh"

// keys[i] = float64 distance bits, vals[i] = retrieval position, i enumerates the local candidates of
// query q in retrieval order (so a stable sort by key reproduces sorted(..., key=dist)).
// dynamic smem: pre[nv+1] int64 (exclusive prefix of local cell sizes over the visits)
__global__ void __launch_bounds__(256)
k_exact_dist(ModelView mv, IndexView ix, PlanView pv, int q, const double* __restrict__ lut64,
             unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals, int64_t nloc) {
    extern __shared__ int64_t sm_pre[];
    const int nv = pv.nvis[q];
    const int64_t o = (int64_t)q * pv.maxvis;
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int t = 0; t < nv; ++t) {
            sm_pre[t] = acc;
            if (pv.vis_pbase[o + t] >= 0) acc += ix.lsize[pv.vis_cell[o + t]];
        }
        sm_pre[nv] = acc;
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = nv;                 // last t with pre[t] <= i (skipping empty visits: pre[t+1] > i)
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm_pre[mid] <= i) lo = mid; else hi = mid; }
        const int t = lo;
        const int64_t idx = i - sm_pre[t];
        const int cell = pv.vis_cell[o + t];
        const uint8_t* code = ix.codes + (ix.cell_start[cell] + idx) * mv.MP;
        const double* l0 = lut64 + (int64_t)pv.vis_lut0[o + t] * mv.m * mv.K;
        const double* l1 = lut64 + (int64_t)pv.vis_lut1[o + t] * mv.m * mv.K;
        double acc = 0.0;
        for (int j = 0; j < mv.M; ++j) {
            const int cb = code_byte(code, idx, j, mv.SW);
            const double e = (j < mv.m) ? l0[j * mv.K + cb] : l1[(j - mv.m) * mv.K + cb];
            acc = (j == 0) ? e : __dadd_rn(acc, e);
        }
        keys[i] = (unsigned long long)__double_as_longlong(acc);
        vals[i] = (unsigned int)(pv.vis_base[o + t] + idx);
    }
}

__global__ void __launch_bounds__(256)
k_exact_emit(ModelView mv, IndexView ix, PlanView pv, int q, int qout, int nq_out, int k,
             const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals, int64_t nloc, void* recbuf) {
    RecView rv = rec_view(recbuf, nq_out, k, mv.M);
    const int nv = pv.nvis[q];
    const int64_t o = (int64_t)q * pv.maxvis;
    const int nout = (int)min((int64_t)k, nloc);
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        const unsigned int pos = vals[i];
        int v = -1;
        for (int t = 0; t < nv; ++t) {
            if (pv.vis_pbase[o + t] >= 0) {
                const int64_t b = pv.vis_base[o + t];
                if ((int64_t)pos >= b && (int64_t)pos < b + ix.lsize[pv.vis_cell[o + t]]) { v = t; break; }
            }
        }
        const int cell = pv.vis_cell[o + v];
        const int64_t incell = (int64_t)pos - pv.vis_base[o + v];
        const int64_t row = ix.cell_start[cell] + incell;
        const int64_t e = (int64_t)qout * k + i;
        rv.d64[e] = __longlong_as_double((long long)keys[i]);
        rv.pos[e] = pos;
        rv.rowid[e] = ix.rowids[row];
        rv.cell[e] = cell;
        for (int j = 0; j < mv.M; ++j) rv.fine[e * mv.M + j] = code_byte(ix.codes + row * mv.MP, incell, j, mv.SW);
    }
    if (threadIdx.x == 0) {
        rv.lb[qout] = __longlong_as_double(0x7FF0000000000000ll);
        rv.count[qout] = nout;
        rv.visited[qout] = nv;
        rv.ncand[qout] = pv.ncand[q];
    }
}
