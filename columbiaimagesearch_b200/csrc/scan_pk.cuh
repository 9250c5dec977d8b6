// Packed ADC code scan: the kernel of scan.cuh with 16-bit quantised tables, two queries per 32-bit LUT word.
//
// Same work items, same swizzled register-resident code rows, same PRMT-built addresses and the same barrier-free
// lane-minimum bounds as k_scan (see scan.cuh).  The difference is the table: entry (row, column x*32 + g*MP + j)
// holds  code16_{slot A}[j][row] | code16_{slot B}[j][row] << 16  for the slot pair (A, B) = (2p, 2p+1), p = x*G + g,
// with the 16-bit codes of plan.cuh (k_lut_quant; a sum of M codes fits 16 bits).  One LDS therefore serves two
// queries and one integer add accumulates both sums: a work item carries NS = 4*G queries and a look-up costs
// PRMT + 2 LDS + 2 IADD per FOUR (code, query) pairs -- half the shared-memory wavefronts of the float32 scan, the
// resource that bounds it.  The integer sum S is a lower-bound key of the float32 distance
// (d >= B[q] + Delta[q] * (S - 0.1)); k_select keeps the KP smallest S per query, re-ranks them in float64 and
// certifies the first k against the bound of everything not kept.  Queries that cannot be certified are re-run
// with the float32 tables, then (ties) exactly.
#pragma once
#include "scan.cuh"

#define SCAN_STAGE 176    // candidates staged in shared memory per slot and work item, 4 bytes each (overflow: direct global append)
template <int MP>
size_t scan_pk_smem_bytes(int E) {
    typedef ScanCfg<MP> C;
    return (size_t)C::LUT_BYTES + (size_t)4 * C::G * SCAN_STAGE * 4 + (size_t)4 * C::G * E * 4 + 7 * 32 * 4 + 256;
}

template <int OFF> __device__ __forceinline__ uint32_t lds_lut_u(uint32_t o) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(o), "n"(SCAN_LUT_SADDR + OFF));
    return v;
}

template <int MP>
__device__ __forceinline__ void adc_block_pk(const uint32_t (&w)[MP / 4], const uint32_t (&cc)[MP / 4], uint32_t& a0, uint32_t& a1) {
#pragma unroll
    for (int T = 0; T < MP / 4; ++T) {
        uint32_t o;
        o = lut_offset<0>(w[T], cc[T]);
        if (T == 0) { a0 = lds_lut_u<0>(o); a1 = lds_lut_u<128>(o); } else { a0 += lds_lut_u<0>(o); a1 += lds_lut_u<128>(o); }
        o = lut_offset<1>(w[T], cc[T]); a0 += lds_lut_u<0>(o); a1 += lds_lut_u<128>(o);
        o = lut_offset<2>(w[T], cc[T]); a0 += lds_lut_u<0>(o); a1 += lds_lut_u<128>(o);
        o = lut_offset<3>(w[T], cc[T]); a0 += lds_lut_u<0>(o); a1 += lds_lut_u<128>(o);
    }
}

// Bounds from the (unsigned) lane-minimum tables.  Every table entry is the sum of a distinct real candidate (or a
// "nothing yet" pattern above 65535), so the KP-th smallest entry is an upper bound on the slot's KP-th best sum.  The
// sums are 16-bit integers: the warp finds that entry exactly by a 16-step bisection on the value (count of entries
// <= pivot, one warp-wide integer reduction per step) -- about 3.5x tighter than the max-of-group-minima bound of the
// float32 kernel, i.e. 3.5x fewer candidates appended and handed to k_select.
template <int NS>
__device__ __forceinline__ void refresh_bounds_u(const ScanArgs& a, const unsigned int* tab, unsigned int* s_thr, const int* s_q,
                                                 int lane, int warp, unsigned int& gpre, bool exact) {
    const int E = a.E, epl = E / 32;           // entries per lane (contiguous), 1..16
    const int gs = E / a.KP;                   // cheap variant: KP groups of gs entries, bound = max of the group minima
    // the slots are dealt round-robin to the warps: every warp passes the same checkpoints, so together they refresh all
#pragma unroll 1
    for (int sl = warp; sl < NS; sl += SCAN_WARPS) {
        if (s_q[sl] < 0) continue;             // uniform
        const unsigned int* t = tab + sl * E + lane * epl;
        unsigned int v[16];
        if ((epl & 3) == 0) {                  // 16-byte loads (scalar loads at this lane stride are epl-way bank conflicts)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4)
                if (e4 * 4 < epl) { const uint4 f = *(const uint4*)(t + e4 * 4); v[e4 * 4] = f.x; v[e4 * 4 + 1] = f.y; v[e4 * 4 + 2] = f.z; v[e4 * 4 + 3] = f.w; }
        } else for (int e = 0; e < epl; ++e) v[e] = t[e];
        unsigned int lo = 0u, hi = 0x10000u;   // smallest T in [0, 65536] with count(entries <= T) >= KP; 65536: fewer than KP sums
        if (!exact) {
            unsigned int g;
            if (gs <= epl) {
                g = 0u;
                for (int e0 = 0; e0 < epl; e0 += gs) {
                    unsigned int mn = v[e0];
                    for (int e = 1; e < gs; ++e) mn = min(mn, v[e0 + e]);
                    g = max(g, mn);
                }
            } else {
                g = v[0];
                for (int e = 1; e < epl; ++e) g = min(g, v[e]);
                for (int o = 1; o < gs / epl; o <<= 1) g = min(g, __shfl_xor_sync(0xffffffffu, g, o));
            }
            g = __reduce_max_sync(0xffffffffu, g);
            hi = min(g, 0x10000u);
            lo = hi;
        }
#pragma unroll 1
        for (int step = 0; step < 17; ++step) {
            if (lo >= hi) break;
            const unsigned int mid = (lo + hi) >> 1;
            int c = 0;
            if (epl == 4) c = (v[0] <= mid) + (v[1] <= mid) + (v[2] <= mid) + (v[3] <= mid);
            else for (int e = 0; e < epl; ++e) c += (v[e] <= mid);
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= a.KP) hi = mid; else lo = mid + 1;
        }
        if (lane == 0 && hi < 0x10000u) {
            const unsigned int old = atomicMin(&s_thr[sl], hi);
            if (hi < old) atomicMin(&a.gthr[s_q[sl]], hi);         // every bound a lane may filter with is published
        }
    }
    const int ps = warp + SCAN_WARPS * lane;   // this lane pulls what other blocks proved for one of the warp's slots
    if (ps < NS && s_q[ps] >= 0) {
        atomicMin(&s_thr[ps], gpre);
        gpre = *(volatile unsigned int*)&a.gthr[s_q[ps]];
    }
    __syncwarp();
}

template <int MP>
__global__ void __launch_bounds__(SCAN_THREADS, 3)
k_scan_pk(ScanArgs a) {
    typedef ScanCfg<MP> C;
    constexpr int G = C::G, W = C::W, U = C::U, CHUNK = C::CHUNK;
    constexpr int NS = 4 * G, PR = 2 * G;                          // query slots / slot pairs per work item
    extern __shared__ __align__(256) unsigned char smem[];
    uint32_t* lut = (uint32_t*)smem;
    unsigned int* stage = (unsigned int*)(smem + C::LUT_BYTES);    // [NS][SCAN_STAGE] candidates of this item: S << 16 | index in segment
    unsigned int* tab = stage + NS * SCAN_STAGE;                   // [NS][E]
    unsigned int* s_thr = tab + NS * a.E;                          // [32]
    int* s_q = (int*)(s_thr + 32);                                 // [32] query of the slot, -1 = empty
    unsigned int* s_posbase = (unsigned int*)(s_q + 32);           // [32]
    int* s_lut0 = (int*)(s_posbase + 32);                          // [32] table of split 0 (-1 = empty)
    int* s_lut1 = s_lut0 + 32;                                     // [32]
    unsigned int* s_nst = (unsigned int*)(s_lut1 + 32);            // [32] keys staged per slot
    unsigned int* s_mrg = s_nst + 32;                              // [32] append margin of the slot's query (code units)
    unsigned int* s_item = s_mrg + 32;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane / MP, jl = lane % MP;
    const PlanView& pv = a.pv;
    if (smem_u32(smem) != SCAN_LUT_SADDR) __trap();
    const unsigned int INFU = 0xFFFFFFFFu;

    uint32_t cc[W];
#pragma unroll
    for (int T = 0; T < W; ++T) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(4 * (g * MP + (jl ^ (4 * T + b)))) << (8 * b);
        cc[T] = v;
    }
    const int ent = warp * MP + jl;
    // the four slots of this lane: (x, half-word) -> 2*(x*G + g) + hh
    const int sl0 = 2 * g, sl1 = 2 * g + 1, sl2 = 2 * (G + g), sl3 = 2 * (G + g) + 1;

    for (int e = tid; e < B2L_LUT_ROWS * 64; e += SCAN_THREADS) lut[e] = 0u;     // padding columns stay zero
    __syncthreads();

    const int m = a.m;
    const unsigned int fillw = (unsigned int)a.qfill | ((unsigned int)a.qfill << 16);
    // vector fill: a task copies 8 consecutive sub-quantizers (16 bytes of each of the two tables of a pair)
    const int cpv = m / 8;                                      // 16-byte chunks per half row (0: element-wise fill)
    const int tpr = PR * 2 * (cpv > 0 ? cpv : 1);               // tasks per LUT row
    const bool fill_vec = (m % 8) == 0 && (SCAN_THREADS % tpr) == 0;
    const int f_t = tid % tpr, f_p = f_t / (2 * (cpv > 0 ? cpv : 1)), f_r2 = f_t - f_p * 2 * (cpv > 0 ? cpv : 1);
    const int f_half = f_r2 / (cpv > 0 ? cpv : 1), f_ch = f_r2 - f_half * (cpv > 0 ? cpv : 1);
    const int f_row0 = tid / tpr, f_rstep = SCAN_THREADS / tpr;
    const int f_dst = ((f_p / G) * 32 + (f_p % G) * MP + f_half * m + f_ch * 8) * 4;

    const unsigned int n_items = pv.cnt->n_items;
    unsigned int nxt = 0;
    if (tid == 0) *s_item = atomicAdd(&pv.cnt->next_item, 1u);
    while (true) {
        __syncthreads();
        const unsigned int item = *s_item;
        if (item >= n_items) break;
        if (tid == 0) nxt = atomicAdd(&pv.cnt->next_item, 1u);

        int lo = 0, hi = a.nflat;             // item_base[lo] <= item < item_base[hi], f = seg * ncell + cell
        if (item < pv.item_cap) lo = (int)pv.item_f[item];
        else {
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pv.item_base[mid] <= item) lo = mid; else hi = mid;
            }
        }
        const unsigned int seg = (unsigned)(lo / a.ncell);
        const int cell = lo - (int)seg * a.ncell;
        const unsigned int qc = pv.cell_qcount[cell];
        const unsigned int group = item - pv.item_base[lo];
        const int64_t first = (int64_t)seg * pv.segc;
        const int count = (int)min((int64_t)pv.segc, a.lsize[cell] - first);
        const unsigned char* src0 = a.codes + (a.cell_start[cell] + first) * MP;

        // ---- slot descriptors
        if (tid < NS) {
            const unsigned int pi = group * NS + tid;
            if (pi < qc) {
                const int2 qv = pv.cellq[pv.cellq_off[cell] + pi];
                const int64_t o = (int64_t)qv.x * pv.maxvis + qv.y;
                s_lut0[tid] = pv.vis_lut0[o]; s_lut1[tid] = pv.vis_lut1[o];
                s_posbase[tid] = (unsigned int)(pv.vis_base[o] + first);
                s_q[tid] = qv.x;
                s_thr[tid] = *(volatile unsigned int*)&a.gthr[qv.x];
                s_mrg[tid] = a.qmargin ? a.qmargin[qv.x] : 0u;
                s_nst[tid] = 0u;
            } else {
                s_nst[tid] = 0u;
                s_mrg[tid] = 0u;
                s_lut0[tid] = -1; s_lut1[tid] = -1; s_posbase[tid] = 0; s_q[tid] = -1;
                s_thr[tid] = 0u;                                   // an empty slot sums to M * QMAX > 0: nothing passes
            }
        }
        __syncthreads();
        // ---- super-LUT fill: interleave the 16-bit tables of each slot pair into 32-bit words
        if (fill_vec) {
            const int sa = f_half ? s_lut1[2 * f_p] : s_lut0[2 * f_p];
            const int sb = f_half ? s_lut1[2 * f_p + 1] : s_lut0[2 * f_p + 1];
            const unsigned short* pa = a.lut16 + (size_t)(sa < 0 ? 0 : sa) * B2L_LUT_ROWS * m + f_ch * 8;
            const unsigned short* pb = a.lut16 + (size_t)(sb < 0 ? 0 : sb) * B2L_LUT_ROWS * m + f_ch * 8;
            unsigned char* dstl = (unsigned char*)lut + f_dst;
#pragma unroll 8
            for (int row = f_row0; row < B2L_LUT_ROWS; row += f_rstep) {
                uint4 A = make_uint4(fillw, fillw, fillw, fillw), Bv = A;
                if (sa >= 0) A = __ldg((const uint4*)(pa + (size_t)row * m));
                if (sb >= 0) Bv = __ldg((const uint4*)(pb + (size_t)row * m));
                uint4 o0, o1;
                o0.x = __byte_perm(A.x, Bv.x, 0x5410); o0.y = __byte_perm(A.x, Bv.x, 0x7632);
                o0.z = __byte_perm(A.y, Bv.y, 0x5410); o0.w = __byte_perm(A.y, Bv.y, 0x7632);
                o1.x = __byte_perm(A.z, Bv.z, 0x5410); o1.y = __byte_perm(A.z, Bv.z, 0x7632);
                o1.z = __byte_perm(A.w, Bv.w, 0x5410); o1.w = __byte_perm(A.w, Bv.w, 0x7632);
                *(uint4*)(dstl + row * 256) = o0;
                *(uint4*)(dstl + row * 256 + 16) = o1;
            }
        } else {
            for (int e = tid; e < B2L_LUT_ROWS * PR * a.M; e += SCAN_THREADS) {
                const int row = e / (PR * a.M), r = e - row * (PR * a.M);
                const int p = r / a.M, j = r - p * a.M;
                const int half = j >= m, jj = j - half * m;
                const int sa = half ? s_lut1[2 * p] : s_lut0[2 * p];
                const int sb = half ? s_lut1[2 * p + 1] : s_lut0[2 * p + 1];
                const unsigned int va = sa >= 0 ? a.lut16[((size_t)sa * B2L_LUT_ROWS + row) * m + jj] : a.qfill;
                const unsigned int vb = sb >= 0 ? a.lut16[((size_t)sb * B2L_LUT_ROWS + row) * m + jj] : a.qfill;
                lut[row * 64 + (p / G) * 32 + (p % G) * MP + j] = va | (vb << 16);
            }
        }
        for (int e = tid; e < NS * a.E; e += SCAN_THREADS) {
            const int sl = e / a.E;
            const int q = s_q[sl];
            tab[e] = (q >= 0) ? ((const unsigned int*)a.gtab)[(size_t)q * a.E + (e - sl * a.E)] : INFU;
        }
        __syncthreads();

        const int nchunk = (count + CHUNK - 1) / CHUNK;
        unsigned int mn0 = INFU, mn1 = INFU, mn2 = INFU, mn3 = INFU;     // lane minima of slots sl0..sl3
        unsigned int* tab0 = tab + sl0 * a.E;
        unsigned int* tab1 = tab + sl1 * a.E;
        unsigned int* tab2 = tab + sl2 * a.E;
        unsigned int* tab3 = tab + sl3 * a.E;
        unsigned int gpre = (warp + SCAN_WARPS * lane < NS) ? s_thr[warp + SCAN_WARPS * lane] : 0u;

        // ---- dry run when some slot has no bound yet: lane minima only
        const bool nobound = (tid < NS) && s_q[tid] >= 0 && s_thr[tid] >= SCAN_NO_BOUND;
        if (__syncthreads_or(nobound)) {
            const int ndry = a.GEN > 1 ? 1 : SCAN_DRY;
            for (int c = warp, n = 0; c < nchunk && n < ndry; c += SCAN_WARPS, ++n) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t w[W];
                    const int idx = c * CHUNK + u * MP + jl;
                    load_row<W>(src0 + (size_t)idx * MP, w);
                    uint32_t a0, a1;
                    adc_block_pk<MP>(w, cc, a0, a1);
                    if (idx < count) {
                        mn0 = min(mn0, a0 & 0xFFFFu); mn1 = min(mn1, a0 >> 16);
                        mn2 = min(mn2, a1 & 0xFFFFu); mn3 = min(mn3, a1 >> 16);
                    }
                }
            }
            tab0[ent] = min(tab0[ent], mn0); tab1[ent] = min(tab1[ent], mn1);
            tab2[ent] = min(tab2[ent], mn2); tab3[ent] = min(tab3[ent], mn3);
            __syncthreads();
            refresh_bounds_u<NS>(a, tab, s_thr, s_q, lane, warp, gpre, true);
            __syncthreads();
            mn0 = INFU; mn1 = INFU; mn2 = INFU; mn3 = INFU;
        }

        int it = 0, gen = 0;
        const unsigned int mg0 = s_mrg[sl0], mg1 = s_mrg[sl1], mg2 = s_mrg[sl2], mg3 = s_mrg[sl3];
        // candidates are staged per slot in shared memory (one shared-memory atomic, 4 bytes each) and moved to the query's
        // global list once per work item; only a full staging buffer falls back to the direct global append
        auto append = [&](unsigned int v, int sl, int idx) {
            const unsigned int n = atomicAdd(&s_nst[sl], 1u);
            if (n < SCAN_STAGE) stage[sl * SCAN_STAGE + n] = (v << 16) | (unsigned int)idx;      // S < 2^16, idx < segc <= 2^14
            else {
                const int q = s_q[sl];
                const unsigned int gn = atomicAdd(&a.cand_cnt[q], 1u);
                if (gn < SCAN_CAND_CAP)
                    a.cand[(size_t)q * SCAN_CAND_CAP + gn] = ((unsigned long long)v << 32) | (unsigned long long)(s_posbase[sl] + (unsigned)idx);
            }
        };
        const uint8_t* lane_src = src0 + (size_t)jl * MP;
        auto load_chunk = [&](uint32_t (&w)[U][W], int c) {
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<W>(lane_src + ((size_t)c * CHUNK + u * MP) * MP, w[u]);
        };
        // One chunk: the table sums, then -- as soon as the code registers are dead -- the loads of the warp's next chunk,
        // whose latency hides behind the bound / candidate logic below (three blocks per SM cover the rest).
        auto eval_chunk = [&](uint32_t (&w)[U][W], int c) {
            // append threshold = bound + margin (the bounds themselves are refreshed without it); the lane's slots are two
            // adjacent pairs (2g, 2g+1) and (2(G+g), 2(G+g)+1): two 8-byte loads per chunk, the margins ride in registers
            const uint2 ta = *(const uint2*)&s_thr[sl0];
            const uint2 tb = *(const uint2*)&s_thr[sl2];
            const unsigned int t0 = ta.x + mg0, t1 = ta.y + mg1, t2 = tb.x + mg2, t3 = tb.y + mg3;
            uint32_t a0[U], a1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) adc_block_pk<MP>(w[u], cc, a0[u], a1[u]);
            if (c + SCAN_WARPS < nchunk) load_chunk(w, c + SCAN_WARPS);
            const int base = c * CHUNK + jl;
            unsigned int v0[U], v1[U], v2[U], v3[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { v0[u] = a0[u] & 0xFFFFu; v1[u] = a0[u] >> 16; v2[u] = a1[u] & 0xFFFFu; v3[u] = a1[u] >> 16; }
            if (base - jl + CHUNK > count) {                      // last, partial chunk of the segment (warp-uniform)
#pragma unroll
                for (int u = 0; u < U; ++u) if (base + u * MP >= count) { v0[u] = INFU; v1[u] = INFU; v2[u] = INFU; v3[u] = INFU; }
            }
            unsigned int c0 = v0[0], c1 = v1[0], c2 = v2[0], c3 = v3[0];
#pragma unroll
            for (int u = 1; u < U; ++u) { c0 = min(c0, v0[u]); c1 = min(c1, v1[u]); c2 = min(c2, v2[u]); c3 = min(c3, v3[u]); }
            mn0 = min(mn0, c0); mn1 = min(mn1, c1); mn2 = min(mn2, c2); mn3 = min(mn3, c3);
            if (__any_sync(0xffffffffu, (c0 <= t0) | (c1 <= t1) | (c2 <= t2) | (c3 <= t3))) {
                // (some lane passes in about two chunks out of three, but rarely more than one of its four slots does: test
                //  the slot minimum first, so the warp walks the U codes of a slot only when a lane has a hit there)
                if (c0 <= t0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (v0[u] <= t0) append(v0[u], sl0, base + u * MP);
                }
                if (c1 <= t1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (v1[u] <= t1) append(v1[u], sl1, base + u * MP);
                }
                if (c2 <= t2) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (v2[u] <= t2) append(v2[u], sl2, base + u * MP);
                }
                if (c3 <= t3) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (v3[u] <= t3) append(v3[u], sl3, base + u * MP);
                }
            }
            const int cx = it + 1, cy = cx & (cx - 1);              // checkpoints after 1, 2, 3, 4, 6, 8, 12, 16, 24, ... chunks
            if (cy == 0 || ((cy & (cy - 1)) == 0 && (cx - cy) * 2 == cy)) {
                // checkpoint: every warp flushes its lane minima and recomputes the bounds of its share of the slots; the
                // other slots' bounds are picked up from s_thr
                const int e = ent + C::LPS * (gen & (a.GEN - 1));
                tab0[e] = min(tab0[e], mn0); tab1[e] = min(tab1[e], mn1); tab2[e] = min(tab2[e], mn2); tab3[e] = min(tab3[e], mn3);
                if (a.GEN > 1) { mn0 = INFU; mn1 = INFU; mn2 = INFU; mn3 = INFU; }
                ++gen;
                __syncwarp();
                refresh_bounds_u<NS>(a, tab, s_thr, s_q, lane, warp, gpre, it < 2);
            }
            ++it;
        };
        uint32_t wa[U][W];
        if (warp < nchunk) load_chunk(wa, warp);
        for (int c = warp; c < nchunk; c += SCAN_WARPS) eval_chunk(wa, c);
        if (warp < nchunk) {
            const int e = ent + C::LPS * (gen & (a.GEN - 1));
            tab0[e] = min(tab0[e], mn0); tab1[e] = min(tab1[e], mn1); tab2[e] = min(tab2[e], mn2); tab3[e] = min(tab3[e], mn3);
        }
        __syncthreads();
        refresh_bounds_u<NS>(a, tab, s_thr, s_q, lane, warp, gpre, true);       // what this item proved, for the other items
        __syncthreads();
        // staged candidates -> the queries' global lists (one global atomic per slot)
        for (int sl = warp; sl < NS; sl += SCAN_WARPS) {
            const int q = s_q[sl];
            const unsigned int n = min(s_nst[sl], (unsigned int)SCAN_STAGE);
            if (q < 0 || n == 0) continue;
            unsigned int gbase = 0;
            if (lane == 0) gbase = atomicAdd(&a.cand_cnt[q], n);
            gbase = __shfl_sync(0xffffffffu, gbase, 0);
            for (unsigned int i = lane; i < n; i += 32)
                if (gbase + i < SCAN_CAND_CAP) {
                    const unsigned int e = stage[sl * SCAN_STAGE + i];
                    a.cand[(size_t)q * SCAN_CAND_CAP + gbase + i] =
                        ((unsigned long long)(e >> 16) << 32) | (unsigned long long)(s_posbase[sl] + (e & 0xFFFFu));
                }
        }
        __syncthreads();
        for (int e = tid; e < NS * a.E; e += SCAN_THREADS) {
            const int sl = e / a.E;
            const int q = s_q[sl];
            const unsigned int v = tab[e];
            if (q >= 0 && v <= s_thr[sl]) atomicMin((unsigned int*)&a.gtab[(size_t)q * a.E + (e - sl * a.E)], v);
        }
        if (tid == 0) *s_item = nxt;
    }
}
