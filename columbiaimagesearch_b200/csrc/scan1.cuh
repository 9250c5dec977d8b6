// Low-batch ADC scan: ONE query per work item -- the HBM-bound regime (a single query, or a handful: no cross-query reuse
// of a code row, so every ranked code byte comes from HBM once and the roofline is the HBM read bandwidth; SURVEY 8d,
// BASELINE config 3's "ADC scan HBM-roofline run").
//
// Differences from k_scan / k_scan_pk (many queries per cell), which would idle 3/4 .. 7/8 of their table look-ups here:
//   * the 32 lanes of a warp rank 32 DIFFERENT code rows per step (lane = row, all lanes the same query); a look-up is
//     PRMT + LDS + FADD, one conflict-free shared-memory wavefront per 32 (row, sub-quantizer) pairs -- the pipe needs
//     1 clock/SM per 32 code bytes, i.e. <= 9 TB/s at 148 SMs x 1.9 GHz, above the HBM peak, so HBM can be the bound;
//   * float32 tables straight from the LUT kernel (no 16-bit quantisation pass: less latency in front of a tiny batch);
//     table row = 256 bytes, of which columns [0, 32) hold buffer 0 and [32, 64) buffer 1: the tables of the block's NEXT
//     item are fetched (cp.async) into the other half-row while the current item is scanned, so item boundaries cost one
//     barrier, not a table refill;
//   * code rows go global -> registers (coalesced 16-byte loads, two chunks in flight per warp); shared memory carries only
//     table gathers;
//   * candidates (float32 distance <= the running bound) are staged in shared memory and, at the end of an item, filtered by
//     the item's FINAL bound before they are appended to the query's global list: hundreds of blocks start on the same
//     query at once without any bound, and only their best ~KP each are worth keeping.
// Same bound logic as k_scan: lane minima -> KP groups -> max of group minima is an upper bound of the KP-th best; bounds
// are shared across blocks through gthr[q] / gtab[q].
#pragma once
#include "scan.cuh"

#define SCAN1_STAGE 1024          // staged candidates per item (8 bytes each); overflow: direct global append
#define SCAN1_CAND_CAP (1 << 18)  // candidate keys per query in the low-batch regime (nq <= SCAN1_MAX_NQ)
#define SCAN1_MAX_NQ 8
#define SCAN1_PF 4                // L2 prefetch distance, in chunks of the warp

template <int MP>
size_t scan1_smem_bytes(int E) {
    return (size_t)B2L_LUT_ROWS * 256 + (size_t)SCAN1_STAGE * 8 + (size_t)E * 4 + 64 * 4 + 256;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int MP, int OFF>
__device__ __forceinline__ float adc_row1(const uint32_t (&w)[MP / 4], const uint32_t (&cc)[MP / 4]) {
    float a = 0.0f;
#pragma unroll
    for (int T = 0; T < MP / 4; ++T) {
        uint32_t o;
        o = lut_offset<0>(w[T], cc[T]);
        if (T == 0) a = lds_lut<OFF>(o); else a += lds_lut<OFF>(o);
        o = lut_offset<1>(w[T], cc[T]); a += lds_lut<OFF>(o);
        o = lut_offset<2>(w[T], cc[T]); a += lds_lut<OFF>(o);
        o = lut_offset<3>(w[T], cc[T]); a += lds_lut<OFF>(o);
    }
    return a;
}

// -DSCAN1_TRACE (developer builds only, profiles/dev/): %globaltimer stamps of every block's phases, first item only
#ifdef SCAN1_TRACE
__device__ unsigned long long g_scan1_trace[1024 * 8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TR(i) do { if (threadIdx.x == 0 && blockIdx.x < 1024 && !tr_done[i]) { g_scan1_trace[blockIdx.x * 8 + (i)] = gtimer(); tr_done[i] = true; } } while (0)
#else
#define TR(i) do {} while (0)
#endif

struct Item1 {                     // decoded work item
    int cell, q, lut0, lut1, count;
    unsigned int posbase;
    const unsigned char* src;
};

template <int MP>
__global__ void __launch_bounds__(SCAN_THREADS, 3)
k_scan1(ScanArgs a) {
    constexpr int W = MP / 4, U = 4, CHUNK = U * 32, G = 32 / MP;
    extern __shared__ __align__(256) unsigned char smem[];
    float* lut = (float*)smem;                                                   // [256 rows][64 floats]: two interleaved buffers
    unsigned long long* stage = (unsigned long long*)(smem + B2L_LUT_ROWS * 256); // [SCAN1_STAGE] dist bits << 32 | index in segment
    float* tab = (float*)(stage + SCAN1_STAGE);                                  // [E]
    unsigned int* s_misc = (unsigned int*)(tab + a.E);                           // [0] thr bits, [1] staged, [2] item, [3] next item
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int jl = lane % MP, g = lane / MP;
    const PlanView& pv = a.pv;
    if (smem_u32(smem) != SCAN_LUT_SADDR) __trap();
    const float INF = __int_as_float(0x7f800000);
#ifdef SCAN1_TRACE
    bool tr_done[8] = {false, false, false, false, false, false, false, false};
#endif
    TR(0);

    uint32_t cc[W];                                  // byte b of cc[T] = 4 * (g*MP + (jl ^ (4T+b))): column of this lane's look-up
#pragma unroll
    for (int T = 0; T < W; ++T) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(4 * (g * MP + (jl ^ (4 * T + b)))) << (8 * b);
        cc[T] = v;
    }
    const int ent = warp * 32 + lane;                // entry of this lane in the bound table (generation 0)
    const int LPS = 32 * SCAN_WARPS;

    if (a.M < MP) for (int e = tid; e < B2L_LUT_ROWS * 64; e += SCAN_THREADS) lut[e] = 0.0f;   // padding columns of both buffers stay zero
    const unsigned int n_items = pv.cnt->n_items;
    if (tid == 0) { s_misc[2] = atomicAdd(&pv.cnt->next_item, 1u); s_misc[3] = 0xFFFFFFFFu; }
    __syncthreads();

    const int m = a.m, hb = m * 4;                   // bytes of a half row in the k-major tables
    const int CB = (hb % 16 == 0) ? 16 : ((hb % 8 == 0) ? 8 : 4);
    const int cph = hb / CB, cpr = G * 2 * cph;      // chunks per half row / per table row (G copies x 2 halves)

    auto decode = [&](unsigned int item, Item1& it) {
        int lo = 0, hi = a.nflat;
        if (item < pv.item_cap) lo = (int)pv.item_f[item];
        else while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pv.item_base[mid] <= item) lo = mid; else hi = mid; }
        const unsigned int seg = (unsigned)(lo / a.ncell);
        it.cell = lo - (int)seg * a.ncell;
        const unsigned int pi = item - pv.item_base[lo];                  // one query per item
        const int2 qv = pv.cellq[pv.cellq_off[it.cell] + pi];
        const int64_t o = (int64_t)qv.x * pv.maxvis + qv.y;
        const int sc = pv.cell_segc ? pv.cell_segc[it.cell] : pv.segc;
        const int64_t first = (int64_t)seg * sc;
        it.q = qv.x; it.lut0 = pv.vis_lut0[o]; it.lut1 = pv.vis_lut1[o];
        it.count = (int)min((int64_t)sc, a.lsize[it.cell] - first);
        it.posbase = (unsigned int)(pv.vis_base[o] + first);
        it.src = a.codes + (a.cell_start[it.cell] + first) * MP;
    };
    auto fetch_lut = [&](const Item1& it, int buf) {                      // cp.async the item's two half tables into buffer `buf`
        for (int e = tid; e < B2L_LUT_ROWS * cpr; e += SCAN_THREADS) {
            const int row = e / cpr, r = e - row * cpr;
            const int gg = r / (2 * cph), r2 = r - gg * 2 * cph;
            const int half = r2 / cph, part = r2 - half * cph;
            const int slot = half ? it.lut1 : it.lut0;
            cp_async((unsigned char*)lut + row * 256 + buf * 128 + (gg * MP + half * m) * 4 + part * CB,
                     (const unsigned char*)(a.lut32 + ((size_t)slot * B2L_LUT_ROWS + row) * m) + part * CB, CB);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    Item1 cur;
    unsigned int item = s_misc[2];
    int buf = 0, tab_q = -1;                          // tab_q: the query whose lane minima the block's bound table holds
    if (item < n_items) { decode(item, cur); fetch_lut(cur, 0); }
    TR(1);
    while (item < n_items) {
        if (tid == 0) s_misc[3] = atomicAdd(&pv.cnt->next_item, 1u);
        // bound table: what finished items of the query left in gtab -- unless the block's previous item was the same query
        // (then its own table, which only gets better, stays: hundreds of blocks share one query in this regime, and a
        // read-modify-write of the query's global table per item would serialise them on the same 256 addresses)
        if (tab_q != cur.q) for (int e = tid; e < a.E; e += SCAN_THREADS) tab[e] = a.gtab[(size_t)cur.q * a.E + e];
        if (tid == 0) { s_misc[0] = *(volatile unsigned int*)&a.gthr[cur.q]; s_misc[1] = 0u; }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                                  // tables of `cur` landed; s_misc published
        TR(2);
        const unsigned int nxt = s_misc[3];
        Item1 nx;
        if (nxt < n_items) { decode(nxt, nx); fetch_lut(nx, buf ^ 1); }   // overlaps the scan below

        const int count = cur.count;
        const int nchunk = (count + CHUNK - 1) / CHUNK;
        float mn = INF;
        int it_n = 0, gen = 0;
        unsigned int gpre = s_misc[0];
        const uint8_t* lane_src = cur.src + (size_t)lane * MP;

        auto refresh = [&]() {                                             // whole warp: bound from the lane-minimum table
            const int E = a.E, gs = E / a.KP, epl = E / 32;             // epl = 8 * GEN: a multiple of 4
            // the lane's run of the table comes in with 16-byte loads (scalar loads at a 32-byte lane stride would be 8-way
            // bank conflicts: as many wavefronts per refresh as 64 rows of look-ups)
            float tv[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
                if (e4 * 4 < epl) {
                    const float4 f = *(const float4*)(tab + lane * epl + e4 * 4);
                    tv[e4 * 4] = f.x; tv[e4 * 4 + 1] = f.y; tv[e4 * 4 + 2] = f.z; tv[e4 * 4 + 3] = f.w;
                }
            }
            const float* t = tv;
            float v;
            if (gs <= epl) {
                v = 0.0f;
                for (int e0 = 0; e0 < epl; e0 += gs) {
                    float mnv = t[e0];
                    for (int e = 1; e < gs; ++e) mnv = fminf(mnv, t[e0 + e]);
                    v = fmaxf(v, mnv);
                }
            } else {
                v = t[0];
                for (int e = 1; e < epl; ++e) v = fminf(v, t[e]);
                for (int o = 1; o < gs / epl; o <<= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) {
                if (v < 3.0e38f) {
                    const unsigned int bits = __float_as_uint(v);
                    const unsigned int old = atomicMin(&s_misc[0], bits);
                    if (bits < old) atomicMin(&a.gthr[cur.q], bits);
                }
                atomicMin(&s_misc[0], gpre);                               // what other blocks proved (loaded at the previous refresh)
                gpre = *(volatile unsigned int*)&a.gthr[cur.q];
            }
            __syncwarp();
        };
        auto load_chunk = [&](uint32_t (&w)[U][W], int c) {
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<W>(lane_src + ((size_t)c * CHUNK + u * 32) * MP, w[u]);
        };
        auto eval_chunk = [&](uint32_t (&w)[U][W], int c) {
            const float thr = __uint_as_float(s_misc[0]);
            float d[U];
            if (buf) {
#pragma unroll
                for (int u = 0; u < U; ++u) d[u] = adc_row1<MP, 128>(w[u], cc);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) d[u] = adc_row1<MP, 0>(w[u], cc);
            }
            if (c + SCAN_WARPS < nchunk) load_chunk(w, c + SCAN_WARPS);   // registers are dead: next chunk's rows in flight
            // ... and the warp's chunks after that are pulled into L2 (one 128-byte line per lane), so that the register loads
            // above find them there: with one chunk per warp in flight a single query would be bound by the HBM latency
            {
                const int cp = c + SCAN1_PF * SCAN_WARPS;
                if (cp < nchunk && lane * 128 < CHUNK * MP) prefetch_l2(cur.src + (size_t)cp * CHUNK * MP + lane * 128);
            }
            const int base = c * CHUNK + lane;
            if (c * CHUNK + CHUNK > count) {
#pragma unroll
                for (int u = 0; u < U; ++u) if (base + u * 32 >= count) d[u] = INF;
            }
            float cm = d[0];
#pragma unroll
            for (int u = 1; u < U; ++u) cm = fminf(cm, d[u]);
            mn = fminf(mn, cm);
            if (__any_sync(0xffffffffu, cm <= thr)) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (d[u] <= thr) {
                        const unsigned int n = atomicAdd(&s_misc[1], 1u);
                        const unsigned long long key = ((unsigned long long)__float_as_uint(d[u]) << 32) | (unsigned long long)(unsigned)(base + u * 32);
                        if (n < SCAN1_STAGE) stage[n] = key;
                        else {
                            const unsigned int gn = atomicAdd(&a.cand_cnt[cur.q], 1u);
                            if (gn < (unsigned)a.cand_cap) a.cand[(size_t)cur.q * a.cand_cap + gn] = key + cur.posbase;
                        }
                    }
                }
            }
            const int cx = it_n + 1, cy = cx & (cx - 1);                  // checkpoints after 1, 2, 3, 4, 6, 8, 12, 16, ... chunks
            if (cy == 0 || ((cy & (cy - 1)) == 0 && (cx - cy) * 2 == cy)) {
                const int e = ent + LPS * (gen & (a.GEN - 1));
                tab[e] = fminf(tab[e], mn);
                if (a.GEN > 1) { mn = INF; ++gen; }
                __syncwarp();
                refresh();
            }
            ++it_n;
        };
        uint32_t wa[U][W];
        if (warp < nchunk) load_chunk(wa, warp);
#ifdef SCAN1_TRACE
        if (tid == 0 && !tr_done[3]) { volatile uint32_t sink = wa[0][0]; (void)sink; }
#endif
        TR(3);
#pragma unroll
        for (int pf = 1; pf < SCAN1_PF; ++pf) {
            const int cp = warp + pf * SCAN_WARPS;
            if (cp < nchunk && lane * 128 < CHUNK * MP) prefetch_l2(cur.src + (size_t)cp * CHUNK * MP + lane * 128);
        }
        if (s_misc[0] >= SCAN_NO_BOUND) {                                  // (block-uniform: read after the barrier above)
            // no bound yet (the first items of a query all start at once): lane minima of the warp's first chunk only, then a
            // first bound; the chunk is evaluated again by the main loop, which then appends against that bound
            if (warp < nchunk) {
                float d0 = INF;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float dv = buf ? adc_row1<MP, 128>(wa[u], cc) : adc_row1<MP, 0>(wa[u], cc);
                    if (warp * CHUNK + u * 32 + lane < count) d0 = fminf(d0, dv);
                }
                tab[ent] = fminf(tab[ent], d0);
            }
            __syncthreads();
            refresh();
            __syncthreads();
        }
        TR(4);
        for (int c = warp; c < nchunk; c += SCAN_WARPS) eval_chunk(wa, c);
        if (warp < nchunk) {
            const int e = ent + LPS * (gen & (a.GEN - 1));
            tab[e] = fminf(tab[e], mn);
        }
        __syncthreads();
        TR(5);
        refresh();                                                         // the item's final bound (every warp computes the same)
        __syncthreads();
        TR(6);
        // staged candidates at or below the final bound -> the query's global list
        {
            const unsigned int thr = s_misc[0];
            const unsigned int ns = min(s_misc[1], (unsigned int)SCAN1_STAGE);
            for (unsigned int i0 = 0; i0 < ns; i0 += SCAN_THREADS) {
                const unsigned int i = i0 + tid;
                const bool keep = i < ns && (unsigned int)(stage[i] >> 32) <= thr;
                const unsigned int bal = __ballot_sync(0xffffffffu, keep);
                unsigned int gb = 0;
                if (lane == 0 && bal) gb = atomicAdd(&a.cand_cnt[cur.q], (unsigned)__popc(bal));
                gb = __shfl_sync(0xffffffffu, gb, 0) + __popc(bal & ((1u << lane) - 1u));
                if (keep && gb < (unsigned)a.cand_cap) a.cand[(size_t)cur.q * a.cand_cap + gb] = stage[i] + cur.posbase;
            }
            // leave the table to other blocks only when this block moves on to another query (or stops)
            // (and only if some item can still start later: with at most one item per resident block nobody would read it,
            //  and hundreds of blocks finishing together would queue up on the query's 256 table words)
            if (!(nxt < n_items && nx.q == cur.q) && n_items > gridDim.x) {
                for (int e = tid; e < a.E; e += SCAN_THREADS) {
                    const float v = tab[e];
                    if (__float_as_uint(v) <= thr) atomicMin((unsigned int*)&a.gtab[(size_t)cur.q * a.E + e], __float_as_uint(v));
                }
            }
        }
        __syncthreads();                                                   // stage / tab / s_misc free for the next item
        tab_q = cur.q;
        item = nxt;
        cur = nx;
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    TR(7);
}
