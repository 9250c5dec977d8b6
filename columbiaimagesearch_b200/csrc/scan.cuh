// ADC code scan (LOPQSearcherBase.compute_distances, search.py:137-177, + the top-`limit` cut of
// search.py:210-215) -- the hot kernel.
//
// Work item = (cell segment of <= segc codes, group of G = 32/MP queries visiting that cell).
// Shared-memory "super LUT": 256 rows (code byte) x 32 columns (bank = column):
//     column g*MP + j  holds  LUT_{query g}[j][row]      (j < M; columns j >= M are zero padding)
// Lane l of a warp owns query slot g = l / MP and, per pass, code `jl = l % MP` of a block of MP
// codes; at step s it looks up sub-quantizer j = jl ^ s.  So at every step the 32 lanes of a warp
// hit 32 distinct columns = 32 distinct banks (conflict free), and the code words they fetch from
// the TMA-staged tile are conflict free as well (word (jl>>2)^T of code jl; lanes of different
// query slots read the same word = broadcast).  Each lane accumulates its own (code, query) sum in
// a register: no shuffles, no reductions.
// Code tiles are streamed global -> shared with cp.async.bulk (TMA 1-D bulk copy) behind an
// mbarrier ring.  Survivors (dist <= running k'-th best of the slot) are appended to a per-slot
// candidate buffer that is compacted by an in-block bitonic sort when it exceeds 2k'.
// Keys are (float32 dist bits << 32 | retrieval position): order = (dist, retrieval order), the
// order the reference's stable sort produces (ties keep retrieval order, search.py:210).
#pragma once
#include "common.cuh"
#include "plan.cuh"

#define SCAN_THREADS 256
#define SCAN_WARPS 8
#define SCAN_U 4

struct ScanArgs {
    const uint8_t* codes;           // [rows][MP]
    const int64_t* cell_start;      // [ncell]
    const int64_t* lsize;           // [ncell]
    const float* lut32;             // [slots][256][m]
    unsigned long long* partial;    // [n_partial][KP]
    PlanView pv;
    unsigned int* gthr;             // [nq] float bits: smallest k'-th best distance any finished segment of the query reported
    int ncell, nflat, KP, cap, m, M, use_tau;
    unsigned int n_items;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int MP> struct ScanCfg {
    static constexpr int G = 32 / MP;                       // query slots per group
    static constexpr int W = MP / 4;                        // 32-bit words per code row
    static constexpr int TILE = 512;                        // codes per pipeline stage
    static constexpr int NSTAGE = 4;
    static constexpr int CPW = TILE / SCAN_WARPS;           // codes per warp per tile (64)
    static constexpr int U = (CPW / MP) < SCAN_U ? (CPW / MP) : SCAN_U;   // passes interleaved per iteration
    static constexpr int ITERS = CPW / (U * MP);
    static constexpr int STAGE_BYTES = TILE * MP;
    static constexpr int LUT_BYTES = B2L_LUT_ROWS * 32 * 4;
    static constexpr int GROUPS = TILE / MP;                // MP-code groups of one slot in a tile (dry-run threshold)
    static_assert(ITERS >= 1 && CPW % (U * MP) == 0, "tile shape");
};

template <int MP>
size_t scan_smem_bytes(int cap) {
    typedef ScanCfg<MP> C;
    return (size_t)C::LUT_BYTES + (size_t)C::NSTAGE * C::STAGE_BYTES + (size_t)C::G * cap * 8 + 256;
}

__device__ __forceinline__ void cp_async(void* dst, const void* src, int bytes) {
    const uint32_t d = smem_u32(dst);
    if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

// One pass of a warp over its share of a staged tile.  DRY: no appends; returns, per lane, the maximum over
// this lane's MP-code groups of the group minimum (an upper bound on the GROUPS-th best distance of the tile).
template <int MP, bool DRY>
__device__ __forceinline__ float scan_tile(const uint32_t* __restrict__ words, const unsigned char* __restrict__ lutc,
                                           const uint32_t (&lut_off)[MP], const uint32_t (&sel)[4], int warp, int jl, int g,
                                           int tile_first, int count, float thr, unsigned int posbase, int KP2,
                                           int* s_cnt, unsigned long long* cand_g, int& over) {
    typedef ScanCfg<MP> C;
    constexpr int W = C::W, U = C::U;
    float gmax = 0.0f;
#pragma unroll 1
    for (int it = 0; it < C::ITERS; ++it) {
        const int cbase = warp * C::CPW + it * (U * MP) + jl;     // this lane's code of pass 0
        float acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = 0.0f;
#pragma unroll
        for (int T = 0; T < W; ++T) {
            uint32_t wd[U];
#pragma unroll
            for (int u = 0; u < U; ++u) wd[u] = words[(cbase + u * MP) * W + ((jl >> 2) ^ T)];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t c = __byte_perm(wd[u], 0u, sel[b]);
                    acc[u] += *(const float*)(lutc + (lut_off[T * 4 + b] + (c << 7)));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = tile_first + cbase + u * MP;                  // index inside the segment
            if (DRY) {
                // minimum over the MP lanes of this slot (same code block), then running maximum
                unsigned int v = (idx < count) ? __float_as_uint(acc[u]) : 0x7f800000u;
                if (MP == 32) v = __reduce_min_sync(0xffffffffu, v);
                else {
#pragma unroll
                    for (int o = MP / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
                }
                gmax = fmaxf(gmax, __uint_as_float(v));
            } else if (idx < count && acc[u] <= thr) {
                const int slot = atomicAdd(s_cnt, 1);
                over |= (slot >= KP2);
                cand_g[slot] = ((unsigned long long)__float_as_uint(acc[u]) << 32) | (unsigned long long)(posbase + (unsigned)idx);
            }
        }
    }
    return gmax;
}

template <int MP>
__global__ void __launch_bounds__(SCAN_THREADS, 2)
k_scan(ScanArgs a) {
    typedef ScanCfg<MP> C;
    constexpr int G = C::G, TILE = C::TILE, NSTAGE = C::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem[];
    float* lut = (float*)smem;
    unsigned char* stages = smem + C::LUT_BYTES;
    unsigned long long* cand = (unsigned long long*)(stages + NSTAGE * C::STAGE_BYTES);
    uint64_t* bars = (uint64_t*)(cand + (size_t)G * a.cap);       // [NSTAGE]
    int* s_cnt = (int*)(bars + NSTAGE);                            // [G]
    float* s_thr = (float*)(s_cnt + 8);                            // [G]
    unsigned int* s_posbase = (unsigned int*)(s_thr + 8);          // [G]
    int* s_pslot = (int*)(s_posbase + 8);                          // [G]  (-1 = empty slot)
    int* s_q = s_pslot + 8;                                        // [G]  query of the slot
    unsigned int* s_tau = (unsigned int*)(s_q + 8);                // [G]  dry-run threshold (float bits)
    unsigned int* s_item = s_tau + 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane / MP, jl = lane % MP;
    const int KP = a.KP, cap = a.cap;
    const int trigger = cap - TILE;           // compaction when a slot's buffer could overflow in the next tile
    const PlanView& pv = a.pv;

    // per-lane constants of the conflict-free mapping
    uint32_t lut_off[MP];                     // byte offset of column (g*MP + (jl ^ s)) inside a LUT row
#pragma unroll
    for (int s = 0; s < MP; ++s) lut_off[s] = 4u * (uint32_t)(g * MP + (jl ^ s));
    const unsigned char* lutc = (const unsigned char*)lut;
    uint32_t sel[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) sel[b] = 0x4440u | (uint32_t)((jl & 3) ^ b);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < B2L_LUT_ROWS * 32; e += SCAN_THREADS) lut[e] = 0.0f;   // padding columns stay zero
    __syncthreads();

    // LUT staging geometry: a half row (m floats of one (query, split) table) is copied in CB-byte chunks
    const int m = a.m;
    const int hb = m * 4;
    const int CB = (hb % 16 == 0) ? 16 : ((hb % 8 == 0) ? 8 : 4);
    const int cph = hb / CB;                  // chunks per half row
    const int cpr = G * 2 * cph;              // chunks per LUT row

    uint32_t tiles_done = 0;                  // tiles consumed by this block so far (ring position)

    while (true) {
        if (tid == 0) *s_item = atomicAdd(&pv.cnt->next_item, 1u);
        __syncthreads();
        const unsigned int item = *s_item;
        if (item >= a.n_items) break;

        // ---- decode the item: (segment, cell) by binary search over item_base, then the query group
        int lo = 0, hi = a.nflat;             // item_base[lo] <= item < item_base[hi], f = seg * ncell + cell
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pv.item_base[mid] <= item) lo = mid; else hi = mid;
        }
        const unsigned int seg = (unsigned)(lo / a.ncell);
        const int cell = lo - (int)seg * a.ncell;
        const unsigned int qc = pv.cell_qcount[cell];
        const unsigned int group = item - pv.item_base[lo];
        const int64_t first = (int64_t)seg * pv.segc;
        const int count = (int)min((int64_t)pv.segc, a.lsize[cell] - first);
        const int ntiles = (count + TILE - 1) / TILE;
        const unsigned char* src0 = a.codes + (a.cell_start[cell] + first) * MP;

        // ---- producer prologue: fill the ring
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int t = 0; t < min(ntiles, NSTAGE); ++t) {
                const uint32_t st = (tiles_done + t) % NSTAGE;
                const uint32_t bytes = (uint32_t)((min(TILE, count - t * TILE) * MP + 15) & ~15);
                mbar_expect_tx(&bars[st], bytes);
                tma_load_1d(stages + st * C::STAGE_BYTES, src0 + (size_t)t * C::STAGE_BYTES, bytes, &bars[st]);
            }
        }
        // ---- slot descriptors
        int lut0[G], lut1[G];
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            const unsigned int pi = group * G + gg;
            lut0[gg] = -1; lut1[gg] = -1;
            if (pi < qc) {
                const int2 qv = pv.cellq[pv.cellq_off[cell] + pi];
                const int64_t o = (int64_t)qv.x * pv.maxvis + qv.y;
                lut0[gg] = pv.vis_lut0[o]; lut1[gg] = pv.vis_lut1[o];
                if (tid == gg) {
                    s_posbase[gg] = (unsigned int)(pv.vis_base[o] + first);
                    s_pslot[gg] = pv.pbase[qv.x] + pv.vis_pbase[o] + (int)seg;
                    s_cnt[gg] = 0;
                    s_q[gg] = qv.x;
                    s_thr[gg] = __uint_as_float(min(a.gthr[qv.x], 0x7f800000u));     // best k'-th distance seen so far for the query
                    s_tau[gg] = 0u;
                }
            } else if (tid == gg) {
                s_posbase[gg] = 0; s_pslot[gg] = -1; s_cnt[gg] = 0; s_q[gg] = -1; s_thr[gg] = -1.0f; s_tau[gg] = 0u;
            }
        }
        // ---- super-LUT fill with cp.async: chunk e -> (row, slot g, split half, part)
        for (int e = tid; e < B2L_LUT_ROWS * cpr; e += SCAN_THREADS) {
            const int row = e / cpr, r = e - row * cpr;
            const int gg = r / (2 * cph), r2 = r - gg * 2 * cph;
            const int half = r2 / cph, part = r2 - half * cph;
            int slot = -1;
#pragma unroll
            for (int t = 0; t < G; ++t) if (t == gg) slot = half ? lut1[t] : lut0[t];
            if (slot >= 0)
                cp_async((unsigned char*)lut + row * 128 + (gg * MP + half * m) * 4 + part * CB,
                         (const unsigned char*)(a.lut32 + ((size_t)slot * B2L_LUT_ROWS + row) * m) + part * CB, CB);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const unsigned int posbase = s_posbase[g];
        unsigned long long* cand_g = cand + (size_t)g * cap;
        int over = 0;                          // this thread pushed a slot's buffer past the trigger

        // ---- dry run of tile 0: threshold = max over MP-code groups of the group minimum (>= k' groups)
        if (a.use_tau) {
            mbar_wait(&bars[tiles_done % NSTAGE], (tiles_done / NSTAGE) & 1u);
            const float gm = scan_tile<MP, true>((const uint32_t*)(stages + (tiles_done % NSTAGE) * C::STAGE_BYTES), lutc, lut_off, sel,
                                                 warp, jl, g, 0, count, 0.0f, posbase, 0, nullptr, nullptr, over);
            if (jl == 0 && s_pslot[g] >= 0) atomicMax(&s_tau[g], __float_as_uint(gm));
            __syncthreads();
            if (tid < G && s_pslot[tid] >= 0) s_thr[tid] = fminf(s_thr[tid], __uint_as_float(s_tau[tid]));
            __syncthreads();
        }
        float thr = s_thr[g];

        // ---- main loop over the tiles of the segment
        for (int t = 0; t < ntiles; ++t) {
            const uint32_t n = tiles_done + t;
            const uint32_t st = n % NSTAGE;
            mbar_wait(&bars[st], (n / NSTAGE) & 1u);
            scan_tile<MP, false>((const uint32_t*)(stages + st * C::STAGE_BYTES), lutc, lut_off, sel, warp, jl, g, t * TILE, count,
                                 thr, posbase, trigger, &s_cnt[g], cand_g, over);
            // every warp is done with stage st; the OR makes the compaction decision uniform (a slot
            // count read after the barrier could already include appends of warps that ran ahead)
            const int any = __syncthreads_or(over);
            if (tid == 0 && t + NSTAGE < ntiles) {
                const int tn = t + NSTAGE;
                const uint32_t bytes = (uint32_t)((min(TILE, count - tn * TILE) * MP + 15) & ~15);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[st], bytes);
                tma_load_1d(stages + st * C::STAGE_BYTES, src0 + (size_t)tn * C::STAGE_BYTES, bytes, &bars[st]);
            }
            // ---- compaction of candidate buffers that could overflow during the next tile
            if (any) {
                over = 0;
#pragma unroll 1
                for (int gg = 0; gg < G; ++gg) {
                    const int cn = s_cnt[gg];
                    if (cn > trigger) {
                        unsigned long long* buf = cand + (size_t)gg * cap;
                        const int np2 = next_pow2_dev(cn);
                        for (int i = cn + tid; i < np2; i += SCAN_THREADS) buf[i] = B2L_KEY_EMPTY;
                        __syncthreads();
                        bitonic_sort_u64(buf, np2);
                        if (tid == 0) { s_cnt[gg] = KP; s_thr[gg] = fminf(s_thr[gg], __uint_as_float((unsigned)(buf[KP - 1] >> 32))); }
                    }
                }
                __syncthreads();
                thr = s_thr[g];
            }
        }
        tiles_done += ntiles;

        // ---- final selection of the segment's k' best per slot -> partial lists
#pragma unroll 1
        for (int gg = 0; gg < G; ++gg) {
            const int ps = s_pslot[gg];
            if (ps < 0) continue;
            const int cn = s_cnt[gg];
            unsigned long long* buf = cand + (size_t)gg * cap;
            const int np2 = next_pow2_dev(cn < 1 ? 1 : cn);
            for (int i = cn + tid; i < np2; i += SCAN_THREADS) buf[i] = B2L_KEY_EMPTY;
            __syncthreads();
            if (np2 > 1) bitonic_sort_u64(buf, np2);
            unsigned long long* out = a.partial + (size_t)ps * KP;
            for (int i = tid; i < KP; i += SCAN_THREADS) out[i] = (i < cn) ? buf[i] : B2L_KEY_EMPTY;
            // publish the k'-th best distance of this segment: a valid pruning bound for every other segment of the query
            if (tid == 0 && cn >= KP) atomicMin(&a.gthr[s_q[gg]], (unsigned int)(buf[KP - 1] >> 32));
        }
        __syncthreads();                                       // s_* and cand are reused by the next item
    }
}
