// ADC code scan (LOPQSearcherBase.compute_distances, search.py:137-177, + the top-`limit` cut of
// search.py:210-215) -- the hot kernel.
//
// Work item = (cell segment of <= segc codes, group of NS = 2*G queries visiting that cell), G = 32/MP.
//
// Shared-memory "super LUT": 256 rows (code byte) x 64 float columns (row = 256 B, bank = column % 32):
//     column x*32 + g*MP + j  holds  LUT_{slot x*G+g}[j][row]     (j < M; columns j >= M stay zero)
// Lane l of a warp owns slot pair (g = l / MP; x = 0, 1) and, per pass, code `jl = l % MP` of a block of MP
// consecutive codes.  At step s it looks up sub-quantizer j = jl ^ s, so the 32 lanes of a warp hit 32 distinct
// banks at every step (conflict free), and the x = 1 look-up is the same address + 128 B.
//
// Code rows are stored pre-swizzled in HBM (index.cuh): stored byte s of in-cell row i = code byte (i % MP) ^ s.
// A lane therefore loads ITS code row straight from global memory into registers (one coalesced 16-byte load
// for M = 16), and step s consumes byte s of those registers -- a compile-time position.  One PRMT builds the
// complete shared-memory offset  (code byte << 8) | column offset  from the code word and a per-lane constant
// register, so a look-up is PRMT + LDS + FADD for the first slot and LDS + FADD for the second:
// 2.5 instructions and one conflict-free shared-memory wavefront per 32 look-ups x 2.  Shared memory carries
// only LUT gathers (the resource that bounds this kernel); codes never pass through it.
//
// Top-k' preselection without sorting or block barriers: every lane tracks the minimum distance it has seen per
// slot; the lanes of a block that serve one slot see disjoint codes, so after grouping the lane minima into
// KP groups, max(group minima) is an upper bound on the slot's KP-th best distance.  Warps refresh that bound
// at exponentially spaced checkpoints, share it through shared memory and -- across blocks working on other
// segments / cells of the same query -- through gthr[q] in global memory (atomicMin).  A (code, query) pair
// whose float32 distance is <= the current bound is appended to the query's candidate list in global memory
// (a few hundred per query).  k_select (select.cuh) keeps the KP smallest and re-ranks them in float64.
// Keys are (float32 dist bits << 32 | retrieval position): order = (dist, retrieval order), the order the
// reference's stable sort produces (ties keep retrieval order, search.py:210).
#pragma once
#include "common.cuh"
#include "plan.cuh"

#define SCAN_THREADS 256
#define SCAN_WARPS 8
#define SCAN_CAND_CAP 8192          // candidate keys per query (overflow => the query is re-ranked by the exact path)
#define SCAN_DRY 8                  // chunks per warp evaluated without appending when a query has no bound yet
#define SCAN_NO_BOUND 0x7f7f7f7fu   // "no bound yet" (3.39e38), the memset pattern of gthr

struct ScanArgs {
    const uint8_t* codes;           // [rows][MP], swizzled
    const int64_t* cell_start;      // [ncell]
    const int64_t* lsize;           // [ncell]
    const float* lut32;             // [slots][256][m]
    const unsigned short* lut16;    // [slots][256][m] quantised tables (packed scan, scan_pk.cuh)
    unsigned int qfill;             // code an empty slot is filled with (QMAX)
    unsigned long long* cand;       // [nq][SCAN_CAND_CAP]
    unsigned int* cand_cnt;         // [nq]
    PlanView pv;
    unsigned int* gthr;             // [nq] float bits: smallest proven upper bound on the query's KP-th best distance
    float* gtab;                    // [nq][E] lane-minimum table of the query merged over finished work items
    int ncell, nflat, KP, m, M;
    int E, GEN;                     // bound table: E = MP*SCAN_WARPS*GEN entries per slot (>= KP), GEN generations per lane
    int cand_cap;                   // capacity of a query's candidate list (k_scan1; the others use SCAN_CAND_CAP)
    const unsigned int* qmargin;    // [nq] packed scan: candidates are appended up to bound + qmargin[q] (NULL: 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async(void* dst, const void* src, int bytes) {
    const uint32_t d = smem_u32(dst);
    if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

template <int MP> struct ScanCfg {
    static constexpr int G = 32 / MP;                       // lane groups per warp
    static constexpr int NS = 2 * G;                        // query slots per work item
    static constexpr int W = MP / 4;                        // 32-bit words per code row
    static constexpr int U = (64 / MP) < 4 ? (64 / MP) : 4; // MP-code blocks interleaved per iteration
    static constexpr int CHUNK = U * MP;                    // codes per warp per iteration
    static constexpr int LPS = MP * SCAN_WARPS;             // lanes of a block serving one slot
    static constexpr int LUT_BYTES = B2L_LUT_ROWS * 256;
};

template <int MP>
size_t scan_smem_bytes(int E) {
    typedef ScanCfg<MP> C;
    return (size_t)C::LUT_BYTES + (size_t)C::NS * E * 4 + 64 * 4 + 256;
}

// one code row -> registers
template <int W> __device__ __forceinline__ void load_row(const uint8_t* p, uint32_t (&w)[W]) {
    if (W == 1) { w[0] = __ldg((const unsigned int*)p); }
    else if (W == 2) { const uint2 v = __ldg((const uint2*)p); w[0] = v.x; w[1] = v.y; }
    else {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            const uint4 v = __ldg((const uint4*)p + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
}

// shared-memory byte offset of a look-up: byte 1 <- code byte b of `word`, byte 0 <- byte b of `cc` (column offset,
// always < 128, so its replicated sign gives the two zero upper bytes)
template <int B> __device__ __forceinline__ uint32_t lut_offset(uint32_t word, uint32_t cc) {
    uint32_t r;
    constexpr uint32_t sel = (4u + B) | ((uint32_t)B << 4) | ((12u + B) << 8) | ((12u + B) << 12);
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(cc), "n"(sel));
    return r;
}

// The super LUT sits at offset 0 of the kernel's dynamic shared memory, which (no static shared memory, no
// cluster) is shared-window address SCAN_LUT_SADDR: the 1 KB the system reserves per CTA on sm_100.  k_scan checks
// this at entry and traps otherwise.  Knowing it at compile time lets the base ride in the LDS immediate, so the
// PRMT result is the complete address register.
#define SCAN_LUT_SADDR 1024
template <int OFF> __device__ __forceinline__ float lds_lut(uint32_t o) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(o), "n"(SCAN_LUT_SADDR + OFF));
    return v;
}

template <int MP>
__device__ __forceinline__ void adc_block(const uint32_t (&w)[MP / 4], const uint32_t (&cc)[MP / 4], float& a0, float& a1) {
#pragma unroll
    for (int T = 0; T < MP / 4; ++T) {
        uint32_t o;
        o = lut_offset<0>(w[T], cc[T]);
        if (T == 0) { a0 = lds_lut<0>(o); a1 = lds_lut<128>(o); } else { a0 += lds_lut<0>(o); a1 += lds_lut<128>(o); }
        o = lut_offset<1>(w[T], cc[T]); a0 += lds_lut<0>(o); a1 += lds_lut<128>(o);
        o = lut_offset<2>(w[T], cc[T]); a0 += lds_lut<0>(o); a1 += lds_lut<128>(o);
        o = lut_offset<3>(w[T], cc[T]); a0 += lds_lut<0>(o); a1 += lds_lut<128>(o);
    }
}

// Refresh the per-slot bounds from the lane-minimum table (whole warp; result lands in s_thr / gthr).
// tab[sl][E]: every entry is +inf or the distance of a real candidate of slot sl, distinct entries <-> distinct codes.
template <int MP>
__device__ __forceinline__ void refresh_bounds(const ScanArgs& a, const float* tab, unsigned int* s_thr, const int* s_q, int lane,
                                               unsigned int& gpre) {
    typedef ScanCfg<MP> C;
    const int E = a.E, gs = E / a.KP;          // gs entries per group (power of two), KP groups
    const int epl = E / 32;                    // entries per lane (contiguous)
#pragma unroll 1
    for (int sl = 0; sl < C::NS; ++sl) {
        if (s_q[sl] < 0) continue;             // uniform
        const float* t = tab + sl * E + lane * epl;
        float v;
        if (epl == 4) {                        // common shapes: one 16-byte load per lane
            const float4 f = *(const float4*)t;
            if (gs == 1) v = fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w));
            else if (gs == 2) v = fmaxf(fminf(f.x, f.y), fminf(f.z, f.w));
            else {
                v = fminf(fminf(f.x, f.y), fminf(f.z, f.w));
                for (int o = 1; o < gs / 4; o <<= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
            }
        } else if (gs <= epl) {                // whole groups inside the lane's run: max of group minima
            v = 0.0f;
            for (int e0 = 0; e0 < epl; e0 += gs) {
                float mn = t[e0];
                for (int e = 1; e < gs; ++e) mn = fminf(mn, t[e0 + e]);
                v = fmaxf(v, mn);
            }
        } else {                               // a group spans gs/epl lanes
            v = t[0];
            for (int e = 1; e < epl; ++e) v = fminf(v, t[e]);
            for (int o = 1; o < gs / epl; o <<= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0 && v < 3.0e38f) {
            const unsigned int bits = __float_as_uint(v);
            const unsigned int old = atomicMin(&s_thr[sl], bits);
            if (bits < old) atomicMin(&a.gthr[s_q[sl]], bits);     // every bound a lane may filter with is published
        }
    }
    // pull in what other blocks proved for the same queries: the value loaded at the previous refresh is merged now,
    // the next one is requested (its latency hides behind the chunks evaluated until the next refresh)
    if (lane < C::NS && s_q[lane] >= 0) {
        atomicMin(&s_thr[lane], gpre);
        gpre = *(volatile unsigned int*)&a.gthr[s_q[lane]];
    }
    __syncwarp();
}

template <int MP>
__global__ void __launch_bounds__(SCAN_THREADS, 2)
k_scan(ScanArgs a) {
    typedef ScanCfg<MP> C;
    constexpr int G = C::G, NS = C::NS, W = C::W, U = C::U, CHUNK = C::CHUNK;
    extern __shared__ __align__(256) unsigned char smem[];
    float* lut = (float*)smem;
    float* tab = (float*)(smem + C::LUT_BYTES);                    // [NS][E]
    unsigned int* s_thr = (unsigned int*)(tab + NS * a.E);         // [NS] float bits (16 slots max)
    int* s_q = (int*)(s_thr + 16);                                 // [NS] query of the slot, -1 = empty
    unsigned int* s_posbase = (unsigned int*)(s_q + 16);           // [NS]
    unsigned int* s_item = s_posbase + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane / MP, jl = lane % MP;
    const PlanView& pv = a.pv;
    if (smem_u32(smem) != SCAN_LUT_SADDR) __trap();
    const float INF = __int_as_float(0x7f800000);

    // per-lane column offsets: byte b of cc[T] = 4 * (g*MP + (jl ^ (4T+b)))
    uint32_t cc[W];
#pragma unroll
    for (int T = 0; T < W; ++T) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(4 * (g * MP + (jl ^ (4 * T + b)))) << (8 * b);
        cc[T] = v;
    }
    const int ent = warp * MP + jl;           // this lane's entry in a slot's bound table (generation 0)

    for (int e = tid; e < B2L_LUT_ROWS * 64; e += SCAN_THREADS) lut[e] = 0.0f;   // padding columns stay zero
    __syncthreads();

    // LUT staging geometry: a half row (m floats of one (query, split) table) is copied in CB-byte chunks
    const int m = a.m;
    const int hb = m * 4;
    const int CB = (hb % 16 == 0) ? 16 : ((hb % 8 == 0) ? 8 : 4);
    const int cph = hb / CB;                  // chunks per half row
    const int cpr = NS * 2 * cph;             // chunks per LUT row
    const bool fill_fast = (SCAN_THREADS % cpr) == 0;      // a thread then copies one fixed column chunk of every rstep-th row
    const int f_r = tid % cpr, f_sl = f_r / (2 * cph), f_r2 = f_r - f_sl * 2 * cph;
    const int f_half = f_r2 / cph, f_part = f_r2 - f_half * cph;
    const int f_row0 = tid / cpr, f_rstep = SCAN_THREADS / cpr;
    const int f_dst = ((f_sl / G) * 32 + (f_sl % G) * MP + f_half * m) * 4 + f_part * CB;

    const unsigned int n_items = pv.cnt->n_items;      // written by k_plan (no host round trip)
    unsigned int nxt = 0;
    if (tid == 0) *s_item = atomicAdd(&pv.cnt->next_item, 1u);
    while (true) {
        __syncthreads();                       // previous item fully consumed (LUT, tables, slot descriptors); s_item published
        const unsigned int item = *s_item;
        if (item >= n_items) break;
        if (tid == 0) nxt = atomicAdd(&pv.cnt->next_item, 1u);     // next item: the round trip overlaps this item

        // ---- decode the item: (segment, cell) by binary search over item_base, then the query group
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

        // ---- slot descriptors + super-LUT fill with cp.async: chunk -> (row, slot, split half, part)
        int lut0[NS], lut1[NS], sq[NS];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const unsigned int pi = group * NS + sl;
            lut0[sl] = -1; lut1[sl] = -1; sq[sl] = -1;
            if (pi < qc) {
                const int2 qv = pv.cellq[pv.cellq_off[cell] + pi];
                const int64_t o = (int64_t)qv.x * pv.maxvis + qv.y;
                lut0[sl] = pv.vis_lut0[o]; lut1[sl] = pv.vis_lut1[o]; sq[sl] = qv.x;
                if (tid == sl) {
                    s_posbase[sl] = (unsigned int)(pv.vis_base[o] + first);
                    s_q[sl] = qv.x;
                    s_thr[sl] = *(volatile unsigned int*)&a.gthr[qv.x];
                }
            } else if (tid == sl) {
                s_posbase[sl] = 0; s_q[sl] = -1; s_thr[sl] = 0xBF800000u;   // -1.0f: nothing passes
            }
        }
        if (fill_fast) {
            int slot = -1;
#pragma unroll
            for (int t = 0; t < NS; ++t) if (t == f_sl) slot = f_half ? lut1[t] : lut0[t];
            if (slot >= 0) {
                const unsigned char* srcl = (const unsigned char*)(a.lut32 + (size_t)slot * B2L_LUT_ROWS * m) + f_part * CB;
                unsigned char* dstl = (unsigned char*)lut + f_dst;
                for (int row = f_row0; row < B2L_LUT_ROWS; row += f_rstep) cp_async(dstl + row * 256, srcl + (size_t)row * hb, CB);
            }
        } else {
            for (int e = tid; e < B2L_LUT_ROWS * cpr; e += SCAN_THREADS) {
                const int row = e / cpr, r = e - row * cpr;
                const int sl = r / (2 * cph), r2 = r - sl * 2 * cph;
                const int half = r2 / cph, part = r2 - half * cph;
                int slot = -1;
#pragma unroll
                for (int t = 0; t < NS; ++t) if (t == sl) slot = half ? lut1[t] : lut0[t];
                if (slot >= 0)
                    cp_async((unsigned char*)lut + row * 256 + ((sl / G) * 32 + (sl % G) * MP + half * m) * 4 + part * CB,
                             (const unsigned char*)(a.lut32 + ((size_t)slot * B2L_LUT_ROWS + row) * m) + part * CB, CB);
            }
        }
        // bound tables start from what finished items of the same queries left in gtab (disjoint codes: still valid)
        for (int e = tid; e < NS * a.E; e += SCAN_THREADS) {
            const int sl = e / a.E;
            int q = -1;
#pragma unroll
            for (int t = 0; t < NS; ++t) if (t == sl) q = sq[t];
            tab[e] = (q >= 0) ? a.gtab[(size_t)q * a.E + (e - sl * a.E)] : INF;
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        const int nchunk = (count + CHUNK - 1) / CHUNK;           // chunk c is scanned by warp c % SCAN_WARPS
        float mn0 = INF, mn1 = INF;                               // lane minima of slots g and G+g
        float* tab0 = tab + g * a.E;
        float* tab1 = tab + (G + g) * a.E;

        unsigned int gpre = (lane < NS) ? s_thr[lane] : 0u;        // (slots' gthr as just loaded)
        // ---- dry run of the first SCAN_DRY chunks of every warp when some slot has no bound yet: lane minima only
        const bool nobound = (tid < NS) && s_q[tid] >= 0 && s_thr[tid] >= SCAN_NO_BOUND;
        if (__syncthreads_or(nobound)) {
            // (with several generations per lane only the first chunk: it is the one generation 0 will hold)
            const int ndry = a.GEN > 1 ? 1 : SCAN_DRY;
            for (int c = warp, n = 0; c < nchunk && n < ndry; c += SCAN_WARPS, ++n) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t w[W];
                    const int idx = c * CHUNK + u * MP + jl;
                    load_row<W>(src0 + (size_t)idx * MP, w);
                    float a0, a1;
                    adc_block<MP>(w, cc, a0, a1);
                    if (idx < count) { mn0 = fminf(mn0, a0); mn1 = fminf(mn1, a1); }
                }
            }
            tab0[ent] = fminf(tab0[ent], mn0); tab1[ent] = fminf(tab1[ent], mn1);
            __syncthreads();
            refresh_bounds<MP>(a, tab, s_thr, s_q, lane, gpre);
            mn0 = INF; mn1 = INF;
        }

        // ---- main loop: this warp's chunks; the code rows of the next chunk are in flight (register ping-pong)
        // while the current one is evaluated.  Common case per chunk: no lane passes -> one vote, no append code.
        int it = 0, gen = 0;
        auto eval_chunk = [&](const uint32_t (&w)[U][W], int c) {
            const float thr0 = __uint_as_float(s_thr[g]), thr1 = __uint_as_float(s_thr[G + g]);
            float a0[U], a1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) adc_block<MP>(w[u], cc, a0[u], a1[u]);
            const int base = c * CHUNK + jl;
            if (base - jl + CHUNK > count) {                      // last, partial chunk of the segment (warp-uniform)
#pragma unroll
                for (int u = 0; u < U; ++u) if (base + u * MP >= count) { a0[u] = INF; a1[u] = INF; }
            }
            float c0 = a0[0], c1 = a1[0];
#pragma unroll
            for (int u = 1; u < U; ++u) { c0 = fminf(c0, a0[u]); c1 = fminf(c1, a1[u]); }
            mn0 = fminf(mn0, c0); mn1 = fminf(mn1, c1);
            if (__any_sync(0xffffffffu, (c0 <= thr0) | (c1 <= thr1))) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (a0[u] <= thr0) {
                        const int q = s_q[g];
                        const unsigned int n = atomicAdd(&a.cand_cnt[q], 1u);
                        if (n < SCAN_CAND_CAP)
                            a.cand[(size_t)q * SCAN_CAND_CAP + n] = ((unsigned long long)__float_as_uint(a0[u]) << 32) |
                                                                    (unsigned long long)(s_posbase[g] + (unsigned)(base + u * MP));
                    }
                    if (a1[u] <= thr1) {
                        const int q = s_q[G + g];
                        const unsigned int n = atomicAdd(&a.cand_cnt[q], 1u);
                        if (n < SCAN_CAND_CAP)
                            a.cand[(size_t)q * SCAN_CAND_CAP + n] = ((unsigned long long)__float_as_uint(a1[u]) << 32) |
                                                                    (unsigned long long)(s_posbase[G + g] + (unsigned)(base + u * MP));
                    }
                }
            }
            // checkpoints after iterations 0, 1, 3, 7, 15, ... : flush the lane minima (generation slot), refresh bounds
            if (((it + 1) & it) == 0) {
                const int e = ent + C::LPS * (gen & (a.GEN - 1));
                tab0[e] = fminf(tab0[e], mn0); tab1[e] = fminf(tab1[e], mn1);
                if (a.GEN > 1) { mn0 = INF; mn1 = INF; ++gen; }
                __syncwarp();
                refresh_bounds<MP>(a, tab, s_thr, s_q, lane, gpre);
            }
            ++it;
        };
        const uint8_t* lane_src = src0 + (size_t)jl * MP;
        auto load_chunk = [&](uint32_t (&w)[U][W], int c) {
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<W>(lane_src + ((size_t)c * CHUNK + u * MP) * MP, w[u]);
        };
        uint32_t wa[U][W], wb[U][W];
        if (warp < nchunk) load_chunk(wa, warp);
        for (int c = warp; c < nchunk; c += 2 * SCAN_WARPS) {
            const bool more = c + SCAN_WARPS < nchunk;
            if (more) load_chunk(wb, c + SCAN_WARPS);
            eval_chunk(wa, c);
            if (more) {
                if (c + 2 * SCAN_WARPS < nchunk) load_chunk(wa, c + 2 * SCAN_WARPS);
                eval_chunk(wb, c + SCAN_WARPS);
            }
        }
        // final flush: what this warp learned helps the other segments of the queries
        if (warp < nchunk) {
            const int e = ent + C::LPS * (gen & (a.GEN - 1));
            tab0[e] = fminf(tab0[e], mn0); tab1[e] = fminf(tab1[e], mn1);
            __syncwarp();
            refresh_bounds<MP>(a, tab, s_thr, s_q, lane, gpre);
        }
        __syncthreads();
        // leave the useful part of the table (entries at or below the bound) to later items of the same queries
        for (int e = tid; e < NS * a.E; e += SCAN_THREADS) {
            const int sl = e / a.E;
            const int q = s_q[sl];
            const float v = tab[e];
            if (q >= 0 && __float_as_uint(v) <= s_thr[sl]) atomicMin((unsigned int*)&a.gtab[(size_t)q * a.E + (e - sl * a.E)], __float_as_uint(v));
        }
        if (tid == 0) *s_item = nxt;
    }
}
