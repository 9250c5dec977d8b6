// Fine argmin on the 5th-generation tensor cores: predict_fine (model.py:575-602 -> utils.py:33-53) for a whole batch.
//
// The argmin of |p - c_k|^2 over the K sub-centroids is the argmin of the score  s_k = |c_k|^2 / 2 - p.c_k , and the
// scores of 128 rows against all 256 centroids of a sub-quantizer are ONE dense product  S[128, 256] = A[128, KK] . B[256, KK]^T :
//
//     A row = ( p_hi | p_hi | p_lo | 1 1 1 0 | 0.. )          B row k = ( -c_hi | -c_lo | -c_hi | h_hi h_lo h_lo2 0 | 0.. )
//
// with p = p_hi + p_lo, c = c_hi + c_lo split into TF32 pieces (round to nearest) and h = |c_k|^2 / 2 in three pieces: the
// "3 x TF32" scheme, so the product carries ~22 significant bits although the tensor cores multiply 11-bit mantissas.
// It runs as tcgen05.mma.kind::tf32 (UTCHMMA in SASS) with both operands in shared memory (K-major, no swizzle: 8-row x
// 16-byte core matrices) and the 128 x 256 float32 accumulator in TENSOR MEMORY (256 columns per block).  Two blocks share
// an SM (all 512 columns): one reduces its scores while the product of the other runs.  The centroid operand of a
// sub-quantizer (prebuilt on the host in its shared-memory image) arrives by one bulk copy (cp.async.bulk + mbarrier
// complete_tx, UBLKCP; double-buffered where two blocks still fit the SM) and serves FTC_T row tiles; the product signals
// its completion through tcgen05.commit on an mbarrier.
//
// Reduction (tcgen05.ld 32x32b.x32, one accumulator row per thread, two warps per lane quarter each taking 128 columns in
// 32-column pieces, the next piece in flight while the current one is reduced):
//   pass 1: m = min_k s_k                               (FMNMX3: half an instruction per score, four independent chains)
//   pass 2: acc = sum_k [s_k < m + 3E] * (1024 + k)       (FFMA.SAT + FFMA per score, both on the FMA pipe, four sums)
// [.] is evaluated as sat((thr - s) * 2^64), exactly 0 or 1 whenever 2^-39 <= |thr| < 2^62 (two distinct floats that close
// to thr differ by >= 2^-63).  acc in [1024, 1024 + 127] in exactly one column half <=> exactly one score lies below
// m + 3E: that centroid is the float64 argmin, because every score is within
//     E = 16 * 2^-24 * (|p| + max_k |c_k|)^2
// of its exact value (input rounding to float32 2^-24, the TF32 split 2^-22 per operand, the dropped lo.lo term 2^-22, the
// float32 accumulation of the tensor core, one truncation per instruction: together below 8 * 2^-24 (|p| + |c|)^2 -- emulated
// in tests/test_error_bounds.py; the largest error measured on the hardware is 1.41 * 2^-24 (|p| + max|c|)^2 over 1.5 million
// scores of six models and 1.84 * 2^-24 on adversarial rows whose products all have one sign (profiles/dev/
// ftc_adversarial_probe.py: the bias of a truncating accumulator shows), and tests/test_gpu_surface.py keeps it below E / 4).  Everything
// else -- near ties, exact ties, NaNs, out-of-range thresholds -- goes to a list that k_fine_redo evaluates in float64 in
// NumPy's order with the first-minimum rule: the codes are the reference's, bit for bit.
#pragma once
#include "common.cuh"

#define FTC_THREADS 256
#define FTC_TILE 128
#define FTC_T 4                 // row tiles that share one load of a sub-quantizer's centroid operand
#define FTC_ECONST 16.0f

template <int DS> struct FtcGeo {
    static constexpr int KSTEPS = (3 * DS + 8 + 7) / 8;       // kind::tf32 contracts 8 elements per instruction
    static constexpr int NCH = 2 * KSTEPS;                    // 16-byte chunks per operand row
    static constexpr int BIAS_CH = 3 * DS / 4;                // chunk holding the half-norm pieces / the ones
    static constexpr int PARTS = DS / 4;                      // 32-byte pieces of a sub-vector (one per staging thread)
    static constexpr int LBO = 128;                           // bytes between the two K chunks of an instruction
    static constexpr int SBO = NCH * 128;                     // bytes between 8-row groups
    static constexpr int A_BYTES = FTC_TILE * NCH * 16;
    static constexpr int B_BYTES = 256 * NCH * 16;
    static constexpr int NPRE = FTC_TILE * PARTS / FTC_THREADS;                 // staging pieces per thread
    static_assert(FTC_TILE * PARTS % FTC_THREADS == 0, "whole staging pieces per thread");
    // two blocks per SM (256 accumulator columns each: one works on its scores while the product of the other runs);
    // the centroid operand is double-buffered when two such blocks fit the shared memory of an SM
    static constexpr int BBUF = (A_BYTES + 2 * B_BYTES <= 100 * 1024) ? 2 : 1;
    // A | B[BBUF] | p2part[2][PARTS][128] | xmin[2][128] | xacc[2][128] | cmax[64] | barriers + tensor-memory slot
    static constexpr int OFF_B = A_BYTES, OFF_P2 = OFF_B + BBUF * B_BYTES, OFF_XMIN = OFF_P2 + 2 * PARTS * 128 * 4,
                         OFF_XACC = OFF_XMIN + 2 * 128 * 4, OFF_CMAX = OFF_XACC + 2 * 128 * 4, OFF_BAR = OFF_CMAX + 64 * 4,
                         SMEM = OFF_BAR + 64;
};
// byte offset of (row r, chunk c) in an operand image
__host__ __device__ inline int ftc_off(int r, int c, int nch) { return (r >> 3) * (nch * 128) + c * 128 + (r & 7) * 16; }

// round to the nearest TF32 (ties away from zero), as cvt.rna.tf32.f32 does
inline float ftc_tf32_host(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
// Host: the centroid operand images [M][B_BYTES / 4] of a model (subs [M][K][ds] float64, K <= 256; unused rows get a
// huge half norm so that they never win)
template <int DS> void ftc_build_tables(int M, int K, const double* subs, float* out) {
    using G = FtcGeo<DS>;
    for (int j = 0; j < M; ++j) {
        float* img = out + (size_t)j * (G::B_BYTES / 4);
        for (size_t i = 0; i < (size_t)G::B_BYTES / 4; ++i) img[i] = 0.0f;
        for (int k = 0; k < 256; ++k) {
            auto at = [&](int c, int e) -> float& { return img[(ftc_off(k, c, G::NCH) >> 2) + e]; };
            if (k >= K) { at(G::BIAS_CH, 0) = ftc_tf32_host(1e30f); continue; }
            const double* c64 = subs + ((size_t)j * K + k) * DS;
            double n2 = 0.0;
            for (int d = 0; d < DS; ++d) {
                n2 += c64[d] * c64[d];
                const float c32 = (float)c64[d];
                const float hi = ftc_tf32_host(c32), lo = ftc_tf32_host(c32 - hi);
                at(0 * G::PARTS + d / 4, d % 4) = -hi;
                at(1 * G::PARTS + d / 4, d % 4) = -lo;
                at(2 * G::PARTS + d / 4, d % 4) = -hi;
            }
            double hres = 0.5 * n2;
            for (int e = 0; e < 3; ++e) { const float piece = ftc_tf32_host((float)hres); at(G::BIAS_CH, e) = piece; hres -= (double)piece; }
        }
    }
}

struct FtcArgs {
    const double* PX;             // [n][D] float64 projections
    int64_t n;
    uint8_t* fine;                // [n][M]
    const float* tabs;            // [M] centroid operand images
    unsigned long long* redo;     // (row << 8 | j) of the sub-vectors the tensor-core stage could not decide
    unsigned int* nredo;
    unsigned int redo_cap;
    unsigned long long* nguard;
    float* dbg;                   // diagnostic: scores [128][256] of sub-quantizer dbg_j for rows 0..127 (or NULL)
    int dbg_j;
};

namespace ftc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt) : "memory"); }
// bounded wait: a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    unsigned long long t0 = 0;
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (!t0) t0 = t;
        else if (t - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#define FTC_R8(r, o) "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define FTC_W8(r, o) "+r"(r[o]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])
// 32 consecutive accumulator columns of this thread's row
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : FTC_R8(r, 0), FTC_R8(r, 8), FTC_R8(r, 16), FTC_R8(r, 24) : "r"(taddr) : "memory");
}
// wait for the loads in flight; the registers are operands so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : FTC_W8(r, 0), FTC_W8(r, 8), FTC_W8(r, 16), FTC_W8(r, 24)::"memory");
}
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float tf32_rna(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }
__device__ __forceinline__ void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int C>
__device__ __forceinline__ void dump_chunk(float* dst, const uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[C * 32 + i] = __uint_as_float(r[i]);
}
}  // namespace ftc

// unit of a block = (row tile, sub-quantizer): tiles in groups of FTC_T, all M sub-quantizers per group, the tiles of the
// group innermost (they share the centroid operand).  Walked incrementally (no divisions in the loop).
struct FtcUnit {
    int tile, j, jj, t, g0, tg;          // tile (block-local), sub-quantizer, operand sequence number, position in group, group base, group size
    __device__ __forceinline__ bool last_of_j() const { return t == tg - 1; }
    __device__ __forceinline__ void first(int nt) { tile = 0; j = 0; jj = 0; t = 0; g0 = 0; tg = nt < FTC_T ? nt : FTC_T; }
    __device__ __forceinline__ void next(int nt, int M) {
        if (++t < tg) { ++tile; return; }
        t = 0; ++jj;
        if (++j < M) { tile = g0; return; }
        j = 0; g0 += tg; tile = g0; tg = nt - g0 < FTC_T ? nt - g0 : FTC_T;
    }
};

template <int DS>
__global__ void __launch_bounds__(FTC_THREADS, 2) k_fine_tc(ModelView mv, FtcArgs a) {
    using G = FtcGeo<DS>;
    extern __shared__ __align__(1024) unsigned char sm_ftc[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = mv.M;
    // contiguous range of row tiles of this block
    const int64_t ntile = (a.n + FTC_TILE - 1) / FTC_TILE;
    const int64_t t_begin = ntile * blockIdx.x / gridDim.x, t_end = ntile * (blockIdx.x + 1) / gridDim.x;
    const int nt = (int)(t_end - t_begin);
    if (nt == 0) return;
    const int U = nt * M;
    const int njj = ((nt + FTC_T - 1) / FTC_T) * M;

    const uint32_t sbase = ftc::smem_u32(sm_ftc);
    float* p2part = (float*)(sm_ftc + G::OFF_P2);
    float* xmin = (float*)(sm_ftc + G::OFF_XMIN);
    float* xacc = (float*)(sm_ftc + G::OFF_XACC);
    float* cmax = (float*)(sm_ftc + G::OFF_CMAX);
    const uint32_t bar_mma = sbase + G::OFF_BAR, bar_b = bar_mma + 16;     // one product barrier; one barrier per centroid buffer
    uint32_t* tmem_slot = (uint32_t*)(sm_ftc + G::OFF_BAR + 32);

    if (tid == 0) {
        ftc::mbar_init(bar_mma, 1);
        ftc::mbar_init(bar_b, 1); ftc::mbar_init(bar_b + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ftc::smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // constant chunks of the row operand: zeros, and (1, 1, 1, 0) against the three half-norm pieces
    for (int e = tid; e < FTC_TILE * G::NCH; e += FTC_THREADS) {
        const int r = e / G::NCH, c = e % G::NCH;
        const float one = (c == G::BIAS_CH) ? 1.0f : 0.0f;
        *(float4*)(sm_ftc + ftc_off(r, c, G::NCH)) = make_float4(one, one, one, 0.0f);
    }
    for (int j = tid; j < M && j < 64; j += FTC_THREADS) cmax[j] = sqrtf(mv.c2max[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // centroid operand of sequence element jj: buffer jj % BBUF (thread 0)
    auto table_load = [&](int jj) {
        const int b = jj % G::BBUF;
        ftc::bulk_g2s(sbase + G::OFF_B + b * G::B_BYTES, a.tabs + (size_t)(jj % M) * (G::B_BYTES / 4), G::B_BYTES, bar_b + 8 * b);
    };
    if (tid == 0) { table_load(0); if (G::BBUF > 1 && njj > 1) table_load(1); }

    // staging of the row operand: projections float64 -> float32 -> TF32 pieces, in registers one unit ahead
    double pre[G::NPRE][4];
    int s_off[G::NPRE];                     // byte offset of (row, piece) of this thread in the operand image
#pragma unroll
    for (int i = 0; i < G::NPRE; ++i) { const int e = tid + i * FTC_THREADS; s_off[i] = ftc_off(e & 127, e >> 7, G::NCH); }
    auto rows_fetch = [&](const FtcUnit& w) {
#pragma unroll
        for (int i = 0; i < G::NPRE; ++i) {
            const int e = tid + i * FTC_THREADS;
            int64_t row = (t_begin + w.tile) * FTC_TILE + (e & 127);
            if (row >= a.n) row = a.n - 1;
            const double2* src = (const double2*)(a.PX + row * (int64_t)mv.D + (w.j * DS + (e >> 7) * 4));
            const double2 v0 = __ldg(src), v1 = __ldg(src + 1);
            pre[i][0] = v0.x; pre[i][1] = v0.y; pre[i][2] = v1.x; pre[i][3] = v1.y;
        }
    };
    auto rows_stage = [&](int u) {
        float* p2 = p2part + (u & 1) * (G::PARTS * 128);
#pragma unroll
        for (int i = 0; i < G::NPRE; ++i) {
            float hi[4], lo[4], s2 = 0.0f;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float v = (float)pre[i][d];
                s2 = fmaf(v, v, s2);
                hi[d] = ftc::tf32_rna(v);
                lo[d] = ftc::tf32_rna(v - hi[d]);
            }
            const float4 vh = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *(float4*)(sm_ftc + s_off[i]) = vh;
            *(float4*)(sm_ftc + s_off[i] + 1 * G::PARTS * 128) = vh;
            *(float4*)(sm_ftc + s_off[i] + 2 * G::PARTS * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            p2[tid + i * FTC_THREADS] = s2;               // [piece][row]
        }
    };
    // instruction descriptor (cute::UMMA::InstrDescriptor): D float32, A and B TF32, both K-major, N = 256, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da0 = ftc::smem_desc(sbase, G::LBO, G::SBO);

    const int q = warp & 3, ch = warp >> 2;          // tensor-memory lane quarter of this warp; its half of the columns
    const int r_epi = q * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + ch * 128;
    const float UEPS = 5.9604645e-08f;
    const float NBIG = -18446744073709551616.0f;
    unsigned int guards = 0;
    FtcUnit w0, w1;                      // units u and u + 1
    w0.first(nt);
    w1 = w0; w1.next(nt, M);
    rows_fetch(w0);
    for (int u = 0; u < U; ++u) {
        rows_stage(u);                                // the product of unit u-1 (the last reader of the operand) is complete
        if (u + 1 < U) rows_fetch(w1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                              // operand staged; every thread has drained the accumulator of unit u-1
        if (tid == 0) {
            const int b = w0.jj % G::BBUF;
            ftc::mbar_wait(bar_b + 8 * b, (w0.jj / G::BBUF) & 1);              // the centroid operand has landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t db = ftc::smem_desc(sbase + G::OFF_B + b * G::B_BYTES, G::LBO, G::SBO);
#pragma unroll
            for (int ks = 0; ks < G::KSTEPS; ++ks)
                ftc::mma_tf32(tmem, da0 + (uint64_t)(ks * 2 * G::LBO >> 4), db + (uint64_t)(ks * 2 * G::LBO >> 4), idesc, ks > 0);
            ftc::mma_commit(bar_mma);
        }
        __syncwarp();
        // meanwhile: the error bound of this row
        const float* p2 = p2part + (u & 1) * (G::PARTS * 128);
        float pn2 = 0.0f;
#pragma unroll
        for (int pp = 0; pp < G::PARTS; ++pp) pn2 += p2[pp * 128 + r_epi];
        const float sp = sqrtf(pn2) + cmax[w0.j];
        const float E3 = 3.0f * (FTC_ECONST * UEPS * sp * sp * 1.01f + 1e-30f);

        ftc::mbar_wait(bar_mma, u & 1);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // every product that read centroid buffer jj % BBUF is complete: fetch the next operand that goes there
        if (tid == 0 && w0.last_of_j() && w0.jj + G::BBUF < njj) table_load(w0.jj + G::BBUF);
        __syncwarp();

        uint32_t ra[32], rb[32];
        // pass 1: the smallest score (independent chains)
        float m0 = 3.0e38f, m1 = 3.0e38f, m2 = 3.0e38f, m3 = 3.0e38f;
        auto pass1 = [&](const uint32_t (&r)[32]) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                m0 = ftc::min3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = ftc::min3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                m2 = ftc::min3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                m3 = ftc::min3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
            }
        };
        ftc::tmem_ld32(tbase, ra); ftc::tmem_ld32(tbase + 32, rb);
        ftc::tmem_wait(ra);
        pass1(ra);
        ftc::tmem_ld32(tbase + 64, ra);
        ftc::tmem_wait(rb);
        pass1(rb);
        ftc::tmem_ld32(tbase + 96, rb);
        ftc::tmem_wait(ra);
        pass1(ra);
        ftc::tmem_ld32(tbase, ra);                    // (first chunk of pass 2 already under way)
        ftc::tmem_wait(rb);
        pass1(rb);
        ftc::tmem_ld32(tbase + 32, rb);
        float m = fminf(fminf(m0, m1), fminf(m2, m3));
        xmin[ch * 128 + r_epi] = m;
        ftc::bar_named(1 + q, 64);
        m = fminf(m, xmin[(ch ^ 1) * 128 + r_epi]);
        const float thr = m + E3;
        const float thr_big = thr * 18446744073709551616.0f;
        // pass 2: which scores lie below the threshold (independent sums of small integers: exact)
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        const bool dump = a.dbg && w0.j == a.dbg_j && t_begin + w0.tile == 0;
#define FTC_PASS2(r, C)                                                                                          \
        {                                                                                                        \
            if (dump) ftc::dump_chunk<C>(a.dbg + r_epi * 256 + ch * 128, r);                                       \
            _Pragma("unroll") for (int i = 0; i < 32; i += 4) {                                                    \
                a0 = fmaf(ftc::fma_sat(__uint_as_float(r[i]), NBIG, thr_big), (float)(1024 + C * 32 + i), a0);           \
                a1 = fmaf(ftc::fma_sat(__uint_as_float(r[i + 1]), NBIG, thr_big), (float)(1024 + C * 32 + i + 1), a1);   \
                a2 = fmaf(ftc::fma_sat(__uint_as_float(r[i + 2]), NBIG, thr_big), (float)(1024 + C * 32 + i + 2), a2);   \
                a3 = fmaf(ftc::fma_sat(__uint_as_float(r[i + 3]), NBIG, thr_big), (float)(1024 + C * 32 + i + 3), a3);   \
            }                                                                                                    \
        }
        ftc::tmem_wait(ra);
        FTC_PASS2(ra, 0)
        ftc::tmem_ld32(tbase + 64, ra);
        ftc::tmem_wait(rb);
        FTC_PASS2(rb, 1)
        ftc::tmem_ld32(tbase + 96, rb);
        ftc::tmem_wait(ra);
        FTC_PASS2(ra, 2)
        ftc::tmem_wait(rb);
        FTC_PASS2(rb, 3)
#undef FTC_PASS2
        const float acc = (a0 + a1) + (a2 + a3);
        xacc[ch * 128 + r_epi] = acc;
        ftc::bar_named(1 + q, 64);
        if ((lane >> 4) == ch) {                          // each of the two warps of a lane quarter settles 16 of its rows
            const int64_t row = (t_begin + w0.tile) * FTC_TILE + r_epi;
            if (row < a.n) {
                const float athr = fabsf(thr);
                const bool in_range = (athr >= 1.8189894e-12f) && (athr < 4.6116860e18f);
                const float c0 = ch ? xacc[r_epi] : acc, c1 = ch ? acc : xacc[128 + r_epi];
                const float tot = c0 + c1;
                int code = -1;
                if (in_range && tot >= 1024.0f && tot < 1152.0f)      // exactly one score below the threshold
                    code = (c0 != 0.0f ? 0 : 128) + (int)tot - 1024;
                if (code >= 0) a.fine[row * (int64_t)M + w0.j] = (uint8_t)code;
                else {
                    const unsigned int slot = atomicAdd(a.nredo, 1u);
                    if (slot < a.redo_cap) a.redo[slot] = ((unsigned long long)row << 8) | (unsigned long long)w0.j;
                    else {                                                 // list full: exact evaluation here and now
                        const double* p64 = a.PX + row * (int64_t)mv.D + (int64_t)w0.j * DS;
                        double pj[DS];
#pragma unroll
                        for (int d = 0; d < DS; ++d) pj[d] = p64[d];
                        double b64 = 1e300; int bk = 0;
                        for (int k = 0; k < mv.K; ++k) {
                            const double d64 = sqdist_np<double>(pj, mv.subs + ((int64_t)w0.j * mv.K + k) * DS, DS);
                            if (d64 < b64) { b64 = d64; bk = k; }
                        }
                        a.fine[row * (int64_t)M + w0.j] = (uint8_t)bk;
                        ++guards;
                    }
                }
            }
        }
        __syncwarp();
        w0 = w1; w1.next(nt, M);
    }
    if (guards && a.nguard) atomicAdd(a.nguard, (unsigned long long)guards);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// the sub-vectors the tensor-core stage left undecided: float64, NumPy's summation order, first minimum (utils.py:33-53).
// One warp per entry, lane l takes centroids l, l + 32, ...
template <int DS>
__global__ void __launch_bounds__(256) k_fine_redo(ModelView mv, FtcArgs a) {
    const unsigned int cnt = min(*a.nredo, a.redo_cap);
    const int lane = threadIdx.x & 31;
    const unsigned int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    unsigned int done = 0;
    for (unsigned int e = wid; e < cnt; e += nw) {
        const unsigned long long rec = a.redo[e];
        const int64_t row = (int64_t)(rec >> 8);
        const int j = (int)(rec & 255);
        const double* p64 = a.PX + row * (int64_t)mv.D + (int64_t)j * DS;
        double pj[DS];
#pragma unroll
        for (int d = 0; d < DS; ++d) pj[d] = p64[d];
        double best = 1e300; int bk = 0x7fffffff;
        for (int k = lane; k < mv.K; k += 32) {
            const double d64 = sqdist_np<double>(pj, mv.subs + ((int64_t)j * mv.K + k) * DS, DS);
            if (d64 < best) { best = d64; bk = k; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best, off);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
            if (od < best || (od == best && ok < bk)) { best = od; bk = ok; }
        }
        if (lane == 0) { a.fine[row * (int64_t)mv.M + j] = (uint8_t)(bk == 0x7fffffff ? 0 : bk); ++done; }
    }
    if (lane == 0 && done && a.nguard) atomicAdd(a.nguard, (unsigned long long)done);
}
