// libb200lopq: C-ABI implementation (host orchestration of the kernels in *.cuh).
// See include/b200lopq.h for the contract and the reference functions each entry point replaces.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200lopq.h"
#include "common.cuh"
#include "comm.cuh"
#include "encode.cuh"
#include "fine_tc.cuh"
#include "exact.cuh"
#include "index.cuh"
#include "plan.cuh"
#include "scan.cuh"
#include "scan_pk.cuh"
#include "scan1.cuh"
#include "largev.cuh"
#include "train.cuh"
#include "select.cuh"

#define B2L_ABI_VERSION 3

namespace {

std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool borrowed = false;             // a sibling handle's view of its parent's buffer: never freed or grown here
    void borrow(const DevBuf& o) { p = o.p; cap = o.cap; borrowed = true; }
    cudaError_t reserve(size_t bytes, bool keep = false, cudaStream_t st = 0) {
        if (bytes <= cap) return cudaSuccess;
        if (borrowed) return cudaErrorInvalidValue;
        size_t ncap = std::max(bytes, cap + cap / 2);
        ncap = (ncap + 255) & ~(size_t)255;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) {
            e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (p) cudaFree(p);
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
    template <typename T> T* as() const { return (T*)p; }
};

}  // namespace

struct b2l_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    // per-call records (CUDA events + pinned plan counters): a ring, so that several asynchronous searches can be in
    // flight and still be timed individually
    struct CallRec {
        cudaEvent_t ev[5] = {};
        PlanCounters* h_pc = nullptr;      // pinned
        bool pending = false, has_pc = false;
        int64_t launch0 = 0;
        int nq = 0, segc = 0;              // shape of the call (segment-length feedback)
        b2l_stats st = {};
    };
    static const int NREC = 8;
    CallRec ring[NREC];
    CallRec* cr = nullptr;
    uint64_t seq = 0;
    std::string err;
    std::mutex mu;
    // model
    bool has_model = false, has_pca = false;
    ModelView mv = {};
    DevBuf dCs, dmus, dRt, dsubs, dsubs32, dsubs32T, dc2max, dP, dpmu, dftc, dCs32;
    // index: master copy in insertion order
    int64_t n_items = 0;
    DevBuf m_coarse, m_fine, m_rowid;
    // index: cell-major layout
    bool dirty = true, global_set = false;
    int64_t rows_padded = 0;
    DevBuf codes, rowids, cell_start, lsize, gsize, sorted_first;
    std::vector<int64_t> h_lsize, h_gsize, h_cell_start;
    // large V (V > B2L_MAX_V): sparse cell directory instead of the dense per-cell arrays (largev.cuh)
    DevBuf d_ucell, d_ustart, d_hkeys, d_hvals, w_walk, w_walk2;
    std::vector<unsigned int> h_ucell, h_ustart;
    unsigned int nu = 0, hmask = 0, max_run = 0;
    // workspaces
    DevBuf w_q, w_xq, w_px, w_coarse, w_fine, w_lut32, w_lut64, w_p64, w_cellq, w_cand, w_gtab, w_lut16, w_quant, w_plan, w_sort_a, w_sort_b,
        w_sort_tmp, w_rec, w_rec2, w_out, w_out2, w_misc, w_bkt, w_perm, w_need2, w_segc;
    PlanView pv = {};
    unsigned int* gthr = nullptr;      // [nq] per-query pruning bound shared by the scan blocks
    unsigned int* cand_cnt = nullptr;  // [nq] candidates the scan appended
    b2l_stats stats = {};
    int64_t launches = 0;
    bool async_mode = false;
    int fb_nq = 0, fb_segc = 0;        // last collected fast search: batch size, segment length, work items it produced
    int64_t fb_items = 0;
    float c2m = 0.0f;                  // max_j max_k |subs[j][k]|^2 (upper bound), for the float32 table error model
    float c2sum = 0.0f;                // sum_j max_k |subs[j][k]|^2 (upper bound): |c|^2 of any code (float32 preselection at large V)
    int force_redo = 0;                // test knob: bit 0 / 1 = treat every query as uncertified after the first / second stage
    int fine_mode = 0;                 // fine argmin: 0 tensor-core stage (fine_tc.cuh) where the model allows, else as 2; 1: float64 only;
                                       // 2: float32 SIMT stage + float64 guard
    const float* ftc_tabs = nullptr;   // centroid operand images of the tensor-core stage (NULL: model shape not covered)
    DevBuf w_redo, w_ftc_dbg;          // undecided sub-vectors of the tensor-core stage; diagnostic score dump
    unsigned int* d_nredo = nullptr;
    cudaStream_t copy_stream = nullptr;        // host-resident encode: rows of block i+1 arrive while block i is encoded
    cudaEvent_t ev_in[2] = {}, ev_free[2] = {};
    int ftc_dbg_j = -1;
    unsigned long long* d_nguard = nullptr;   // sub-vectors the guard re-evaluated in float64 (device counter)
    int scan_mode = 0;                 // 0: packed 16-bit tables first (default), 1: float32 tables only
    int kp_min = 0;                    // lower limit of the preselection width KP (0: k + 8 rounded up to a power of two)
    void* h_out = nullptr;             // pinned staging of the search outputs
    size_t h_out_cap = 0;
    // sibling handles (b2l_create_sibling): own stream and workspaces, the parent's model and index
    b2l_ctx* parent = nullptr;
    int n_siblings = 0;
    // multi-GPU exchange (comm.cuh): this rank's window and the peers' windows
    struct Comm {
        int world = 0, rank = 0;
        int64_t max_home = 0;          // largest home slice (queries per rank and batch)
        int max_k = 0;
        size_t qrow = 0;               // bytes of one query row (float32 or float64, D0 wide)
        size_t q_region = 0, r_region = 0;
        DevBuf window;
        unsigned char* peer[COMM_MAX_WORLD] = {};
        bool opened[COMM_MAX_WORLD] = {};
        bool connected = false;
        unsigned long long seq = 0;
        int* d_err = nullptr;          // device word set by a wait that timed out
        unsigned int* d_unc = nullptr; // queries of the current batch this rank could not certify
        DevBuf cnt_out;                // [world] payloads gathered by the last wait of a batch
    } comm;
};

namespace {

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            h->err = b__;                                                                          \
            return B2L_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)
#define FAIL(code, ...)                                      \
    do {                                                     \
        char b__[512];                                       \
        snprintf(b__, sizeof b__, __VA_ARGS__);              \
        h->err = b__;                                        \
        return code;                                         \
    } while (0)
#define LAUNCHED() do { ++h->launches; CU(cudaGetLastError()); } while (0)

inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
inline int grid_for(int64_t n, int threads, int maxblocks = 148 * 16) {
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, maxblocks));
}

template <int MP> int launch_scan(b2l_handle h, const ScanArgs& a) {
    const size_t smem = scan_smem_bytes<MP>(a.E);
    if (smem > 227 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "scan shared memory %zu too large", smem);
    CU(cudaFuncSetAttribute(k_scan<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan<MP>, SCAN_THREADS, smem));
    if (occ < 1) occ = 1;
    const unsigned grid = (unsigned)(h->num_sms * occ);        // persistent; blocks without work leave at once
    k_scan<MP><<<grid, SCAN_THREADS, smem, h->stream>>>(a);
    LAUNCHED();
    return B2L_OK;
}

template <int MP> int launch_scan_pk(b2l_handle h, const ScanArgs& a) {
    const size_t smem = scan_pk_smem_bytes<MP>(a.E);
    if (smem > 227 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "scan shared memory %zu too large", smem);
    CU(cudaFuncSetAttribute(k_scan_pk<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan_pk<MP>, SCAN_THREADS, smem));
    if (occ < 1) occ = 1;
    const unsigned grid = (unsigned)(h->num_sms * occ);
    k_scan_pk<MP><<<grid, SCAN_THREADS, smem, h->stream>>>(a);
    LAUNCHED();
    return B2L_OK;
}

template <int MP> int launch_scan1(b2l_handle h, const ScanArgs& a) {
    const size_t smem = scan1_smem_bytes<MP>(a.E);
    if (smem > 227 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "scan shared memory %zu too large", smem);
    static thread_local size_t cfg_smem = 0;          // attribute + occupancy query once per shared-memory size (host latency
    static thread_local int cfg_occ = 0, cfg_dev = -1; //  in front of a single-query scan is comparable to the scan itself)
    if (cfg_smem != smem || cfg_dev != h->device) {
        CU(cudaFuncSetAttribute(k_scan1<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int o = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_scan1<MP>, SCAN_THREADS, smem));
        cfg_occ = o < 1 ? 1 : o; cfg_smem = smem; cfg_dev = h->device;
    }
    const int occ = cfg_occ;
    const unsigned grid = (unsigned)(h->num_sms * occ);
    k_scan1<MP><<<grid, SCAN_THREADS, smem, h->stream>>>(a);
    LAUNCHED();
    return B2L_OK;
}

// grouped rotation GEMM for h > 64 (encode.cuh): the pipelined kernel, or the plain one when the input is not 16-byte aligned
template <int MODE>
int launch_rotate_g(b2l_handle h, const void* x, int xf64, int64_t n, const unsigned int* cnt, const unsigned int* base,
                    const unsigned int* tile_base, const unsigned int* perm, const int32_t* desc, double* out, unsigned tiles) {
    const ModelView& mv = h->mv;
    const dim3 g2(tiles, (unsigned)(mv.h / 64));
    const bool aligned = ((uintptr_t)x % 16 == 0) && (((size_t)mv.D * (xf64 ? 8 : 4)) % 16 == 0);
    if (aligned) {
        if (xf64) { const size_t sm = rotate_g_smem_bytes<double>();
            CU(cudaFuncSetAttribute(k_rotate_dmma_g<double, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_rotate_dmma_g<double, MODE><<<g2, 128, sm, h->stream>>>(mv, (const double*)x, n, cnt, base, tile_base, perm, desc, out); }
        else { const size_t sm = rotate_g_smem_bytes<float>();
            CU(cudaFuncSetAttribute(k_rotate_dmma_g<float, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_rotate_dmma_g<float, MODE><<<g2, 128, sm, h->stream>>>(mv, (const float*)x, n, cnt, base, tile_base, perm, desc, out); }
    } else {
        const size_t smr = (size_t)2 * 64 * ROT_LD * 8 + 64 * 4;
        if (xf64) { CU(cudaFuncSetAttribute(k_rotate_dmma_g0<double, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
            k_rotate_dmma_g0<double, MODE><<<g2, 128, smr, h->stream>>>(mv, (const double*)x, n, cnt, base, tile_base, perm, desc, out); }
        else { CU(cudaFuncSetAttribute(k_rotate_dmma_g0<float, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
            k_rotate_dmma_g0<float, MODE><<<g2, 128, smr, h->stream>>>(mv, (const float*)x, n, cnt, base, tile_base, perm, desc, out); }
    }
    return B2L_OK;
}

// copy helper honouring on_device
int copy_in(b2l_handle h, void* dst, const void* src, size_t bytes, int on_device) {
    CU(cudaMemcpyAsync(dst, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    return B2L_OK;
}
int copy_out(b2l_handle h, void* dst, const void* src, size_t bytes, int on_device) {
    if (!dst) return B2L_OK;
    CU(cudaMemcpyAsync(dst, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    return B2L_OK;
}

// ---- encode pipeline on device-resident input ----------------------------------------------------
// X device [n][D0|D]; outputs device.  want_px: also keep the projection in w_px (n rows).
int encode_device(b2l_handle h, const void* dX, int x_is_f64, int64_t n, const int32_t* d_coarse_in,
                  int32_t* d_coarse, uint8_t* d_fine) {
    const ModelView& mv = h->mv;
    const void* x = dX;
    int xf64 = x_is_f64;
    if (h->has_pca) {
        CU(h->w_xq.reserve((size_t)n * mv.D * 4));
        const size_t smem = (size_t)(mv.D0 + mv.D) * 8;
        if (x_is_f64) { CU(cudaFuncSetAttribute(k_pca<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pca<double><<<(unsigned)n, 128, smem, h->stream>>>(mv, (const double*)dX, n, h->w_xq.as<float>()); }
        else { CU(cudaFuncSetAttribute(k_pca<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pca<float><<<(unsigned)n, 128, smem, h->stream>>>(mv, (const float*)dX, n, h->w_xq.as<float>()); }
        LAUNCHED();
        x = h->w_xq.p; xf64 = 0;
    }
    CU(h->w_px.reserve((size_t)n * mv.D * 8));
    const unsigned blocks = (unsigned)((n + ENC_WARPS - 1) / ENC_WARPS);
    const size_t smem = (size_t)ENC_WARPS * mv.h * 8;
    // batch encode of the common shape: coarse assignment, then the rotation as a grouped float64 tensor-core GEMM
    const bool gemm = d_fine && !d_coarse_in && d_coarse && mv.h % 64 == 0 && n >= 2048 && n < ((int64_t)1 << 31) &&
                      ((uintptr_t)x % 16 == 0);    // (the h = 64 kernel copies raw rows with 16-byte cp.async)
    // coarse assignment only (utils.predict_cluster over rows; the assignment step of k-means training): no projection
    const bool coarse_only = !d_fine && !d_coarse_in && d_coarse;
    double* px_out = (gemm || coarse_only) ? nullptr : h->w_px.as<double>();
    const size_t cb_coarse = (!xf64 && mv.coarse_f32) ? coarse_c_bytes<float>(mv.V, mv.h) : coarse_c_bytes<double>(mv.V, mv.h);
    // (fine mode 2 keeps the warp-per-row kernel for models whose centroids fit its shared memory: the A/B switch of the probes)
    if ((gemm || coarse_only) && (cb_coarse > 48 * 1024 || h->fine_mode == 0) && (mv.h == 32 || mv.h == 64 || mv.h == 128) && h->d_nredo && h->fine_mode != 1 &&
        n < ((int64_t)1 << 30) && (uintptr_t)x % 16 == 0) {
        // one row per thread, float32 scores against centroid chunks streamed through shared memory (any V: the only path for
        // centroids that do not fit the shared memory of k_coarse_assign), exact arithmetic for the listed near ties
        CU(h->w_redo.reserve((size_t)2 * n * 8));
        CU(cudaMemsetAsync(h->d_nredo, 0, 4, h->stream));
        // two rows per thread once the centroids dominate the shared-memory traffic (h <= 64: the rows fit the registers)
        const int rpt = (mv.V >= 256 && mv.h <= 64) ? 2 : 1;
        const unsigned gb = (unsigned)((n + CBIG_THREADS * rpt - 1) / (CBIG_THREADS * rpt));
        unsigned long long* rl = h->w_redo.as<unsigned long long>();
#define CBIG(XTV, HV, RV)                                                                                       \
    do {                                                                                                        \
        CU(cudaFuncSetAttribute(k_coarse_big<XTV, HV, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbig_smem_bytes<HV>())); \
        k_coarse_big<XTV, HV, RV><<<gb, CBIG_THREADS, cbig_smem_bytes<HV>(), h->stream>>>(mv, (const XTV*)x, n, d_coarse, rl, h->d_nredo); \
    } while (0)
#define CBIG_H(XTV)                                                                                             \
    do {                                                                                                        \
        if (mv.h == 32) { if (rpt == 2) CBIG(XTV, 32, 2); else CBIG(XTV, 32, 1); }                              \
        else if (mv.h == 64) { if (rpt == 2) CBIG(XTV, 64, 2); else CBIG(XTV, 64, 1); }                         \
        else CBIG(XTV, 128, 1);                                                                                 \
    } while (0)
        if (xf64) CBIG_H(double); else CBIG_H(float);
#undef CBIG_H
#undef CBIG
        LAUNCHED();
        if (xf64) k_coarse_redo<double><<<h->num_sms * 2, 256, 0, h->stream>>>(mv, (const double*)x, d_coarse, rl, h->d_nredo);
        else k_coarse_redo<float><<<h->num_sms * 2, 256, 0, h->stream>>>(mv, (const float*)x, d_coarse, rl, h->d_nredo);
    } else
    if ((gemm || coarse_only) && mv.h % 8 == 0 && mv.h <= 128) {
        const size_t tsz = (!xf64 && mv.coarse_f32) ? 4 : 8, xsz = xf64 ? 8 : 4;
        const size_t cb = tsz == 4 ? coarse_c_bytes<float>(mv.V, mv.h) : coarse_c_bytes<double>(mv.V, mv.h);
        const int c_smem = cb <= 48 * 1024 ? 1 : 0;
        const size_t smc = (c_smem ? cb : 0) + (size_t)COARSE_WARPS * mv.D * xsz;
        const unsigned cgrid = (unsigned)std::min<int64_t>((n + COARSE_WARPS - 1) / COARSE_WARPS, (int64_t)h->num_sms * 8);
        if (xf64) { CU(cudaFuncSetAttribute(k_coarse_assign<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
            k_coarse_assign<double><<<cgrid, COARSE_WARPS * 32, smc, h->stream>>>(mv, (const double*)x, n, d_coarse, c_smem); }
        else { CU(cudaFuncSetAttribute(k_coarse_assign<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
            k_coarse_assign<float><<<cgrid, COARSE_WARPS * 32, smc, h->stream>>>(mv, (const float*)x, n, d_coarse, c_smem); }
    } else if (xf64) { CU(cudaFuncSetAttribute(k_coarse_project<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_coarse_project<double><<<blocks, ENC_WARPS * 32, smem, h->stream>>>(mv, (const double*)x, n, d_coarse_in, d_coarse, px_out); }
    else { CU(cudaFuncSetAttribute(k_coarse_project<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_coarse_project<float><<<blocks, ENC_WARPS * 32, smem, h->stream>>>(mv, (const float*)x, n, d_coarse_in, d_coarse, px_out); }
    LAUNCHED();
    if (gemm) {
        const int nb = 2 * mv.V;
        CU(h->w_sort_a.reserve((size_t)2 * n * 4));                       // perm[2][n]
        CU(h->w_misc.reserve((size_t)(4 * nb + 4) * 4));                   // cnt | base | cursor | tile_base
        unsigned int* cnt = h->w_misc.as<unsigned int>();
        unsigned int* base = cnt + nb; unsigned int* cursor = base + nb; unsigned int* tile_base = cursor + nb;
        unsigned int* perm = h->w_sort_a.as<unsigned int>();
        CU(cudaMemsetAsync(cnt, 0, (size_t)nb * 4, h->stream));
        k_enc_hist<<<grid_for(n, 256), 256, 0, h->stream>>>(d_coarse, n, mv.V, cnt);
        LAUNCHED();
        k_enc_offsets<<<1, 1024, 0, h->stream>>>(mv.V, cnt, base, cursor, tile_base);
        LAUNCHED();
        k_enc_scatter<<<grid_for(n, 256), 256, 0, h->stream>>>(d_coarse, n, mv.V, base, cursor, perm);
        LAUNCHED();
        const unsigned tiles = (unsigned)(2 * ((n + 63) / 64) + nb);       // upper bound; surplus blocks leave at once
        if (mv.h != ROT_H) {             // large models (2048-d: h = 1024): 64 x 64 output tiles, contraction in chunks of 64
            int rc = launch_rotate_g<0>(h, x, xf64, n, cnt, base, tile_base, perm, nullptr, h->w_px.as<double>(), tiles);
            if (rc) return rc;
        } else
        {
            // consecutive tiles per block: enough blocks for ~6 per resident slot, at most 8 tiles each
            const int tpb = (int)std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)tiles / ((int64_t)h->num_sms * 18)));
            const unsigned g1 = (tiles + (unsigned)tpb - 1) / (unsigned)tpb;
            if (xf64) { const size_t sm1 = rotate_smem_bytes<double>();
                CU(cudaFuncSetAttribute(k_rotate_dmma<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
                k_rotate_dmma<double><<<g1, 128, sm1, h->stream>>>(mv, (const double*)x, n, cnt, base, tile_base, perm, h->w_px.as<double>(), tpb); }
            else { const size_t sm1 = rotate_smem_bytes<float>();
                CU(cudaFuncSetAttribute(k_rotate_dmma<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
                k_rotate_dmma<float><<<g1, 128, sm1, h->stream>>>(mv, (const float*)x, n, cnt, base, tile_base, perm, h->w_px.as<double>(), tpb); }
        }
        LAUNCHED();
    }
    if (d_fine) {
        int kchunk = mv.K;
        while ((size_t)kchunk * mv.ds * 8 > 96 * 1024 && kchunk > 1) kchunk = (kchunk + 1) / 2;
        const size_t sm2 = (size_t)kchunk * mv.ds * 8;
        const unsigned b2 = (unsigned)((n + FINE_THREADS - 1) / FINE_THREADS);
#define FINE(DSV)                                                                                               \
    do {                                                                                                        \
        CU(cudaFuncSetAttribute(k_fine_argmin<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));     \
        k_fine_argmin<DSV><<<b2, FINE_THREADS, sm2, h->stream>>>(mv, h->w_px.as<double>(), n, d_fine, kchunk);   \
    } while (0)
#define FINE32(DSV, RV)                                                                                         \
    do {                                                                                                        \
        const size_t sma = (size_t)mv.M * mv.K * (DSV + 1) * 4;                                                 \
        if (DSV <= 16 && sma <= 200 * 1024) {                                       \
            CU(cudaFuncSetAttribute(k_fine_argmin32_all<DSV, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sma)); \
            k_fine_argmin32_all<DSV, RV><<<h->num_sms, FINE_ALL_THREADS, sma, h->stream>>>(mv, h->w_px.as<double>(), n, d_fine, h->d_nguard); \
            break;                                                                                              \
        }                                                                                                       \
        const size_t sm3 = (size_t)256 * (DSV + 1) * 4 * 2;                                                     \
        const unsigned b3 = (unsigned)((n + (int64_t)FINE_THREADS * RV - 1) / ((int64_t)FINE_THREADS * RV));     \
        if (sm3 > 48 * 1024) CU(cudaFuncSetAttribute(k_fine_argmin32<DSV, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3)); \
        k_fine_argmin32<DSV, RV><<<b3, FINE_THREADS, sm3, h->stream>>>(mv, h->w_px.as<double>(), n, d_fine, h->d_nguard); \
    } while (0)
        const bool f32stage = h->fine_mode != 1 && n >= 2048;    // float32 first stage + float64 guard (same codes)
        // tensor-core stage (tcgen05, fine_tc.cuh) + float64 list pass for what it cannot decide (same codes)
        if (h->fine_mode == 0 && h->ftc_tabs && h->d_nredo && (n >= 2048 || h->ftc_dbg_j >= 0) && n < ((int64_t)1 << 31)) {
            FtcArgs fa;
            fa.PX = h->w_px.as<double>(); fa.n = n; fa.fine = d_fine; fa.tabs = h->ftc_tabs;
            fa.redo_cap = (unsigned int)std::max<int64_t>(4096, n * mv.M / 64);
            CU(h->w_redo.reserve((size_t)fa.redo_cap * 8));
            fa.redo = h->w_redo.as<unsigned long long>(); fa.nredo = h->d_nredo; fa.nguard = h->d_nguard;
            fa.dbg = h->ftc_dbg_j >= 0 ? h->w_ftc_dbg.as<float>() : nullptr; fa.dbg_j = h->ftc_dbg_j;
            CU(cudaMemsetAsync(h->d_nredo, 0, 4, h->stream));
            const int64_t ntile = (n + FTC_TILE - 1) / FTC_TILE;
            const unsigned g = (unsigned)std::min<int64_t>(ntile, 2 * (int64_t)h->num_sms);     // two blocks per SM
            if (mv.ds == 8) {
                CU(cudaFuncSetAttribute(k_fine_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtcGeo<8>::SMEM));
                k_fine_tc<8><<<g, FTC_THREADS, FtcGeo<8>::SMEM, h->stream>>>(mv, fa);
                LAUNCHED();
                k_fine_redo<8><<<h->num_sms, 256, 0, h->stream>>>(mv, fa);
            } else {
                CU(cudaFuncSetAttribute(k_fine_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtcGeo<16>::SMEM));
                k_fine_tc<16><<<g, FTC_THREADS, FtcGeo<16>::SMEM, h->stream>>>(mv, fa);
                LAUNCHED();
                k_fine_redo<16><<<h->num_sms, 256, 0, h->stream>>>(mv, fa);
            }
            LAUNCHED();
            return B2L_OK;
        }
        switch (mv.ds) {
            case 2: if (f32stage) FINE32(2, 4); else FINE(2); break;
            case 4: if (f32stage) FINE32(4, 4); else FINE(4); break;
            case 8: if (f32stage) FINE32(8, 4); else FINE(8); break;
            case 16: if (f32stage) FINE32(16, 2); else FINE(16); break;
            case 32: if (f32stage) FINE32(32, 1); else FINE(0); break;
            case 64: if (f32stage) FINE32(64, 1); else FINE(0); break;
            default:
                if (mv.ds > 128) FAIL(B2L_ERR_UNSUPPORTED, "sub-vector length D/M = %d > 128 not supported", mv.ds);
                FINE(0);
        }
#undef FINE32
#undef FINE
        LAUNCHED();
    }
    return B2L_OK;
}

// ---- (re)build the cell-major layout ---------------------------------------------------------------
// large V: rows sorted by cell, run-length encoded into the sparse directory (no dense V*V array anywhere)
int ensure_index_sparse(b2l_handle h) {
    const ModelView& mv = h->mv;
    const int64_t n = h->n_items;
    h->nu = 0; h->max_run = 0; h->hmask = 0;
    h->h_ucell.clear(); h->h_ustart.assign(1, 0u);
    h->rows_padded = n + 256;
    CU(h->codes.reserve((size_t)h->rows_padded * mv.MP));
    CU(h->rowids.reserve((size_t)h->rows_padded * 8));
    CU(cudaMemsetAsync(h->codes.p, 0, (size_t)h->rows_padded * mv.MP, h->stream));
    if (n > 0) {
        if (n >= (int64_t)1 << 31) FAIL(B2L_ERR_UNSUPPORTED, "more than 2^31 - 1 rows per shard");
        CU(h->w_sort_a.reserve((size_t)n * 8));
        CU(h->w_sort_b.reserve((size_t)n * 8));
        CU(h->w_misc.reserve(64));
        unsigned int* cell = h->w_sort_a.as<unsigned int>();
        unsigned int* order = cell + n;
        unsigned int* scell = h->w_sort_b.as<unsigned int>();
        unsigned int* ssrc = scell + n;
        int* bad = h->w_misc.as<int>();
        CU(cudaMemsetAsync(bad, 0, 64, h->stream));
        k_cell_ids<<<grid_for(n, 256), 256, 0, h->stream>>>(h->m_coarse.as<int32_t>(), n, mv.V, cell, order, nullptr, bad);
        LAUNCHED();
        int bits = 1;
        while (((int64_t)1 << bits) < (int64_t)mv.V * mv.V) ++bits;
        size_t tmp = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, cell, scell, order, ssrc, (int)n, 0, bits, h->stream));
        CU(h->w_sort_tmp.reserve(tmp));
        CU(cub::DeviceRadixSort::SortPairs(h->w_sort_tmp.p, tmp, cell, scell, order, ssrc, (int)n, 0, bits, h->stream));
        ++h->launches;
        // runs of equal cell ids: unique cells + counts (the unsorted cell / order arrays are free now)
        unsigned int* ucell_tmp = cell;
        unsigned int* counts = order;
        int* d_nruns = bad + 4;
        size_t t2 = 0;
        CU(cub::DeviceRunLengthEncode::Encode(nullptr, t2, scell, ucell_tmp, counts, d_nruns, (int)n, h->stream));
        CU(h->w_sort_tmp.reserve(t2));
        CU(cub::DeviceRunLengthEncode::Encode(h->w_sort_tmp.p, t2, scell, ucell_tmp, counts, d_nruns, (int)n, h->stream));
        ++h->launches;
        int hb[8] = {};
        CU(cudaMemcpyAsync(hb, bad, 32, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (hb[0]) FAIL(B2L_ERR_ARG, "coarse code out of range [0, V) in the index");
        const unsigned int nu = (unsigned int)hb[4];
        CU(h->d_ucell.reserve((size_t)nu * 4));
        CU(h->d_ustart.reserve((size_t)(nu + 1) * 4));
        CU(cudaMemcpyAsync(h->d_ucell.p, ucell_tmp, (size_t)nu * 4, cudaMemcpyDeviceToDevice, h->stream));
        k_run_starts<<<1, 1024, 0, h->stream>>>(counts, nu, h->d_ustart.as<unsigned int>());
        LAUNCHED();
        unsigned int hs = 64;
        while (hs < 2 * nu) hs <<= 1;
        CU(h->d_hkeys.reserve((size_t)hs * 4));
        CU(h->d_hvals.reserve((size_t)hs * 4));
        CU(cudaMemsetAsync(h->d_hkeys.p, 0xFF, (size_t)hs * 4, h->stream));
        k_dir_build<<<grid_for(nu, 256), 256, 0, h->stream>>>(h->d_ucell.as<unsigned int>(), nu, h->d_hkeys.as<unsigned int>(),
                                                               h->d_hvals.as<unsigned int>(), hs - 1);
        LAUNCHED();
        k_scatter_runs<<<grid_for(nu, 128), 128, 0, h->stream>>>(ssrc, h->d_ustart.as<unsigned int>(), nu, h->m_fine.as<uint8_t>(),
                                                                  h->m_rowid.as<int64_t>(), mv.M, mv.MP, mv.SW, h->codes.as<uint8_t>(),
                                                                  h->rowids.as<int64_t>());
        LAUNCHED();
        h->h_ucell.resize(nu);
        h->h_ustart.resize(nu + 1);
        CU(cudaMemcpyAsync(h->h_ucell.data(), h->d_ucell.p, (size_t)nu * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->h_ustart.data(), h->d_ustart.p, (size_t)(nu + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->nu = nu; h->hmask = hs - 1;
        for (unsigned int r = 0; r < nu; ++r) h->max_run = std::max(h->max_run, h->h_ustart[r + 1] - h->h_ustart[r]);
    }
    h->dirty = false;
    return B2L_OK;
}

int ensure_index(b2l_handle h) {
    if (!h->dirty) return B2L_OK;
    if (h->mv.V > B2L_MAX_V) return ensure_index_sparse(h);
    const ModelView& mv = h->mv;
    const int ncell = mv.V * mv.V;
    const int64_t n = h->n_items;
    h->h_lsize.assign(ncell, 0);
    h->h_cell_start.assign(ncell, 0);
    CU(h->lsize.reserve((size_t)ncell * 8));
    CU(h->gsize.reserve((size_t)ncell * 8));
    CU(h->cell_start.reserve((size_t)ncell * 8));
    CU(h->sorted_first.reserve((size_t)ncell * 8));
    std::vector<int64_t> first(ncell, 0);
    if (n > 0) {
        if (n >= (int64_t)1 << 31) FAIL(B2L_ERR_UNSUPPORTED, "more than 2^31 - 1 rows per shard");
        CU(h->w_sort_a.reserve((size_t)n * 8));      // cell[n] | order[n]
        CU(h->w_sort_b.reserve((size_t)n * 8));      // sorted_cell[n] | sorted_src[n]
        CU(h->w_misc.reserve((size_t)ncell * 8 + 64));
        unsigned int* cell = h->w_sort_a.as<unsigned int>();
        unsigned int* order = cell + n;
        unsigned int* scell = h->w_sort_b.as<unsigned int>();
        unsigned int* ssrc = scell + n;
        unsigned long long* hist = h->w_misc.as<unsigned long long>();
        int* bad = (int*)(hist + ncell);
        CU(cudaMemsetAsync(h->w_misc.p, 0, (size_t)ncell * 8 + 64, h->stream));
        k_cell_ids<<<grid_for(n, 256), 256, 0, h->stream>>>(h->m_coarse.as<int32_t>(), n, mv.V, cell, order, hist, bad);
        LAUNCHED();
        int bits = 1;
        while ((1 << bits) < ncell) ++bits;
        size_t tmp = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, cell, scell, order, ssrc, (int)n, 0, bits, h->stream));
        CU(h->w_sort_tmp.reserve(tmp));
        CU(cub::DeviceRadixSort::SortPairs(h->w_sort_tmp.p, tmp, cell, scell, order, ssrc, (int)n, 0, bits, h->stream));
        ++h->launches;
        std::vector<unsigned long long> hh(ncell);
        int hbad = 0;
        CU(cudaMemcpyAsync(hh.data(), hist, (size_t)ncell * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (hbad) FAIL(B2L_ERR_ARG, "coarse code out of range [0, V) in the index");
        for (int c = 0; c < ncell; ++c) h->h_lsize[c] = (int64_t)hh[c];
    }
    int64_t start = 0, acc = 0;
    for (int c = 0; c < ncell; ++c) {
        h->h_cell_start[c] = start;
        first[c] = acc;
        acc += h->h_lsize[c];
        start += (h->h_lsize[c] + 15) & ~(int64_t)15;
    }
    h->rows_padded = start + 256;          // the scans read whole chunks (up to 128 rows) past the end of the last cell
    CU(h->codes.reserve((size_t)h->rows_padded * mv.MP));
    CU(h->rowids.reserve((size_t)h->rows_padded * 8));
    CU(cudaMemsetAsync(h->codes.p, 0, (size_t)h->rows_padded * mv.MP, h->stream));
    CU(cudaMemcpyAsync(h->lsize.p, h->h_lsize.data(), (size_t)ncell * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->cell_start.p, h->h_cell_start.data(), (size_t)ncell * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->sorted_first.p, first.data(), (size_t)ncell * 8, cudaMemcpyHostToDevice, h->stream));
    if (!h->global_set) h->h_gsize = h->h_lsize;
    if ((int)h->h_gsize.size() != ncell) FAIL(B2L_ERR_STATE, "global cell sizes have the wrong length");
    CU(cudaMemcpyAsync(h->gsize.p, h->h_gsize.data(), (size_t)ncell * 8, cudaMemcpyHostToDevice, h->stream));
    if (n > 0) {
        unsigned int* scell = h->w_sort_b.as<unsigned int>();
        k_scatter_rows<<<grid_for(n, 256), 256, 0, h->stream>>>(scell, scell + n, n, h->sorted_first.as<int64_t>(),
                                                                 h->cell_start.as<int64_t>(), h->m_fine.as<uint8_t>(),
                                                                 h->m_rowid.as<int64_t>(), mv.M, mv.MP, mv.SW, h->codes.as<uint8_t>(),
                                                                 h->rowids.as<int64_t>());
        LAUNCHED();
    }
    CU(cudaStreamSynchronize(h->stream));   // `first` is a local host buffer
    h->dirty = false;
    return B2L_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// carve the per-batch plan arrays out of one allocation
int setup_plan(b2l_handle h, int nq, int segc, int nsegmax = 1, int nseg_cap = 1) {
    const ModelView& mv = h->mv;
    const int ncell = mv.V * mv.V, maxvis = ncell;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_cnt = take(sizeof(PlanCounters));
    const size_t o_qc = take((size_t)ncell * 4);            // cell_qcount (zeroed together with the counters)
    const size_t o_ccnt = take((size_t)nq * 4);             // candidates appended per query
    const size_t zero_bytes = off;
    const size_t o_fill = take((size_t)ncell * 4);
    const size_t o_coff = take((size_t)(ncell + 1) * 4);
    const size_t o_ib = take(((size_t)std::max(nsegmax, nseg_cap) * ncell + 1) * 4);    // (sized for the shortest segments)
    const size_t item_cap = std::min<size_t>((size_t)1 << 20,
                                             (size_t)std::max(nsegmax, nseg_cap) * ((size_t)nq * ncell / 2 + ncell) + 1);
    const size_t o_if = take(item_cap * 4);
    const size_t o_nvis = take((size_t)nq * 4);
    const size_t o_ncand = take((size_t)nq * 8);
    const size_t o_ncl = take((size_t)nq * 8);
    const size_t o_npart = take((size_t)nq * 4);
    const size_t o_pbase = take((size_t)nq * 4);
    const size_t o_vcell = take((size_t)nq * maxvis * 4);
    const size_t o_vbase = take((size_t)nq * maxvis * 8);
    const size_t o_vl0 = take((size_t)nq * maxvis * 4);
    const size_t o_vl1 = take((size_t)nq * maxvis * 4);
    const size_t o_vpb = take((size_t)nq * maxvis * 4);
    const size_t o_desc = take((size_t)nq * 2 * mv.V * 3 * 4);
    const size_t o_gthr = take((size_t)nq * 4);
    CU(h->w_plan.reserve(off));
    unsigned char* b = h->w_plan.as<unsigned char>();
    PlanView& pv = h->pv;
    pv.nq = nq; pv.maxvis = maxvis; pv.segc = segc;
    pv.cnt = (PlanCounters*)(b + o_cnt);
    pv.cell_qcount = (unsigned int*)(b + o_qc);
    pv.cell_fill = (unsigned int*)(b + o_fill);
    pv.cellq_off = (unsigned int*)(b + o_coff);
    pv.item_base = (unsigned int*)(b + o_ib);
    pv.item_f = (unsigned int*)(b + o_if); pv.item_cap = (unsigned int)item_cap;
    pv.nvis = (int32_t*)(b + o_nvis);
    pv.ncand = (int64_t*)(b + o_ncand);
    pv.ncand_local = (int64_t*)(b + o_ncl);
    pv.npart = (int32_t*)(b + o_npart);
    pv.pbase = (int32_t*)(b + o_pbase);
    pv.vis_cell = (int32_t*)(b + o_vcell);
    pv.vis_base = (int64_t*)(b + o_vbase);
    pv.vis_lut0 = (int32_t*)(b + o_vl0);
    pv.vis_lut1 = (int32_t*)(b + o_vl1);
    pv.vis_pbase = (int32_t*)(b + o_vpb);
    pv.lut_desc = (int32_t*)(b + o_desc);
    pv.cellq = nullptr;
    pv.cell_segc = nullptr;
    pv.vis_dist = nullptr;
    h->gthr = (unsigned int*)(b + o_gthr);
    h->cand_cnt = (unsigned int*)(b + o_ccnt);
    CU(cudaMemsetAsync(b, 0, zero_bytes, h->stream));
    return B2L_OK;
}

// Turn one finished call record into statistics: h->stats = that call, accumulators += that call.
int collect_call(b2l_handle h, b2l_ctx::CallRec& r) {
    if (!r.pending) return B2L_OK;
    r.pending = false;
    CU(cudaEventSynchronize(r.ev[4]));
    if (r.has_pc) {
        const PlanCounters& pc = *r.h_pc;
        r.st.lut_slots = pc.n_lut;
        r.st.codes_scanned = (int64_t)pc.cand_local;
        r.st.scan_bytes = (int64_t)pc.cand_local * h->mv.M;
        r.st.work_items = pc.n_items;
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, r.ev[0], r.ev[2])); r.st.plan_ms = ms;
    CU(cudaEventElapsedTime(&ms, r.ev[2], r.ev[3])); r.st.scan_ms = ms;
    CU(cudaEventElapsedTime(&ms, r.ev[3], r.ev[4])); r.st.select_ms = ms;
    CU(cudaEventElapsedTime(&ms, r.ev[0], r.ev[4])); r.st.total_ms = ms;
    if (r.segc > 0 && r.st.work_items > 0) { h->fb_nq = r.nq; h->fb_segc = r.segc; h->fb_items = r.st.work_items; }
    b2l_stats acc = h->stats;                       // carries the accumulators
    b2l_stats cur = r.st;
    cur.acc_calls = acc.acc_calls + 1;
    cur.acc_scan_ms = acc.acc_scan_ms + cur.scan_ms; cur.acc_plan_ms = acc.acc_plan_ms + cur.plan_ms;
    cur.acc_select_ms = acc.acc_select_ms + cur.select_ms; cur.acc_total_ms = acc.acc_total_ms + cur.total_ms;
    cur.acc_codes_scanned = acc.acc_codes_scanned + cur.codes_scanned; cur.acc_scan_bytes = acc.acc_scan_bytes + cur.scan_bytes;
    cur.acc_work_items = acc.acc_work_items + cur.work_items; cur.acc_kernel_launches = acc.acc_kernel_launches + cur.kernel_launches;
    cur.acc_exact_queries = acc.acc_exact_queries + cur.exact_queries;
    cur.acc_rescan_queries = acc.acc_rescan_queries + cur.rescan_queries;
    h->stats = cur;
    return B2L_OK;
}

// collect every call record that is still pending, oldest first (waits for them)
int finish_stats(b2l_handle h) {
    for (uint64_t i = 0; i < b2l_ctx::NREC; ++i) {
        b2l_ctx::CallRec& r = h->ring[(h->seq + i) % b2l_ctx::NREC];      // h->seq % NREC is the oldest slot
        int rc = collect_call(h, r);
        if (rc) return rc;
    }
    return B2L_OK;
}

// carve the per-sub-batch arrays of the large-V plan out of w_walk
int setup_walk(b2l_handle h, int nqc, int segcap, size_t desc_cap, int viscap, WalkView& wv) {
    const ModelView& mv = h->mv;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o_cnt = take(sizeof(WalkCounters)), o_nvis = take((size_t)nqc * 4), o_nseg = take((size_t)nqc * 4),
                 o_ncand = take((size_t)nqc * 4), o_seg = take((size_t)nqc * segcap * 16), o_s0 = take((size_t)nqc * mv.V * 4),
                 o_s1 = take((size_t)nqc * mv.V * 4), o_desc = take(desc_cap * 12),
                 o_vc = take((size_t)nqc * viscap * 4), o_vd = take((size_t)nqc * viscap * 8);
    CU(h->w_walk.reserve(off));
    unsigned char* b = h->w_walk.as<unsigned char>();
    wv = WalkView();
    wv.nq = nqc; wv.segcap = segcap;
    wv.cnt = (WalkCounters*)(b + o_cnt);
    wv.nvis = (int32_t*)(b + o_nvis); wv.nseg = (int32_t*)(b + o_nseg); wv.ncand = (unsigned int*)(b + o_ncand);
    wv.seg = (uint4*)(b + o_seg); wv.slot0 = (int32_t*)(b + o_s0); wv.slot1 = (int32_t*)(b + o_s1);
    wv.lut_desc = (int32_t*)(b + o_desc);
    wv.viscap = viscap;
    wv.vis_cells = viscap ? (int32_t*)(b + o_vc) : nullptr;
    wv.vis_dists = viscap ? (double*)(b + o_vd) : nullptr;
    wv.max_visit = ((long long)1) << 62;
    CU(cudaMemsetAsync(wv.cnt, 0, sizeof(WalkCounters), h->stream));
    return B2L_OK;
}

SparseDir sparse_dir(b2l_handle h) {
    SparseDir d;
    d.hkeys = h->d_hkeys.as<unsigned int>(); d.hvals = h->d_hvals.as<unsigned int>(); d.hmask = h->hmask;
    d.ustart = h->d_ustart.as<unsigned int>(); d.ucell = h->d_ucell.as<unsigned int>(); d.nu = h->nu;
    return d;
}

int launch_walk(b2l_handle h, const void* x, int xf64, int nqc, int64_t quota, const WalkView& wv) {
    const ModelView& mv = h->mv;
    const size_t smem = walk_smem_bytes(mv.V);
    if (smem > 227 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "V=%d too large for the traversal kernel", mv.V);
    SparseDir dir = sparse_dir(h);
    if (h->nu == 0) {           // empty index: a one-entry empty hash table
        CU(h->d_hkeys.reserve(256)); CU(h->d_hvals.reserve(256)); CU(h->d_ustart.reserve(256)); CU(h->d_ucell.reserve(256));
        CU(cudaMemsetAsync(h->d_hkeys.p, 0xFF, 256, h->stream));
        dir = sparse_dir(h);
        dir.hmask = 63;
    }
    if (xf64) { CU(cudaFuncSetAttribute(k_walk<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_walk<double><<<nqc, WALK_THREADS, smem, h->stream>>>(mv, (const double*)x, quota, dir, wv); }
    else { CU(cudaFuncSetAttribute(k_walk<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_walk<float><<<nqc, WALK_THREADS, smem, h->stream>>>(mv, (const float*)x, quota, dir, wv); }
    LAUNCHED();
    return B2L_OK;
}

// Large-V search (V > B2L_MAX_V): traversal by k_walk, projections by the grouped GEMM, exact float64 ADC of every
// retrieved code, stable segmented sort, records.  Synchronous planning (two small read-backs per sub-batch).
int search_large_impl(b2l_handle h, const void* x, int xf64, int nq, int64_t quota, int k, void* d_records) {
    const ModelView& mv = h->mv;
    if (h->global_set) FAIL(B2L_ERR_UNSUPPORTED, "a cell-sharded index is not supported at V > %d", B2L_MAX_V);
    const size_t esz = xf64 ? 8 : 4;
    const int64_t nu = h->nu;
    const int segcap = (int)std::min<int64_t>(std::max<int64_t>(quota, 1), nu) + 2;
    // sub-batches: bound the segment lists (16 bytes per non-empty visited cell) to ~2 GB
    const int qchunk = (int)std::max<int64_t>(1, std::min<int64_t>(nq, ((int64_t)2 << 30) / ((int64_t)segcap * 16 + (int64_t)mv.V * 8 + 64)));
    CU(cudaEventRecord(h->cr->ev[1], h->stream));
    CU(cudaEventRecord(h->cr->ev[2], h->stream));
    int64_t cand_sum = 0, lut_sum = 0;
    for (int qa0 = 0; qa0 < nq; qa0 += qchunk) {
        const int nqc = std::min(qchunk, nq - qa0);
        const void* xs = (const char*)x + (size_t)qa0 * mv.D * esz;
        WalkView wv;
        const size_t desc_cap = (size_t)nqc * 2 * (size_t)std::min<int64_t>(mv.V, segcap);
        int rc = setup_walk(h, nqc, segcap, desc_cap, 0, wv);
        if (rc) return rc;
        if ((rc = launch_walk(h, xs, xf64, nqc, quota, wv))) return rc;
        WalkCounters wc;
        std::vector<unsigned int> ncand(nqc);
        CU(cudaMemcpyAsync(&wc, wv.cnt, sizeof wc, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(ncand.data(), wv.ncand, (size_t)nqc * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (wc.err == 1) FAIL(B2L_ERR_UNSUPPORTED, "more than %d cells at one and the same coarse distance: the traversal order is degenerate", WALK_CAP);
        if (wc.err) FAIL(B2L_ERR_STATE, "internal: segment list overflow in the traversal");
        cand_sum += (int64_t)wc.cand_total; lut_sum += wc.n_lut;
        // ---- projections of every (query, split, coarse code) in use
        const size_t nl = wc.n_lut;
        CU(h->w_p64.reserve(std::max<size_t>(1, nl) * mv.h * 8));
        if (nl) {
            if (mv.h % 64 == 0) {
                const int nb2 = 2 * mv.V;
                CU(h->w_perm.reserve((size_t)2 * nl * 4));
                CU(h->w_bkt.reserve((size_t)(4 * nb2 + 4) * 4));
                unsigned int* bc = h->w_bkt.as<unsigned int>();
                unsigned int* bbase = bc + nb2; unsigned int* bcur = bbase + nb2; unsigned int* btile = bcur + nb2;
                unsigned int* perm = h->w_perm.as<unsigned int>();
                CU(cudaMemsetAsync(bc, 0, (size_t)nb2 * 4, h->stream));
                const unsigned sg = (unsigned)std::min<size_t>((nl + 255) / 256, 2048);
                k_slot_hist<<<sg, 256, 0, h->stream>>>(wv.lut_desc, &wv.cnt->n_lut, mv.V, bc);
                LAUNCHED();
                k_enc_offsets<<<1, 1024, 0, h->stream>>>(mv.V, bc, bbase, bcur, btile);
                LAUNCHED();
                k_slot_scatter<<<sg, 256, 0, h->stream>>>(wv.lut_desc, &wv.cnt->n_lut, mv.V, nl, bbase, bcur, perm);
                LAUNCHED();
                if ((rc = launch_rotate_g<1>(h, xs, xf64, (int64_t)nl, bc, bbase, btile, perm, wv.lut_desc, h->w_p64.as<double>(), (unsigned)(nl / 64 + nb2)))) return rc;
                LAUNCHED();
            } else {
                // shapes without the grouped GEMM: one block per slot computes the projection (k_lut with no table output)
                CU(h->w_misc.reserve(sizeof(PlanCounters) + 64));
                PlanCounters pcs = {};
                pcs.n_lut = (unsigned)nl;
                CU(cudaMemcpyAsync(h->w_misc.p, &pcs, sizeof pcs, cudaMemcpyHostToDevice, h->stream));
                const size_t smem = (size_t)(2 * mv.h + LUT_THREADS) * 8;
                const unsigned lgrid = (unsigned)std::min<size_t>(nl, (size_t)h->num_sms * 8);
                if (xf64) { CU(cudaFuncSetAttribute(k_lut<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_lut<double, 0><<<lgrid, LUT_THREADS, smem, h->stream>>>(mv, (const double*)xs, wv.lut_desc, (const PlanCounters*)h->w_misc.p, h->w_p64.as<double>(), nullptr, nullptr, 2); }
                else { CU(cudaFuncSetAttribute(k_lut<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_lut<float, 0><<<lgrid, LUT_THREADS, smem, h->stream>>>(mv, (const float*)xs, wv.lut_desc, (const PlanCounters*)h->w_misc.p, h->w_p64.as<double>(), nullptr, nullptr, 2); }
                LAUNCHED();
                CU(cudaStreamSynchronize(h->stream));     // pcs is a local
            }
        }
        // ---- float32 copies of the projections + their norms (float32 preselection of the retrieved codes)
        const bool presel_shape = k <= PRS_SCAP / 2 && nl > 0 && h->scan_mode != 1;     // (scan mode 1: full float64 evaluation, the A/B switch)
        float* P32 = nullptr; float* n2s = nullptr;
        if (presel_shape) {
            CU(h->w_lut32.reserve(align256(nl * mv.h * 4) + nl * 4));                     // (the dense-plan tables are not used at large V)
            P32 = h->w_lut32.as<float>(); n2s = (float*)(h->w_lut32.as<unsigned char>() + align256(nl * mv.h * 4));
            k_slot_prep<<<(unsigned)std::min<size_t>((nl + 7) / 8, (size_t)h->num_sms * 8), 256, 0, h->stream>>>(h->w_p64.as<double>(), &wv.cnt->n_lut, mv.h, P32, n2s);
            LAUNCHED();
        }
        // ---- groups of queries with at most ~48M candidates: exact distances, stable segmented sort, emit
        const int64_t budget = (int64_t)48 << 20;
        for (int ga = 0; ga < nqc;) {
            int gb = ga;
            int64_t tot = 0;
            while (gb < nqc && (gb == ga || tot + ncand[gb] <= budget)) { tot += ncand[gb]; ++gb; }
            const int ng = gb - ga;
            if (tot >= ((int64_t)1 << 31)) FAIL(B2L_ERR_UNSUPPORTED, "a single query retrieves %lld codes (>= 2^31)", (long long)tot);
            std::vector<unsigned long long> qoff(ng + 1, 0ull);
            for (int g = 0; g < ng; ++g) qoff[g + 1] = qoff[g] + ncand[ga + g];
            const size_t T = (size_t)std::max<int64_t>(tot, 1);
            size_t woff = 0;
            auto wt = [&](size_t bytes) { size_t o = woff; woff += align256(bytes); return o; };
            const size_t o_k1 = wt(T * 8), o_k2 = wt(T * 8), o_v1 = wt(T * 4), o_v2 = wt(T * 4), o_qo = wt((size_t)(ng + 1) * 8);
            CU(h->w_walk2.reserve(woff));
            unsigned char* wb = h->w_walk2.as<unsigned char>();
            unsigned long long* k1 = (unsigned long long*)(wb + o_k1); unsigned long long* k2 = (unsigned long long*)(wb + o_k2);
            unsigned int* v1 = (unsigned int*)(wb + o_v1); unsigned int* v2 = (unsigned int*)(wb + o_v2);
            unsigned long long* dqo = (unsigned long long*)(wb + o_qo);
            CU(cudaMemcpyAsync(dqo, qoff.data(), (size_t)(ng + 1) * 8, cudaMemcpyHostToDevice, h->stream));
            const unsigned long long* kout = k1;
            const unsigned int* vout = v1;
            unsigned int nmax = 0;
            for (int g = 0; g < ng; ++g) nmax = std::max(nmax, ncand[ga + g]);
            const bool by_select = k <= SELK_MAXK && nmax <= SELK_MAXN;        // first k by selection instead of a full sort
            bool done = false;
            if (presel_shape && nmax <= PRS_MAXN && tot > 0) {                // float32 preselection, exact distances of the survivors only
                CU(cudaMemsetAsync(&wv.cnt->presel_fallback, 0, 4, h->stream));
                const size_t sms = presel_smem_bytes(nmax);
                CU(cudaFuncSetAttribute(k_presel_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms));
                k_presel_emit<<<ng, PRS_THREADS, sms, h->stream>>>(mv, h->codes.as<uint8_t>(), h->rowids.as<int64_t>(), wv, ga, h->w_p64.as<double>(),
                                                                   P32, n2s, h->c2sum, nmax, k, d_records, nq, qa0);
                LAUNCHED();
                unsigned int fb = 0;
                CU(cudaMemcpyAsync(&fb, &wv.cnt->presel_fallback, 4, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                done = (fb == 0);                                             // else: massive near ties somewhere -- the whole group in full
            }
            if (done) { ga = gb; continue; }
            if (tot > 0) {
                k_cand_dist<<<grid_for(tot, 256, h->num_sms * 16), 256, 0, h->stream>>>(mv, h->codes.as<uint8_t>(), wv, ga, ng, dqo,
                                                                                          h->w_p64.as<double>(), k1, v1);
                LAUNCHED();
                if (!by_select) {
                    size_t tmp = 0;
                    CU(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp, k1, k2, v1, v2, (int)tot, ng, dqo, dqo + 1, 0, 64, h->stream));
                    CU(h->w_sort_tmp.reserve(tmp));
                    CU(cub::DeviceSegmentedRadixSort::SortPairs(h->w_sort_tmp.p, tmp, k1, k2, v1, v2, (int)tot, ng, dqo, dqo + 1, 0, 64, h->stream));
                    ++h->launches;
                    kout = k2; vout = v2;
                }
            }
            if (by_select) {
                const size_t sms = selk_smem_bytes(std::max(nmax, 1u));
                CU(cudaFuncSetAttribute(k_select_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms));
                k_select_emit<<<ng, SELK_THREADS, sms, h->stream>>>(mv, h->codes.as<uint8_t>(), h->rowids.as<int64_t>(), wv, ga, dqo, k1,
                                                                    std::max(nmax, 1u), k, d_records, nq, qa0);
            } else
            k_emit_sorted<<<ng, 128, 0, h->stream>>>(mv, h->codes.as<uint8_t>(), h->rowids.as<int64_t>(), wv, ga, dqo, kout, vout, k,
                                                     d_records, nq, qa0);
            LAUNCHED();
            CU(cudaStreamSynchronize(h->stream));         // qoff is a local; the work buffers are reused by the next group
            ga = gb;
        }
    }
    CU(cudaEventRecord(h->cr->ev[3], h->stream));
    CU(cudaEventRecord(h->cr->ev[4], h->stream));
    h->cr->has_pc = false;
    h->cr->st.lut_slots = lut_sum;
    h->cr->st.codes_scanned = cand_sum;
    h->cr->st.scan_bytes = cand_sum * mv.M;
    h->cr->st.work_items = 0;
    h->cr->st.exact_queries = nq;
    h->cr->st.kernel_launches = h->launches - h->cr->launch0;
    h->cr->pending = true;
    CU(cudaStreamSynchronize(h->stream));
    return finish_stats(h);
}

// finish = false: return with the work queued on the stream (the caller synchronises and calls finish_stats)
int search_local_impl(b2l_handle h, const void* Q, int q_is_f64, int nq, int on_device, int64_t quota, int k, int exact,
                      void* d_records, bool finish = true, const RecRoute* route_in = nullptr) {
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (nq < 1 || k < 1 || !Q || !d_records) FAIL(B2L_ERR_ARG, "bad search arguments (nq=%d k=%d)", nq, k);
    if (h->mv.V > B2L_MAX_V_SPARSE) FAIL(B2L_ERR_UNSUPPORTED, "V=%d > %d", h->mv.V, B2L_MAX_V_SPARSE);
    const bool largev = h->mv.V > B2L_MAX_V;
    if (largev && route_in) FAIL(B2L_ERR_UNSUPPORTED, "the in-library exchange is not available at V > %d", B2L_MAX_V);
    int rc = ensure_index(h);
    if (rc) return rc;
    const ModelView& mv = h->mv;
    const int ncell = largev ? 1 : mv.V * mv.V;
    int64_t gtotal = 0;
    if (largev) gtotal = h->n_items;
    else for (int c = 0; c < ncell; ++c) gtotal += h->h_gsize[c];
    if (gtotal >= ((int64_t)1 << 32)) FAIL(B2L_ERR_UNSUPPORTED, "more than 2^32 indexed codes");

    { cudaError_t pre = cudaGetLastError(); if (pre != cudaSuccess) FAIL(B2L_ERR_CUDA, "stale CUDA error at search entry: %s", cudaGetErrorString(pre)); }
    {   // next call record; if the ring wrapped onto a call that was never collected, collect it now
        b2l_ctx::CallRec& r = h->ring[h->seq % b2l_ctx::NREC];
        int rc2 = collect_call(h, r);
        if (rc2) return rc2;
        ++h->seq;
        h->cr = &r;
        memset(&r.st, 0, sizeof r.st);
        r.has_pc = false;
        r.launch0 = h->launches;
        r.nq = nq; r.segc = 0;
    }
    CU(cudaEventRecord(h->cr->ev[0], h->stream));
    // queries -> device, PCA
    const int Din = h->has_pca ? mv.D0 : mv.D;
    const size_t esz = q_is_f64 ? 8 : 4;
    const void* dq = Q;
    if (!on_device) {
        CU(h->w_q.reserve((size_t)nq * Din * esz));
        rc = copy_in(h, h->w_q.p, Q, (size_t)nq * Din * esz, 0);
        if (rc) return rc;
        dq = h->w_q.p;
    }
    const void* x = dq;
    int xf64 = q_is_f64;
    if (h->has_pca) {
        CU(h->w_xq.reserve((size_t)nq * mv.D * 4));
        const size_t smem = (size_t)(mv.D0 + mv.D) * 8;
        if (q_is_f64) { CU(cudaFuncSetAttribute(k_pca<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pca<double><<<nq, 128, smem, h->stream>>>(mv, (const double*)dq, nq, h->w_xq.as<float>()); }
        else { CU(cudaFuncSetAttribute(k_pca<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pca<float><<<nq, 128, smem, h->stream>>>(mv, (const float*)dq, nq, h->w_xq.as<float>()); }
        LAUNCHED();
        x = h->w_xq.p; xf64 = 0;
    }
    if (largev) return search_large_impl(h, x, xf64, nq, quota, k, d_records);
    // fast path eligibility: the bound table of the scan needs >= KP entries per slot
    // KP = how many candidates of the preselection are re-ranked in float64 (power of two >= k + 8).  A wider KP certifies
    // more queries at the first stage when distances are concentrated (high-dimensional data, coarse 16-bit tables at
    // M = 32) at the price of a longer selection; b2l_set_preselect overrides the default.
    int KP = std::max(16, next_pow2(k + 8));
    if (h->kp_min > 0) KP = std::min(512, std::max(KP, next_pow2(h->kp_min)));
    const bool lowb_shape = nq <= SCAN1_MAX_NQ && h->scan_mode != 1 && h->scan_mode != 2 && exact == 0 && mv.MP >= 8 && mv.G > 0;
    const int LPS = lowb_shape ? 32 * SCAN_WARPS : mv.MP * SCAN_WARPS;
    const int GEN = std::max(1, KP / std::max(1, LPS));
    // exact: 0 = default fast scan (16-bit packed tables unless the handle is set to float32), 1 = float64 full sort,
    // 2 = fast scan with float32 tables
    const bool fast = exact != 1 && mv.G > 0 && KP <= 512;
    if (route_in && !fast) FAIL(B2L_ERR_UNSUPPORTED, "the in-library exchange needs the fast scan (M <= 32, k <= 504)");
    // low-batch regime (a handful of queries: no cross-query reuse of a code row, the scan is HBM-bound): one query per
    // work item, float32 tables, 32 rows per warp step (scan1.cuh)
    const bool lowb = fast && exact != 2 && nq <= SCAN1_MAX_NQ && h->scan_mode != 1 && h->scan_mode != 2 && mv.MP >= 8;
    const bool packed = fast && !lowb && exact == 0 && h->scan_mode != 1 && 65535 / mv.M >= 255;
    const int NS = lowb ? 1 : (packed ? 4 : 2) * mv.G;
    // segment length: a multiple of 64 codes, sized so the batch yields enough work items
    int64_t maxcell = 0;
    for (int c = 0; c < ncell; ++c) maxcell = std::max(maxcell, h->h_lsize[c]);
    // Start from 16K codes; the last finished search of the same batch size tells whether that gave the persistent grid
    // enough items (few queries, or a small shard of a multi-GPU index: shorter segments) -- results do not depend on it.
    int segc = 16 * 1024;
    if (fast && h->fb_nq == nq && h->fb_segc > 0) {
        segc = h->fb_segc;
        if (lowb) {
            segc = 16 * 1024;          // (set below from the expected candidate count, no feedback needed)
        } else {
        const int64_t grid = (int64_t)h->num_sms * 2;
        if (h->fb_items < 3 * grid || h->fb_items > 24 * grid) {           // aim at ~6 items per block, in one jump
            const double want = (double)segc * (double)h->fb_items / (6.0 * (double)grid);
            int s2 = 2048;
            while (s2 * 2 <= want * 1.42 && s2 < 16 * 1024) s2 *= 2;        // nearest power of two
            segc = s2;
        }
        }
    }
    std::vector<int32_t> cell_segc;
    int nseg_low = 1;
    if (lowb) {
        // One query per item, the next item's tables prefetched.  The items of a query all compete for the resident blocks
        // (3 per SM) at once, so their sizes decide the tail: every cell is cut into n_c equal segments with n_c proportional
        // to its size, such that the expected number of items is a whole number of waves over the resident blocks and all
        // items have (nearly) the same length, whatever the sizes of the cells.
        const int64_t est = std::min<int64_t>(gtotal, (quota < gtotal ? quota : gtotal) + maxcell / 2);   // codes ranked per query
        const double grid = (double)h->num_sms * 3.0;
        const double total = (double)est * nq;
        const double waves = std::max(1.0, std::ceil(total / (grid * 32768.0)));
        const double tgt = std::max(2048.0, total / (grid * waves));           // codes per item
        cell_segc.assign(ncell, 1024);
        std::vector<std::pair<double, int>> rem;
        int64_t sumn = 0;
        for (int c = 0; c < ncell; ++c) {
            const double x = (double)h->h_lsize[c] / tgt;
            const int64_t n = std::max<int64_t>(h->h_lsize[c] > 0 ? 1 : 0, (int64_t)std::floor(x));
            cell_segc[c] = (int32_t)n;                                         // (segments for now)
            sumn += n;
            if (h->h_lsize[c] > 0) rem.push_back({x - std::floor(x), c});
        }
        if (est >= gtotal) {                                                   // whole index per query: hit the wave count exactly
            int64_t want = (int64_t)(grid * waves) / std::max(1, nq) - sumn;  // (items = segments x queries)
            std::sort(rem.begin(), rem.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
            for (size_t i = 0; i < rem.size() && want > 0; ++i, --want) ++cell_segc[rem[i].second];
        } else {
            for (auto& r : rem) if (r.first >= 0.5) ++cell_segc[r.second];
        }
        for (int c = 0; c < ncell; ++c) {
            const int64_t n = std::max<int64_t>(1, cell_segc[c]);
            nseg_low = (int)std::max<int64_t>(nseg_low, h->h_lsize[c] > 0 ? n : 1);
            const int64_t len = (h->h_lsize[c] + n - 1) / n;
            cell_segc[c] = (int32_t)std::max<int64_t>(128, ((len + 127) / 128) * 128);
        }
        segc = 16 * 1024;
    }
    if (maxcell > 0 && maxcell < segc) segc = (int)(((maxcell + 63) / 64) * 64);
    const int nsegmax = lowb ? nseg_low + 1 : (int)std::max<int64_t>(1, (maxcell + segc - 1) / segc);
    rc = setup_plan(h, nq, segc, nsegmax, (int)((maxcell + 2047) / 2048) + 1);
    if (rc) return rc;
    if (fast) h->cr->segc = segc;
    PlanView& pv = h->pv;
    if (lowb) {
        CU(h->w_segc.reserve((size_t)ncell * 4));
        CU(cudaMemcpyAsync(h->w_segc.p, cell_segc.data(), (size_t)ncell * 4, cudaMemcpyHostToDevice, h->stream));   // (pageable: staged at once)
        pv.cell_segc = h->w_segc.as<int32_t>();
    }
    // per-query scratch of the fast path (bounds, bound tables, quantiser ranges): initialised by k_coarse_order
    QuantView qv = {};
    InitView iv = {};
    iv.gthr = h->gthr;
    if (fast) {
        CU(h->w_gtab.reserve((size_t)nq * LPS * GEN * 4));
        iv.gtab = h->w_gtab.as<unsigned int>(); iv.E = LPS * GEN;
    }
    if (packed) {
        const size_t o_qmin = 0, o_qmax = align256((size_t)nq * mv.M * 4), o_B = o_qmax + align256((size_t)nq * 4),
                     o_dl = o_B + align256((size_t)nq * 8), o_sl = o_dl + align256((size_t)nq * 8), o_mg = o_sl + align256((size_t)nq * 8),
                     qbytes = o_mg + align256((size_t)nq * 4);
        CU(h->w_quant.reserve(qbytes));
        unsigned char* qb = h->w_quant.as<unsigned char>();
        qv.qmin = (unsigned int*)(qb + o_qmin); qv.qmax = (unsigned int*)(qb + o_qmax);
        qv.B = (double*)(qb + o_B); qv.delta = (double*)(qb + o_dl); qv.slack = (double*)(qb + o_sl); qv.qmax_code = 65535 / mv.M;
        qv.margin = (unsigned int*)(qb + o_mg);
        qv.ds = mv.ds; qv.c2m = h->c2m;
        // float32 table entries (their error is part of the certification bound): the register-resident kernel of the
        // headline shape, or the generic one behind the grouped rotation GEMM of large models
        qv.f32_entries = ((mv.ds == 8 && mv.m == 8) || (mv.h % 64 == 0 && mv.h > 64)) ? 1 : 0;
        iv.qmin = qv.qmin; iv.qmax = qv.qmax; iv.M = mv.M;
    }
    {
        const size_t smem = (size_t)(3 * mv.V + 2) * 8 + (size_t)(7 * mv.V + 4) * 4 + 16;
        if (xf64) k_coarse_order<double><<<nq, 32, smem, h->stream>>>(mv, (const double*)x, quota, h->gsize.as<int64_t>(), h->lsize.as<int64_t>(), pv, iv);
        else k_coarse_order<float><<<nq, 32, smem, h->stream>>>(mv, (const float*)x, quota, h->gsize.as<int64_t>(), h->lsize.as<int64_t>(), pv, iv);
        LAUNCHED();
    }
    if (fast) {
        k_plan<<<1, 1024, 0, h->stream>>>(ncell, nsegmax, NS, segc, h->lsize.as<int64_t>(), pv);
        LAUNCHED();
        k_item_table<<<(nsegmax * ncell + 255) / 256, 256, 0, h->stream>>>(nsegmax * ncell, pv);
        LAUNCHED();
    }
    // Buffer sizes: the fast path sizes everything for the worst case the plan can produce (2V tables and V*V visited
    // cells per query), so nothing has to come back to the host between the kernels; the counters are fetched
    // with the results.  The exact path (and worst cases beyond B2L_ASYNC_WS bytes) reads the counters first.
    const size_t lut_row_bytes = (size_t)B2L_LUT_ROWS * mv.m * 4 + (size_t)mv.h * 8;
    const size_t worst_lut = (size_t)nq * 2 * mv.V;
    const bool nosync = fast && worst_lut * lut_row_bytes <= ((size_t)2 << 30);
    PlanCounters pc = {};
    size_t cap_lut = worst_lut, cap_pairs = (size_t)nq * ncell;
    if (!nosync) {
        CU(cudaMemcpyAsync(&pc, pv.cnt, sizeof pc, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        cap_lut = std::max(1u, pc.n_lut); cap_pairs = std::max(1u, pc.n_pairs);
    }
    h->cr->has_pc = nosync;

    // LUT build
    CU(h->w_p64.reserve(cap_lut * mv.h * 8));
    float* lut32 = nullptr;
    double* lut64 = nullptr;
    if (fast) { CU(h->w_lut32.reserve(cap_lut * B2L_LUT_ROWS * mv.m * 4)); lut32 = h->w_lut32.as<float>(); }
    else { CU(h->w_lut64.reserve(cap_lut * mv.m * mv.K * 8)); lut64 = h->w_lut64.as<double>(); }
    if (packed) CU(h->w_lut16.reserve(cap_lut * B2L_LUT_ROWS * mv.m * 2));
    if (nosync || pc.n_lut) {
        const size_t smem = (size_t)(2 * mv.h + LUT_THREADS) * 8;
        const unsigned lgrid = (unsigned)std::min<size_t>(cap_lut, (size_t)h->num_sms * 8);
        bool ranged = false;
#define LUTK(XT, DSV)                                                                                                   \
    do {                                                                                                                \
        CU(cudaFuncSetAttribute(k_lut<XT, DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        k_lut<XT, DSV><<<lgrid, LUT_THREADS, smem, h->stream>>>(mv, (const XT*)x, pv.lut_desc, pv.cnt, h->w_p64.as<double>(), \
                                                                lut32, lut64);                                          \
    } while (0)
#define LUTD(XT)                                  \
    switch (mv.ds) {                              \
        case 2: LUTK(XT, 2); break;               \
        case 4: LUTK(XT, 4); break;               \
        case 8: LUTK(XT, 8); break;               \
        case 16: LUTK(XT, 16); break;             \
        default: LUTK(XT, 0);                     \
    }
        if (mv.h % 64 == 0 && mv.h > 64) {
            // large models (2048-d: R[c] is 8 MB): the projections of all slots as ONE grouped float64 tensor-core GEMM
            // (slots bucketed by (split, coarse code)), then the table entries from the finished projections
            const int nb2 = 2 * mv.V;
            CU(h->w_perm.reserve((size_t)2 * cap_lut * 4));                    // perm[2][cap_lut]
            CU(h->w_bkt.reserve((size_t)(4 * nb2 + 4) * 4));
            unsigned int* bc = h->w_bkt.as<unsigned int>();
            unsigned int* bbase = bc + nb2; unsigned int* bcur = bbase + nb2; unsigned int* btile = bcur + nb2;
            unsigned int* perm = h->w_perm.as<unsigned int>();
            CU(cudaMemsetAsync(bc, 0, (size_t)nb2 * 4, h->stream));
            const unsigned sg = (unsigned)std::min<size_t>((cap_lut + 255) / 256, 1024);
            k_slot_hist<<<sg, 256, 0, h->stream>>>(pv.lut_desc, &pv.cnt->n_lut, mv.V, bc);
            LAUNCHED();
            k_enc_offsets<<<1, 1024, 0, h->stream>>>(mv.V, bc, bbase, bcur, btile);
            LAUNCHED();
            k_slot_scatter<<<sg, 256, 0, h->stream>>>(pv.lut_desc, &pv.cnt->n_lut, mv.V, cap_lut, bbase, bcur, perm);
            LAUNCHED();
            if ((rc = launch_rotate_g<1>(h, x, xf64, (int64_t)cap_lut, bc, bbase, btile, perm, pv.lut_desc, h->w_p64.as<double>(), (unsigned)(cap_lut / 64 + nb2)))) return rc;
            LAUNCHED();
            if (packed) {
                constexpr int SB = 4;
                const size_t sme = (size_t)SB * mv.h * 4;
                CU(cudaFuncSetAttribute(k_lut_entries_f32<SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sme));
                const unsigned eg = (unsigned)std::min<size_t>(cap_lut / SB + 2, (size_t)h->num_sms * 4);
                k_lut_entries_f32<SB><<<eg, 256, sme, h->stream>>>(mv, h->w_p64.as<double>(), perm, bbase, bc, cap_lut, lut32);
            } else if (xf64) {
                CU(cudaFuncSetAttribute(k_lut<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_lut<double, 0><<<lgrid, LUT_THREADS, smem, h->stream>>>(mv, (const double*)x, pv.lut_desc, pv.cnt, h->w_p64.as<double>(), lut32, lut64, 1);
            } else {
                CU(cudaFuncSetAttribute(k_lut<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_lut<float, 0><<<lgrid, LUT_THREADS, smem, h->stream>>>(mv, (const float*)x, pv.lut_desc, pv.cnt, h->w_p64.as<double>(), lut32, lut64, 1);
            }
        } else if (packed && qv.f32_entries) {
            // headline shape, packed scan: float64 projection, table entries in float32 from the register-resident float32
            // codebook (their evaluation error is part of the certification bound), column ranges reduced on the way
            const size_t sm32 = (size_t)(2 * mv.h + 256) * 8 + (size_t)mv.h * 4;
            const unsigned rgrid = (unsigned)(2 * h->num_sms);
            if (xf64) k_lut_f32<double, 8, 8><<<rgrid, 256, sm32, h->stream>>>(mv, (const double*)x, pv.lut_desc, pv.cnt,
                                                                              h->w_p64.as<double>(), lut32, qv.qmin, qv.qmax);
            else k_lut_f32<float, 8, 8><<<rgrid, 256, sm32, h->stream>>>(mv, (const float*)x, pv.lut_desc, pv.cnt,
                                                                          h->w_p64.as<double>(), lut32, qv.qmin, qv.qmax);
            ranged = true;
        } else if (fast && mv.ds == 8 && mv.m == 8 && mv.K <= 256) {
            // headline shape: sub-quantizer codebook register-resident, one block per SM bound to one coarse split;
            // the column ranges the 16-bit tables need are reduced on the way
            const size_t smr = (size_t)(2 * mv.h + LUTR_THREADS) * 8;
            const unsigned rgrid = (unsigned)(2 * std::max(1, h->num_sms / 2));
            if (xf64) k_lut_reg<double, 8, 4><<<rgrid, LUTR_THREADS, smr, h->stream>>>(mv, (const double*)x, pv.lut_desc, pv.cnt,
                                                                                     h->w_p64.as<double>(), lut32, qv.qmin, qv.qmax);
            else k_lut_reg<float, 8, 4><<<rgrid, LUTR_THREADS, smr, h->stream>>>(mv, (const float*)x, pv.lut_desc, pv.cnt,
                                                                                 h->w_p64.as<double>(), lut32, qv.qmin, qv.qmax);
            ranged = true;
        } else if (xf64) { LUTD(double); } else { LUTD(float); }
#undef LUTD
#undef LUTK
        LAUNCHED();
        if (packed) {
            const unsigned qgrid = (unsigned)std::min<size_t>(cap_lut, (size_t)h->num_sms * 8);
            if (!ranged) {
                k_lut_range<<<qgrid, 256, 0, h->stream>>>(mv.m, pv.lut_desc, pv.cnt, lut32, mv.K, qv, mv.M);
                LAUNCHED();
            }
            k_lut_quant<<<qgrid, 256, 0, h->stream>>>(mv.m, pv.lut_desc, pv.cnt, lut32, qv, mv.M, h->w_lut16.as<unsigned short>());
            LAUNCHED();
        }
    }
    IndexView ix = {h->codes.as<uint8_t>(), h->rowids.as<int64_t>(), h->cell_start.as<int64_t>(), h->lsize.as<int64_t>()};
    CU(cudaEventRecord(h->cr->ev[1], h->stream));
    if (fast) {
        CU(h->w_cellq.reserve(cap_pairs * 8));
        const int cand_cap = lowb ? SCAN1_CAND_CAP : SCAN_CAND_CAP;
        CU(h->w_cand.reserve((size_t)nq * cand_cap * 8));

        pv.cellq = h->w_cellq.as<int2>();
        k_fill<<<(nq + 255) / 256, 256, 0, h->stream>>>(pv);
        LAUNCHED();
        CU(cudaEventRecord(h->cr->ev[2], h->stream));
        if (nosync || pc.n_items) {
            ScanArgs a;
            a.codes = ix.codes; a.cell_start = ix.cell_start; a.lsize = ix.lsize; a.lut32 = lut32;
            a.cand = h->w_cand.as<unsigned long long>(); a.cand_cnt = h->cand_cnt;
            a.pv = pv; a.ncell = ncell; a.nflat = nsegmax * ncell; a.KP = KP; a.m = mv.m; a.M = mv.M;
            a.GEN = GEN; a.E = LPS * GEN;
            a.gthr = h->gthr; a.gtab = h->w_gtab.as<float>();
            a.lut16 = packed ? h->w_lut16.as<unsigned short>() : nullptr; a.qfill = (unsigned)qv.qmax_code;
            a.cand_cap = cand_cap;
            a.qmargin = packed ? qv.margin : nullptr;
            if (lowb) {
                switch (mv.MP) {
                    case 8: rc = launch_scan1<8>(h, a); break;
                    case 16: rc = launch_scan1<16>(h, a); break;
                    case 32: rc = launch_scan1<32>(h, a); break;
                    default: FAIL(B2L_ERR_UNSUPPORTED, "no low-batch scan for code stride %d", mv.MP);
                }
            } else
            switch (mv.MP) {
                case 4: rc = packed ? launch_scan_pk<4>(h, a) : launch_scan<4>(h, a); break;
                case 8: rc = packed ? launch_scan_pk<8>(h, a) : launch_scan<8>(h, a); break;
                case 16: rc = packed ? launch_scan_pk<16>(h, a) : launch_scan<16>(h, a); break;
                case 32: rc = packed ? launch_scan_pk<32>(h, a) : launch_scan<32>(h, a); break;
                default: FAIL(B2L_ERR_UNSUPPORTED, "no scan instantiation for code stride %d", mv.MP);
            }
            if (rc) return rc;
        }
        CU(cudaEventRecord(h->cr->ev[3], h->stream));
        const double eps_rel = (double)(mv.M + 4) * 2.0 * ldexp(1.0, -24);
        const size_t smem = select_smem_bytes(KP);
        CU(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RecRoute route = {};
        if (route_in) route = *route_in; else { route.base[0] = (unsigned char*)d_records; route.nq_home = nq; }
        CU(h->w_need2.reserve((size_t)nq));
        k_select<<<nq, SEL_THREADS, smem, h->stream>>>(mv, ix, pv, h->w_cand.as<unsigned long long>(), h->cand_cnt, h->gthr,
                                                       cand_cap, h->w_p64.as<double>(), KP, k, eps_rel, route,
                                                       packed ? 1 : 0, qv.B, qv.delta, qv.slack, h->w_need2.as<uint8_t>(),
                                                       packed ? qv.margin : nullptr);
        LAUNCHED();
        if (k <= SEL2_LIST) {
            // second chance of the queries the KP best could not certify (near-ties): every appended candidate, exactly
            const size_t sm2 = select2_smem_bytes();
            CU(cudaFuncSetAttribute(k_select2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            k_select2<<<nq, 256, sm2, h->stream>>>(mv, ix, pv, h->w_cand.as<unsigned long long>(), h->cand_cnt, h->gthr, cand_cap,
                                                   h->w_p64.as<double>(), k, eps_rel, route, packed ? 1 : 0, qv.B, qv.delta, qv.slack,
                                                   h->w_need2.as<uint8_t>(), packed ? qv.margin : nullptr);
            LAUNCHED();
        }
        h->cr->st.packed = packed ? 1 : 0;
        CU(cudaEventRecord(h->cr->ev[4], h->stream));
    } else {
        CU(cudaEventRecord(h->cr->ev[2], h->stream));
        CU(cudaEventRecord(h->cr->ev[3], h->stream));
        std::vector<int64_t> ncl(nq);
        CU(cudaMemcpyAsync(ncl.data(), pv.ncand_local, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        int64_t nmax = 1;
        for (int q = 0; q < nq; ++q) nmax = std::max(nmax, ncl[q]);
        CU(h->w_sort_a.reserve((size_t)nmax * 12));
        CU(h->w_sort_b.reserve((size_t)nmax * 12));
        unsigned long long* ka = h->w_sort_a.as<unsigned long long>();
        unsigned int* va = (unsigned int*)(ka + nmax);
        unsigned long long* kb = h->w_sort_b.as<unsigned long long>();
        unsigned int* vb = (unsigned int*)(kb + nmax);
        size_t tmp = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ka, kb, va, vb, (int)nmax, 0, 64, h->stream));
        CU(h->w_sort_tmp.reserve(tmp));
        for (int q = 0; q < nq; ++q) {
            const int64_t n = ncl[q];
            if (n > 0) {
                const size_t smem = (size_t)(pv.maxvis + 1) * 8;
                k_exact_dist<<<grid_for(n, 256), 256, smem, h->stream>>>(mv, ix, pv, q, lut64, ka, va, n);
                LAUNCHED();
                size_t t2 = tmp;
                CU(cub::DeviceRadixSort::SortPairs(h->w_sort_tmp.p, t2, ka, kb, va, vb, (int)n, 0, 64, h->stream));
                ++h->launches;
            }
            k_exact_emit<<<1, 256, 0, h->stream>>>(mv, ix, pv, q, q, nq, k, kb, vb, n, d_records);
            LAUNCHED();
        }
        h->cr->st.exact_queries = nq;
        CU(cudaEventRecord(h->cr->ev[4], h->stream));
    }
    if (h->cr->has_pc) {
        CU(cudaMemcpyAsync(h->cr->h_pc, pv.cnt, sizeof(PlanCounters), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->cr->ev[4], h->stream));        // (re-recorded after the copy: the record is complete when it fires)
    } else {
        h->cr->st.lut_slots = pc.n_lut;
        h->cr->st.codes_scanned = (int64_t)pc.cand_local;
        h->cr->st.scan_bytes = (int64_t)pc.cand_local * mv.M;
        h->cr->st.work_items = fast ? pc.n_items : 0;
    }
    h->cr->st.kernel_launches = h->launches - h->cr->launch0;
    h->cr->pending = true;
    if (!finish) return B2L_OK;
    CU(cudaStreamSynchronize(h->stream));
    return finish_stats(h);
}

// single-rank copy-out of records (any k); multi-rank merge through k_final
int merge_impl(b2l_handle h, const void* d_recs, int nranks, int nq, int k, int on_device, int64_t* rowid, double* dist,
               int32_t* coarse, uint8_t* fine, int32_t* count, int32_t* visited, uint8_t* certified, bool finish = true,
               unsigned int* d_unc = nullptr) {
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (nranks < 1 || nq < 1 || k < 1 || !count) FAIL(B2L_ERR_ARG, "bad merge arguments");
    const ModelView& mv = h->mv;
    const int n = next_pow2(nranks * k);
    const size_t smem = (size_t)n * 16 + 16;
    if (smem > 200 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "nranks*k = %d too large for the final merge", nranks * k);
    // device-side outputs
    int64_t* d_rowid = rowid; double* d_dist = dist; int32_t* d_coarse = coarse; uint8_t* d_fine = fine;
    int32_t* d_count = count; int32_t* d_visited = visited; uint8_t* d_cert = certified;
    if (!on_device) {
        const size_t nk = (size_t)nq * k;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
        const size_t o1 = take(nk * 8), o2 = take(nk * 8), o3 = take(nk * 8), o4 = take(nk * mv.M), o5 = take((size_t)nq * 4),
                     o6 = take((size_t)nq * 4), o7 = take((size_t)nq);
        CU(h->w_out.reserve(off));
        unsigned char* b = h->w_out.as<unsigned char>();
        d_rowid = (int64_t*)(b + o1); d_dist = (double*)(b + o2); d_coarse = (int32_t*)(b + o3); d_fine = b + o4;
        d_count = (int32_t*)(b + o5); d_visited = (int32_t*)(b + o6); d_cert = b + o7;
    }
    CU(cudaFuncSetAttribute(k_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_final<<<nq, 128, smem, h->stream>>>(mv.V, mv.M, d_recs, nranks, nq, k, n, d_rowid, d_dist, d_coarse, d_fine, d_count,
                                          d_visited, d_cert, d_unc, (h->force_redo & 4) ? 1 : 0);
    LAUNCHED();
    if (!on_device) {
        const size_t nk = (size_t)nq * k;
        int rc;
        if ((rc = copy_out(h, rowid, d_rowid, nk * 8, 0))) return rc;
        if ((rc = copy_out(h, dist, d_dist, nk * 8, 0))) return rc;
        if ((rc = copy_out(h, coarse, d_coarse, nk * 8, 0))) return rc;
        if ((rc = copy_out(h, fine, d_fine, nk * mv.M, 0))) return rc;
        if ((rc = copy_out(h, count, d_count, (size_t)nq * 4, 0))) return rc;
        if ((rc = copy_out(h, visited, d_visited, (size_t)nq * 4, 0))) return rc;
        if ((rc = copy_out(h, certified, d_cert, (size_t)nq, 0))) return rc;
    }
    if (finish) CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

}  // namespace

// ===================================================================================================
extern "C" {

int b2l_version(void) { return B2L_ABI_VERSION; }

const char* b2l_last_error(b2l_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int b2l_create(int device, b2l_handle* out) {
    if (!out) { g_create_error = "b2l_create: out is NULL"; return B2L_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("b2l_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return B2L_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "b2l_create: bad device ordinal"; return B2L_ERR_ARG; }
    b2l_ctx* h = new b2l_ctx();
    h->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (int r = 0; r < b2l_ctx::NREC && e == cudaSuccess; ++r) {
        for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ring[r].ev[i]);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->ring[r].h_pc, sizeof(PlanCounters), cudaHostAllocDefault);
    }
    h->cr = &h->ring[0];
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        g_create_error = std::string("b2l_create: ") + cudaGetErrorString(e);
        delete h;
        return B2L_ERR_CUDA;
    }
    h->num_sms = prop.multiProcessorCount;
    if (cudaMalloc((void**)&h->d_nguard, 8) == cudaSuccess) cudaMemset(h->d_nguard, 0, 8);
    if (cudaMalloc((void**)&h->d_nredo, 4) == cudaSuccess) cudaMemset(h->d_nredo, 0, 4);
    if (prop.major < 10) {
        g_create_error = "b2l_create: device is not sm_100 class (the library is built for sm_100a only)";
        delete h;
        return B2L_ERR_UNSUPPORTED;
    }
    *out = h;
    return B2L_OK;
}

int b2l_destroy(b2l_handle h) {
    if (!h) return B2L_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->parent) { std::lock_guard<std::mutex> lk(h->parent->mu); --h->parent->n_siblings; }
    DevBuf* bufs[] = {&h->dCs, &h->dmus, &h->dRt, &h->dsubs, &h->dsubs32, &h->dsubs32T, &h->dc2max, &h->dP, &h->dpmu, &h->dftc, &h->dCs32, &h->w_redo, &h->w_ftc_dbg, &h->m_coarse, &h->m_fine, &h->m_rowid, &h->codes,
                      &h->rowids, &h->cell_start, &h->lsize, &h->gsize, &h->sorted_first, &h->w_q, &h->w_xq, &h->w_px,
                      &h->w_coarse, &h->w_fine, &h->w_lut32, &h->w_lut64, &h->w_p64, &h->w_cellq, &h->w_cand, &h->w_gtab, &h->w_lut16, &h->w_quant, &h->w_plan,
                      &h->w_sort_a, &h->w_sort_b, &h->w_sort_tmp, &h->w_rec, &h->w_rec2, &h->w_out, &h->w_out2, &h->w_misc, &h->w_bkt, &h->w_perm, &h->w_need2, &h->w_segc, &h->d_ucell, &h->d_ustart, &h->d_hkeys, &h->d_hvals, &h->w_walk, &h->w_walk2};
    for (DevBuf* b : bufs) b->release();
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->d_nguard) cudaFree(h->d_nguard);
    if (h->d_nredo) cudaFree(h->d_nredo);
    for (int r = 0; r < COMM_MAX_WORLD; ++r) if (h->comm.opened[r]) cudaIpcCloseMemHandle(h->comm.peer[r]);
    h->comm.window.release(); h->comm.cnt_out.release();
    if (h->comm.d_err) cudaFree(h->comm.d_err);
    if (h->comm.d_unc) cudaFree(h->comm.d_unc);
    for (int r = 0; r < b2l_ctx::NREC; ++r) {
        for (int i = 0; i < 5; ++i) if (h->ring[r].ev[i]) cudaEventDestroy(h->ring[r].ev[i]);
        if (h->ring[r].h_pc) cudaFreeHost(h->ring[r].h_pc);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int b = 0; b < 2; ++b) { if (h->ev_in[b]) cudaEventDestroy(h->ev_in[b]); if (h->ev_free[b]) cudaEventDestroy(h->ev_free[b]); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return B2L_OK;
}

/* A second handle on the same device that SHARES the parent's model and (finalised) index but has its own stream and
 * workspaces: batches issued alternately on a handle and its siblings overlap on the GPU (the plan / table / selection
 * kernels and the exchange waits of one batch run under the scan of another).  While siblings exist, model and index of
 * the family are frozen (mutating calls fail with B2L_ERR_STATE).  Destroy siblings before the parent. */
int b2l_create_sibling(b2l_handle p, b2l_handle* out) {
    if (!p || !out) return B2L_ERR_ARG;
    *out = nullptr;
    b2l_handle h = p;                                           // (error macros report on the parent)
    std::lock_guard<std::mutex> lk(p->mu);
    CU(cudaSetDevice(p->device));
    if (!p->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (p->parent) FAIL(B2L_ERR_STATE, "create siblings from the parent handle");
    int rc = ensure_index(p);
    if (rc) return rc;
    b2l_handle s = nullptr;
    rc = b2l_create(p->device, &s);
    if (rc) { p->err = g_create_error; return rc; }
    s->has_model = true; s->has_pca = p->has_pca; s->mv = p->mv; s->c2m = p->c2m; s->c2sum = p->c2sum;
    s->fine_mode = p->fine_mode; s->ftc_tabs = p->ftc_tabs; s->scan_mode = p->scan_mode; s->kp_min = p->kp_min; s->force_redo = p->force_redo;
    DevBuf* src[] = {&p->dCs, &p->dmus, &p->dRt, &p->dsubs, &p->dsubs32, &p->dsubs32T, &p->dc2max, &p->dP, &p->dpmu, &p->dftc, &p->dCs32, &p->m_coarse, &p->m_fine,
                     &p->m_rowid, &p->codes, &p->rowids, &p->cell_start, &p->lsize, &p->gsize, &p->sorted_first, &p->d_ucell, &p->d_ustart,
                     &p->d_hkeys, &p->d_hvals};
    DevBuf* dst[] = {&s->dCs, &s->dmus, &s->dRt, &s->dsubs, &s->dsubs32, &s->dsubs32T, &s->dc2max, &s->dP, &s->dpmu, &s->dftc, &s->dCs32, &s->m_coarse, &s->m_fine,
                     &s->m_rowid, &s->codes, &s->rowids, &s->cell_start, &s->lsize, &s->gsize, &s->sorted_first, &s->d_ucell, &s->d_ustart,
                     &s->d_hkeys, &s->d_hvals};
    for (size_t i = 0; i < sizeof(src) / sizeof(src[0]); ++i) dst[i]->borrow(*src[i]);
    s->n_items = p->n_items; s->rows_padded = p->rows_padded; s->dirty = false; s->global_set = p->global_set;
    s->h_lsize = p->h_lsize; s->h_gsize = p->h_gsize; s->h_cell_start = p->h_cell_start;
    s->h_ucell = p->h_ucell; s->h_ustart = p->h_ustart; s->nu = p->nu; s->hmask = p->hmask; s->max_run = p->max_run;
    s->parent = p;
    ++p->n_siblings;
    *out = s;
    return B2L_OK;
}

void* b2l_stream(b2l_handle h) { return h ? (void*)h->stream : nullptr; }

int b2l_debug_candidates(b2l_handle h, int nq, uint32_t* appended, uint32_t* bound_bits) {
    if (!h || nq < 1) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->cand_cnt || !h->gthr || nq > h->pv.nq) FAIL(B2L_ERR_STATE, "no fast-path search of >= %d queries has run", nq);
    if (appended) CU(cudaMemcpyAsync(appended, h->cand_cnt, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    if (bound_bits) CU(cudaMemcpyAsync(bound_bits, h->gthr, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

int b2l_get_stats(b2l_handle h, b2l_stats* out) {
    if (!h || !out) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    int rc = finish_stats(h);                  // waits for searches still in flight
    if (rc) return rc;
    *out = h->stats;
    return B2L_OK;
}

int b2l_reset_stats(b2l_handle h) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    int rc = finish_stats(h);
    if (rc) return rc;
    memset(&h->stats, 0, sizeof h->stats);
    return B2L_OK;
}

int b2l_debug_force_redo(b2l_handle h, int mask) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->force_redo = mask & 7;
    return B2L_OK;
}

int b2l_set_fine_mode(b2l_handle h, int mode) {
    if (!h || mode < 0 || mode > 2) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->fine_mode = mode;
    return B2L_OK;
}

int64_t b2l_encode_guard_count(b2l_handle h, int reset) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    unsigned long long v = 0;
    if (h->d_nguard) {
        CU(cudaMemcpyAsync(&v, h->d_nguard, 8, cudaMemcpyDeviceToHost, h->stream));
        if (reset) CU(cudaMemsetAsync(h->d_nguard, 0, 8, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return (int64_t)v;
}

int b2l_set_scan_mode(b2l_handle h, int mode) {
    if (!h || mode < 0 || mode > 2) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->scan_mode = mode;
    return B2L_OK;
}

int b2l_set_preselect(b2l_handle h, int kp_min) {
    if (!h || kp_min < 0 || kp_min > 512) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->kp_min = kp_min;
    return B2L_OK;
}

int b2l_set_async(b2l_handle h, int enabled) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->async_mode = enabled != 0;
    return B2L_OK;
}

int b2l_sync(b2l_handle h) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return finish_stats(h);
}

int b2l_set_model(b2l_handle h, int D, int V, int M, int K, int coarse_is_f32, const double* Cs, const double* Rs,
                  const double* mus, const double* subs) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!Cs || !Rs || !mus || !subs) FAIL(B2L_ERR_ARG, "NULL model parameter");
    if (D < 2 || (D % 2) || M < 2 || (M % 2) || (D % M) || V < 1 || K < 1)
        FAIL(B2L_ERR_ARG, "bad model shape D=%d V=%d M=%d K=%d (need D%%2==0, M%%2==0, D%%M==0)", D, V, M, K);
    if (K > B2L_MAX_K) FAIL(B2L_ERR_UNSUPPORTED, "subquantizer_clusters=%d > 256 (fine codes are bytes)", K);
    if (V > 65536) FAIL(B2L_ERR_UNSUPPORTED, "V=%d > 65536", V);
    if (h->n_items) FAIL(B2L_ERR_STATE, "clear the index before changing the model");
    if (h->parent || h->n_siblings) FAIL(B2L_ERR_STATE, "the model is shared with sibling handles: destroy them first");
    ModelView& mv = h->mv;
    mv = ModelView();
    mv.D = D; mv.V = V; mv.M = M; mv.K = K; mv.h = D / 2; mv.m = M / 2; mv.ds = D / M; mv.coarse_f32 = coarse_is_f32 ? 1 : 0;
    if (M <= 32) { mv.MP = std::max(4, next_pow2(M)); mv.G = 32 / mv.MP; mv.SW = mv.MP - 1; }
    else { mv.MP = (M + 15) & ~15; mv.G = 0; mv.SW = 0; }
    const size_t hh = (size_t)mv.h;
    const size_t nC = 2 * (size_t)V * hh, nR = nC * hh, nS = (size_t)M * K * mv.ds;
    CU(h->dCs.reserve(nC * 8)); CU(h->dmus.reserve(nC * 8)); CU(h->dRt.reserve(nR * 8)); CU(h->dsubs.reserve(nS * 8));
    std::vector<double> rt(nR);
    for (size_t sc = 0; sc < 2 * (size_t)V; ++sc) {
        const double* R = Rs + sc * hh * hh;
        double* T = rt.data() + sc * hh * hh;
        for (size_t t = 0; t < hh; ++t)
            for (size_t d = 0; d < hh; ++d) T[d * hh + t] = R[t * hh + d];
    }
    CU(cudaMemcpyAsync(h->dCs.p, Cs, nC * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dmus.p, mus, nC * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dRt.p, rt.data(), nR * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dsubs.p, subs, nS * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<float> s32(nS + (size_t)M * K), c2(M);       // float32 centroids, then their half norms [M][K] (k_fine_argmin32)
    for (int j = 0; j < M; ++j) {
        double mx = 0.0;
        for (int k = 0; k < K; ++k) {
            double n2 = 0.0;
            for (int d = 0; d < mv.ds; ++d) { const double v = subs[((size_t)j * K + k) * mv.ds + d]; n2 += v * v; s32[((size_t)j * K + k) * mv.ds + d] = (float)v; }
            mx = std::max(mx, n2);
        }
        for (int k = 0; k < K; ++k) {
            float hn = 0.0f;
            for (int d = 0; d < mv.ds; ++d) { const float v = s32[((size_t)j * K + k) * mv.ds + d]; hn = std::fmaf(v, v, hn); }
            s32[nS + (size_t)j * K + k] = 0.5f * hn;
        }
        c2[j] = (float)(mx * 1.001 + 1e-30);
        h->c2m = std::max(j ? h->c2m : 0.0f, c2[j]);
        h->c2sum = (j ? h->c2sum : 0.0f) + c2[j] * 1.0001f;
    }
    std::vector<float> s32t(nS);
    for (int j = 0; j < M; ++j)
        for (int k = 0; k < K; ++k)
            for (int d = 0; d < mv.ds; ++d) s32t[((size_t)j * mv.ds + d) * K + k] = s32[((size_t)j * K + k) * mv.ds + d];
    CU(h->dsubs32.reserve((nS + (size_t)M * K) * 4)); CU(h->dsubs32T.reserve(nS * 4)); CU(h->dc2max.reserve((size_t)M * 4));
    CU(cudaMemcpyAsync(h->dsubs32T.p, s32t.data(), nS * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dsubs32.p, s32.data(), (nS + (size_t)M * K) * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dc2max.p, c2.data(), (size_t)M * 4, cudaMemcpyHostToDevice, h->stream));
    {   // float32 coarse centroids + half norms + max norm per split (k_coarse_big)
        std::vector<float> c32(nC + 2 * (size_t)V + 2);
        for (int sp = 0; sp < 2; ++sp) {
            float mx = 0.0f;
            for (int v = 0; v < V; ++v) {
                float q = 0.0f;
                for (size_t d = 0; d < hh; ++d) { const float c = (float)Cs[((size_t)sp * V + v) * hh + d]; c32[((size_t)sp * V + v) * hh + d] = c; q = std::fmaf(c, c, q); }
                c32[nC + (size_t)sp * V + v] = 0.5f * q;
                mx = std::max(mx, q);
            }
            c32[nC + 2 * (size_t)V + sp] = std::sqrt(mx) * 1.0001f;
        }
        CU(h->dCs32.reserve(c32.size() * 4));
        CU(cudaMemcpyAsync(h->dCs32.p, c32.data(), c32.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));             // (c32 goes out of scope)
        mv.Cs32 = h->dCs32.as<float>(); mv.Chn32 = mv.Cs32 + nC; mv.Cmax32 = mv.Chn32 + 2 * (size_t)V;
    }
    h->ftc_tabs = nullptr;
    std::vector<float> ftab;
    if (M > 64) {}                     // (the kernel keeps per-sub-quantizer constants for M <= 64)
    else if (mv.ds == 8) { ftab.resize((size_t)M * (FtcGeo<8>::B_BYTES / 4)); ftc_build_tables<8>(M, K, subs, ftab.data()); }
    else if (mv.ds == 16) { ftab.resize((size_t)M * (FtcGeo<16>::B_BYTES / 4)); ftc_build_tables<16>(M, K, subs, ftab.data()); }
    if (!ftab.empty()) {
        CU(h->dftc.reserve(ftab.size() * 4));
        CU(cudaMemcpyAsync(h->dftc.p, ftab.data(), ftab.size() * 4, cudaMemcpyHostToDevice, h->stream));
        h->ftc_tabs = h->dftc.as<float>();
    }
    CU(cudaStreamSynchronize(h->stream));
    mv.subs32 = h->dsubs32.as<float>(); mv.subs32T = h->dsubs32T.as<float>(); mv.c2max = h->dc2max.as<float>();
    mv.Cs = h->dCs.as<double>(); mv.mus = h->dmus.as<double>(); mv.Rt = h->dRt.as<double>(); mv.subs = h->dsubs.as<double>();
    h->has_model = true; h->has_pca = false; h->dirty = true; h->global_set = false;
    return B2L_OK;
}

int b2l_set_pca(b2l_handle h, int D0, const double* P, const double* mu, int renorm) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "set the model before the PCA");
    if (h->parent || h->n_siblings) FAIL(B2L_ERR_STATE, "the model is shared with sibling handles: destroy them first");
    if (D0 < 1 || !P || !mu) FAIL(B2L_ERR_ARG, "bad PCA arguments");
    if ((size_t)(D0 + h->mv.D) * 8 > 200 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "PCA input dimension %d too large", D0);
    CU(h->dP.reserve((size_t)D0 * h->mv.D * 8)); CU(h->dpmu.reserve((size_t)D0 * 8));
    CU(cudaMemcpyAsync(h->dP.p, P, (size_t)D0 * h->mv.D * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dpmu.p, mu, (size_t)D0 * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->mv.D0 = D0; h->mv.renorm = renorm ? 1 : 0; h->mv.P = h->dP.as<double>(); h->mv.pmu = h->dpmu.as<double>();
    h->has_pca = true;
    return B2L_OK;
}

int b2l_encode(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, int32_t* coarse, uint8_t* fine) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (n < 0 || (n && (!X || !coarse))) FAIL(B2L_ERR_ARG, "bad encode arguments");
    const ModelView& mv = h->mv;
    const int Din = h->has_pca ? mv.D0 : mv.D;
    const size_t esz = x_is_f64 ? 8 : 4;
    h->launches = 0;
    // chunk so that the float64 projection workspace stays <= 1 GiB
    int64_t chunk = std::max<int64_t>(1024, ((int64_t)1 << 30) / ((int64_t)mv.D * 8));
    chunk = std::min<int64_t>(chunk, (int64_t)1 << 22);
    if (on_device) {
        for (int64_t a = 0; a < n; a += chunk) {
            const int64_t c = std::min(chunk, n - a);
            int rc = encode_device(h, (const char*)X + (size_t)a * Din * esz, x_is_f64, c, nullptr, coarse + a * 2, fine ? fine + a * mv.M : nullptr);
            if (rc) return rc;
        }
    } else if (n) {
        // host rows: blocks of <= 256K rows through two staging buffers; the rows of block i+1 travel on a second stream
        // while block i is encoded (pinned host memory makes the copies asynchronous; pageable memory still works)
        const int64_t blk = std::min<int64_t>(chunk, (int64_t)1 << 18);
        if (!h->copy_stream) {
            CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) { CU(cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&h->ev_free[b], cudaEventDisableTiming)); }
        }
        const size_t qb = (((size_t)blk * Din * esz) + 255) & ~(size_t)255, cb = (size_t)blk * 8, fb = (((size_t)blk * mv.M) + 255) & ~(size_t)255;
        CU(h->w_q.reserve(2 * qb)); CU(h->w_coarse.reserve(2 * cb)); CU(h->w_fine.reserve(2 * fb));
        CU(cudaStreamSynchronize(h->stream));              // earlier work on the staging buffers is over
        // the copy of block i+1 is enqueued BEFORE block i's results are read back: a read-back into pageable host memory
        // blocks the calling thread until block i is done, and the next rows travel meanwhile
        auto rows_in = [&](int i) -> int {
            const int b = i & 1;
            const int64_t a = (int64_t)i * blk, c = std::min(blk, n - a);
            if (i >= 2) CU(cudaStreamWaitEvent(h->copy_stream, h->ev_free[b], 0));      // block i-2 is through with this buffer
            CU(cudaMemcpyAsync((char*)h->w_q.p + b * qb, (const char*)X + (size_t)a * Din * esz, (size_t)c * Din * esz, cudaMemcpyHostToDevice, h->copy_stream));
            CU(cudaEventRecord(h->ev_in[b], h->copy_stream));
            return B2L_OK;
        };
        const int nblk = (int)((n + blk - 1) / blk);
        { int rc = rows_in(0); if (rc) return rc; }
        for (int i = 0; i < nblk; ++i) {
            const int b = i & 1;
            const int64_t a = (int64_t)i * blk, c = std::min(blk, n - a);
            char* dq = (char*)h->w_q.p + b * qb;
            int32_t* dco = (int32_t*)((char*)h->w_coarse.p + b * cb);
            uint8_t* dfi = fine ? (uint8_t*)h->w_fine.p + b * fb : nullptr;
            if (i + 1 < nblk) { int rc = rows_in(i + 1); if (rc) return rc; }
            CU(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
            int rc = encode_device(h, dq, x_is_f64, c, nullptr, dco, dfi);
            if (rc) { cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream); return rc; }
            CU(cudaEventRecord(h->ev_free[b], h->stream));
            CU(cudaMemcpyAsync(coarse + a * 2, dco, (size_t)c * 8, cudaMemcpyDeviceToHost, h->stream));
            if (fine) CU(cudaMemcpyAsync(fine + a * mv.M, dfi, (size_t)c * mv.M, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->copy_stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    h->stats.kernel_launches = h->launches;
    return B2L_OK;
}

int b2l_debug_fine_scores(b2l_handle h, const void* X, int x_is_f64, int64_t n, int j, float* scores, double* px) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    const ModelView& mv = h->mv;
    if (!h->ftc_tabs || h->fine_mode != 0) FAIL(B2L_ERR_UNSUPPORTED, "the tensor-core fine argmin does not cover this model (ds = %d) or is switched off", mv.ds);
    if (!X || !scores || n < 128 || n > ((int64_t)1 << 20) || j < 0 || j >= mv.M) FAIL(B2L_ERR_ARG, "bad arguments (need 128 <= n <= 2^20 host rows, 0 <= j < M)");
    const int Din = h->has_pca ? mv.D0 : mv.D;
    const size_t esz = x_is_f64 ? 8 : 4;
    CU(h->w_q.reserve((size_t)n * Din * esz)); CU(h->w_coarse.reserve((size_t)n * 8)); CU(h->w_fine.reserve((size_t)n * mv.M));
    CU(h->w_ftc_dbg.reserve((size_t)128 * 256 * 4));
    CU(cudaMemcpyAsync(h->w_q.p, X, (size_t)n * Din * esz, cudaMemcpyHostToDevice, h->stream));
    h->ftc_dbg_j = j;
    const int rc = encode_device(h, h->w_q.p, x_is_f64, n, nullptr, h->w_coarse.as<int32_t>(), h->w_fine.as<uint8_t>());
    h->ftc_dbg_j = -1;
    if (rc) return rc;
    CU(cudaMemcpyAsync(scores, h->w_ftc_dbg.p, (size_t)128 * 256 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (px) CU(cudaMemcpyAsync(px, h->w_px.p, (size_t)128 * mv.D * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

static int apply_pca_impl(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, float* Y, double* Y64) {
    if (!h->has_pca) FAIL(B2L_ERR_STATE, "no PCA set");
    if (n < 1 || !X || (!Y && !Y64)) FAIL(B2L_ERR_ARG, "bad apply_pca arguments");
    const ModelView& mv = h->mv;
    const size_t esz = x_is_f64 ? 8 : 4;
    const size_t osz = Y64 ? 8 : 4;
    const void* dx = X;
    if (!on_device) {
        CU(h->w_q.reserve((size_t)n * mv.D0 * esz));
        CU(cudaMemcpyAsync(h->w_q.p, X, (size_t)n * mv.D0 * esz, cudaMemcpyHostToDevice, h->stream));
        dx = h->w_q.p;
    }
    void* dy = Y64 ? (void*)Y64 : (void*)Y;
    if (!on_device) { CU(h->w_xq.reserve((size_t)n * mv.D * osz)); dy = h->w_xq.p; }
    float* dy32 = Y64 ? nullptr : (float*)dy;
    double* dy64 = Y64 ? (double*)dy : nullptr;
    const size_t smem = (size_t)(mv.D0 + mv.D) * 8;
    if (x_is_f64) { CU(cudaFuncSetAttribute(k_pca<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pca<double><<<(unsigned)n, 128, smem, h->stream>>>(mv, (const double*)dx, n, dy32, dy64); }
    else { CU(cudaFuncSetAttribute(k_pca<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pca<float><<<(unsigned)n, 128, smem, h->stream>>>(mv, (const float*)dx, n, dy32, dy64); }
    LAUNCHED();
    if (!on_device) CU(cudaMemcpyAsync(Y64 ? (void*)Y64 : (void*)Y, dy, (size_t)n * mv.D * osz, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

int b2l_apply_pca(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, float* Y) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    return apply_pca_impl(h, X, x_is_f64, n, on_device, Y, nullptr);
}

int b2l_apply_pca64(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, double* Y) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    return apply_pca_impl(h, X, x_is_f64, n, on_device, nullptr, Y);
}

int b2l_project_lut(b2l_handle h, const void* X, int x_is_f64, int64_t n, const int32_t* coarse, double* px, double* lut) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (n < 1 || !X || !coarse) FAIL(B2L_ERR_ARG, "bad project arguments");
    const ModelView& mv = h->mv;
    const size_t esz = x_is_f64 ? 8 : 4;
    CU(h->w_q.reserve((size_t)n * mv.D * esz)); CU(h->w_coarse.reserve((size_t)n * 8));
    CU(cudaMemcpyAsync(h->w_q.p, X, (size_t)n * mv.D * esz, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->w_coarse.p, coarse, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    // project only (no PCA here: x is already D-dimensional)
    const bool pca = h->has_pca;
    h->has_pca = false;
    int rc = encode_device(h, h->w_q.p, x_is_f64, n, h->w_coarse.as<int32_t>(), nullptr, nullptr);
    h->has_pca = pca;
    if (rc) return rc;
    if (px) CU(cudaMemcpyAsync(px, h->w_px.p, (size_t)n * mv.D * 8, cudaMemcpyDeviceToHost, h->stream));
    if (lut) {
        CU(h->w_lut64.reserve((size_t)n * mv.M * mv.K * 8));
        k_lut64_probe<<<(unsigned)n, 256, 0, h->stream>>>(mv, h->w_px.as<double>(), n, h->w_lut64.as<double>());
        LAUNCHED();
        CU(cudaMemcpyAsync(lut, h->w_lut64.p, (size_t)n * mv.M * mv.K * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

int b2l_index_add(b2l_handle h, const int32_t* coarse, const uint8_t* fine, int64_t n, const int64_t* rowids, int on_device) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (n < 0 || (n && (!coarse || !fine))) FAIL(B2L_ERR_ARG, "bad index_add arguments");
    if (h->parent || h->n_siblings) FAIL(B2L_ERR_STATE, "the index is shared with sibling handles: destroy them first");
    if (n == 0) return B2L_OK;
    const ModelView& mv = h->mv;
    const int64_t tot = h->n_items + n;
    if (tot >= ((int64_t)1 << 31)) FAIL(B2L_ERR_UNSUPPORTED, "more than 2^31 - 1 rows per shard (%lld)", (long long)tot);
    CU(h->m_coarse.reserve((size_t)tot * 8, true, h->stream));
    CU(h->m_fine.reserve((size_t)tot * mv.M, true, h->stream));
    CU(h->m_rowid.reserve((size_t)tot * 8, true, h->stream));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpyAsync(h->m_coarse.as<int32_t>() + h->n_items * 2, coarse, (size_t)n * 8, kind, h->stream));
    CU(cudaMemcpyAsync(h->m_fine.as<uint8_t>() + h->n_items * mv.M, fine, (size_t)n * mv.M, kind, h->stream));
    if (rowids) CU(cudaMemcpyAsync(h->m_rowid.as<int64_t>() + h->n_items, rowids, (size_t)n * 8, kind, h->stream));
    else { k_iota64<<<grid_for(n, 256), 256, 0, h->stream>>>(h->m_rowid.as<int64_t>() + h->n_items, h->n_items, n); LAUNCHED(); }
    // the batch is committed only if every coarse code is in range (the reference logs and skips a bad item,
    // search.py:343-367; here the whole batch is refused and the index stays as it was)
    CU(h->w_misc.reserve(64));
    CU(cudaMemsetAsync(h->w_misc.p, 0, 4, h->stream));
    k_check_coarse<<<grid_for(2 * n, 256), 256, 0, h->stream>>>(h->m_coarse.as<int32_t>() + h->n_items * 2, n, mv.V, h->w_misc.as<int>());
    LAUNCHED();
    int hbad = 0;
    CU(cudaMemcpyAsync(&hbad, h->w_misc.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (hbad) FAIL(B2L_ERR_ARG, "index_add: coarse code out of range [0, %d); nothing was added", mv.V);
    h->n_items = tot;
    h->dirty = true;
    return B2L_OK;
}

int b2l_index_clear(b2l_handle h) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->parent || h->n_siblings) FAIL(B2L_ERR_STATE, "the index is shared with sibling handles: destroy them first");
    h->n_items = 0; h->dirty = true; h->global_set = false;
    return B2L_OK;
}

int64_t b2l_index_size(b2l_handle h) { return h ? h->n_items : -1; }

int b2l_index_cell_sizes(b2l_handle h, int64_t* sizes) {
    if (!h || !sizes) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    int rc = ensure_index(h);
    if (rc) return rc;
    if (h->mv.V > B2L_MAX_V) {             // sparse directory -> the dense array the caller asked for
        memset(sizes, 0, (size_t)h->mv.V * h->mv.V * 8);
        for (unsigned int r = 0; r < h->nu; ++r) sizes[h->h_ucell[r]] = (int64_t)(h->h_ustart[r + 1] - h->h_ustart[r]);
        return B2L_OK;
    }
    memcpy(sizes, h->h_lsize.data(), h->h_lsize.size() * 8);
    return B2L_OK;
}

int b2l_index_set_global_cell_sizes(b2l_handle h, const int64_t* sizes) {
    if (!h || !sizes) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (h->mv.V > B2L_MAX_V) FAIL(B2L_ERR_UNSUPPORTED, "a cell-sharded index is not supported at V > %d", B2L_MAX_V);
    if (h->parent || h->n_siblings) FAIL(B2L_ERR_STATE, "the index is shared with sibling handles: destroy them first");
    const int ncell = h->mv.V * h->mv.V;
    h->h_gsize.assign(sizes, sizes + ncell);
    h->global_set = true;
    if (!h->dirty) {
        CU(cudaMemcpyAsync(h->gsize.p, h->h_gsize.data(), (size_t)ncell * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return B2L_OK;
}

int64_t b2l_index_get_cell(b2l_handle h, int c0, int c1, int64_t cap, int64_t* rowids, uint8_t* fine) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    const ModelView& mv = h->mv;
    if (c0 < 0 || c0 >= mv.V || c1 < 0 || c1 >= mv.V) return 0;
    int rc = ensure_index(h);
    if (rc) return rc;
    const int cell = c0 * mv.V + c1;
    int64_t n, s0 = 0;
    if (mv.V > B2L_MAX_V) {
        auto it = std::lower_bound(h->h_ucell.begin(), h->h_ucell.end(), (unsigned int)cell);
        if (it == h->h_ucell.end() || *it != (unsigned int)cell) return 0;
        const size_t r = it - h->h_ucell.begin();
        s0 = h->h_ustart[r]; n = (int64_t)h->h_ustart[r + 1] - s0;
    } else { n = h->h_lsize[cell]; s0 = n ? h->h_cell_start[cell] : 0; }
    const int64_t take = std::min(n, cap);
    if (take > 0) {
        const int64_t s = s0;
        if (rowids) CU(cudaMemcpyAsync(rowids, h->rowids.as<int64_t>() + s, (size_t)take * 8, cudaMemcpyDeviceToHost, h->stream));
        if (fine) {
            CU(h->w_fine.reserve((size_t)take * mv.M));
            k_unswizzle_rows<<<grid_for(take * mv.M, 256), 256, 0, h->stream>>>(h->codes.as<uint8_t>() + s * mv.MP, take, mv.M, mv.MP,
                                                                                  mv.SW, h->w_fine.as<uint8_t>());
            LAUNCHED();
            CU(cudaMemcpyAsync(fine, h->w_fine.p, (size_t)take * mv.M, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
    }
    return n;
}

int b2l_cell_order(b2l_handle h, const void* Q, int q_is_f64, int nq, int64_t quota, int32_t* cells, double* dists, int32_t* nvis) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (nq < 1 || !Q || !nvis) FAIL(B2L_ERR_ARG, "bad cell_order arguments");
    if (h->mv.V > B2L_MAX_V) FAIL(B2L_ERR_UNSUPPORTED, "V=%d > %d: use b2l_cell_order_prefix (the full V*V order is not materialised)", h->mv.V, B2L_MAX_V);
    int rc = ensure_index(h);
    if (rc) return rc;
    const ModelView& mv = h->mv;
    const int maxvis = mv.V * mv.V;
    const size_t esz = q_is_f64 ? 8 : 4;
    CU(h->w_q.reserve((size_t)nq * mv.D * esz));
    CU(cudaMemcpyAsync(h->w_q.p, Q, (size_t)nq * mv.D * esz, cudaMemcpyHostToDevice, h->stream));
    rc = setup_plan(h, nq, 1024);
    if (rc) return rc;
    PlanView& pv = h->pv;
    CU(h->w_misc.reserve((size_t)nq * maxvis * 8));
    pv.vis_dist = h->w_misc.as<double>();
    const size_t smem = (size_t)(3 * mv.V + 2) * 8 + (size_t)(7 * mv.V + 4) * 4 + 16;
    if (q_is_f64) k_coarse_order<double><<<nq, 32, smem, h->stream>>>(mv, h->w_q.as<double>(), quota, h->gsize.as<int64_t>(), h->lsize.as<int64_t>(), pv, InitView());
    else k_coarse_order<float><<<nq, 32, smem, h->stream>>>(mv, h->w_q.as<float>(), quota, h->gsize.as<int64_t>(), h->lsize.as<int64_t>(), pv, InitView());
    LAUNCHED();
    if (cells) CU(cudaMemcpyAsync(cells, pv.vis_cell, (size_t)nq * maxvis * 4, cudaMemcpyDeviceToHost, h->stream));
    if (dists) CU(cudaMemcpyAsync(dists, pv.vis_dist, (size_t)nq * maxvis * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(nvis, pv.nvis, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    pv.vis_dist = nullptr;
    return B2L_OK;
}

int b2l_cell_order_prefix(b2l_handle h, const void* Q, int q_is_f64, int nq, int64_t quota, int max_cells,
                          int32_t* cells, double* dists, int32_t* nvis) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (nq < 1 || !Q || !nvis || max_cells < 1) FAIL(B2L_ERR_ARG, "bad cell_order_prefix arguments");
    const ModelView& mv = h->mv;
    if (mv.V > B2L_MAX_V_SPARSE) FAIL(B2L_ERR_UNSUPPORTED, "V=%d > %d", mv.V, B2L_MAX_V_SPARSE);
    if (h->global_set) FAIL(B2L_ERR_UNSUPPORTED, "b2l_cell_order_prefix works on an unsharded index");
    if (mv.V <= B2L_MAX_V) FAIL(B2L_ERR_UNSUPPORTED, "V=%d <= %d: use b2l_cell_order", mv.V, B2L_MAX_V);
    int rc = ensure_index(h);
    if (rc) return rc;
    const size_t esz = q_is_f64 ? 8 : 4;
    CU(h->w_q.reserve((size_t)nq * mv.D * esz));
    CU(cudaMemcpyAsync(h->w_q.p, Q, (size_t)nq * mv.D * esz, cudaMemcpyHostToDevice, h->stream));
    WalkView wv;
    const int segcap = (int)std::min<int64_t>(std::min<int64_t>(std::max<int64_t>(quota, 1), (int64_t)h->nu), (int64_t)max_cells) + 2;
    if ((rc = setup_walk(h, nq, segcap, (size_t)nq * 2 * std::min(mv.V, segcap), max_cells, wv))) return rc;
    wv.max_visit = max_cells;
    if ((rc = launch_walk(h, h->w_q.p, q_is_f64, nq, quota, wv))) return rc;
    WalkCounters wc;
    CU(cudaMemcpyAsync(&wc, wv.cnt, sizeof wc, cudaMemcpyDeviceToHost, h->stream));
    if (cells) CU(cudaMemcpyAsync(cells, wv.vis_cells, (size_t)nq * max_cells * 4, cudaMemcpyDeviceToHost, h->stream));
    if (dists) CU(cudaMemcpyAsync(dists, wv.vis_dists, (size_t)nq * max_cells * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(nvis, wv.nvis, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (wc.err == 1) FAIL(B2L_ERR_UNSUPPORTED, "more than %d cells at one and the same coarse distance: the traversal order is degenerate", WALK_CAP);
    return B2L_OK;
}

int64_t b2l_records_bytes(b2l_handle h, int nq, int k) {
    if (!h || !h->has_model || nq < 1 || k < 1) return -1;
    return (int64_t)rec_bytes(nq, k, h->mv.M);
}

int b2l_search_local(b2l_handle h, const void* Q, int q_is_f64, int nq, int on_device, int64_t quota, int k, int exact,
                     void* d_records) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    return search_local_impl(h, Q, q_is_f64, nq, on_device, quota, k, exact, d_records, !h->async_mode);
}

int b2l_search_merge(b2l_handle h, const void* d_records_all, int nranks, int nq, int k, int on_device, int64_t* rowid,
                     double* dist, int32_t* coarse, uint8_t* fine, int32_t* count, int32_t* visited, uint8_t* certified) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    return merge_impl(h, d_records_all, nranks, nq, k, on_device, rowid, dist, coarse, fine, count, visited, certified,
                      !h->async_mode);
}

int64_t b2l_merge_block_bytes(b2l_handle h, int nq, int k) {
    if (!h || !h->has_model || nq < 1 || k < 1) return -1;
    const size_t nk = (size_t)nq * k;
    return (int64_t)(3 * align256(nk * 8) + align256(nk * h->mv.M) + 2 * align256((size_t)nq * 4) + align256((size_t)nq));
}

int b2l_search_merge_block(b2l_handle h, const void* d_records_all, int nranks, int nq, int k, void* block, int on_device) {
    if (!h || !block) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    const ModelView& mv = h->mv;
    const size_t nk = (size_t)nq * k;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o1 = take(nk * 8), o2 = take(nk * 8), o3 = take(nk * 8), o4 = take(nk * mv.M), o5 = take((size_t)nq * 4),
                 o6 = take((size_t)nq * 4), o7 = take((size_t)nq);
    unsigned char* b = (unsigned char*)block;
    if (!on_device) { CU(h->w_out.reserve(off)); b = h->w_out.as<unsigned char>(); }
    int rc = merge_impl(h, d_records_all, nranks, nq, k, 1, (int64_t*)(b + o1), (double*)(b + o2), (int32_t*)(b + o3), b + o4,
                        (int32_t*)(b + o5), (int32_t*)(b + o6), b + o7, false);
    if (rc) return rc;
    if (!on_device) CU(cudaMemcpyAsync(block, b, off, cudaMemcpyDeviceToHost, h->stream));
    if (!h->async_mode) CU(cudaStreamSynchronize(h->stream));
    return B2L_OK;
}

int b2l_search(b2l_handle h, const void* Q, int q_is_f64, int nq, int on_device, int64_t quota, int k, int64_t* rowid,
               double* dist, int32_t* coarse, uint8_t* fine, int32_t* count, int32_t* visited) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "no model set");
    if (nq < 1 || k < 1 || !Q || !count) FAIL(B2L_ERR_ARG, "bad search arguments (nq=%d k=%d)", nq, k);
    const ModelView& mv = h->mv;
    CU(h->w_rec.reserve(rec_bytes(nq, k, mv.M)));
    // everything up to the certification flags is queued without a host round trip of its own
    int rc = search_local_impl(h, Q, q_is_f64, nq, on_device, quota, k, 0, h->w_rec.p, false);
    if (rc) return rc;
    // final outputs: one device block [rowid | dist | coarse | fine | count | visited | certified]
    const size_t nk = (size_t)nq * k;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o1 = take(nk * 8), o2 = take(nk * 8), o3 = take(nk * 8), o4 = take(nk * mv.M), o5 = take((size_t)nq * 4),
                 o6 = take((size_t)nq * 4), o7 = take((size_t)nq);
    CU(h->w_out2.reserve(off));
    unsigned char* b = h->w_out2.as<unsigned char>();
    int64_t* d_rowid = (int64_t*)(b + o1); double* d_dist = (double*)(b + o2); int32_t* d_coarse = (int32_t*)(b + o3);
    uint8_t* d_fine = b + o4; int32_t* d_count = (int32_t*)(b + o5); int32_t* d_visited = (int32_t*)(b + o6); uint8_t* d_cert = b + o7;
    rc = merge_impl(h, h->w_rec.p, 1, nq, k, 1, d_rowid, d_dist, d_coarse, d_fine, d_count, d_visited, d_cert, false);
    if (rc) return rc;
    // host side: the whole block comes back in one copy into pinned staging (device-resident callers: flags only)
    if (h->h_out_cap < off) {
        if (h->h_out) cudaFreeHost(h->h_out);
        h->h_out = nullptr; h->h_out_cap = 0;
        CU(cudaHostAlloc(&h->h_out, off + off / 2, cudaHostAllocDefault));
        h->h_out_cap = off + off / 2;
    }
    unsigned char* hb = (unsigned char*)h->h_out;
    if (on_device) CU(cudaMemcpyAsync(hb + o7, d_cert, nq, cudaMemcpyDeviceToHost, h->stream));
    else CU(cudaMemcpyAsync(hb, b, off, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if ((rc = finish_stats(h))) return rc;
    ++h->stats.kernel_launches; ++h->stats.acc_kernel_launches;      // k_final
    const b2l_stats st = h->stats;
    const uint8_t* cert = hb + o7;
    std::vector<int> redo;
    for (int q = 0; q < nq; ++q) if (!cert[q] || (h->force_redo & 1)) redo.push_back(q);
    // Uncertified queries go down the chain: float32 tables (if the first pass used the 16-bit ones), then the float64
    // full sort.  Each stage re-runs the subset, merges it into a scratch block and patches the output rows.
    int64_t n_rescan = 0, n_exact = 0, extra_launches = 0;
    for (int stage = (st.packed ? 2 : 1); !redo.empty(); stage = 1) {
        const int ns = (int)redo.size();
        const int Din = h->has_pca ? mv.D0 : mv.D;
        const size_t qrow = (size_t)Din * (q_is_f64 ? 8 : 4);
        CU(h->w_misc.reserve((size_t)ns * qrow));
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        for (int i = 0; i < ns; ++i)
            CU(cudaMemcpyAsync(h->w_misc.as<char>() + i * qrow, (const char*)Q + redo[i] * qrow, qrow, kind, h->stream));
        CU(h->w_rec2.reserve(rec_bytes(ns, k, mv.M)));
        rc = search_local_impl(h, h->w_misc.p, q_is_f64, ns, 1, quota, k, stage, h->w_rec2.p, false);
        if (rc) return rc;
        const size_t sk = (size_t)ns * k;
        size_t so = 0;
        auto stake = [&](size_t bytes) { size_t o = so; so += align256(bytes); return o; };
        const size_t s1 = stake(sk * 8), s2 = stake(sk * 8), s3 = stake(sk * 8), s4 = stake(sk * mv.M), s5 = stake((size_t)ns * 4),
                     s6 = stake((size_t)ns * 4), s7 = stake((size_t)ns);
        CU(h->w_out.reserve(so));
        unsigned char* sb = h->w_out.as<unsigned char>();
        rc = merge_impl(h, h->w_rec2.p, 1, ns, k, 1, (int64_t*)(sb + s1), (double*)(sb + s2), (int32_t*)(sb + s3), sb + s4,
                        (int32_t*)(sb + s5), (int32_t*)(sb + s6), sb + s7, false);
        if (rc) return rc;
        std::vector<uint8_t> c2(ns);
        CU(cudaMemcpyAsync(c2.data(), sb + s7, ns, cudaMemcpyDeviceToHost, h->stream));
        for (int i = 0; i < ns; ++i) {
            const size_t q = (size_t)redo[i];
            CU(cudaMemcpyAsync(d_rowid + q * k, (int64_t*)(sb + s1) + (size_t)i * k, (size_t)k * 8, cudaMemcpyDeviceToDevice, h->stream));
            CU(cudaMemcpyAsync(d_dist + q * k, (double*)(sb + s2) + (size_t)i * k, (size_t)k * 8, cudaMemcpyDeviceToDevice, h->stream));
            CU(cudaMemcpyAsync(d_coarse + q * k * 2, (int32_t*)(sb + s3) + (size_t)i * k * 2, (size_t)k * 8, cudaMemcpyDeviceToDevice, h->stream));
            CU(cudaMemcpyAsync(d_fine + q * k * mv.M, sb + s4 + (size_t)i * k * mv.M, (size_t)k * mv.M, cudaMemcpyDeviceToDevice, h->stream));
            CU(cudaMemcpyAsync(d_count + q, (int32_t*)(sb + s5) + i, 4, cudaMemcpyDeviceToDevice, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        if ((rc = finish_stats(h))) return rc;
        extra_launches += h->stats.kernel_launches + 1;
        if (stage == 2) n_rescan += ns; else n_exact += ns;
        std::vector<int> next;
        if (stage == 2) for (int i = 0; i < ns; ++i) if (!c2[i] || (h->force_redo & 2)) next.push_back(redo[i]);
        redo.swap(next);
    }
    if (n_rescan || n_exact) {
        const b2l_stats s2 = h->stats;           // carries the accumulators of the reruns
        h->stats = st;
        h->stats.kernel_launches = st.kernel_launches + extra_launches;
        h->stats.exact_queries = n_exact; h->stats.rescan_queries = n_rescan;
        h->stats.acc_kernel_launches = s2.acc_kernel_launches + (n_rescan ? 1 : 0) + (n_exact ? 1 : 0);
        h->stats.acc_exact_queries = st.acc_exact_queries + n_exact;
        h->stats.acc_rescan_queries = st.acc_rescan_queries + n_rescan;
        if (!on_device) CU(cudaMemcpyAsync(hb, b, off, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (on_device) {
        if ((rc = copy_out(h, rowid, d_rowid, nk * 8, 1))) return rc;
        if ((rc = copy_out(h, dist, d_dist, nk * 8, 1))) return rc;
        if ((rc = copy_out(h, coarse, d_coarse, nk * 8, 1))) return rc;
        if ((rc = copy_out(h, fine, d_fine, nk * mv.M, 1))) return rc;
        if ((rc = copy_out(h, count, d_count, (size_t)nq * 4, 1))) return rc;
        if ((rc = copy_out(h, visited, d_visited, (size_t)nq * 4, 1))) return rc;
        CU(cudaStreamSynchronize(h->stream));
    } else {
        if (rowid) memcpy(rowid, hb + o1, nk * 8);
        if (dist) memcpy(dist, hb + o2, nk * 8);
        if (coarse) memcpy(coarse, hb + o3, nk * 8);
        if (fine) memcpy(fine, hb + o4, nk * mv.M);
        memcpy(count, hb + o5, (size_t)nq * 4);
        if (visited) memcpy(visited, hb + o6, (size_t)nq * 4);
    }
    return B2L_OK;
}

// ---- multi-GPU exchange inside the library (comm.cuh) -------------------------------------------------------------
static size_t comm_slot_q_off(const b2l_ctx::Comm& c, int slot) { return COMM_FLAG_BYTES + (size_t)slot * (c.q_region + c.r_region); }
static size_t comm_slot_r_off(const b2l_ctx::Comm& c, int slot) { return comm_slot_q_off(c, slot) + c.q_region; }
static size_t comm_flag_off(const b2l_ctx::Comm& c, int slot, int kind) { return ((size_t)(slot * 3 + kind) * COMM_MAX_WORLD) * 8; }

int b2l_comm_init(b2l_handle h, int world, int rank, int max_nq_home, int max_k, int q_is_f64) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->has_model) FAIL(B2L_ERR_STATE, "set the model before b2l_comm_init");
    if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || max_nq_home < 1 || max_k < 1)
        FAIL(B2L_ERR_ARG, "bad comm arguments (world=%d rank=%d)", world, rank);
    b2l_ctx::Comm& c = h->comm;
    if (c.connected || c.window.p) FAIL(B2L_ERR_STATE, "the exchange is already initialised on this handle");
    const ModelView& mv = h->mv;
    c.world = world; c.rank = rank; c.max_home = max_nq_home; c.max_k = max_k;
    c.qrow = (size_t)(h->has_pca ? mv.D0 : mv.D) * (q_is_f64 ? 8 : 4);
    c.q_region = align256((size_t)world * max_nq_home * c.qrow);
    c.r_region = (size_t)world * rec_bytes(max_nq_home, max_k, mv.M);
    const size_t total = COMM_FLAG_BYTES + (size_t)COMM_SLOTS * (c.q_region + c.r_region);
    CU(c.window.reserve(total));
    CU(cudaMemsetAsync(c.window.p, 0, COMM_FLAG_BYTES, h->stream));
    if (!c.d_err) { CU(cudaMalloc((void**)&c.d_err, 4)); CU(cudaMalloc((void**)&c.d_unc, 4)); }
    CU(cudaMemsetAsync(c.d_err, 0, 4, h->stream));
    CU(c.cnt_out.reserve(256));
    CU(cudaStreamSynchronize(h->stream));
    c.peer[rank] = c.window.as<unsigned char>();
    c.seq = 0;
    return B2L_OK;
}

int b2l_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int b2l_comm_get_handle(b2l_handle h, void* ipc_handle_out, void** local_ptr_out) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->comm.window.p) FAIL(B2L_ERR_STATE, "b2l_comm_init first");
    if (local_ptr_out) *local_ptr_out = h->comm.window.p;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t mh;
        CU(cudaIpcGetMemHandle(&mh, h->comm.window.p));
        memcpy(ipc_handle_out, &mh, sizeof mh);
    }
    return B2L_OK;
}

int b2l_comm_connect(b2l_handle h, const void* handles, int same_process) {
    if (!h || !handles) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    b2l_ctx::Comm& c = h->comm;
    if (!c.window.p) FAIL(B2L_ERR_STATE, "b2l_comm_init first");
    for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) continue;
        if (same_process) {
            c.peer[r] = (unsigned char*)((void* const*)handles)[r];
            cudaPointerAttributes at;
            CU(cudaPointerGetAttributes(&at, c.peer[r]));
            if (at.device != h->device) {                       // another GPU of this process: map it
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, h->device, at.device));
                if (!can) FAIL(B2L_ERR_UNSUPPORTED, "device %d cannot access device %d", h->device, at.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
                cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, (const char*)handles + (size_t)r * sizeof mh, sizeof mh);
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
            c.peer[r] = (unsigned char*)p; c.opened[r] = true;
        }
    }
    c.connected = true;
    return B2L_OK;
}

int64_t b2l_sharded_block_bytes(b2l_handle h, int nq_home, int k) {
    const int64_t b = b2l_merge_block_bytes(h, nq_home, k);
    return b < 0 ? b : b + 256;
}

/* One batch of the cell-sharded search with the exchange inside the library.  Every rank passes ITS home slice of the
 * batch (nq_home queries, the same number on every rank); the global batch is rank-major.  Asynchronous: everything is
 * enqueued on the handle's stream; `block` (pinned host, or device) receives the merge block of the home queries followed by
 * [world] int32 counts of queries each rank could not certify (all zero: the results are final). */
int b2l_search_sharded(b2l_handle h, const void* Qhome, int q_is_f64, int nq_home, int on_device, int64_t quota, int k,
                       void* block, int block_on_device) {
    if (!h || !Qhome || !block) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    b2l_ctx::Comm& c = h->comm;
    if (!c.connected && c.world != 1) FAIL(B2L_ERR_STATE, "b2l_comm_connect first");
    if (!c.window.p) FAIL(B2L_ERR_STATE, "b2l_comm_init first");
    const ModelView& mv = h->mv;
    const size_t qrow = (size_t)(h->has_pca ? mv.D0 : mv.D) * (q_is_f64 ? 8 : 4);
    if (nq_home < 1 || nq_home > c.max_home || k < 1 || k > c.max_k || qrow != c.qrow)
        FAIL(B2L_ERR_ARG, "search_sharded: nq_home=%d k=%d outside what b2l_comm_init reserved (%lld, %d) or query type changed",
             nq_home, k, (long long)c.max_home, c.max_k);
    if ((nq_home * qrow) % 16) FAIL(B2L_ERR_ARG, "home slice of %zu bytes is not a multiple of 16", nq_home * qrow);
    const unsigned long long seq = ++c.seq;
    const int slot = (int)(seq % COMM_SLOTS), W = c.world, nq = nq_home * W;
    PeerPtrs peers;
    for (int r = 0; r < COMM_MAX_WORLD; ++r) peers.p[r] = r < W ? c.peer[r] : nullptr;
    unsigned char* win = c.window.as<unsigned char>();
    // 1. queries: my slice into everybody's query mailbox
    const size_t slice = (size_t)nq_home * qrow;
    const uint4* src = (const uint4*)Qhome;
    if (!on_device) {
        CU(h->w_q.reserve(slice));
        CU(cudaMemcpyAsync(h->w_q.p, Qhome, slice, cudaMemcpyHostToDevice, h->stream));
        src = h->w_q.as<uint4>();
    }
    k_comm_put<<<std::min<unsigned>(h->num_sms, (unsigned)(slice / 16 / 256 + 1)), 256, 0, h->stream>>>(
        peers, W, comm_slot_q_off(c, slot) + (size_t)c.rank * slice, src, slice);
    LAUNCHED();
    k_comm_signal<<<1, 32, 0, h->stream>>>(peers, W, c.rank, comm_flag_off(c, slot, 0), seq, nullptr);
    LAUNCHED();
    k_comm_wait<<<1, 32, 0, h->stream>>>(win, W, comm_flag_off(c, slot, 0), seq, c.d_err, nullptr);
    LAUNCHED();
    // 2. rank the whole batch on the local cells; records go straight to the home ranks' mailboxes
    RecRoute route = {};
    route.nq_home = nq_home;
    const size_t rb = rec_bytes(nq_home, k, mv.M);
    for (int r = 0; r < W; ++r) route.base[r] = c.peer[r] + comm_slot_r_off(c, slot) + (size_t)c.rank * rb;
    int rc = search_local_impl(h, win + comm_slot_q_off(c, slot), q_is_f64, nq, 1, quota, k, 0, win /*unused*/, false, &route);
    if (rc) return rc;
    k_comm_signal<<<1, 32, 0, h->stream>>>(peers, W, c.rank, comm_flag_off(c, slot, 1), seq, nullptr);
    LAUNCHED();
    k_comm_wait<<<1, 32, 0, h->stream>>>(win, W, comm_flag_off(c, slot, 1), seq, c.d_err, nullptr);
    LAUNCHED();
    // 3. merge my home queries
    const size_t nk = (size_t)nq_home * k;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    const size_t o1 = take(nk * 8), o2 = take(nk * 8), o3 = take(nk * 8), o4 = take(nk * mv.M), o5 = take((size_t)nq_home * 4),
                 o6 = take((size_t)nq_home * 4), o7 = take((size_t)nq_home);
    const size_t o8 = off; off += 256;
    unsigned char* b = (unsigned char*)block;
    if (!block_on_device) { CU(h->w_out.reserve(off)); b = h->w_out.as<unsigned char>(); }
    CU(cudaMemsetAsync(c.d_unc, 0, 4, h->stream));
    rc = merge_impl(h, win + comm_slot_r_off(c, slot), W, nq_home, k, 1, (int64_t*)(b + o1), (double*)(b + o2), (int32_t*)(b + o3), b + o4,
                    (int32_t*)(b + o5), (int32_t*)(b + o6), b + o7, false, c.d_unc);
    if (rc) return rc;
    k_comm_signal<<<1, 32, 0, h->stream>>>(peers, W, c.rank, comm_flag_off(c, slot, 2), seq, c.d_unc);
    LAUNCHED();
    k_comm_wait<<<1, 32, 0, h->stream>>>(win, W, comm_flag_off(c, slot, 2), seq, c.d_err, (int32_t*)(b + o8));
    LAUNCHED();
    if (!block_on_device) CU(cudaMemcpyAsync(block, b, off, cudaMemcpyDeviceToHost, h->stream));
    if (!h->async_mode) { CU(cudaStreamSynchronize(h->stream)); return finish_stats(h); }
    return B2L_OK;
}

int b2l_kmeans(b2l_handle h, const double* X, int64_t n, int d, int k, int iters, double* C, const int64_t* reseed,
               int32_t* assign, double* cost) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!X || !C || n < 1 || d < 1 || k < 1 || iters < 0 || (iters > 0 && !reseed)) FAIL(B2L_ERR_ARG, "bad kmeans arguments");
    if ((size_t)KM_WARPS * d * 8 > 200 * 1024) FAIL(B2L_ERR_UNSUPPORTED, "kmeans: dimension %d too large", d);
    DevBuf dX, dC, dS, dN, dA, dR, dCost;
    auto freeall = [&]() { dX.release(); dC.release(); dS.release(); dN.release(); dA.release(); dR.release(); dCost.release(); };
#define KM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { freeall(); char b_[256]; snprintf(b_, sizeof b_, "kmeans: %s", cudaGetErrorString(e_)); h->err = b_; return B2L_ERR_CUDA; } } while (0)
    KM(dX.reserve((size_t)n * d * 8)); KM(dC.reserve((size_t)k * d * 8)); KM(dS.reserve((size_t)k * d * 8)); KM(dN.reserve((size_t)k * 8));
    KM(dA.reserve((size_t)n * 4)); KM(dR.reserve((size_t)std::max(1, iters) * k * 8)); KM(dCost.reserve(8));
    KM(cudaMemcpyAsync(dX.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, h->stream));
    KM(cudaMemcpyAsync(dC.p, C, (size_t)k * d * 8, cudaMemcpyHostToDevice, h->stream));
    if (iters > 0) KM(cudaMemcpyAsync(dR.p, reseed, (size_t)iters * k * 8, cudaMemcpyHostToDevice, h->stream));
    const size_t smem = (size_t)KM_WARPS * d * 8;
    KM(cudaFuncSetAttribute(k_km_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned ablocks = (unsigned)((n + KM_WARPS - 1) / KM_WARPS);
    for (int it = 0; it <= iters; ++it) {                     // `iters` updates, then the final assignment and cost
        KM(cudaMemsetAsync(dCost.p, 0, 8, h->stream));
        k_km_assign<<<ablocks, KM_WARPS * 32, smem, h->stream>>>(dX.as<double>(), n, d, dC.as<double>(), k, dA.as<int32_t>(), dCost.as<double>());
        KM(cudaGetLastError());
        if (it == iters) break;
        KM(cudaMemsetAsync(dS.p, 0, (size_t)k * d * 8, h->stream));
        KM(cudaMemsetAsync(dN.p, 0, (size_t)k * 8, h->stream));
        k_km_accum<<<grid_for(n * d, 256), 256, 0, h->stream>>>(dX.as<double>(), n, d, dA.as<int32_t>(), dS.as<double>(), dN.as<unsigned long long>());
        KM(cudaGetLastError());
        k_km_update<<<grid_for((int64_t)k * d, 256), 256, 0, h->stream>>>(dX.as<double>(), d, k, dS.as<double>(), dN.as<unsigned long long>(),
                                                                         dR.as<int64_t>() + (size_t)it * k, dC.as<double>());
        KM(cudaGetLastError());
    }
    KM(cudaMemcpyAsync(C, dC.p, (size_t)k * d * 8, cudaMemcpyDeviceToHost, h->stream));
    if (assign) KM(cudaMemcpyAsync(assign, dA.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (cost) KM(cudaMemcpyAsync(cost, dCost.p, 8, cudaMemcpyDeviceToHost, h->stream));
    KM(cudaStreamSynchronize(h->stream));
#undef KM
    freeall();
    return B2L_OK;
}

/* 0 = no wait of the exchange has timed out; r + 1 = the wait for rank r did */
int b2l_comm_error(b2l_handle h) {
    if (!h) return B2L_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    if (!h->comm.d_err) return 0;
    int e = 0;
    CU(cudaMemcpy(&e, h->comm.d_err, 4, cudaMemcpyDeviceToHost));
    return e;
}

}  // extern "C"

#ifdef SCAN1_TRACE
// developer builds only (profiles/dev/scan1_trace_probe.py): the block phase stamps of the last k_scan1 launch
extern "C" int b2l_debug_scan1_trace(unsigned long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_scan1_trace, (size_t)n * 8);
}
#endif
