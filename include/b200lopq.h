/*
 * b200lopq.h -- C-ABI of the B200-native LOPQ hot path (libb200lopq.so).
 *
 * The reference (ColumbiaDVMM/ColumbiaImageSearch) has no FFI on this path: everything below
 * replaces pure-NumPy/Python code in lopq/lopq/{model,search,utils}.py.  Each entry point cites
 * the reference function whose arithmetic it takes over (paths relative to
 * /root/reference/lopq/lopq/).  The Python host layer (columbiaimagesearch_b200/lopq/) keeps the
 * reference class/method names and calls these through ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every call returns 0 on success, a negative code on failure; b2l_last_error() gives the text.
 *   - all buffers are caller-owned, dense, row-major.  Pointers are HOST pointers unless the
 *     argument is named d_* or the call has an `on_device` flag set to 1 (then they are device
 *     pointers of the handle's device and the call is asynchronous on the handle's stream up to
 *     its own final synchronisation).
 *   - one in-flight call per handle (internal mutex); handles are independent; a handle owns one
 *     CUDA stream; the CUDA context is created lazily by b2l_create (call it after fork()).
 *   - no CPU fallback exists: without a CUDA device b2l_create fails.
 */
#ifndef B200LOPQ_H
#define B200LOPQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2l_ctx* b2l_handle;

/* error codes */
#define B2L_OK            0
#define B2L_ERR_CUDA     -1
#define B2L_ERR_ARG      -2
#define B2L_ERR_STATE    -3
#define B2L_ERR_UNSUPPORTED -4

/* ---- lifecycle ---------------------------------------------------------------------------- */
int         b2l_create(int device, b2l_handle* out);
int         b2l_destroy(b2l_handle h);
/* A second handle on the same device sharing the parent's model and (finalised) index, with its own stream and workspaces:
 * batches issued alternately on a handle and its siblings overlap on the GPU.  While siblings exist the family's model and
 * index are frozen (mutating calls return B2L_ERR_STATE).  Destroy siblings before the parent. */
int         b2l_create_sibling(b2l_handle parent, b2l_handle* out);
/* h may be NULL: returns the text of the last failure of b2l_create. */
const char* b2l_last_error(b2l_handle h);
/* library ABI version (bumped on any signature change) */
int         b2l_version(void);

/* ---- model parameters ------------------------------------------------------------------------
 * LOPQModel attributes Cs, Rs, mus, subquantizers (model.py:463-473), all passed as float64
 * (float32 parameters widen exactly).  `coarse_is_f32` says the coarse centroids were float32 in
 * the model object: NumPy then evaluates coarse distances, cell ordering and the residual
 * x - C[c] in float32 for float32 queries, and the device code reproduces that.
 *   Cs   [2][V][h]        h = D/2
 *   Rs   [2][V][h][h]     R[s][c] applied as R . v (model.py:635-638, no transpose)
 *   mus  [2][V][h]
 *   subs [M][K][ds]       ds = D/M; subquantizers[0] then subquantizers[1]; K <= 256
 */
int b2l_set_model(b2l_handle h, int D, int V, int M, int K, int coarse_is_f32,
                  const double* Cs, const double* Rs, const double* mus, const double* subs);
/* LOPQModelPCA.apply_PCA (model.py:961-978): y = (x - mu) . P, optional L2 renorm, cast float32.
 *   P [D0][D], mu [D0].  Call after b2l_set_model (D must match). */
int b2l_set_pca(b2l_handle h, int D0, const double* P, const double* mu, int renorm);

/* ---- batch encode ----------------------------------------------------------------------------
 * LOPQModel.predict over rows (model.py:543-602, 980-1003; utils.py:203-218 compute_codes_*).
 *   X       [n][D0 or D]  float32 (x_is_f64 = 0) or float64 (1)
 *   coarse  [n][2] int32, fine [n][M] uint8
 * on_device = 1: X, coarse, fine are device pointers.  fine = NULL: coarse assignment only
 * (predict_coarse, model.py:563-573; utils.predict_cluster, utils.py:33-53). */
int b2l_encode(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device,
               int32_t* coarse, uint8_t* fine);
/* Fine-argmin arithmetic of b2l_encode (predict_fine, model.py:575-602; the codes are the same in every mode).
 * 0 (default): sub-vector lengths 8 and 16 are scored on the tensor cores (tcgen05 kind::tf32, three TF32 pieces per
 * float32 product, accumulators in tensor memory) and a centroid is accepted only when it is the single score below
 * min + 3E, E a bound on the evaluation error; other shapes as mode 2.  1: float64 only.  2: float32 on the SIMT pipe with
 * the same kind of guard.  What a guard cannot decide is redone in float64 in NumPy's summation order.
 * b2l_encode_guard_count: sub-vectors redone in float64 so far. */
int     b2l_set_fine_mode(b2l_handle h, int mode);
int64_t b2l_encode_guard_count(b2l_handle h, int reset);
/* Diagnostic of the tensor-core stage: encodes the n host rows X (128 <= n <= 2^20) as b2l_encode does and returns the
 * float32 scores |c_k|^2 / 2 - p.c_k [128][256] the tensor cores produced for rows 0..127 against sub-quantizer j, and
 * (px != NULL) the float64 projections [128][D] of those rows (project, model.py:604-641).  The tests bound
 * |score - exact| by the E of the guard. */
int     b2l_debug_fine_scores(b2l_handle h, const void* X, int x_is_f64, int64_t n, int j, float* scores, double* px);
/* apply_PCA alone (model.py:961-978): Y [n][D] float32. */
int b2l_apply_pca(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, float* Y);
/* the same before the final cast (apply_PCA(x, dtype=numpy.float64), model.py:961): Y [n][D] float64. */
int b2l_apply_pca64(b2l_handle h, const void* X, int x_is_f64, int64_t n, int on_device, double* Y);
/* LOPQModel.project (model.py:604-641) and get_subquantizer_distances (model.py:673-704) for
 * explicit (vector, coarse pair) inputs; float64 out.  px [n][D], lut [n][M][K] (either may be
 * NULL).  x is the (post-PCA) D-dim vector. */
int b2l_project_lut(b2l_handle h, const void* X, int x_is_f64, int64_t n, const int32_t* coarse,
                    double* px, double* lut);

/* ---- inverted index ---------------------------------------------------------------------------
 * LOPQSearcher.add_codes (search.py:325-369), layout only: rows are appended in call order and
 * kept in insertion order inside each cell (ties in search are broken by retrieval order).
 * Per-cell id de-duplication stays on the host.  rowids may be NULL (then insertion index). */
int     b2l_index_add(b2l_handle h, const int32_t* coarse, const uint8_t* fine, int64_t n,
                      const int64_t* rowids, int on_device);
int     b2l_index_clear(b2l_handle h);
int64_t b2l_index_size(b2l_handle h);
/* local per-cell sizes [V*V] (int64) of this handle's shard */
int     b2l_index_cell_sizes(b2l_handle h, int64_t* sizes);
/* multi-GPU: the index is sharded by cell; every rank must know the GLOBAL cell sizes because the
 * quota cut (search.py:128-133) counts all retrieved items.  Default = local sizes. */
int     b2l_index_set_global_cell_sizes(b2l_handle h, const int64_t* sizes);
/* read back one cell in insertion order (LOPQSearcher.get_cell, search.py:372-382).
 * Returns the number of rows in the cell; fills at most `cap` rows. */
int64_t b2l_index_get_cell(b2l_handle h, int c0, int c1, int64_t cap, int64_t* rowids, uint8_t* fine);

/* ---- cell order -------------------------------------------------------------------------------
 * search.multisequence (search.py:13-82) cut by the quota rule of get_result_quota (search.py:110-135)
 * against the handle's global cell sizes; quota = INT64_MAX gives the full V*V order.
 *   Q [nq][D] (post-PCA vectors), cells [nq][V*V] int32 (c0*V + c1), dists [nq][V*V] float64,
 *   nvis [nq] cells visited.  Host pointers; cells / dists may be NULL. */
int b2l_cell_order(b2l_handle h, const void* Q, int q_is_f64, int nq, int64_t quota,
                   int32_t* cells, double* dists, int32_t* nvis);

/* The same for V > 64 (the product's V = 2048 / 4096 models), where the V*V order is never materialised: the first
 * min(visited, max_cells) cells of the traversal of every query, stopping at the quota cut or after max_cells cells.
 *   cells [nq][max_cells], dists [nq][max_cells] (either may be NULL), nvis [nq]. */
int b2l_cell_order_prefix(b2l_handle h, const void* Q, int q_is_f64, int nq, int64_t quota, int max_cells,
                          int32_t* cells, double* dists, int32_t* nvis);

/* ---- search ------------------------------------------------------------------------------------
 * LOPQSearcherBase.search (search.py:179-224) for a batch of queries:
 *   multisequence cell order (search.py:13-82) -> whole cells until >= quota (110-135) ->
 *   ADC distances over all retrieved codes (137-177) -> stable ascending order -> first k.
 * Outputs [nq][k] (rows beyond count[q] are padding):
 *   rowid int64, dist float64, coarse [nq][k][2] int32, fine [nq][k][M] uint8,
 *   count [nq] int32 (= min(k, retrieved)), visited [nq] int32 (cells visited, incl. empty).
 * Any output pointer except count may be NULL.  Queries: [nq][D0 or D] float32/float64, host.
 * on_device = 1: Q and all outputs are device pointers.
 * V <= 64: dense V*V plan, quantised ADC scan + certified float64 re-rank.  64 < V <= 4096: sparse cell directory, the
 * multi-sequence traversal on the device, every retrieved code ranked in float64 (csrc/largev.cuh). */
int b2l_search(b2l_handle h, const void* Q, int q_is_f64, int nq, int on_device,
               int64_t quota, int k,
               int64_t* rowid, double* dist, int32_t* coarse, uint8_t* fine,
               int32_t* count, int32_t* visited);

/* Two-phase form for an index sharded by cell across ranks (one handle per rank):
 *   1. every rank: b2l_search_local -> this rank's best k candidates per query (exact float64
 *      distances, retrieval positions) in a device record buffer of b2l_records_bytes(nq,k) bytes;
 *   2. the host layer all-gathers the record buffers (NCCL), rank-major;
 *   3. every rank: b2l_search_merge over the gathered [nranks] buffers -> final outputs and a
 *      per-query `certified` flag.  certified[q] = 0 means the float32 scan could not prove the
 *      float64 order of the first k results (ties at the k-th place); rerun those queries with
 *      exact = 1 (b2l_search does this internally).
 * exact = 1 ranks every retrieved code in float64 (slow, any k / any M); exact = 2 forces the float32-table
 * fast scan (the middle stage of the chain  16-bit tables -> float32 tables -> float64 full sort). */
int64_t b2l_records_bytes(b2l_handle h, int nq, int k);
int b2l_search_local(b2l_handle h, const void* Q, int q_is_f64, int nq, int on_device,
                     int64_t quota, int k, int exact, void* d_records);
int b2l_search_merge(b2l_handle h, const void* d_records_all, int nranks, int nq, int k, int on_device,
                     int64_t* rowid, double* dist, int32_t* coarse, uint8_t* fine,
                     int32_t* count, int32_t* visited, uint8_t* certified);

/* b2l_search_merge with all outputs in ONE block (one device-to-host copy instead of seven): fields in the order
 * rowid [nq][k] i64 | dist [nq][k] f64 | coarse [nq][k][2] i32 | fine [nq][k][M] u8 | count [nq] i32 | visited [nq] i32 |
 * certified [nq] u8, each starting on a 256-byte boundary; b2l_merge_block_bytes gives the total.  `block` is a host
 * (pinned, in asynchronous mode) or device buffer. */
int64_t b2l_merge_block_bytes(b2l_handle h, int nq, int k);
int b2l_search_merge_block(b2l_handle h, const void* d_records_all, int nranks, int nq, int k, void* block, int on_device);

/* ---- multi-GPU exchange inside the library ------------------------------------------------------------------------
 * The reference has no multi-GPU path; SURVEY.md 8(e) defines it: inverted lists sharded by coarse cell, queries
 * replicated, one exchange of per-rank top-k records, merge.  These calls run that exchange without NCCL and without the
 * host: every rank owns a device "window" (query and record mailboxes + flags) that the other ranks map through CUDA IPC
 * (one process per GPU) or plain pointers (handles of one process: same_process = 1); queries are put into every rank's
 * mailbox, the selection kernel writes the k records of a query directly into the mailbox of the query's HOME rank over
 * NVLink, and release/acquire flags at system scope order it (csrc/comm.cuh).
 *   b2l_comm_init        reserve the window for home slices of <= max_nq_home queries, top-<= max_k, queries float32/64
 *   b2l_comm_get_handle  the 64-byte IPC handle of the window (b2l_comm_handle_bytes) and/or its local address
 *   b2l_comm_connect     handles = [world] IPC handles (64 bytes each), or [world] window addresses when same_process
 *   b2l_search_sharded   one batch: Qhome [nq_home][D0 or D] is THIS rank's slice of the global batch (rank-major, the same
 *                        nq_home on every rank); block receives b2l_sharded_block_bytes(nq_home, k) bytes: the merge block
 *                        of the home queries (layout of b2l_search_merge_block) followed, 256-byte aligned, by [world] int32 =
 *                        queries each rank could not certify (any non-zero: rerun those through the fallback chain).
 *                        Honours b2l_set_async.  Every rank must call it the same number of times.
 *   b2l_comm_error       0, or 1 + the rank a bounded device-side wait gave up on */
int     b2l_comm_init(b2l_handle h, int world, int rank, int max_nq_home, int max_k, int q_is_f64);
int     b2l_comm_handle_bytes(void);
int     b2l_comm_get_handle(b2l_handle h, void* ipc_handle_out, void** local_ptr_out);
int     b2l_comm_connect(b2l_handle h, const void* handles, int same_process);
int64_t b2l_sharded_block_bytes(b2l_handle h, int nq_home, int k);
int     b2l_search_sharded(b2l_handle h, const void* Qhome, int q_is_f64, int nq_home, int on_device, int64_t quota, int k,
                           void* block, int block_on_device);
int     b2l_comm_error(b2l_handle h);

/* ---- training support ------------------------------------------------------------------------------------------------
 * The k-means fits of training (model.py:290-336: coarse quantizers and sub-quantizers; the reference uses sklearn's
 * MiniBatchKMeans) as Lloyd iterations on the device.  X [n][d] float64; C [k][d] float64: initial centroids in, trained
 * centroids out; assignment = utils.predict_cluster over rows (utils.py:33-53: direct-form squared L2 in NumPy order, first
 * minimum); an empty cluster takes row reseed[it][c] of X (reseed: int64 [iters][k]).  `iters` updates are followed by a
 * final assignment: assign [n] int32 and the summed squared error *cost (either may be NULL).  Host pointers.  Needs no
 * model on the handle.  Training has no parity contract (models are inputs of the hot path). */
int b2l_kmeans(b2l_handle h, const double* X, int64_t n, int d, int k, int iters, double* C, const int64_t* reseed,
               int32_t* assign, double* cost);

/* ---- introspection (counters of the most recent b2l_search / b2l_search_local) ------------------ */
typedef struct b2l_stats {
    double  scan_ms;          /* device time of the ADC scan kernel (CUDA events)            */
    double  plan_ms;          /* cell order + work list + LUT build                          */
    double  select_ms;        /* partial top-k merge + float64 re-rank                       */
    double  total_ms;         /* whole call on the device stream                             */
    int64_t codes_scanned;    /* sum over queries of retrieved codes ranked on this rank     */
    int64_t scan_bytes;       /* algorithmic bytes of the scan = M * codes_scanned           */
    int64_t work_items;       /* (cell segment, query group) items processed                 */
    int64_t lut_slots;        /* (query, split, coarse code) distance tables built           */
    int64_t kernel_launches;  /* kernels of this library launched by the call                */
    int64_t exact_queries;    /* queries that went through the float64 full-sort path        */
    /* sums of the fields above over every search collected since b2l_reset_stats (asynchronous searches are timed
     * individually, each with its own CUDA events) */
    int64_t acc_calls;
    double  acc_scan_ms, acc_plan_ms, acc_select_ms, acc_total_ms;
    int64_t acc_codes_scanned, acc_scan_bytes, acc_work_items, acc_kernel_launches, acc_exact_queries;
    int64_t packed;           /* 1: the scan used the 16-bit packed tables                  */
    int64_t rescan_queries;   /* queries re-run with float32 tables (not certifiable from the 16-bit ones) */
    int64_t acc_rescan_queries;
} b2l_stats;
/* waits for searches still in flight on the handle, then returns the statistics */
int b2l_get_stats(b2l_handle h, b2l_stats* out);
int b2l_reset_stats(b2l_handle h);
/* Kernel choice of the fast scan.  0 (default): batches of <= 8 queries take the one-query-per-item scan with float32 tables
 * (scan1.cuh: no cross-query reuse, HBM-bound); larger batches take 16-bit quantised tables, two queries per shared-memory
 * word; queries whose float64 order cannot be proven are re-run with float32 tables, then exactly.  1: float32 tables only
 * (batched kernel).  2: as 0 without the low-batch kernel. */
int b2l_set_scan_mode(b2l_handle h, int mode);
/* Width of the preselection: at least kp_min (<= 512) candidates per query are re-ranked in float64 (default 0: k + 8 rounded
 * up to a power of two).  Results do not depend on it; a wider preselection certifies more queries at the first stage when
 * distances are concentrated (few go down the fallback chain) and costs a longer selection kernel. */
int b2l_set_preselect(b2l_handle h, int kp_min);
/* Asynchronous mode (pipelined / multi-GPU searches).  While enabled, b2l_search_local and b2l_search_merge only
 * enqueue their work (including the copies from / to host buffers, which must then be PINNED and stay alive) on the
 * handle's stream and return; the caller orders other work against that stream (b2l_stream) and waits for it
 * (an event recorded on the stream, or b2l_sync) before reading results.  Default: disabled. */
int b2l_set_async(b2l_handle h, int enabled);
int b2l_sync(b2l_handle h);
/* test knob for b2l_search: bit 0 sends every query of a batch to the second stage of the certification chain
 * (float32 tables) even if the first certified it, bit 1 sends every query of that stage on to the float64 full sort.
 * Bit 2 (any merge: b2l_search_merge*, b2l_search_sharded): every fifth query is reported uncertified, so the callers'
 * fallback chains run. */
int b2l_debug_force_redo(b2l_handle h, int mask);
/* diagnostics of the most recent fast-path search: per query, candidates the scan appended and the final
 * pruning bound (float32 bits).  Either pointer may be NULL. */
int b2l_debug_candidates(b2l_handle h, int nq, uint32_t* appended, uint32_t* bound_bits);
/* the stream the handle launches on (cudaStream_t as void*), for CUDA-event timing by the caller */
void* b2l_stream(b2l_handle h);

#ifdef __cplusplus
}
#endif
#endif /* B200LOPQ_H */
