// Multi-GPU exchange of the cell-sharded search, inside the library: peer-mapped "windows" (one cudaMalloc per rank,
// opened by the other ranks through CUDA IPC, or plain pointers when the ranks are handles of one process) and
// flag-based put / signal / wait over NVLink -- no NCCL call and no host round trip on the search path.
//
// Window of a rank:   flags [R slots][3 kinds][world] u64  |  per slot: query mailbox [world][nq_home x row bytes]
//                                                          |  per slot: record mailbox [world][rec_bytes(nq_home, k, M)]
// A batch (sequence number seq, slot seq % R):
//   1. every rank PUTS its home slice of the queries into the query mailbox of every rank (k_comm_put, peer stores),
//      then SIGNALS kind 0 (st.release.sys of seq into flags[slot][0][rank] of every rank) and WAITS for all ranks;
//   2. every rank ranks the whole batch against the cells it owns; k_select writes the k records of query q straight
//      into the record mailbox of q's HOME rank (block = the writer's rank) -- the all-to-all is fused into the
//      selection kernel; SIGNAL / WAIT kind 1;
//   3. every rank merges the world record blocks of its home queries (k_final), counts the queries it could not certify,
//      SIGNALS kind 2 with that count as payload and WAITS: every rank then knows whether anybody needs the fallback chain.
// A flag carries seq << 20 | payload and only ever grows, so slots need no reset.  A slot is reused after R batches; a
// rank can only be R - 1 batches ahead of a peer if that peer has signalled the batches in between, which it does after
// consuming the slot (stream order), so R = 4 is ample.  Waits are bounded (COMM_TIMEOUT_NS): on expiry the error word of
// the window is set and the host raises.
#pragma once
#include "common.cuh"

#define COMM_MAX_WORLD 8
#define COMM_SLOTS 4
#define COMM_FLAG_BYTES 4096           // R * 3 * world * 8 <= 768
#define COMM_TIMEOUT_NS 20000000000ull

struct PeerPtrs { unsigned char* p[COMM_MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// copy `bytes` (multiple of 16) from src to peers.p[r] + dst_off for every rank r < world (the own window included)
__global__ void k_comm_put(PeerPtrs peers, int world, size_t dst_off, const uint4* __restrict__ src, size_t bytes) {
    const size_t n = bytes / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int r = 0; r < world; ++r) ((uint4*)(peers.p[r] + dst_off))[i] = v;
    }
}

// flags[.][rank] of every rank <- seq << 20 | payload  (payload read from device memory when payload_ptr != NULL)
__global__ void k_comm_signal(PeerPtrs peers, int world, int rank, size_t flag_off, unsigned long long seq,
                              const unsigned int* __restrict__ payload_ptr) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const unsigned long long pay = payload_ptr ? (unsigned long long)min(*payload_ptr, 0xFFFFFu) : 0ull;
    __threadfence_system();
    st_release_sys((unsigned long long*)(peers.p[r] + flag_off) + rank, (seq << 20) | pay);
}

// wait until every rank has signalled `seq` in this window's flags; payloads -> out[world] (may be NULL)
__global__ void k_comm_wait(const unsigned char* __restrict__ window, int world, size_t flag_off, unsigned long long seq,
                            int* __restrict__ err, int32_t* __restrict__ out) {
    const int r = threadIdx.x;
    if (r >= world) return;                                       // (exited threads count as arrived at the barrier below)
    const unsigned long long* f = (const unsigned long long*)(window + flag_off) + r;
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long v = ld_acquire_sys(f);
    while ((v >> 20) < seq) {
        if (globaltimer_ns() - t0 > COMM_TIMEOUT_NS) { atomicExch(err, 1 + r); break; }
        __nanosleep(200);
        v = ld_acquire_sys(f);
    }
    if (out) out[r] = (int32_t)(v & 0xFFFFFull);
    __threadfence_system();
    __syncthreads();
    if (out && r == 0) out[COMM_MAX_WORLD] = *(volatile int*)err;      // 0, or 1 + the rank some wait gave up on
}
