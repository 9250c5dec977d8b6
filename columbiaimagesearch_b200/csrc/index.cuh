// Inverted-index layout kernels (LOPQSearcher.add_codes, search.py:325-369 -- layout only):
// rows appended in insertion order are regrouped cell-major (stable, so the in-cell order stays the
// insertion order the reference's per-cell lists have), code rows padded to MP bytes and byte-swizzled
// by the in-cell row index (stored byte s = code byte (i & SW) ^ s, consumed by scan.cuh), cell starts
// aligned to 16 rows so every row load of the scan is naturally aligned.
#pragma once
#include "common.cuh"

__global__ void k_cell_ids(const int32_t* __restrict__ coarse, int64_t n, int V, unsigned int* __restrict__ cell,
                           unsigned int* __restrict__ order, unsigned long long* __restrict__ hist, int* __restrict__ bad) {
    // (warp-aggregated histogram: the lanes of a warp that fall into the same cell issue one atomic)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) / 32) * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < n;
        unsigned int c = 0xFFFFFFFFu;
        if (live) {
            const int c0 = coarse[2 * i], c1 = coarse[2 * i + 1];
            c = 0;
            if (c0 < 0 || c0 >= V || c1 < 0 || c1 >= V) atomicExch(bad, 1);
            else c = (unsigned)(c0 * V + c1);
            cell[i] = c;
            order[i] = (unsigned int)i;
        }
        if (!hist) continue;                 // large V: no dense per-cell histogram (sparse directory, largev.cuh)
        const unsigned int peers = __match_any_sync(0xffffffffu, c);
        if (live && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[c], (unsigned long long)__popc(peers));
    }
}

// sorted_cell/sorted_src: stable sort of (cell, insertion index); sorted_first[c] = first sorted slot of cell c
__global__ void k_scatter_rows(const unsigned int* __restrict__ sorted_cell, const unsigned int* __restrict__ sorted_src, int64_t n,
                               const int64_t* __restrict__ sorted_first, const int64_t* __restrict__ cell_start,
                               const uint8_t* __restrict__ fine_in, const int64_t* __restrict__ rowid_in, int M, int MP, int SW,
                               uint8_t* __restrict__ codes, int64_t* __restrict__ rowids) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned int c = sorted_cell[i];
        const int64_t src = sorted_src[i];
        const int64_t incell = i - sorted_first[c];
        const int64_t dst = cell_start[c] + incell;
        const int sw = (int)incell & SW;
        for (int s = 0; s < MP; ++s) {
            const int j = s ^ sw;
            codes[dst * MP + s] = (j < M) ? fine_in[src * M + j] : (uint8_t)0;
        }
        rowids[dst] = rowid_in[src];
    }
}

// coarse codes of a batch about to be appended: *bad = 1 if any is outside [0, V)
__global__ void k_check_coarse(const int32_t* __restrict__ coarse, int64_t n, int V, int* __restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = coarse[i];
        if (c < 0 || c >= V) atomicExch(bad, 1);
    }
}

__global__ void k_iota64(int64_t* p, int64_t base, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = base + i;
}

// plain [n][M] code rows of one cell out of the swizzled layout (LOPQSearcher.get_cell, search.py:372-382)
__global__ void k_unswizzle_rows(const uint8_t* __restrict__ codes, int64_t n, int M, int MP, int SW, uint8_t* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * M; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / M;
        const int j = (int)(e - i * M);
        out[e] = codes[i * MP + (((int)i & SW) ^ j)];
    }
}
