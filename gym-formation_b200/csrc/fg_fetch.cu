// fg_fetch.cu -- device -> pinned-host transfer of one step's observations for the VecEnv adapter (sm_100a).
//
// The reference's trainers read obs[E,N,D] as a host array after every step (train/maddpg-v2/utils/env_wrappers.py:68-72).
// For formation_hd_env a row is [p_vel | p_j - p_i (N-1) | comm zeros (N-1) | ideal_shape (N) | ideal_vel]
// (formation_hd_env.py:52-59): only the first N of its 3N float2 items change from step to step; the rest changes when
// the env is reset.  The host array PERSISTS between steps, so a step only has to ship the dynamic prefix of every row
// -- and whole rows for the envs whose episode just ended -- to leave the host array byte-identical to the device
// tensor: a third of the PCIe bytes.
//
// mode 0: the whole tensor, one contiguous cudaMemcpyAsync (the round-1 path; also used when every env was reset).
// mode 1: cudaMemcpy2DAsync of the dynamic prefix of every row (width = dyn bytes, pitch = row bytes): copy engine,
//         no SM work; rows of envs that were reset are NOT refreshed (the caller falls back to mode 0 for such steps).
// mode 2: zero-copy scatter kernel: SM stores straight into the mapped pinned host array (UVA pointer), dynamic
//         prefix of every row, whole rows for envs with done[e*N] != 0 (done is read on the device: no host round trip).
// mode 3: pack kernel into a device staging buffer [rows, dyn] followed by one contiguous cudaMemcpyAsync into a pinned
//         host staging buffer (the host scatters; diagnostic only).
// mode 4: zero-copy kernel that writes WHOLE 64-byte host cache lines: for every row the lines its dynamic prefix touches
//         (a 72-byte prefix at a 216-byte pitch straddles two lines), 16 bytes per thread.  The bytes written beyond the
//         prefix are static row items whose device values equal what the host array already holds, so the result is the
//         same; the point is that the root complex receives full-line writes instead of partial ones (modes 1 and 2
//         stall at ~23 GB/s on the read-modify-write of partial lines).  Rows of envs with done[row] != 0 are copied
//         whole, rounded out to lines.  obs_dev and obs_host must be 64-byte aligned.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "../../include/formation_gym_b200.h"

namespace {

template <typename V>
__global__ void __launch_bounds__(256) k_rows_to_host(const V* __restrict__ src, V* __restrict__ dst,
                                                      const uint8_t* __restrict__ done, uint32_t rows, uint32_t n_agents,
                                                      uint32_t row_items, uint32_t dyn_items) {
    // one warp per row: lanes walk the row's items; a row whose env just ended is copied whole
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    for (uint32_t row = blockIdx.x * wpc + wib; row < rows; row += gridDim.x * wpc) {
        const uint32_t n = (done && done[row]) ? row_items : dyn_items;       // done[E,N]: one flag per row
        const V* s = src + (size_t)row * row_items;
        V* d = dst + (size_t)row * row_items;
        for (uint32_t k = lane; k < n; k += 32) d[k] = s[k];
    }
    (void)n_agents;
}

template <typename V>
__global__ void __launch_bounds__(256) k_rows_flat_to_host(const V* __restrict__ src, V* __restrict__ dst,
                                                           const uint8_t* __restrict__ done, uint32_t rows,
                                                           uint32_t row_items, uint32_t dyn_items, uint32_t magic_dyn) {
    // short rows (dyn_items < 32): consecutive threads <-> consecutive (row, item) of the dynamic prefixes, so a warp's
    // store covers several rows' prefixes; rows of ended episodes get their static tail from the same thread group
    const uint32_t total = rows * dyn_items;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const uint32_t row = magic_dyn ? __umulhi(q, magic_dyn) : q;
        const uint32_t k = q - row * dyn_items;
        const size_t base = (size_t)row * row_items;
        dst[base + k] = src[base + k];
        if (done && done[row])
            for (uint32_t j = dyn_items + k; j < row_items; j += dyn_items) dst[base + j] = src[base + j];
    }
}

// mode 4: thread q <-> 16-byte chunk c = q % cpr of row q / cpr, counted from the 64-byte line in which the row's
// prefix starts; cpr = chunks a row can need (the prefix rounded out to lines; whole rows when `done` is given).
__global__ void __launch_bounds__(256) k_rows_lines_to_host(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                            const uint8_t* __restrict__ done, uint32_t rows,
                                                            uint32_t row_bytes, uint32_t dyn_bytes, uint32_t cpr) {
    const uint64_t total = (uint64_t)rows * cpr, end = (uint64_t)rows * row_bytes;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(q / cpr);
        const uint32_t c = (uint32_t)(q - (uint64_t)row * cpr);
        const uint64_t b0 = (uint64_t)row * row_bytes;
        const uint32_t n = (done && done[row]) ? row_bytes : dyn_bytes;
        const uint64_t lo = b0 & ~(uint64_t)63, hi = min((b0 + n + 63) & ~(uint64_t)63, end);   // (never past the tensor)
        const uint64_t a = lo + (uint64_t)c * 16;
        if (a < hi) dst[a >> 4] = src[a >> 4];
    }
}

template <typename V>
__global__ void __launch_bounds__(256) k_rows_pack(const V* __restrict__ src, V* __restrict__ dst, uint32_t rows,
                                                   uint32_t row_items, uint32_t dyn_items, uint32_t magic_dyn) {
    const uint32_t total = rows * dyn_items;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const uint32_t row = magic_dyn ? __umulhi(q, magic_dyn) : q;
        const uint32_t k = q - row * dyn_items;
        dst[q] = src[(size_t)row * row_items + k];
    }
}

uint32_t magic_for(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((((uint64_t)1) << 32) / d) + 1u; }

template <typename V>
int fetch_impl(const void* obs_dev, void* obs_host, const uint8_t* done_dev, void* staging_dev, uint32_t rows,
               uint32_t n_agents, uint32_t row_items, uint32_t dyn_items, int mode, cudaStream_t st) {
    const size_t isz = sizeof(V);
    cudaError_t err = cudaSuccess;
    if (mode == 0) {
        err = cudaMemcpyAsync(obs_host, obs_dev, (size_t)rows * row_items * isz, cudaMemcpyDeviceToHost, st);
    } else if (mode == 1) {
        err = cudaMemcpy2DAsync(obs_host, row_items * isz, obs_dev, row_items * isz, dyn_items * isz, rows,
                                cudaMemcpyDeviceToHost, st);
    } else if (mode == 2) {
        if (dyn_items >= 32) {
            int grid = (int)((rows + 7) / 8);
            if (grid > 148 * 16) grid = 148 * 16;
            k_rows_to_host<V><<<grid, 256, 0, st>>>((const V*)obs_dev, (V*)obs_host, done_dev, rows, n_agents,
                                                    row_items, dyn_items);
        } else {
            const uint64_t total = (uint64_t)rows * dyn_items;
            if (total >= (1ull << 32)) return FG_ERR_ARG;
            int grid = (int)((total + 255) / 256);
            if (grid > 148 * 16) grid = 148 * 16;
            k_rows_flat_to_host<V><<<grid, 256, 0, st>>>((const V*)obs_dev, (V*)obs_host, done_dev, rows, row_items,
                                                         dyn_items, magic_for(dyn_items));
        }
        err = cudaGetLastError();
    } else if (mode == 3) {
        if (!staging_dev) return FG_ERR_ARG;
        const uint64_t total = (uint64_t)rows * dyn_items;
        if (total >= (1ull << 32)) return FG_ERR_ARG;
        int grid = (int)((total + 255) / 256);
        if (grid > 148 * 16) grid = 148 * 16;
        k_rows_pack<V><<<grid, 256, 0, st>>>((const V*)obs_dev, (V*)staging_dev, rows, row_items, dyn_items,
                                             magic_for(dyn_items));
        err = cudaGetLastError();
        if (err == cudaSuccess)
            err = cudaMemcpyAsync(obs_host, staging_dev, (size_t)total * isz, cudaMemcpyDeviceToHost, st);
    } else if (mode == 4) {
        if ((((uintptr_t)obs_dev) | ((uintptr_t)obs_host)) & 63) return FG_ERR_ARG;
        const uint32_t row_bytes = row_items * (uint32_t)isz, dyn_bytes = dyn_items * (uint32_t)isz;
        if (((uint64_t)rows * row_bytes) & 15) return FG_ERR_ARG;            // 16-byte chunks up to the end of the tensor
        const uint32_t span = done_dev ? row_bytes : dyn_bytes;
        const uint32_t cpr = ((span + 63 + 63) / 64) * 4;                   // worst-case lines touched x 4 chunks
        const uint64_t total = (uint64_t)rows * cpr;
        int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)148 * 16);
        k_rows_lines_to_host<<<grid, 256, 0, st>>>((const uint4*)obs_dev, (uint4*)obs_host, done_dev, rows, row_bytes,
                                                   dyn_bytes, cpr);
        err = cudaGetLastError();
    } else {
        return FG_ERR_ARG;
    }
    return err == cudaSuccess ? FG_OK : FG_ERR_CUDA;
}

}  // namespace

extern "C" int fg_obs_to_host(const void* obs_dev, void* obs_host, const uint8_t* done_dev, void* staging_dev, int E,
                              int N, int row_items, int dyn_items, int item_bytes, int mode, void* stream) {
    if (!obs_dev || !obs_host || E < 1 || N < 1 || row_items < 1 || dyn_items < 1 || dyn_items > row_items)
        return FG_ERR_ARG;
    if ((uint64_t)E * (uint64_t)N >= (1ull << 31)) return FG_ERR_ARG;
    const uint32_t rows = (uint32_t)E * (uint32_t)N;
    cudaStream_t st = (cudaStream_t)stream;
    if (item_bytes == 8)
        return fetch_impl<float2>(obs_dev, obs_host, done_dev, staging_dev, rows, (uint32_t)N, (uint32_t)row_items,
                                  (uint32_t)dyn_items, mode, st);
    if (item_bytes == 16)
        return fetch_impl<double2>(obs_dev, obs_host, done_dev, staging_dev, rows, (uint32_t)N, (uint32_t)row_items,
                                   (uint32_t)dyn_items, mode, st);
    return FG_ERR_ARG;
}
