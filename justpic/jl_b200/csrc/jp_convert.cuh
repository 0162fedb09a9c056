// jp_convert.cuh -- Array(CellArray) / CuArray(CellArray): the layout change between the device
// CellArray (blocklength 0, data[C, S, 1]: element (cell c, component s) at c + s*C) and the host
// CellArray (blocklength 1, data[1, S, C]: at s + c*S), i.e. the reference's
// permutedims(CA.data, (3, 2, 1)) with the optional element conversion of Array(T, CA) / CuArray(T, CA)
// (src/CellArrays/conversion.jl:32-43, ext/JustPICCUDAExt.jl:166-179).  Used for checkpoints
// (test/test_save_load.jl:120-173).  One 32 x 32 tile per CTA through shared memory: reads are
// contiguous along the source's fast index, writes along the destination's.
#pragma once
#include <stdint.h>

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_cellarray_permute(const TS *__restrict__ src, TD *__restrict__ dst, int64_t C, int S, int to_host) {
    __shared__ TD tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
    if (to_host) {
        // read: lanes along cells (src fast index), write: lanes along components (dst fast index)
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int s = s0 + r; const int64_t c = c0 + tx;
            if (s < S && c < C) tile[r][tx] = static_cast<TD>(src[c + (int64_t)s * C]);
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t c = c0 + r; const int s = s0 + tx;
            if (s < S && c < C) dst[s + c * S] = tile[tx][r];
        }
    } else {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t c = c0 + r; const int s = s0 + tx;
            if (s < S && c < C) tile[r][tx] = static_cast<TD>(src[s + c * S]);
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int s = s0 + r; const int64_t c = c0 + tx;
            if (s < S && c < C) dst[c + (int64_t)s * C] = tile[tx][r];
        }
    }
}

template <typename TS, typename TD>
static void launch_cellarray_permute(const void *src, void *dst, int64_t C, int S, int to_host, cudaStream_t st) {
    const dim3 grd((unsigned)((C + 31) / 32), (unsigned)((S + 31) / 32), 1), blk(32, 8, 1);
    k_cellarray_permute<TS, TD><<<grd, blk, 0, st>>>((const TS *)src, (TD *)dst, C, S, to_host);
}
