// Host -> device transport of sparse belief rows.  A belief set that still lives in host memory reaches the GPU over PCIe
// (55 GB/s measured), which bounds the end-to-end backup (1.76 GB of beliefs per step on the bench workload against 3.5 ms of
// kernels).  Belief points of the olfactory model are 30 % dense and their zeros come in runs, so the host side packs every row
// into [bitmap over 4-double chunks | the non-zero chunks], only that crosses the bus, and `unpack_rows_kernel` rebuilds the
// dense rows in HBM -- byte for byte (a chunk is dropped only if all of its 32 bytes are zero bits, so -0.0 survives).
// (the host half, pbvi_pack_rows_host, is plain C++: hostpack_host.cpp)
#include <algorithm>

#include "pbvi_common.cuh"

namespace pbvi {

constexpr int PACK = 4;   // doubles per packed chunk

// warp per row; 128 states (one bitmap word) per iteration in four coalesced passes of 32 doubles.  Rows are packed in slabs of
// slabRows rows: row i belongs to slab i / slabRows, whose chunks start at slab * regionChunks in `packed` and whose row offsets
// are rowStart[slab * (slabRows + 1) + i % slabRows].
__global__ void __launch_bounds__(256) unpack_rows_kernel(const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ rowStart,
                                                          const double* __restrict__ packed, int n, int rowLen, int W, int slabRows,
                                                          long long regionChunks, double* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const int slab = row / slabRows;
    const uint32_t* bm = bitmap + (size_t)row * W;
    const double* src = packed + ((size_t)slab * regionChunks + rowStart[(size_t)slab * (slabRows + 1) + (row - slab * slabRows)]) * PACK;
    double* dst = out + (size_t)row * rowLen;
    int running = 0;
    for (int w = 0; w < W; w++) {
        const uint32_t bits = bm[w];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int s = w * 128 + p * 32 + lane;
            const int cl = p * 8 + (lane >> 2);
            double v = 0.0;
            if ((bits >> cl) & 1u) v = src[(size_t)(running + __popc(bits & ((1u << cl) - 1u))) * PACK + (lane & 3)];
            if (s < rowLen) dst[s] = v;
        }
        running += __popc(bits);
    }
}

int unpack_rows_launch(pbvi_model* m, const uint32_t* d_bitmap, const int32_t* d_row_start, const double* d_packed, int n, int row_len,
                       int slab_rows, long long region_chunks, double* d_out, cudaStream_t st) {
    if (n <= 0) return PBVI_OK;
    const int nC = (row_len + PACK - 1) / PACK, W = (nC + 31) / 32;
    unpack_rows_kernel<<<ceil_div(n, 8), 256, 0, st>>>(d_bitmap, d_row_start, d_packed, n, row_len, W, slab_rows, region_chunks, d_out);
    m->last_launches++;
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}

}  // namespace pbvi

using namespace pbvi;

extern "C" int pbvi_unpack_rows(pbvi_model* m, const uint32_t* d_bitmap, const int32_t* d_row_start, const double* d_packed, int n,
                                int row_len, int slab_rows, int64_t region_chunks, double* d_out, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0 && row_len > 0 && slab_rows > 0 && region_chunks >= 0, "need n >= 0 rows of positive length in slabs of slab_rows > 0 rows");
    m->last_launches = 0;
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_bitmap && d_row_start && d_packed && d_out, "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    const int nC = (row_len + PACK - 1) / PACK, W = (nC + 31) / 32;
    unpack_rows_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(d_bitmap, d_row_start, d_packed, n, row_len, W, slab_rows,
                                                                        (long long)region_chunks, d_out);
    m->last_launches = 1;
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}
