// Host half of the sparse-row transport (see hostpack.cu): packs rows of doubles into [bitmap over 4-double chunks | the chunks
// that hold a non-zero bit].  Plain C++ (no CUDA), thread-safe, one call per slab of rows; callers pack slabs in parallel.
// The AVX2 path is branch-free (store the chunk unconditionally, advance the cursor by the liveness bit) and is selected at run
// time; the portable path is the fallback for CPUs without AVX2.
#include <cstdint>
#include <cstring>
#include <immintrin.h>

#include "../../include/pbvi_b200.h"

namespace {

constexpr int PACK = 4;

inline int pack_tail(const uint64_t* row, int s, int row_len, uint64_t* q) {
    uint64_t v[PACK] = {0, 0, 0, 0};
    for (int j = 0; j < PACK && s + j < row_len; j++) v[j] = row[s + j];
    if (!(v[0] | v[1] | v[2] | v[3])) return 0;
    std::memcpy(q, v, sizeof(v));
    return 1;
}

int64_t pack_row_portable(const uint64_t* row, int row_len, int nC, int W, uint32_t* bm, uint64_t* dst) {
    int64_t cnt = 0;
    const int fullC = row_len / PACK;
    for (int w = 0; w < W; w++) {
        uint32_t bits = 0u;
        const int c1 = (w + 1) * 32 < nC ? (w + 1) * 32 : nC;
        for (int c = w * 32; c < c1; c++) {
            int live;
            if (c < fullC) {
                const uint64_t* p = row + (size_t)c * PACK;
                live = (p[0] | p[1] | p[2] | p[3]) != 0;
                if (live) std::memcpy(dst + cnt * PACK, p, PACK * sizeof(uint64_t));
            } else {
                live = pack_tail(row, c * PACK, row_len, dst + cnt * PACK);
            }
            bits |= (uint32_t)live << (c - w * 32);
            cnt += live;
        }
        bm[w] = bits;
    }
    return cnt;
}

__attribute__((target("avx2"))) int64_t pack_row_avx2(const uint64_t* row, int row_len, int nC, int W, uint32_t* bm, uint64_t* dst) {
    int64_t cnt = 0;
    const int fullC = row_len / PACK;
    for (int w = 0; w < W; w++) {
        uint32_t bits = 0u;
        const int c1 = (w + 1) * 32 < nC ? (w + 1) * 32 : nC;
        int c = w * 32;
        const int cv = c1 < fullC ? c1 : fullC;
        for (; c < cv; c++) {
            // one software prefetch per cache line, 2 KB ahead: on its hardware prefetcher alone (it stops at every 4 KB page) a core
            // scans 5.6 GB/s, with this 9-10 -- the packers of a 16-core host go from 90 to 156 GB/s (tools/e2e_timeline.py)
            if (!(c & 1)) _mm_prefetch(reinterpret_cast<const char*>(row + (size_t)c * PACK) + 2048, _MM_HINT_T0);
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row + (size_t)c * PACK));
            const int live = !_mm256_testz_si256(v, v);
            _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + cnt * PACK), v);     // overwritten by the next live chunk if dead
            bits |= (uint32_t)live << (c - w * 32);
            cnt += live;
        }
        for (; c < c1; c++) {
            const int live = pack_tail(row, c * PACK, row_len, dst + cnt * PACK);
            bits |= (uint32_t)live << (c - w * 32);
            cnt += live;
        }
        bm[w] = bits;
    }
    return cnt;
}

}  // namespace

// h_packed needs room for n * ceil(row_len / 4) + 1 chunks (the branch-free path stores one chunk past the live ones)
extern "C" int pbvi_pack_rows_host(const double* h_rows, int n, int row_len, uint32_t* h_bitmap, int32_t* h_row_start, double* h_packed,
                                   int64_t* h_chunks) {
    if (n < 0 || row_len <= 0 || !h_rows || !h_bitmap || !h_row_start || !h_packed || !h_chunks) return PBVI_ERR_BAD_ARG;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const int nC = (row_len + PACK - 1) / PACK, W = (nC + 31) / 32;
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t* row = reinterpret_cast<const uint64_t*>(h_rows) + (size_t)i * row_len;
        uint32_t* bm = h_bitmap + (size_t)i * W;
        uint64_t* dst = reinterpret_cast<uint64_t*>(h_packed) + (size_t)total * PACK;
        h_row_start[i] = (int32_t)total;
        total += avx2 ? pack_row_avx2(row, row_len, nC, W, bm, dst) : pack_row_portable(row, row_len, nC, W, bm, dst);
        if (total >= (int64_t(1) << 31)) return PBVI_ERR_BAD_ARG;      // int32 row offsets: pack smaller slabs
    }
    h_row_start[n] = (int32_t)total;
    *h_chunks = total;
    return PBVI_OK;
}

// One packer thread: slabs first_slab, first_slab + slab_step, ... of slab_rows rows each.  Slab i writes its bitmap rows, its
// row offsets h_row_start[i * (slab_rows + 1) ..], its chunks at h_packed + i * region_doubles and, LAST, its chunk count to
// h_totals[i] (initialised to -1 by the caller, who polls it): the caller never needs this thread to come back before it
// can ship a finished slab.
extern "C" int pbvi_pack_slabs_host(const double* h_rows, int n, int row_len, int slab_rows, int first_slab, int slab_step, uint32_t* h_bitmap,
                                    int32_t* h_row_start, double* h_packed, int64_t region_doubles, int64_t* h_totals) {
    if (n < 0 || row_len <= 0 || slab_rows <= 0 || first_slab < 0 || slab_step <= 0 || !h_totals) return PBVI_ERR_BAD_ARG;
    const int nC = (row_len + PACK - 1) / PACK, W = (nC + 31) / 32;
    const int n_slabs = (n + slab_rows - 1) / slab_rows;
    for (int i = first_slab; i < n_slabs; i += slab_step) {
        const int r0 = i * slab_rows, r1 = r0 + slab_rows < n ? r0 + slab_rows : n;
        int64_t total = 0;
        const int rc = pbvi_pack_rows_host(h_rows + (size_t)r0 * row_len, r1 - r0, row_len, h_bitmap + (size_t)r0 * W,
                                           h_row_start + (size_t)i * (slab_rows + 1), h_packed + (size_t)i * region_doubles, &total);
        if (rc != PBVI_OK) {
            __atomic_store_n(&h_totals[i], (int64_t)-2, __ATOMIC_RELEASE);
            return rc;
        }
        __atomic_store_n(&h_totals[i], total, __ATOMIC_RELEASE);
    }
    return PBVI_OK;
}
