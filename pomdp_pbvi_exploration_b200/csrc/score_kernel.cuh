// Score kernel of the point-based backup: block-sparse FP64 DMMA GEMM with a gathered B operand and a fused
// first-index argmax (replaces src/pomdp.py:1485-1495 of the reference without materialising Gamma).
//
// GATHER (reachable_state_count == 1, the olfactory / sea-robin class):
//     score[b,a,o,v] = sum_s  (beliefs[b,s] * RTO[a][o][s])  *  alphaT[ reach[a][s] ][v]
//   the A operand is the raw belief tile, scaled by the RTO column when the fragment is loaded; the B operand rows
//   are alphaT rows picked through the reachable-state table (always a coalesced BN-wide row, whatever the dynamics).
// PLAIN:
//     score[b,z,v]   = sum_s  beliefs[b,s] * bmat[z][s][v]
//   bmat is alphaT (z stride 0: max_v b.alpha_v of compute_change / SSGA / GER) or the transposed Gamma projection
//   GammaT[a,o] built by gamma_project_kernel for models with reachable_state_count > 1.
//
// One block owns a BM x BN tile of one z = (a,o) and walks only the K chunks (KC source states) in which some belief
// of the tile is non-zero AND some RTO entry of (a,o) is non-zero (list built by build_chunk_lists_kernel); inside a
// chunk a warp whose RG beliefs are all zero is skipped.  gamma > 0 scales every score equally and is left out
// (argmax invariant; exact-zero rows stay exactly zero, so "first index of the maximum" is preserved).
// Bound: the FP64 pipe (DMMA.8x8x4 runs at the FP64 peak on sm_100a, see profiles/r01_fp64_pipe_microbench.txt).
#pragma once
#include "pbvi_common.cuh"

namespace pbvi {

constexpr int STAGES = 5;
constexpr int LDA = KC + 4;   // 20: row stride 4 mod 16 doubles -> conflict-free 8-byte fragment loads
constexpr int LDB = BN + 4;   // 132

struct __align__(16) ScoreStage {
    double Bs[KC * LDB];
    double As[BM * LDA];
    double Rs[KC];
};
constexpr size_t SCORE_SMEM = sizeof(ScoreStage) * STAGES;
static_assert(sizeof(ScoreStage) % 16 == 0, "stage alignment");
static_assert(SCORE_SMEM <= 227 * 1024, "score pipeline exceeds shared memory");

struct ScoreParams {
    const double* beliefs;     // [nB][S]
    const double* bmat;        // GATHER: alphaT [S][Vp];  PLAIN: [gridDim.z][S][Vp], matrix of block z at blockIdx.z * zStrideB
    size_t zStrideB;
    const int32_t* reachP;     // [A][Sp]          (GATHER)
    const double* rtoP;        // [A*O][Sp]        (GATHER)
    const uint32_t* lists;     // [nMt][nZ][nChunks]  chunk | row-group bits << 24
    const int32_t* listCount;  // [nMt][nZ]
    const int32_t* zOrder;     // [nZ] heavy-first processing order (nullptr: identity)
    double* pval;              // [nNt][nB][nZ]
    int32_t* pidx;             // [nNt][nB][nZ]
    unsigned long long* stats; // visited (chunk, row group) pairs, summed over blocks
    int nB, S, Sp, V, Vp, nChunks, nZ, O;
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// (value, index) combine with NumPy argmax semantics: larger value wins, equal values keep the lower index
__device__ __forceinline__ void argmax_combine(double& v, int& i, double ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <bool GATHER>
__global__ void __launch_bounds__(SCORE_THREADS, 1) score_kernel(const ScoreParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScoreStage* stages = reinterpret_cast<ScoreStage*>(smem_raw);
    __shared__ uint32_t s_meta[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int warp_m = warp & 3, warp_n = warp >> 2;
    const int nt = blockIdx.x, mt = blockIdx.y;
    const int z = p.zOrder ? p.zOrder[blockIdx.z] : (int)blockIdx.z;
    const int a = GATHER ? z / p.O : 0;
    const int m0 = mt * BM, n0 = nt * BN;
    const uint32_t* __restrict__ list = p.lists + ((size_t)mt * p.nZ + z) * p.nChunks;
    const int nAct = p.listCount[mt * p.nZ + z];
    const double* __restrict__ bsrc = (GATHER ? p.bmat : p.bmat + (size_t)blockIdx.z * p.zStrideB) + n0;
    const int32_t* __restrict__ reach = GATHER ? p.reachP + (size_t)a * p.Sp : nullptr;
    const double* __restrict__ rto = GATHER ? p.rtoP + (size_t)z * p.Sp : nullptr;

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int n = 0; n < 8; n++) { acc[i][n][0] = 0.0; acc[i][n][1] = 0.0; }

    // ---- producer state, software-pipelined so that no thread waits on a dependent global load:
    //      e0/rows0 describe the chunk issued by the NEXT call of issue(); e1 the one after it.
    const int bkk = tid >> 6;            // B tile: this thread copies piece (tid & 63) of rows bkk + 4*i
    const int bpiece = (tid & 63) * 2;
    uint32_t e0 = 0, e1 = 0;
    int rows0[4] = {0, 0, 0, 0};
    int fetched = 0;                     // list position whose entry sits in e1
    auto load_rows = [&](uint32_t e, int (&rows)[4]) {
        const int k0 = (int)(e & 0xFFFFFFu) * KC;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = k0 + bkk + 4 * i;
            rows[i] = GATHER ? reach[k] : min(k, p.S - 1);
        }
    };
    if (nAct > 0) {
        e0 = list[0];
        e1 = list[min(1, nAct - 1)];
        fetched = 1;
        load_rows(e0, rows0);
    }

    auto issue = [&](int slot) {
        ScoreStage& st = stages[slot];
        const uint32_t e = e0;
        const int k0 = (int)(e & 0xFFFFFFu) * KC;
        const uint32_t rg = e >> 24;
        if (tid == 0) s_meta[slot] = e;
#pragma unroll
        for (int i = 0; i < BM * KC / SCORE_THREADS; i++) {
            const int idx = tid + i * SCORE_THREADS;
            const int m = idx >> 4, kk = idx & 15;
            if (!((rg >> (m >> 5)) & 1u)) continue;          // rows of an all-zero row group are never read
            const int k = k0 + kk, row = m0 + m;
            const bool valid = (row < p.nB) && (k < p.S);
            const double* src = valid ? p.beliefs + (size_t)row * p.S + k : p.beliefs;
            cp_async8(&st.As[m * LDA + kk], src, valid);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int kk = bkk + 4 * i;
            const bool valid = (k0 + kk) < p.S;
            cp_async16(&st.Bs[kk * LDB + bpiece], bsrc + (size_t)rows0[i] * p.Vp + bpiece, valid);
        }
        if (GATHER && tid < KC) cp_async8(&st.Rs[tid], rto + k0 + tid, true);
        // advance the prefetch registers (consumed by the next issue, one full chunk of math from now)
        e0 = e1;
        load_rows(e0, rows0);
        fetched = min(fetched + 1, nAct - 1);
        e1 = list[fetched];
    };

    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nAct) issue(s);
        cp_async_commit();
    }

    unsigned long long visited = 0;
    for (int it = 0; it < nAct; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (it + STAGES - 1 < nAct) issue((it + STAGES - 1) % STAGES);
        cp_async_commit();

        const ScoreStage& st = stages[it % STAGES];
        const uint32_t rg = s_meta[it % STAGES] >> 24;
        if (tid == 0) visited += __popc(rg);
        if ((rg >> warp_m) & 1u) {
            const double* Ab = st.As + (warp_m * RG + g) * LDA + t;
            const double* Bb = st.Bs + t * LDB + warp_n * 64 + g;
#pragma unroll
            for (int ks = 0; ks < KC / 4; ks++) {
                double af[4], bf[8];
#pragma unroll
                for (int i = 0; i < 4; i++) af[i] = Ab[i * 8 * LDA + ks * 4];
                if (GATHER) {
                    const double r = st.Rs[ks * 4 + t];
#pragma unroll
                    for (int i = 0; i < 4; i++) af[i] *= r;
                }
#pragma unroll
                for (int n = 0; n < 8; n++) bf[n] = Bb[ks * 4 * LDB + n * 8];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int n = 0; n < 8; n++) dmma884(acc[i][n], af[i], bf[n]);
            }
        }
    }
    cp_async_wait<0>();
    if (tid == 0 && p.stats && visited) atomicAdd(p.stats, visited);

    // ---- fused argmax: ascending columns per thread, then the quad (disjoint columns of the same rows), then the
    //      two column-warps through shared memory.  An empty list leaves acc == 0: every score is 0, first column wins.
    double best[4];
    int bidx[4];
    const int cbase = n0 + warp_n * 64 + 2 * t;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        best[i] = -INFINITY;
        bidx[i] = 0x7fffffff;
#pragma unroll
        for (int n = 0; n < 8; n++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int col = cbase + n * 8 + j;
                const double v = acc[i][n][j];
                if (col < p.V && v > best[i]) { best[i] = v; bidx[i] = col; }
            }
#pragma unroll
        for (int off = 1; off <= 2; off <<= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best[i], off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], off);
            argmax_combine(best[i], bidx[i], ov, oi);
        }
    }
    __syncthreads();   // all warps are done with the stage buffers
    double* sval = reinterpret_cast<double*>(smem_raw);                         // [2][BM]
    int* sidx = reinterpret_cast<int*>(smem_raw + sizeof(double) * 2 * BM);     // [2][BM]
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = warp_m * RG + i * 8 + g;
            sval[warp_n * BM + row] = best[i];
            sidx[warp_n * BM + row] = bidx[i];
        }
    }
    __syncthreads();
    if (tid < BM && m0 + tid < p.nB) {
        double v = sval[tid];
        int i = sidx[tid];
        argmax_combine(v, i, sval[BM + tid], sidx[BM + tid]);
        const size_t out = ((size_t)nt * p.nB + (m0 + tid)) * p.nZ + z;
        p.pval[out] = v;
        p.pidx[out] = i;
    }
}

}  // namespace pbvi
