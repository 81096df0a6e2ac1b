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
// One block owns a BM x BN = 64 x 256 tile of one z = (a,o) and walks only the K chunks (KC = 8 source states) in which
// some belief of the tile is non-zero AND some RTO entry of (a,o) is non-zero AND some alpha of the tile is non-zero at a
// state the chunk lands on (list built by build_chunk_lists_kernel): every skipped term is an exact zero.
//
// Warp specialisation (N_CONSUMER_WARPS + 1 warps):
//   last warp   producer.  Per chunk it arms the stage's `full` mbarrier with the byte count and issues the stage as
//               bulk async copies (cp.async.bulk, the TMA engine): one 2 KB copy per gathered alphaT row (KC of them), ONE
//               1 KB copy per live row group (belief_mask_kernel stores the belief tiles as ready-made, bank-swizzled
//               shared-memory images) and the RTO chunk.  List entries and gathered row indices are prefetched one chunk
//               ahead, so the producer never waits on a dependent global load.
//   the rest    consumers: wait on `full`, DMMA, arrive on `empty`.  No block-wide barrier inside the K loop.
//               Warp w owns row group w / 4 (RG beliefs) and column quarter w % 4, so each SM sub-partition
//               (warp id mod 4) holds one warp of EACH row group: a row group that is all-zero on the chunk is skipped
//               and the saving is spread over all four FP64 pipes (the first, 128 x 128 layout stalled on the barrier
//               instead: profiles/r01_score_kernel_v1_ncu_summary.txt).
// gamma > 0 scales every score equally and is left out (argmax invariant; exact-zero rows stay exactly zero, so "first
// index of the maximum" is preserved).
// Bound: the FP64 pipe (DMMA.8x8x4 runs at the FP64 peak on sm_100a, see profiles/r01_fp64_pipe_microbench.txt).
#pragma once
#include "pbvi_common.cuh"

namespace pbvi {

constexpr int LDB = BN + 4;   // 260
constexpr int NCW = BN / 64;  // 4 column warps, 64 columns each
constexpr int MT = RG / 8;    // m8 row tiles per consumer warp
// One row group of one chunk in beliefsP == its shared-memory image: RG rows of KC doubles, dense, with the columns of row r
// XOR-swizzled by a_swizzle(r) so that the 8-byte fragment loads of a half-warp (4 rows x 4 columns) hit 16 different banks
// whatever KC is (row r starts at 8-byte slot r*KC mod 16).
constexpr int A_GROUP_DOUBLES = RG * KC;
__host__ __device__ constexpr int a_swizzle(int row) { return KC == 16 ? ((row & 3) << 2) : KC == 8 ? (((row >> 1) & 1) << 2) : 0; }
static_assert(KC == 4 || KC == 8 || KC == 16, "a_swizzle covers KC = 4, 8, 16");
constexpr int N_CONSUMER_WARPS = NRG * NCW;                 // 16
constexpr int SCORE_THREADS_TOTAL = (N_CONSUMER_WARPS + 1) * 32;   // + the producer warp

struct __align__(128) ScoreStage {
    double Bs[KC * LDB];
    double As[BM * KC];
    double Rs[KC < 16 ? 16 : KC];
};
constexpr int STAGES = (int)((212 * 1024) / sizeof(ScoreStage)) < 16 ? (int)((212 * 1024) / sizeof(ScoreStage)) : 16;
constexpr size_t SCORE_SMEM = sizeof(ScoreStage) * STAGES;
static_assert(sizeof(ScoreStage) % 128 == 0, "stage alignment");
static_assert(SCORE_SMEM <= 226 * 1024 && STAGES >= 3, "score pipeline exceeds shared memory");
static_assert(BM == NRG * RG && N_CONSUMER_WARPS * 32 == SCORE_THREADS, "warp layout: NRG row groups x NCW column warps");
static_assert(2 * sizeof(double) * NCW * BM <= sizeof(ScoreStage), "argmax staging reuses the first stage");
static_assert(KC + NRG + 1 <= 32, "producer lane mapping: lanes [0,KC) B rows, [KC,KC+NRG) A row groups, KC+NRG the RTO chunk");
static_assert((A_GROUP_DOUBLES * 8) % 16 == 0 && (KC * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");

struct ScoreParams {
    const double* beliefsP;    // [nMt][nChunks][NRG][RG][KC]  belief tiles as swizzled shared-memory images (belief_mask_kernel)
    const double* bmat;        // GATHER: alphaT [S][Vp];  PLAIN: [gridDim.z][S][Vp], matrix of block z at blockIdx.z * zStrideB
    size_t zStrideB;
    const int32_t* reachP;     // [A][Sp]          (GATHER)
    const double* rtoP;        // [A*O][Sp]        (GATHER)
    const uint32_t* lists;     // [nMt][nZ][nNt][nChunks]  chunk | row-group bits << 24 | column-quarter bits << 28
    const int32_t* listCount;  // [nMt][nZ][nNt]
    const int32_t* zOrder;     // [nZ] heavy-first processing order (nullptr: identity)
    double* pval;              // [nNt][nB][nZ]
    int32_t* pidx;             // [nNt][nB][nZ]
    unsigned long long* stats; // visited (chunk, row group, column quarter) triples, summed over blocks
    int nB, S, Sp, V, Vp, nChunks, nZ, O;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

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
__global__ void __launch_bounds__(SCORE_THREADS_TOTAL, 1) score_kernel(const ScoreParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreStage* stages = reinterpret_cast<ScoreStage*>(smem_raw);
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ __align__(8) uint64_t s_empty[STAGES];
    __shared__ uint32_t s_meta[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = blockIdx.x, mt = blockIdx.y;
    const int z = p.zOrder ? p.zOrder[blockIdx.z] : (int)blockIdx.z;
    const int m0 = mt * BM, n0 = nt * BN;
    const size_t listSlot = ((size_t)mt * p.nZ + z) * gridDim.x + nt;
    const uint32_t* __restrict__ list = p.lists + listSlot * p.nChunks;
    const int nAct = p.listCount[listSlot];

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], N_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    double best[MT];
    int bidx[MT];
    const int g = lane >> 2, t = lane & 3;
    // Consumer warp w = 4 m + j sits on SM sub-partition j; it owns row group m and column quarter (j + m) % 4 -- a Latin square, so
    // every sub-partition holds one warp of EACH row group and one of EACH column quarter: a skipped row group or a skipped
    // column quarter takes the same share of work off all four FP64 pipes.
    const int warp_m = warp / NCW, warp_n = (warp + warp_m) % NCW;     // meaningful for consumer warps only

    if (warp == N_CONSUMER_WARPS) {
        // =============================== producer warp ===============================
        const int a = GATHER ? z / p.O : 0;
        const double* __restrict__ bsrc = (GATHER ? p.bmat : p.bmat + (size_t)blockIdx.z * p.zStrideB) + n0;
        const int32_t* __restrict__ reach = GATHER ? p.reachP + (size_t)a * p.Sp : nullptr;
        const double* __restrict__ rto = GATHER ? p.rtoP + (size_t)z * p.Sp : nullptr;
        const double* __restrict__ atile = p.beliefsP + (size_t)mt * p.nChunks * NRG * A_GROUP_DOUBLES;
        auto gathered_row = [&](uint32_t e) -> int {
            const int k = (int)(e & 0xFFFFFFu) * KC + (lane & (KC - 1));
            return GATHER ? reach[k] : min(k, p.S - 1);
        };
        uint32_t e0 = 0, e1 = 0;
        int row0n = 0;
        if (nAct > 0) {
            e0 = list[0];
            e1 = list[min(1, nAct - 1)];
            row0n = gathered_row(e0);
        }
        unsigned long long visited = 0;
        for (int q = 0; q < nAct; q++) {
            const int slot = q % STAGES;
            const unsigned use = (unsigned)(q / STAGES);
            const uint32_t e = e0;
            const int row0 = row0n;
            // prefetch for the next chunk (consumed one iteration from now)
            e0 = e1;
            row0n = gathered_row(e0);
            e1 = list[min(q + 2, nAct - 1)];
            mbar_wait(&s_empty[slot], (use & 1u) ^ 1u);          // first use of a slot passes immediately
            ScoreStage& st = stages[slot];
            const int k0 = (int)(e & 0xFFFFFFu) * KC;
            const uint32_t rg = (e >> 24) & ((1u << NRG) - 1u);
            if (lane == 0) {
                s_meta[slot] = e;
                const unsigned bytes = KC * BN * 8 + (GATHER ? KC * 8 : 0) + __popc(rg) * A_GROUP_DOUBLES * 8;
                mbar_arrive_expect_tx(&s_full[slot], bytes);
                visited += __popc(rg) * __popc(e >> 28);
            }
            __syncwarp();
            if (lane < KC) bulk_g2s(&st.Bs[lane * LDB], bsrc + (size_t)row0 * p.Vp, BN * 8, &s_full[slot]);
            if (lane >= KC && lane < KC + NRG && ((rg >> (lane - KC)) & 1u))
                bulk_g2s(&st.As[(lane - KC) * A_GROUP_DOUBLES], atile + ((size_t)(k0 / KC) * NRG + (lane - KC)) * A_GROUP_DOUBLES,
                         A_GROUP_DOUBLES * 8, &s_full[slot]);
            if (GATHER && lane == KC + NRG) bulk_g2s(&st.Rs[0], rto + k0, KC * 8, &s_full[slot]);
        }
        if (lane == 0 && p.stats && visited) atomicAdd(p.stats, visited);
    } else {
        // =============================== consumer warps ===============================
        double acc[MT][8][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int n = 0; n < 8; n++) { acc[i][n][0] = 0.0; acc[i][n][1] = 0.0; }

        for (int it = 0; it < nAct; it++) {
            const int slot = it % STAGES;
            mbar_wait(&s_full[slot], (unsigned)(it / STAGES) & 1u);
            const ScoreStage& st = stages[slot];
            const uint32_t meta = s_meta[slot];
            if (((meta >> (24 + warp_m)) & 1u) && ((meta >> (28 + warp_n)) & 1u)) {
                const double* Ab = st.As + (warp_m * RG + g) * KC;
                const double* Bb = st.Bs + t * LDB + warp_n * 64 + g;
                const int swz = a_swizzle(g);                 // == a_swizzle(i * 8 + g)
#pragma unroll
                for (int ks = 0; ks < KC / 4; ks++) {
                    double af[MT], bf[8];
#pragma unroll
                    for (int i = 0; i < MT; i++) af[i] = Ab[i * 8 * KC + ((ks * 4 + t) ^ swz)];
                    if (GATHER) {
                        const double r = st.Rs[ks * 4 + t];
#pragma unroll
                        for (int i = 0; i < MT; i++) af[i] *= r;
                    }
#pragma unroll
                    for (int n = 0; n < 8; n++) bf[n] = Bb[ks * 4 * LDB + n * 8];
#pragma unroll
                    for (int i = 0; i < MT; i++)
#pragma unroll
                        for (int n = 0; n < 8; n++) dmma884(acc[i][n], af[i], bf[n]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[slot]);
        }

        // ---- fused argmax, part 1: ascending columns per thread, then the quad (disjoint columns of the same rows).
        //      An empty list leaves acc == 0: every score is 0, first column wins.
        const int cbase = n0 + warp_n * 64 + 2 * t;
#pragma unroll
        for (int i = 0; i < MT; i++) {
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
    }

    // ---- fused argmax, part 2: the column warps through shared memory (every stage has been consumed by now)
    __syncthreads();
    double* sval = reinterpret_cast<double*>(smem_raw);                           // [NCW][BM]
    int* sidx = reinterpret_cast<int*>(smem_raw + sizeof(double) * NCW * BM);     // [NCW][BM]
    if (warp < N_CONSUMER_WARPS && t == 0) {
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const int row = warp_m * RG + i * 8 + g;
            sval[warp_n * BM + row] = best[i];
            sidx[warp_n * BM + row] = bidx[i];
        }
    }
    __syncthreads();
    if (tid < BM && m0 + tid < p.nB) {
        double v = sval[tid];
        int i = sidx[tid];
#pragma unroll
        for (int w = 1; w < NCW; w++) argmax_combine(v, i, sval[w * BM + tid], sidx[w * BM + tid]);
        const size_t out = ((size_t)nt * p.nB + (m0 + tid)) * p.nZ + z;
        p.pval[out] = v;
        p.pidx[out] = i;
    }
}

}  // namespace pbvi
