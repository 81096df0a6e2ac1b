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
// Tiles are BM x BN = 64 beliefs x 256 alphas of one z = (a,o).  A tile touches only the CHUNKS (KC = 4 consecutive source
// states) in which some belief of a row group (RG = 16 beliefs) is non-zero AND some RTO entry of (a,o) is non-zero AND some alpha
// of a column quarter (64 alphas) is non-zero at a state the chunk lands on: every skipped term is an exact zero.  SUB = 4 chunks
// form a pipeline STAGE (16 states): zeros are skipped per chunk, the mbarrier handshake is paid per stage, so a dense
// workload runs like a kernel with 16-state chunks.  build_chunk_lists_kernel writes, per tile, the ordered list of live stages
// with 8 flag bits per chunk (live row groups | live column quarters << 4).
//
// The kernel is persistent (see score_kernel below): one block per SM pulls tiles from a queue.  Warp specialisation
// (N_CONSUMER_WARPS + 1 warps):
//   last warp   producer.  Per live stage it arms the stage's `full` mbarrier with the byte count and issues the stage as
//               bulk async copies (cp.async.bulk, the TMA engine): one 2 KB copy per gathered alphaT row of a live chunk, ONE
//               512-byte copy per live (chunk, row group) (belief_mask_kernel stores the belief tiles as ready-made
//               shared-memory images) and the RTO values of the stage.  List entries and gathered row indices are prefetched
//               one stage ahead, so the producer never waits on a dependent global load.
//   the rest    consumers: wait on `full`, DMMA, arrive on `empty`.  No block-wide barrier anywhere in the stream.
//               Warp 4m + j owns row group m and column quarter (j + m) mod 4 -- a Latin square over the SM sub-partitions
//               (warp id mod 4), so a skipped row group or column quarter takes the same share of work off all four FP64
//               pipes (the first, 128 x 128 layout stalled on the barrier instead: profiles/r01_score_kernel_v1_ncu_summary.txt).
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
    double Bs[SKC * LDB];
    double As[SUB * BM * KC];          // [chunk of the stage][row group][RG rows][KC]
    double Rs[SKC < 16 ? 16 : SKC];
};
constexpr int STAGES = (int)((212 * 1024) / sizeof(ScoreStage)) < 16 ? (int)((212 * 1024) / sizeof(ScoreStage)) : 16;
constexpr size_t SCORE_SMEM = sizeof(ScoreStage) * STAGES;
static_assert(sizeof(ScoreStage) % 128 == 0, "stage alignment");
static_assert(SCORE_SMEM <= 226 * 1024 && STAGES >= 3, "score pipeline exceeds shared memory");
static_assert(BM == NRG * RG && N_CONSUMER_WARPS * 32 == SCORE_THREADS, "warp layout: NRG row groups x NCW column warps");
static_assert(SKC <= 16 && SKC + SUB * NRG <= 32 && NRG == 4 && NCW == 4 && SUB * 8 <= 32,
              "producer lane mapping: lanes [0,SKC) B rows, [SKC, SKC + SUB*NRG) A row groups; flags: 4 + 4 bits per chunk");
static_assert((A_GROUP_DOUBLES * 8) % 16 == 0 && (KC * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");

struct ScoreParams {
    const double* beliefsP;    // [nMt][nChunks][NRG][RG][KC]  belief tiles as swizzled shared-memory images (belief_mask_kernel)
    const double* bmat;        // GATHER: alphaT [S+1][Vp];  PLAIN: [nzLaunch][S+1][Vp], matrix of queue position zi at zi * zStrideB (row S: zeros)
    size_t zStrideB;
    const int32_t* reachP;     // [A][Sp]          (GATHER)
    const double* rtoP;        // [A*O][Sp]        (GATHER)
    const uint2* lists;        // [nMt][nZ][nNt][nStages]  (stage index, flags): flags byte h = live row groups | live column quarters << 4 of chunk h
    const int32_t* listCount;  // [nMt][nZ][nNt]
    const int32_t* zOrder;     // [nzLaunch] z of queue position zi, heavy first (nullptr: identity)
    double* pval;              // [nNt * NCW][nB][nZ]  partial maxima per 64-column quarter, ascending columns
    int32_t* pidx;             // [nNt * NCW][nB][nZ]
    int* tileCounter;          // zeroed before the launch: next tile of the queue
    int nMt, nNt, nzLaunch;    // tile queue = nzLaunch x nMt x nNt, z-major (heavy z first through zOrder)
    unsigned long long* stats; // visited (chunk, row group, column quarter) triples, summed over blocks
    int nB, S, Sp, V, Vp, nChunks, nStages, nZ, O;
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

// (value, index) combine with np.argmax semantics: a NaN counts as the maximum, among equals (and among NaNs) the lower index
// wins.  Index ARG_NONE marks "no column seen" (a thread / column quarter that lies entirely beyond V) and loses to everything.
constexpr int ARG_NONE = 0x7fffffff;
__device__ __forceinline__ void argmax_combine(double& v, int& i, double ov, int oi) {
    const bool on = ov != ov, vn = v != v;
    const bool better = on ? (!vn || oi < i) : (!vn && (ov > v || (ov == v && oi < i)));
    if (oi != ARG_NONE && (i == ARG_NONE || better)) { v = ov; i = oi; }
}
// the same for a candidate that comes AFTER everything seen so far (ascending index): only a strictly larger value or the first NaN replaces
__device__ __forceinline__ void argmax_append(double& v, int& i, double ov, int oi) {
    if (i == ARG_NONE || (v == v && (ov > v || ov != ov))) { v = ov; i = oi; }
}

constexpr uint32_t META_TILE_END = 0xFFFFFFFFu;   // ring item that closes a tile (no stage has this index)
constexpr uint32_t META_STOP = 0xFFFFFFFEu;       // ring item that ends the stream

// Persistent, streaming form: one block per SM walks a queue of (z, belief tile, alpha tile) tiles (heavy z first, handed out by
// an atomic counter).  The producer warp turns the queue into ONE stream of ring items -- the live chunks of a tile followed
// by a TILE_END item -- and runs ahead across tile boundaries, so the first loads of the next tile overlap the last DMMAs
// of the current one and a tile with an empty list costs one ring item instead of a block launch.  A consumer warp that meets
// TILE_END reduces its own 16 x 64 accumulator block to (max, first argmax) per row, writes that partial result straight to
// global memory (the column quarters are merged by combine_tiles_kernel) and clears its accumulators: no block-wide barrier
// anywhere after the mbarrier initialisation.
template <bool GATHER>
__global__ void __launch_bounds__(SCORE_THREADS_TOTAL, 1) score_kernel(const ScoreParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScoreStage* stages = reinterpret_cast<ScoreStage*>(smem_raw);
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ __align__(8) uint64_t s_empty[STAGES];
    __shared__ uint2 s_meta[STAGES];               // (stage index or META_*, flags or tile id)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nTiles = p.nzLaunch * p.nMt * p.nNt;
    const int tilesPerZ = p.nMt * p.nNt;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], N_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == N_CONSUMER_WARPS) {
        // =============================== producer warp ===============================
        int q = 0;                                            // ring items issued so far
        unsigned long long visited = 0;
        auto fetch_tile = [&]() -> int { return lane == 0 ? atomicAdd(p.tileCounter, 1) : 0; };
        struct Hdr { int nAct; uint2 e0, e1; };
        auto slot_of = [&](int t) -> size_t {
            const int zi = t / tilesPerZ, rem = t - zi * tilesPerZ;
            const int z = p.zOrder ? p.zOrder[zi] : zi;
            return ((size_t)(rem / p.nNt) * p.nZ + z) * p.nNt + (rem % p.nNt);
        };
        auto load_hdr = [&](int t) -> Hdr {
            Hdr h{0, make_uint2(0u, 0u), make_uint2(0u, 0u)};
            if (t < nTiles) {
                const size_t listSlot = slot_of(t);
                h.nAct = p.listCount[listSlot];
                const uint2* l = p.lists + listSlot * p.nStages;
                h.e0 = l[0];
                h.e1 = l[1 < p.nStages ? 1 : 0];
            }
            return h;
        };
        int rawNext2 = fetch_tile();
        int tCur = __shfl_sync(0xffffffffu, rawNext2, 0);
        rawNext2 = fetch_tile();
        Hdr hNext = load_hdr(tCur);
        while (tCur < nTiles) {
            const int t = tCur;
            const Hdr h = hNext;
            tCur = __shfl_sync(0xffffffffu, rawNext2, 0);     // requested one tile ago
            rawNext2 = fetch_tile();
            hNext = load_hdr(tCur);                           // consumed one tile from now
            const int zi = t / tilesPerZ, rem = t - zi * tilesPerZ;
            const int mt = rem / p.nNt, nt = rem % p.nNt;
            const int z = p.zOrder ? p.zOrder[zi] : zi;
            const int a = GATHER ? z / p.O : 0;
            const double* __restrict__ bsrc = (GATHER ? p.bmat : p.bmat + (size_t)zi * p.zStrideB) + nt * BN;
            const int32_t* __restrict__ reach = GATHER ? p.reachP + (size_t)a * p.Sp : nullptr;
            const double* __restrict__ rto = GATHER ? p.rtoP + (size_t)z * p.Sp : nullptr;
            const double* __restrict__ atile = p.beliefsP + (size_t)mt * p.nChunks * NRG * A_GROUP_DOUBLES;
            const uint2* __restrict__ list = p.lists + slot_of(t) * p.nStages;
            const int nAct = h.nAct;
            auto gathered_row = [&](uint32_t stage) -> int {
                const int k = (int)stage * SKC + (lane & (SKC - 1));
                return GATHER ? reach[k] : min(k, p.S);         // pad states (k >= S) read the all-zero row S of the B matrix
            };
            uint2 e0 = h.e0, e1 = h.e1;
            int row0n = nAct > 0 ? gathered_row(e0.x) : 0;
            for (int c = 0; c < nAct; c++, q++) {
                const int slot = q % STAGES;
                const unsigned use = (unsigned)(q / STAGES);
                const uint2 e = e0;
                const int row0 = row0n;
                // prefetch for the next stage (consumed one iteration from now)
                e0 = e1;
                row0n = c + 1 < nAct ? gathered_row(e0.x) : 0;
                e1 = list[min(c + 2, nAct - 1)];
                mbar_wait(&s_empty[slot], (use & 1u) ^ 1u);          // first use of a slot passes immediately
                ScoreStage& st = stages[slot];
                const int k0 = (int)e.x * SKC;
                const uint32_t flags = e.y;
                if (lane == 0) {
                    s_meta[slot] = e;
                    unsigned bytes = GATHER ? SKC * 8 : 0;
                    unsigned work = 0;
#pragma unroll
                    for (int hh = 0; hh < SUB; hh++) {
                        const unsigned rg = (flags >> (8 * hh)) & 0xFu, cq = (flags >> (8 * hh + 4)) & 0xFu;
                        bytes += (rg ? KC * BN * 8 : 0) + __popc(rg) * A_GROUP_DOUBLES * 8;
                        work += __popc(rg) * __popc(cq);
                    }
                    mbar_arrive_expect_tx(&s_full[slot], bytes);
                    visited += work;
                }
                __syncwarp();
                if (lane < SKC && ((flags >> (8 * (lane / KC))) & 0xFu))
                    bulk_g2s(&st.Bs[lane * LDB], bsrc + (size_t)row0 * p.Vp, BN * 8, &s_full[slot]);
                if (lane >= SKC && lane < SKC + SUB * NRG && ((flags >> (8 * ((lane - SKC) / NRG) + (lane - SKC) % NRG)) & 1u))
                    bulk_g2s(&st.As[(lane - SKC) * A_GROUP_DOUBLES],
                             atile + ((size_t)(k0 / KC + (lane - SKC) / NRG) * NRG + (lane - SKC) % NRG) * A_GROUP_DOUBLES, A_GROUP_DOUBLES * 8,
                             &s_full[slot]);
                if (GATHER && lane == 31) bulk_g2s(&st.Rs[0], rto + k0, SKC * 8, &s_full[slot]);
            }
            {   // TILE_END
                const int slot = q % STAGES;
                mbar_wait(&s_empty[slot], (((unsigned)(q / STAGES)) & 1u) ^ 1u);
                if (lane == 0) {
                    s_meta[slot] = make_uint2(META_TILE_END, (unsigned)t);
                    mbar_arrive(&s_full[slot]);
                }
                __syncwarp();
                q++;
            }
        }
        {   // STOP
            const int slot = q % STAGES;
            mbar_wait(&s_empty[slot], (((unsigned)(q / STAGES)) & 1u) ^ 1u);
            if (lane == 0) {
                s_meta[slot] = make_uint2(META_STOP, 0u);
                mbar_arrive(&s_full[slot]);
            }
        }
        if (lane == 0 && p.stats && visited) atomicAdd(p.stats, visited);
    } else {
        // =============================== consumer warps ===============================
        const int g = lane >> 2, t = lane & 3;
        // Consumer warp w = 4 m + j sits on SM sub-partition j; it owns row group m and column quarter (j + m) % 4 -- a Latin
        // square, so every sub-partition holds one warp of EACH row group and one of EACH column quarter: a skipped row group or
        // a skipped column quarter takes the same share of work off all four FP64 pipes.
        const int warp_m = warp / NCW, warp_n = (warp + warp_m) % NCW;
        double acc[MT][8][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int n = 0; n < 8; n++) { acc[i][n][0] = 0.0; acc[i][n][1] = 0.0; }

        const int swz = a_swizzle(g);                 // == a_swizzle(i * 8 + g)
        const int aOff = (warp_m * RG + g) * KC, bOff = t * LDB + warp_n * 64 + g;
        for (int it = 0;; it++) {
            const int slot = it % STAGES;
            mbar_wait(&s_full[slot], (unsigned)(it / STAGES) & 1u);
            const ScoreStage& st = stages[slot];
            const uint2 meta = s_meta[slot];
            if (meta.x >= META_STOP) {
                if (meta.x == META_STOP) break;
                // ---- TILE_END: fused argmax of this warp's block (ascending columns per thread, then the quad, which holds disjoint
                //      columns of the same rows).  A tile without live chunks leaves acc == 0: every score is 0, first column wins.
                const int tile = (int)meta.y;
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[slot]);
                const int zi = tile / tilesPerZ, rem = tile - zi * tilesPerZ;
                const int mt = rem / p.nNt, nt = rem % p.nNt;
                const int z = p.zOrder ? p.zOrder[zi] : zi;
                const int cbase = nt * BN + warp_n * 64 + 2 * t;
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    double best = -INFINITY;
                    int bidx = ARG_NONE;
#pragma unroll
                    for (int n = 0; n < 8; n++)
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const int col = cbase + n * 8 + j;
                            if (col < p.V) argmax_append(best, bidx, acc[i][n][j], col);
                            acc[i][n][j] = 0.0;
                        }
#pragma unroll
                    for (int off = 1; off <= 2; off <<= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
                        argmax_combine(best, bidx, ov, oi);
                    }
                    const int row = mt * BM + warp_m * RG + i * 8 + g;
                    if (t == 0 && row < p.nB) {
                        const size_t out = ((size_t)(nt * NCW + warp_n) * p.nB + row) * p.nZ + z;
                        p.pval[out] = best;
                        p.pidx[out] = bidx;
                    }
                }
                continue;
            }
#pragma unroll
            for (int hh = 0; hh < SUB; hh++) {
                if (!(((meta.y >> (8 * hh + warp_m)) & 1u) && ((meta.y >> (8 * hh + 4 + warp_n)) & 1u))) continue;
#pragma unroll
                for (int ks = 0; ks < KC / 4; ks++) {
                    double af[MT], bf[8];
#pragma unroll
                    for (int i = 0; i < MT; i++) af[i] = st.As[hh * BM * KC + aOff + i * 8 * KC + ((ks * 4 + t) ^ swz)];
                    if (GATHER) {
                        const double r = st.Rs[hh * KC + ks * 4 + t];
#pragma unroll
                        for (int i = 0; i < MT; i++) af[i] *= r;
                    }
#pragma unroll
                    for (int n = 0; n < 8; n++) bf[n] = st.Bs[bOff + (hh * KC + ks * 4) * LDB + n * 8];
#pragma unroll
                    for (int i = 0; i < MT; i++)
#pragma unroll
                        for (int n = 0; n < 8; n++) dmma884(acc[i][n], af[i], bf[n]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[slot]);
        }
    }
}

}  // namespace pbvi
