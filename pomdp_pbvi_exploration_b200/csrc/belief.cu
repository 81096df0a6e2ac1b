// Belief update and observation likelihoods (reference Belief.update, src/pomdp.py:382-421; batched twin :1415-1419;
// P(o|b,a) einsum at :1814 / :2046 / :1751).  HBM-bound gather kernels.
//
// Bit-parity with the reference: np.bincount accumulates the weights of one bin in ascending flattened (s,r) order, so
// a CSR over landing states with ascending sources reproduces it exactly; the normaliser is np.sum's float64 pairwise
// summation, reproduced from the precomputed combine tree of the model handle.
#include "pbvi_common.cuh"

namespace pbvi {

// out[i][s'] = sum over predecessors k of s' under action a_i (ascending k) of RTO[a_i][o_i][k] * belief_i[k / R]
__global__ void __launch_bounds__(256) belief_project_kernel(const double* __restrict__ beliefs, const int32_t* __restrict__ actions,
                                                             const int32_t* __restrict__ observations, size_t beliefStride,
                                                             const int32_t* __restrict__ predPtr, const int32_t* __restrict__ predK,
                                                             const double* __restrict__ rtoK, int S, int R, int O,
                                                             double* __restrict__ out, int aConst, int oConst,
                                                             const double* __restrict__ chooseP, double chooseU, int32_t* __restrict__ chosen,
                                                             int zTile) {
    const int i = blockIdx.y;
    const int sp = blockIdx.x * 256 + threadIdx.x;
    int oPick = oConst;
    if (chooseP) {
        // the observation is DRAWN here (Perseus walk, src/pomdp.py:2045-2047): np.random.choice(observations, p=P(o|b,a)) is
        // cdf = cumsum(p) (sequential adds); cdf /= cdf[-1]; searchsorted(cdf, u, side='right') with ONE uniform u, which the host
        // drew in the reference's order.  Every thread repeats the O additions (same order => same doubles => same choice).
        double tot = 0.0;
        for (int o = 0; o < O; o++) tot = __dadd_rn(tot, chooseP[o]);
        double run = 0.0;
        int idx = 0;
        for (int o = 0; o < O; o++) {
            run = __dadd_rn(run, chooseP[o]);
            if (__ddiv_rn(run, tot) <= chooseU) idx++;
        }
        oPick = min(idx, O - 1);
        if (chosen && blockIdx.x == 0 && threadIdx.x == 0) *chosen = oPick;
    }
    if (sp >= S) return;
    // zTile > 0: row i is successor z = i % zTile (a = z / O, o = z % O) of belief i / zTile -- all successors of a belief chunk in one launch
    const int a = zTile > 0 ? (i % zTile) / O : (actions ? actions[i] : aConst);
    const int o = zTile > 0 ? (i % zTile) % O : (observations ? observations[i] : oPick);
    const size_t K = (size_t)S * R;
    const int32_t* ptr = predPtr + (size_t)a * (S + 1);
    const int32_t* pk = predK + (size_t)a * K;
    const double* rto = rtoK + ((size_t)a * O + o) * K;
    const double* b = beliefs + (size_t)(zTile > 0 ? i / zTile : i) * beliefStride;
    double acc = 0.0;   // bincount starts every bin at +0.0 and adds in order
    const int end = ptr[sp + 1];
    for (int j = ptr[sp]; j < end; j++) {
        const int k = pk[j];
        acc = __dadd_rn(acc, __dmul_rn(rto[k], b[R == 1 ? k : k / R]));
    }
    out[(size_t)i * S + sp] = acc;
}

// NumPy pairwise sum of each row (block per row), then optional in-place division by it
__global__ void __launch_bounds__(256) pairwise_normalise_kernel(double* __restrict__ rows, int S, const int2* __restrict__ leaves,
                                                                 int nLeaves, const int2* __restrict__ nodes, int nNodes,
                                                                 int normalise, double* __restrict__ norm) {
    extern __shared__ double s_sum[];   // [nLeaves] leaf sums, then [nNodes] node sums
    __shared__ double s_total;
    double* row = rows + (size_t)blockIdx.x * S;
    // A leaf (<= 128 contiguous elements) is summed by 8 lanes: lane j owns NumPy's accumulator r[j] (every 8th element, in
    // order), then the fixed tree ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and the < 8 trailing elements -- the same additions in
    // the same order as DOUBLE_pairwise_sum, 8 of them at a time.
    const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
    for (int l0 = 0; l0 < nLeaves; l0 += 32) {               // uniform trip count: the shuffles below need the whole warp
        const int l = l0 + grp;
        const bool valid = l < nLeaves;
        const int off = valid ? leaves[l].x : 0, n = valid ? leaves[l].y : 0;
        const double* a = row + off;
        const int lim = n - (n % 8);
        double r = (n >= 8) ? a[j] : 0.0;
        for (int i = 8; i < lim; i += 8) r = __dadd_rn(r, a[i + j]);
        const double p = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
        const double q = __dadd_rn(p, __shfl_down_sync(0xffffffffu, p, 2, 8));
        double res = __dadd_rn(q, __shfl_down_sync(0xffffffffu, q, 4, 8));
        if (valid && j == 0) {
            if (n < 8) {
                res = 0.0;
                for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
            } else {
                for (int i = lim; i < n; i++) res = __dadd_rn(res, a[i]);
            }
            s_sum[l] = res;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* nsum = s_sum + nLeaves;
        for (int j = 0; j < nNodes; j++) {
            const int l = nodes[j].x, r = nodes[j].y;
            const double lv = l < 0 ? s_sum[~l] : nsum[l];
            const double rv = r < 0 ? s_sum[~r] : nsum[r];
            nsum[j] = __dadd_rn(lv, rv);
        }
        s_total = nNodes ? nsum[nNodes - 1] : s_sum[0];
        if (norm) norm[blockIdx.x] = s_total;
    }
    __syncthreads();
    if (normalise) {
        const double tot = s_total;
        for (int s = threadIdx.x; s < S; s += 256) row[s] = row[s] / tot;   // 0/0 = NaN for impossible observations, as in the reference
    }
}

// The same normalisation split in two for a CHAIN of updates (one row per step, every step waiting for the previous one): with
// one block per row the 256 threads of a single SM walk 170 leaves and 22 000 divisions alone (80 us per step on the olfactory
// model, the whole cost of an FSVI expansion).  pairwise_leaf_kernel spreads the leaf sums over ceil(nLeaves / 32) blocks,
// pairwise_finish_kernel walks the (tiny) combine tree in every block and divides that block's slice.  Same additions in the
// same order: bit-identical to pairwise_normalise_kernel.
__global__ void __launch_bounds__(256) pairwise_leaf_kernel(const double* __restrict__ rows, int S, const int2* __restrict__ leaves, int nLeaves,
                                                            double* __restrict__ leafSumsAll) {
    const double* row = rows + (size_t)blockIdx.y * S;             // blockIdx.y: row (a few rows at a time: grid.y <= 64)
    double* leafSums = leafSumsAll + (size_t)blockIdx.y * nLeaves;
    const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
    const int l = blockIdx.x * 32 + grp;
    const bool valid = l < nLeaves;
    const int off = valid ? leaves[l].x : 0, n = valid ? leaves[l].y : 0;
    const double* a = row + off;
    const int lim = n - (n % 8);
    double r = (n >= 8) ? a[j] : 0.0;
    for (int i = 8; i < lim; i += 8) r = __dadd_rn(r, a[i + j]);
    const double p = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
    const double q = __dadd_rn(p, __shfl_down_sync(0xffffffffu, p, 2, 8));
    double res = __dadd_rn(q, __shfl_down_sync(0xffffffffu, q, 4, 8));
    if (valid && j == 0) {
        if (n < 8) {
            res = 0.0;
            for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
        } else {
            for (int i = lim; i < n; i++) res = __dadd_rn(res, a[i]);
        }
        leafSums[l] = res;
    }
}

constexpr int FINISH_SLICE = 2048;   // elements divided per block

__global__ void __launch_bounds__(256) pairwise_finish_kernel(double* __restrict__ rows, int S, const double* __restrict__ leafSumsAll, int nLeaves,
                                                              const int2* __restrict__ nodes, int nNodes, double* __restrict__ normAll, int normalise) {
    extern __shared__ double s_sum[];   // [nLeaves] leaf sums, then [nNodes] node sums
    __shared__ double s_total;
    double* row = rows + (size_t)blockIdx.y * S;
    const double* leafSums = leafSumsAll + (size_t)blockIdx.y * nLeaves;
    double* norm = normAll ? normAll + blockIdx.y : nullptr;
    for (int l = threadIdx.x; l < nLeaves; l += 256) s_sum[l] = leafSums[l];
    __syncthreads();
    if (threadIdx.x == 0) {
        double* nsum = s_sum + nLeaves;
        for (int j = 0; j < nNodes; j++) {
            const int l = nodes[j].x, r = nodes[j].y;
            const double lv = l < 0 ? s_sum[~l] : nsum[l];
            const double rv = r < 0 ? s_sum[~r] : nsum[r];
            nsum[j] = __dadd_rn(lv, rv);
        }
        s_total = nNodes ? nsum[nNodes - 1] : s_sum[0];
        if (norm && blockIdx.x == 0) *norm = s_total;
    }
    __syncthreads();
    if (!normalise) return;
    const double tot = s_total;
    const int s1 = min(S, (int)(blockIdx.x + 1) * FINISH_SLICE);
    for (int s = blockIdx.x * FINISH_SLICE + threadIdx.x; s < s1; s += 256) row[s] = row[s] / tot;   // 0/0 = NaN, as in the reference
}

// out[i][a][o] = sum_k RTO[a][o][k] * belief_i[k / R]; block per (z, i)
__global__ void __launch_bounds__(256) observation_probability_kernel(const double* __restrict__ beliefs, const double* __restrict__ rtoK,
                                                                      int S, int R, int nZ, double* __restrict__ out) {
    __shared__ double sh[8];
    const int z = blockIdx.x, i = blockIdx.y;
    const size_t K = (size_t)S * R;
    const double* rto = rtoK + (size_t)z * K;
    const double* b = beliefs + (size_t)i * S;
    double part = 0.0;
    for (int s = threadIdx.x; s < S; s += 256) {
        const double bs = b[s];
        if (bs != 0.0)
            for (int r = 0; r < R; r++) part = fma(rto[(size_t)s * R + r], bs, part);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) tot += sh[w];
        out[(size_t)i * nZ + z] = tot;
    }
}

}  // namespace pbvi

namespace pbvi {
constexpr size_t CHAIN_SMEM_MAX = 220 * 1024;

// ---- a whole CHAIN of belief updates in ONE launch (the FSVI trajectory, the Perseus walk): one block of 1024 threads keeps the
//      current belief in shared memory and runs the steps back to back, so a step costs a dozen block barriers instead of three or
//      four dependent kernel launches (the multi-launch chain is bound by the launch rate: ~6 us per launch from the host, ~25 us
//      per step).  Every phase repeats the arithmetic of the separate kernels bit for bit:
//        observation draw   P(o|b,a) with observation_probability_kernel's summation shape (256 "virtual threads" per observation:
//                           fma over s = t, t + 256, ..., warp shuffle tree, 8 warp partials in order), then NumPy's choice rule;
//        projection         bincount order over the predecessors of every landing state (belief_project_kernel);
//        normaliser         NumPy's pairwise sum: 8-lane leaf sums (pairwise_leaf_kernel), then the combine tree level by level
//                           (same additions, same association), then one division per state.
//      A step flagged `reset` continues from b0 (FSVI: end state reached).  Used when a belief fits in shared memory (S <= ~26 000).
constexpr int CHAIN_THREADS = 1024;

__global__ void __launch_bounds__(CHAIN_THREADS, 1) belief_chain_kernel(const double* __restrict__ b0, const int32_t* __restrict__ actions,
                                                                        const int32_t* __restrict__ observations, const double* __restrict__ uniforms,
                                                                        const uint8_t* __restrict__ resets, int n, const int32_t* __restrict__ predPtr,
                                                                        const int32_t* __restrict__ predK, const double* __restrict__ rtoK, int S, int R,
                                                                        int O, const int2* __restrict__ leaves, int nLeaves,
                                                                        const int2* __restrict__ nodes, int nNodes,
                                                                        const int32_t* __restrict__ levelNodes, const int32_t* __restrict__ levelPtr,
                                                                        int nLevels, double* __restrict__ out, int32_t* __restrict__ chosenOut) {
    extern __shared__ double chain_smem[];
    double* sb = chain_smem;                     // [S] current belief
    double* s_sum = sb + S;                      // [nLeaves] leaf sums, then [nNodes] node sums
    __shared__ double s_part[4][8];              // warp partials of up to four observation sums at a time
    __shared__ double s_p[32];                   // P(o | b, a)
    const int tid = threadIdx.x, lane = tid & 31;
    const size_t K = (size_t)S * R;
    for (int s = tid; s < S; s += CHAIN_THREADS) sb[s] = b0[s];
    __syncthreads();
    for (int i = 0; i < n; i++) {
        const int a = actions[i];
        int o = observations ? observations[i] : -1;
        if (o < 0) {
            // ---- P(o | b, a) for every o, four observations at a time (thread group g = tid / 256 takes observation o0 + g)
            for (int o0 = 0; o0 < O; o0 += 4) {
                const int g = tid >> 8, t = tid & 255, oo = o0 + g;
                double part = 0.0;
                if (oo < O) {
                    // (observation_probability_kernel skips bs == 0; fma(x, 0, part) == part for the finite x of a model table, so
                    // the unconditional form gives the same bits and lets the loads of consecutive iterations overlap -- one block
                    // has nothing else to hide their latency with)
                    const double* rto = rtoK + ((size_t)a * O + oo) * K;
                    if (R == 1) {
#pragma unroll 8
                        for (int s = t; s < S; s += 256) part = fma(rto[s], sb[s], part);
                    } else {
                        for (int s = t; s < S; s += 256) {
                            const double bs = sb[s];
                            for (int r = 0; r < R; r++) part = fma(rto[(size_t)s * R + r], bs, part);
                        }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
                if (lane == 0) s_part[g][t >> 5] = part;
                __syncthreads();
                if (tid < 4 && o0 + tid < O) {
                    double tot = 0.0;
#pragma unroll
                    for (int w = 0; w < 8; w++) tot += s_part[tid][w];
                    s_p[o0 + tid] = tot;
                }
                __syncthreads();
            }
            // np.random.choice(observations, p=P): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
            const double u = uniforms[i];
            double tot = 0.0;
            for (int x = 0; x < O; x++) tot = __dadd_rn(tot, s_p[x]);
            double run = 0.0;
            int idx = 0;
            for (int x = 0; x < O; x++) {
                run = __dadd_rn(run, s_p[x]);
                if (__ddiv_rn(run, tot) <= u) idx++;
            }
            o = min(idx, O - 1);
        }
        if (tid == 0 && chosenOut) chosenOut[i] = o;
        // ---- projection (bincount order) into the output row
        double* row = out + (size_t)i * S;
        {
            const int32_t* ptr = predPtr + (size_t)a * (S + 1);
            const int32_t* pk = predK + (size_t)a * K;
            const double* rto = rtoK + ((size_t)a * O + o) * K;
            // four landing states per thread at a time: their predecessor chains are independent, so their loads overlap
            constexpr int PU = 4;
            for (int sp0 = tid; sp0 < S; sp0 += PU * CHAIN_THREADS) {
                int beg[PU], len[PU], longest = 0;
                double acc[PU];
#pragma unroll
                for (int u = 0; u < PU; u++) {
                    const int sp = sp0 + u * CHAIN_THREADS;
                    beg[u] = sp < S ? ptr[sp] : 0;
                    len[u] = sp < S ? ptr[sp + 1] - beg[u] : 0;
                    longest = max(longest, len[u]);
                    acc[u] = 0.0;
                }
                for (int x = 0; x < longest; x++) {
#pragma unroll
                    for (int u = 0; u < PU; u++)
                        if (x < len[u]) {
                            const int k = pk[beg[u] + x];
                            acc[u] = __dadd_rn(acc[u], __dmul_rn(rto[k], sb[R == 1 ? k : k / R]));
                        }
                }
#pragma unroll
                for (int u = 0; u < PU; u++) {
                    const int sp = sp0 + u * CHAIN_THREADS;
                    if (sp < S) row[sp] = acc[u];
                }
            }
        }
        __syncthreads();
        // ---- leaf sums: 8 lanes per leaf, NumPy's accumulators r[0..7] + the fixed 3-level tree + the < 8 trailing elements
        for (int q0 = 0; q0 < nLeaves * 8; q0 += CHAIN_THREADS) {            // uniform trip count: the shuffles need whole warps
            const int q = q0 + tid, l = q >> 3, j = q & 7;
            const bool valid = l < nLeaves;
            const int off = valid ? leaves[l].x : 0, len = valid ? leaves[l].y : 0;
            const double* av = row + off;
            const int lim = len - (len % 8);
            // a leaf has at most 128 elements = 16 per lane: load them all first (independent), then add in NumPy's order
            double ev[16];
#pragma unroll
            for (int x = 0; x < 16; x++) ev[x] = (x * 8 < lim) ? av[x * 8 + j] : 0.0;
            double r = (len >= 8) ? ev[0] : 0.0;
#pragma unroll
            for (int x = 1; x < 16; x++)
                if (x * 8 < lim) r = __dadd_rn(r, ev[x]);
            const double p2 = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
            const double q2 = __dadd_rn(p2, __shfl_down_sync(0xffffffffu, p2, 2, 8));
            double res = __dadd_rn(q2, __shfl_down_sync(0xffffffffu, q2, 4, 8));
            if (valid && j == 0) {
                if (len < 8) {
                    res = 0.0;
                    for (int x = 0; x < len; x++) res = __dadd_rn(res, av[x]);
                } else {
                    for (int x = lim; x < len; x++) res = __dadd_rn(res, av[x]);
                }
                s_sum[l] = res;
            }
        }
        __syncthreads();
        // ---- the combine tree, level by level
        double* nsum = s_sum + nLeaves;
        for (int lev = 0; lev < nLevels; lev++) {
            for (int x = levelPtr[lev] + tid; x < levelPtr[lev + 1]; x += CHAIN_THREADS) {
                const int j = levelNodes[x];
                const int l = nodes[j].x, r = nodes[j].y;
                nsum[j] = __dadd_rn(l < 0 ? s_sum[~l] : nsum[l], r < 0 ? s_sum[~r] : nsum[r]);
            }
            __syncthreads();
        }
        const double tot = nNodes ? nsum[nNodes - 1] : s_sum[0];
        // ---- normalise; the next step starts from this row, or from b0 again after a reset
        const bool reset = resets && resets[i];
#pragma unroll 4
        for (int sp = tid; sp < S; sp += CHAIN_THREADS) {
            const double v = row[sp] / tot;                               // 0/0 = NaN for an impossible observation, as in the reference
            row[sp] = v;
            sb[sp] = reset ? b0[sp] : v;
        }
        __syncthreads();
    }
}

// ---- the same chain on a thread-block CLUSTER of 8 CTAs (8 SMs): every CTA keeps a full copy of the current belief in its shared memory
//      (the projection gathers predecessors from anywhere), owns one eighth of the states for the projection and the division and every
//      eighth leaf of the pairwise sum, and the CTAs meet at four cluster barriers per step; rows, leaf sums and the observation
//      probabilities cross CTAs through global memory (L2; read with ld.cg).  Same arithmetic, same association: bitwise the other two
//      forms.  pbvi_set_option("chain_kernel", 2).
constexpr int CHAIN_CLUSTER = 8;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}

__global__ void __launch_bounds__(CHAIN_THREADS, 1) belief_chain_cluster_kernel(
    const double* __restrict__ b0, const int32_t* __restrict__ actions, const int32_t* __restrict__ observations,
    const double* __restrict__ uniforms, const uint8_t* __restrict__ resets, int n, const int32_t* __restrict__ predPtr,
    const int32_t* __restrict__ predK, const double* __restrict__ rtoK, int S, int R, int O, const int2* __restrict__ leaves, int nLeaves,
    const int2* __restrict__ nodes, int nNodes, const int32_t* __restrict__ levelNodes, const int32_t* __restrict__ levelPtr, int nLevels,
    double* out, int32_t* __restrict__ chosenOut, double* leafSumsG, double* obsG) {
    extern __shared__ double chain_smem[];
    double* sb = chain_smem;                     // [S] current belief (full copy per CTA)
    double* s_sum = sb + S;                      // [nLeaves] leaf sums, then [nNodes] node sums
    __shared__ double s_part[8];
    const int tid = threadIdx.x, lane = tid & 31;
    const int rank = (int)cluster_cta_rank();
    const size_t K = (size_t)S * R;
    const int per = (S + CHAIN_CLUSTER - 1) / CHAIN_CLUSTER, lo = min(S, rank * per), hi = min(S, lo + per);
    for (int s = tid; s < S; s += CHAIN_THREADS) sb[s] = b0[s];
    __syncthreads();
    for (int i = 0; i < n; i++) {
        const int a = actions[i];
        int o = observations ? observations[i] : -1;
        if (o < 0) {
            // P(o | b, a): CTA r takes the observations r, r + 8, ...; 256 threads with observation_probability_kernel's summation shape
            for (int oo = rank; oo < O; oo += CHAIN_CLUSTER) {
                double part = 0.0;
                if (tid < 256) {
                    const double* rto = rtoK + ((size_t)a * O + oo) * K;
                    if (R == 1) {
#pragma unroll 8
                        for (int s = tid; s < S; s += 256) part = fma(rto[s], sb[s], part);
                    } else {
                        for (int s = tid; s < S; s += 256) {
                            const double bs = sb[s];
                            for (int r = 0; r < R; r++) part = fma(rto[(size_t)s * R + r], bs, part);
                        }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
                    if (lane == 0) s_part[tid >> 5] = part;
                }
                __syncthreads();
                if (tid == 0) {
                    double tot = 0.0;
#pragma unroll
                    for (int w = 0; w < 8; w++) tot += s_part[w];
                    obsG[oo] = tot;
                }
                __syncthreads();
            }
            cluster_sync_all();
            const double u = uniforms[i];
            double tot = 0.0;
            for (int x = 0; x < O; x++) tot = __dadd_rn(tot, __ldcg(obsG + x));
            double run = 0.0;
            int idx = 0;
            for (int x = 0; x < O; x++) {
                run = __dadd_rn(run, __ldcg(obsG + x));
                if (__ddiv_rn(run, tot) <= u) idx++;
            }
            o = min(idx, O - 1);
        }
        if (rank == 0 && tid == 0 && chosenOut) chosenOut[i] = o;
        // ---- projection of this CTA's slice (bincount order)
        double* row = out + (size_t)i * S;
        {
            const int32_t* ptr = predPtr + (size_t)a * (S + 1);
            const int32_t* pk = predK + (size_t)a * K;
            const double* rto = rtoK + ((size_t)a * O + o) * K;
            for (int sp = lo + tid; sp < hi; sp += CHAIN_THREADS) {
                double acc = 0.0;
                const int end = ptr[sp + 1];
                for (int j = ptr[sp]; j < end; j++) {
                    const int k = pk[j];
                    acc = __dadd_rn(acc, __dmul_rn(rto[k], sb[R == 1 ? k : k / R]));
                }
                row[sp] = acc;
            }
        }
        cluster_sync_all();
        // ---- leaf sums of the leaves rank, rank + 8, ...: 8 lanes per leaf
        for (int q0 = 0; q0 * CHAIN_CLUSTER < nLeaves * 8; q0 += CHAIN_THREADS) {        // uniform trip count over the cluster
            const int q = q0 + tid, l = rank + CHAIN_CLUSTER * (q >> 3), j = q & 7;
            const bool valid = l < nLeaves;
            const int off = valid ? leaves[l].x : 0, len = valid ? leaves[l].y : 0;
            const double* av = row + off;
            const int lim = len - (len % 8);
            double ev[16];
#pragma unroll
            for (int x = 0; x < 16; x++) ev[x] = (x * 8 < lim) ? __ldcg(av + x * 8 + j) : 0.0;
            double r = (len >= 8) ? ev[0] : 0.0;
#pragma unroll
            for (int x = 1; x < 16; x++)
                if (x * 8 < lim) r = __dadd_rn(r, ev[x]);
            const double p2 = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
            const double q2 = __dadd_rn(p2, __shfl_down_sync(0xffffffffu, p2, 2, 8));
            double res = __dadd_rn(q2, __shfl_down_sync(0xffffffffu, q2, 4, 8));
            if (valid && j == 0) {
                if (len < 8) {
                    res = 0.0;
                    for (int x = 0; x < len; x++) res = __dadd_rn(res, __ldcg(av + x));
                } else {
                    for (int x = lim; x < len; x++) res = __dadd_rn(res, __ldcg(av + x));
                }
                leafSumsG[l] = res;
            }
        }
        cluster_sync_all();
        // ---- the combine tree, level by level, redundantly in every CTA
        for (int l = tid; l < nLeaves; l += CHAIN_THREADS) s_sum[l] = __ldcg(leafSumsG + l);
        __syncthreads();
        double* nsum = s_sum + nLeaves;
        for (int lev = 0; lev < nLevels; lev++) {
            for (int x = levelPtr[lev] + tid; x < levelPtr[lev + 1]; x += CHAIN_THREADS) {
                const int j = levelNodes[x];
                const int l = nodes[j].x, r = nodes[j].y;
                nsum[j] = __dadd_rn(l < 0 ? s_sum[~l] : nsum[l], r < 0 ? s_sum[~r] : nsum[r]);
            }
            __syncthreads();
        }
        const double tot = nNodes ? nsum[nNodes - 1] : s_sum[0];
        // ---- division of this CTA's slice, then every CTA reloads the whole row (or b0 after a reset)
        for (int sp = lo + tid; sp < hi; sp += CHAIN_THREADS) row[sp] = __ldcg(row + sp) / tot;
        cluster_sync_all();
        const bool reset = resets && resets[i];
#pragma unroll 4
        for (int s = tid; s < S; s += CHAIN_THREADS) sb[s] = reset ? b0[s] : __ldcg(row + s);
        __syncthreads();
    }
}

int configure_belief_kernels() {
    PBVI_CUDA(cudaFuncSetAttribute(pairwise_normalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PBVI_CUDA(cudaFuncSetAttribute(belief_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM_MAX));
    PBVI_CUDA(cudaFuncSetAttribute(belief_chain_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM_MAX));
    return PBVI_OK;
}

// one-launch chain when a belief (+ the pairwise-sum scratch) fits in shared memory
static bool chain_fits(const pbvi_model* m) {
    return ((size_t)m->S + m->nLeaves + m->nNodes) * sizeof(double) <= CHAIN_SMEM_MAX && m->O <= 32;
}

// steps given in HOST arrays (observations < 0 / NULL: drawn on the device from h_uniforms); uploads them and launches the chain
static int launch_chain(pbvi_model* m, const double* d_b0, const int32_t* h_actions, const int32_t* h_observations, const double* h_uniforms,
                        const uint8_t* h_resets, int n, double* d_out, int32_t* d_chosen, cudaStream_t st) {
    PBVI_TAKE(dA, int32_t, (size_t)n);
    PBVI_CUDA(cudaMemcpyAsync(dA, h_actions, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    int32_t* dO = nullptr;
    double* dU = nullptr;
    uint8_t* dR = nullptr;
    if (h_observations) {
        dO = m->arena.take<int32_t>((size_t)n);
        if (!dO) return PBVI_ERR_OOM;
        PBVI_CUDA(cudaMemcpyAsync(dO, h_observations, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    if (h_uniforms) {
        dU = m->arena.take<double>((size_t)n);
        if (!dU) return PBVI_ERR_OOM;
        PBVI_CUDA(cudaMemcpyAsync(dU, h_uniforms, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (h_resets) {
        dR = m->arena.take<uint8_t>((size_t)n);
        if (!dR) return PBVI_ERR_OOM;
        PBVI_CUDA(cudaMemcpyAsync(dR, h_resets, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    const size_t smem = ((size_t)m->S + m->nLeaves + m->nNodes) * sizeof(double);
    if (m->chain_mode == 2) {
        PBVI_TAKE(leafSumsG, double, (size_t)m->nLeaves);
        PBVI_TAKE(obsG, double, 32);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CHAIN_CLUSTER);
        cfg.blockDim = dim3(CHAIN_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CHAIN_CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, belief_chain_cluster_kernel, d_b0, (const int32_t*)dA, (const int32_t*)dO, (const double*)dU,
                                                 (const uint8_t*)dR, n, (const int32_t*)m->predPtr, (const int32_t*)m->predK, (const double*)m->rtoK,
                                                 m->S, m->R, m->O, (const int2*)m->pwLeaves, m->nLeaves, (const int2*)m->pwNodes, m->nNodes,
                                                 (const int32_t*)m->pwLevelNodes, (const int32_t*)m->pwLevelPtr, m->nLevels, d_out, d_chosen,
                                                 leafSumsG, obsG);
        if (e == cudaSuccess) {
            m->last_launches++;
            return PBVI_OK;
        }
        cudaGetLastError();                 // no room for an 8-block cluster with this much shared memory (partitioned device): one block
        m->chain_mode = 1;
    }
    belief_chain_kernel<<<1, CHAIN_THREADS, smem, st>>>(d_b0, dA, dO, dU, dR, n, m->predPtr, m->predK, m->rtoK, m->S, m->R, m->O, m->pwLeaves,
                                                        m->nLeaves, m->pwNodes, m->nNodes, m->pwLevelNodes, m->pwLevelPtr, m->nLevels, d_out,
                                                        d_chosen);
    m->last_launches++;
    PBVI_CUDA(cudaGetLastError());
    // the host arrays are pageable: the copies above were staged synchronously by the runtime, so the caller may reuse them on return
    return PBVI_OK;
}
}  // namespace pbvi

namespace pbvi {
constexpr int PAIRWISE_SPLIT_ROWS = 64;
// multi-block pairwise normaliser over n <= PAIRWISE_SPLIT_ROWS rows (leaf sums spread over blocks, then per-slice finish): same
// additions in the same order as pairwise_normalise_kernel
static int pairwise_split_launch(pbvi_model* m, double* rows, int n, int normalise, double* d_norm, cudaStream_t st) {
    const size_t smem = (size_t)(m->nLeaves + m->nNodes) * sizeof(double);
    double* leafSums = m->arena.take<double>((size_t)m->nLeaves * n);
    if (!leafSums) return PBVI_ERR_OOM;
    pairwise_leaf_kernel<<<dim3(ceil_div(m->nLeaves, 32), n), 256, 0, st>>>(rows, m->S, m->pwLeaves, m->nLeaves, leafSums);
    pairwise_finish_kernel<<<dim3(ceil_div(m->S, FINISH_SLICE), n), 256, smem, st>>>(rows, m->S, leafSums, m->nLeaves, m->pwNodes, m->nNodes,
                                                                                      d_norm, normalise);
    m->last_launches += 2;
    return PBVI_OK;
}
}  // namespace pbvi

using namespace pbvi;

// belief_stride: doubles between consecutive source beliefs (S for one belief per (a,o) pair, 0 to update one belief n ways)
static int belief_update_impl(pbvi_model* m, const double* d_beliefs, size_t beliefStride, const int32_t* d_actions,
                              const int32_t* d_observations, int n, int normalise, double* d_out, double* d_norm, cudaStream_t st) {
    const size_t smem = (size_t)(m->nLeaves + m->nNodes) * sizeof(double);
    PBVI_REQUIRE(smem <= 200 * 1024, "state space too large for the pairwise-sum kernel");      // opt-in set by configure_belief_kernels
    for (int i0 = 0; i0 < n; i0 += 65535) {
        const int ni = std::min(65535, n - i0);
        double* out = d_out + (size_t)i0 * m->S;
        belief_project_kernel<<<dim3(ceil_div(m->S, 256), ni), 256, 0, st>>>(d_beliefs + (size_t)i0 * beliefStride, d_actions + i0,
                                                                            d_observations + i0, beliefStride, m->predPtr, m->predK,
                                                                            m->rtoK, m->S, m->R, m->O, out, 0, 0, nullptr, 0.0, nullptr, 0);
        m->last_launches++;
        if ((normalise || d_norm) && ni <= PAIRWISE_SPLIT_ROWS && smem <= 48 * 1024) {
            // a few rows (one step of a walk, a single Belief.update, the successors of one belief): the multi-block form of the
            // normaliser -- with one block per row a handful of SMs would walk ~170 leaves and S divisions each
            PBVI_TRY(pairwise_split_launch(m, out, ni, normalise, d_norm ? d_norm + i0 : nullptr, st));
        } else if (normalise || d_norm) {
            pairwise_normalise_kernel<<<ni, 256, smem, st>>>(out, m->S, m->pwLeaves, m->nLeaves, m->pwNodes, m->nNodes, normalise,
                                                            d_norm ? d_norm + i0 : nullptr);
            m->last_launches++;
        }
    }
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}

extern "C" int pbvi_belief_update(pbvi_model* m, const double* d_beliefs, const int32_t* d_actions, const int32_t* d_observations,
                                  int n, int normalise, double* d_out, double* d_norm, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_beliefs && d_actions && d_observations && d_out, "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    m->last_launches = 0;
    return belief_update_impl(m, d_beliefs, (size_t)m->S, d_actions, d_observations, n, normalise, d_out, d_norm, (cudaStream_t)stream);
}

extern "C" int pbvi_belief_trajectory(pbvi_model* m, const double* d_b0, const int32_t* h_actions, const int32_t* h_observations,
                                      const uint8_t* h_reset, int n, double* d_out, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_b0 && h_actions && h_observations && d_out, "NULL pointer argument");
    for (int i = 0; i < n; i++)
        PBVI_REQUIRE(h_actions[i] >= 0 && h_actions[i] < m->A && h_observations[i] >= 0 && h_observations[i] < m->O,
                     "action / observation index out of range");
    PBVI_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(m->nLeaves + m->nNodes) * sizeof(double);
    PBVI_REQUIRE(smem <= 48 * 1024, "state space too large for the chained pairwise-sum kernel");
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    if (chain_fits(m) && !m->no_chain_kernel) return launch_chain(m, d_b0, h_actions, h_observations, nullptr, h_reset, n, d_out, nullptr, st);
    PBVI_TAKE(leafSums, double, (size_t)m->nLeaves);         // reused by every step (stream order)
    const double* src = d_b0;
    for (int i = 0; i < n; i++) {
        double* out = d_out + (size_t)i * m->S;
        belief_project_kernel<<<dim3(ceil_div(m->S, 256), 1), 256, 0, st>>>(src, nullptr, nullptr, 0, m->predPtr, m->predK, m->rtoK, m->S, m->R,
                                                                           m->O, out, h_actions[i], h_observations[i], nullptr, 0.0, nullptr, 0);
        pairwise_leaf_kernel<<<ceil_div(m->nLeaves, 32), 256, 0, st>>>(out, m->S, m->pwLeaves, m->nLeaves, leafSums);
        pairwise_finish_kernel<<<ceil_div(m->S, FINISH_SLICE), 256, smem, st>>>(out, m->S, leafSums, m->nLeaves, m->pwNodes, m->nNodes, nullptr, 1);
        m->last_launches += 3;
        src = (h_reset && h_reset[i]) ? d_b0 : out;
    }
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}

extern "C" int pbvi_perseus_walk(pbvi_model* m, const double* d_b0, const int32_t* h_actions, const double* h_uniforms, int n,
                                 double* d_out, int32_t* d_observations, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_b0 && h_actions && h_uniforms && d_out, "NULL pointer argument");
    for (int i = 0; i < n; i++) PBVI_REQUIRE(h_actions[i] >= 0 && h_actions[i] < m->A, "action index out of range");
    PBVI_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(m->nLeaves + m->nNodes) * sizeof(double);
    PBVI_REQUIRE(smem <= 48 * 1024, "state space too large for the chained pairwise-sum kernel");
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    if (chain_fits(m) && !m->no_chain_kernel) return launch_chain(m, d_b0, h_actions, nullptr, h_uniforms, nullptr, n, d_out, d_observations, st);
    PBVI_TAKE(leafSums, double, (size_t)m->nLeaves);         // reused by every step (stream order)
    PBVI_TAKE(obsProb, double, (size_t)m->O);
    const size_t K = (size_t)m->S * m->R;
    const double* src = d_b0;
    for (int i = 0; i < n; i++) {
        const int a = h_actions[i];
        double* out = d_out + (size_t)i * m->S;
        // P(o | b, a): the blocks of observation_probability_kernel that belong to action a (same arithmetic as
        // pbvi_observation_probabilities, so the walk equals the step-by-step host-driven one bit for bit)
        observation_probability_kernel<<<dim3(m->O, 1), 256, 0, st>>>(src, m->rtoK + (size_t)a * m->O * K, m->S, m->R, m->O, obsProb);
        belief_project_kernel<<<dim3(ceil_div(m->S, 256), 1), 256, 0, st>>>(src, nullptr, nullptr, 0, m->predPtr, m->predK, m->rtoK, m->S, m->R,
                                                                           m->O, out, a, 0, obsProb, h_uniforms[i],
                                                                           d_observations ? d_observations + i : nullptr, 0);
        pairwise_leaf_kernel<<<ceil_div(m->nLeaves, 32), 256, 0, st>>>(out, m->S, m->pwLeaves, m->nLeaves, leafSums);
        pairwise_finish_kernel<<<ceil_div(m->S, FINISH_SLICE), 256, smem, st>>>(out, m->S, leafSums, m->nLeaves, m->pwNodes, m->nNodes, nullptr, 1);
        m->last_launches += 4;
        src = out;
    }
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}

namespace pbvi {
int belief_successors_impl(pbvi_model* m, const double* d_beliefs, int n, int normalise, double* d_out, double* d_norm, cudaStream_t st) {
    const int nZ = m->nZ;
    const size_t smem = (size_t)(m->nLeaves + m->nNodes) * sizeof(double);
    PBVI_REQUIRE(smem <= 200 * 1024, "state space too large for the pairwise-sum kernel");
    // every successor of a slab of beliefs in ONE projection launch (row i = successor i % nZ of belief i / nZ) and one normaliser
    // launch -- the per-belief host loop of the first version cost two launches per belief
    const int slab = std::max(1, 65535 / nZ);
    for (int i0 = 0; i0 < n; i0 += slab) {
        const int nb = std::min(slab, n - i0), rows = nb * nZ;
        double* out = d_out + (size_t)i0 * nZ * m->S;
        belief_project_kernel<<<dim3(ceil_div(m->S, 256), rows), 256, 0, st>>>(d_beliefs + (size_t)i0 * m->S, nullptr, nullptr, (size_t)m->S,
                                                                              m->predPtr, m->predK, m->rtoK, m->S, m->R, m->O, out, 0, 0,
                                                                              nullptr, 0.0, nullptr, nZ);
        if (rows <= PAIRWISE_SPLIT_ROWS && smem <= 48 * 1024) {
            m->last_launches++;
            PBVI_TRY(pairwise_split_launch(m, out, rows, normalise, d_norm ? d_norm + (size_t)i0 * nZ : nullptr, st));
            continue;
        }
        pairwise_normalise_kernel<<<rows, 256, smem, st>>>(out, m->S, m->pwLeaves, m->nLeaves, m->pwNodes, m->nNodes, normalise,
                                                          d_norm ? d_norm + (size_t)i0 * nZ : nullptr);
        m->last_launches += 2;
    }
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}
}  // namespace pbvi

extern "C" int pbvi_belief_successors(pbvi_model* m, const double* d_beliefs, int n, int normalise, double* d_out, double* d_norm,
                                      void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_beliefs && d_out, "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    m->last_launches = 0;
    return belief_successors_impl(m, d_beliefs, n, normalise, d_out, d_norm, (cudaStream_t)stream);
}

extern "C" int pbvi_observation_probabilities(pbvi_model* m, const double* d_beliefs, int n, double* d_out, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_beliefs && d_out, "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    for (int i0 = 0; i0 < n; i0 += 65535) {
        const int ni = std::min(65535, n - i0);
        observation_probability_kernel<<<dim3(m->nZ, ni), 256, 0, (cudaStream_t)stream>>>(d_beliefs + (size_t)i0 * m->S, m->rtoK, m->S, m->R,
                                                                                        m->nZ, d_out + (size_t)i0 * m->nZ);
        m->last_launches++;
    }
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}
