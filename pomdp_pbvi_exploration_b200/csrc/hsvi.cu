// One level of HSVI's exploration on the device (reference PBVI_Solver.expand_hsvi, src/pomdp.py:1803-1855, and the sawtooth upper bound
// BeliefValueMapping.evaluate, :873-895).  The reference -- and this package's first version -- walks a level with a host loop over
// (a, o): successor, upper bound, Q-value, then the lower bound of the chosen action's successors: six host round trips per level, 3 ms
// each on the olfactory model.  pbvi_hsvi_level enqueues the whole level and returns with ONE synchronisation:
//     all A*O successors of b with their masses P(o|b,a)                 (belief.cu, one projection launch + one normaliser launch)
//     their 128-bit keys and the key of b                                (stored beliefs return their stored value, :884-885)
//     the sawtooth upper bound of every successor over the stored beliefs (support lists built once per expansion, below)
//     max_v alpha_v . successor for every successor                      (the score kernel, plain path)
//     Q(a) = b.Rbar[:,a] + gamma * sum_o P(o|b,a) * upper(a,o); a = first argmax; o = first argmax of P(o|b,a) * (upper - lower)
//     and, when the recursion goes on, the append of (key(b), Q(a)) to the stored keys / values
// Stored beliefs are kept as SUPPORT LISTS (pbvi_support_lists: states and values of the non-zero entries, and b_i . corner), built once
// when a belief enters the upper-bound arrays instead of being re-compacted by every level: a level then touches nnz, not S, per stored belief.
#include <algorithm>

#include "pbvi_common.cuh"

namespace pbvi {

// ---- support lists (ELL layout: row i owns idx / val [i*S, i*S + count[i])) + dot[i] = row_i . corner -------------------------------
//      The dot product is summed exactly like sawtooth_terms_kernel did (per-thread strided fma, shuffle tree, 8 partials in order),
//      so the upper bounds do not change by a bit.
__global__ void __launch_bounds__(256) support_lists_kernel(const double* __restrict__ rows, const double* __restrict__ corner, int S,
                                                            int32_t* __restrict__ idx, double* __restrict__ val, int32_t* __restrict__ count,
                                                            double* __restrict__ dot) {
    __shared__ double sh[8];
    __shared__ int scount;
    const int i = blockIdx.x;
    const double* r = rows + (size_t)i * S;
    if (threadIdx.x == 0) scount = 0;
    __syncthreads();
    double d = 0.0;
    for (int s0 = 0; s0 < S; s0 += 256) {
        const int s = s0 + threadIdx.x;
        const double b = s < S ? r[s] : 0.0;
        if (s < S) d = fma(b, corner[s], d);
        const bool nz = b > 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, nz);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && bal) base = atomicAdd(&scount, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (nz) {
            const int at = base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
            idx[(size_t)i * S + at] = s;
            val[(size_t)i * S + at] = b;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        dot[i] = t;
        count[i] = scount;
    }
}

// terms[q][i] = v0[q] + (ubV[i] - dot[i]) * min_j queries[q][idx_ij] / val_ij ; block per stored belief, Q_TILE queries at a time held
// in registers (the minimum is order-free; the division is the exact one of the reference)
constexpr int SAW_QT = 8;

__global__ void __launch_bounds__(256) sawtooth_lists_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val,
                                                             const int32_t* __restrict__ count, const double* __restrict__ dot,
                                                             const double* __restrict__ ubV, int nUb, const double* __restrict__ queries, int nQ,
                                                             int S, const double* __restrict__ v0, const double* __restrict__ qMass,
                                                             double* __restrict__ terms) {
    __shared__ double sh[8][SAW_QT];
    __shared__ unsigned s_hit;
    const int i = blockIdx.x;
    const int32_t* li = idx + (size_t)i * S;
    const double* lv = val + (size_t)i * S;
    const int cnt = count[i];
    const double scale = ubV[i] - dot[i];
    for (int q0 = 0; q0 < nQ; q0 += SAW_QT) {
        double r[SAW_QT];
        unsigned live = 0u;                       // queries of the tile that exist and (qMass given) have positive mass: NaN rows are skipped
#pragma unroll
        for (int k = 0; k < SAW_QT; k++) {
            r[k] = INFINITY;
            if (q0 + k < nQ && (!qMass || qMass[q0 + k] > 0.0)) live |= 1u << k;
        }
        // Queries and stored beliefs are non-negative, so a ratio of 0 is the floor of the minimum: once every query of the tile has
        // hit a state of the support where it is zero -- the rule: two beliefs of a 22 021-state model rarely nest -- the rest of the list
        // cannot change the result.  (A NaN query, the 0/0 successor of an impossible observation, never reaches 0 and walks the whole
        // list unless the caller passes the masses, which mark it as not live.)  s_hit collects, per query of the tile, "some thread holds a 0".
        if (threadIdx.x == 0) s_hit = 0u;
        __syncthreads();
        const unsigned full = (1u << SAW_QT) - 1u;
        for (int j0 = 0; j0 < cnt; j0 += 256 * 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u * 256 + threadIdx.x;
                if (j < cnt) {
                    const int s = li[j];
                    const double b = lv[j];
#pragma unroll
                    for (int k = 0; k < SAW_QT; k++)
                        if ((live >> k) & 1u) r[k] = fmin(r[k], queries[(size_t)(q0 + k) * S + s] / b);
                }
            }
            unsigned hit = ~live & full;
#pragma unroll
            for (int k = 0; k < SAW_QT; k++) hit |= (r[k] == 0.0) ? (1u << k) : 0u;
            hit = __reduce_or_sync(0xffffffffu, hit);
            if ((threadIdx.x & 31) == 0 && hit) atomicOr(&s_hit, hit);
            __syncthreads();
            const bool done = s_hit == full;
            __syncthreads();
            if (done) break;
        }
#pragma unroll
        for (int k = 0; k < SAW_QT; k++) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) r[k] = fmin(r[k], __shfl_down_sync(0xffffffffu, r[k], off));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < SAW_QT; k++) sh[threadIdx.x >> 5][k] = r[k];
        }
        __syncthreads();
        if (threadIdx.x < SAW_QT && q0 + threadIdx.x < nQ) {
            double m = INFINITY;
            for (int w = 0; w < 8; w++) m = fmin(m, sh[w][threadIdx.x]);
            const int q = q0 + threadIdx.x;
            terms[(size_t)q * (nUb + 1) + i] = v0[q] + scale * m;
        }
    }
}

// v0[q] = queries[q] . corner (same reduction shape as sawtooth_v0_kernel); also terms[q][nUb] = v0[q]
__global__ void __launch_bounds__(256) sawtooth_v0_terms_kernel(const double* __restrict__ corner, const double* __restrict__ queries, int S, int nUb,
                                                                double* __restrict__ v0, double* __restrict__ terms) {
    __shared__ double sh[8];
    const double* q = queries + (size_t)blockIdx.x * S;
    double a = 0.0;
    for (int s = threadIdx.x; s < S; s += 256) a = fma(q[s], corner[s], a);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        v0[blockIdx.x] = t;
        terms[(size_t)blockIdx.x * (nUb + 1) + nUb] = t;
    }
}

__global__ void __launch_bounds__(128) row_min_lists_kernel(const double* __restrict__ terms, int n, int width, double* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n) return;
    double best = INFINITY;
    for (int i = 0; i < width; i++) best = fmin(best, terms[(size_t)qi * width + i]);
    out[qi] = best;
}

// rb[a] = b . Rbar[:,a] over the non-zero rewards (warp per action)
__global__ void __launch_bounds__(256) reward_dot_kernel(const double* __restrict__ b, const int32_t* __restrict__ nzPtr, const int32_t* __restrict__ nzIdx,
                                                         const double* __restrict__ nzVal, int A, double* __restrict__ rb) {
    const int a = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= A) return;
    double part = 0.0;
    for (int j = nzPtr[a] + lane; j < nzPtr[a + 1]; j += 32) part = fma(b[nzIdx[j]], nzVal[j], part);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
    if (lane == 0) rb[a] = part;
}

// scores[j][v] = rows[j] . alphas[v] for a FEW rows (the A*O successors of one belief) against all alphas, straight from the row-major
// alphas: block = 8 alphas x 8 rows, thread t owns the states t, t + 256, ... (coalesced loads of both operands), 64 accumulators, fixed
// reduction tree.  The score kernel would run this shape on two SMs (one z, one belief tile: a tile walks its ~1400 pipeline stages
// alone, 0.7 ms); here V / 8 * ceil(n / 8) blocks share the work.  few_rows_max_kernel then takes max_v per row (np.max semantics: NaN wins).
constexpr int FR_A = 8, FR_B = 8;

__global__ void __launch_bounds__(256) few_rows_scores_kernel(const double* __restrict__ rows, int n, const double* __restrict__ alphas, int nV, int S,
                                                              double* __restrict__ scores) {
    __shared__ double sh[8][FR_A * FR_B];
    const int v0 = blockIdx.x * FR_A, j0 = blockIdx.y * FR_B;
    double acc[FR_A][FR_B];
#pragma unroll
    for (int a = 0; a < FR_A; a++)
#pragma unroll
        for (int b = 0; b < FR_B; b++) acc[a][b] = 0.0;
    for (int s = threadIdx.x; s < S; s += 256) {
        double av[FR_A], bv[FR_B];
#pragma unroll
        for (int a = 0; a < FR_A; a++) av[a] = (v0 + a < nV) ? alphas[(size_t)(v0 + a) * S + s] : 0.0;
#pragma unroll
        for (int b = 0; b < FR_B; b++) bv[b] = (j0 + b < n) ? rows[(size_t)(j0 + b) * S + s] : 0.0;
#pragma unroll
        for (int a = 0; a < FR_A; a++)
#pragma unroll
            for (int b = 0; b < FR_B; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < FR_A; a++)
#pragma unroll
        for (int b = 0; b < FR_B; b++) {
            double x = acc[a][b];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][a * FR_B + b] = x;
        }
    __syncthreads();
    if (threadIdx.x < FR_A * FR_B) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w][threadIdx.x];
        const int a = threadIdx.x / FR_B, b = threadIdx.x % FR_B;
        if (v0 + a < nV && j0 + b < n) scores[(size_t)(j0 + b) * nV + v0 + a] = t;
    }
}

__global__ void __launch_bounds__(256) few_rows_max_kernel(const double* __restrict__ scores, int nV, double* __restrict__ out) {
    __shared__ double sh[8];
    __shared__ int snan;
    const double* r = scores + (size_t)blockIdx.x * nV;
    if (threadIdx.x == 0) snan = 0;
    __syncthreads();
    double m = -INFINITY;
    for (int v = threadIdx.x; v < nV; v += 256) {
        const double x = r[v];
        if (x != x) snan = 1;
        m = fmax(m, x);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = -INFINITY;
        for (int w = 0; w < 8; w++) t = fmax(t, sh[w]);
        out[blockIdx.x] = snan ? __longlong_as_double(0x7ff8000000000000ll) : t;
    }
}

struct HsviOut {
    double best_a, best_o, max_qv, best_v_diff;
    long long added, key0, key1, n_possible;
};

// The decisions of one level (one block).  Stored beliefs return their stored value (keys of the successors against the stored keys);
// Q-values and the observation choice in the reference's order and arithmetic (strict >, so the first maximiser wins; observations of
// probability zero are skipped: their successor is 0/0); append of (key(b), Q) when the recursion continues and b is not stored yet.
__global__ void __launch_bounds__(256) hsvi_choose_kernel(const double* __restrict__ mass, const double* __restrict__ upperSaw, const double* __restrict__ lower,
                                                          const unsigned long long* __restrict__ succKeys, const unsigned long long* __restrict__ bKey,
                                                          const double* __restrict__ rb, unsigned long long* __restrict__ storedKeys,
                                                          double* __restrict__ storedVals, int nStored, int storedCap, int A, int O, double gamma,
                                                          double convTerm, int mayContinue, HsviOut* __restrict__ out) {
    extern __shared__ double s_upper[];      // [A*O]
    __shared__ int s_found;
    const int nZ = A * O;
    for (int z = threadIdx.x; z < nZ; z += 256) s_upper[z] = upperSaw[z];
    if (threadIdx.x == 0) s_found = 0;
    __syncthreads();
    const unsigned long long b0 = bKey[0], b1 = bKey[1];
    for (int i = threadIdx.x; i < nStored; i += 256) {
        const unsigned long long k0 = storedKeys[(size_t)i * 2], k1 = storedKeys[(size_t)i * 2 + 1];
        if (k0 == b0 && k1 == b1) s_found = 1;
        for (int z = 0; z < nZ; z++)
            if (succKeys[(size_t)z * 2] == k0 && succKeys[(size_t)z * 2 + 1] == k1) s_upper[z] = storedVals[i];     // keys are unique: one writer
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double maxQ = -INFINITY;
    int bestA = -1, nPossible = 0;
    for (int a = 0; a < A; a++) {
        double acc = 0.0;
        for (int o = 0; o < O; o++) {
            const double p = mass[a * O + o];
            if (p > 0.0) { acc = __dadd_rn(acc, __dmul_rn(p, s_upper[a * O + o])); nPossible++; }
        }
        const double q = __dadd_rn(rb[a], __dmul_rn(gamma, acc));
        if (q > maxQ) { maxQ = q; bestA = a; }
    }
    double maxOVal = -INFINITY, bestDiff = -INFINITY;
    int bestO = -1;
    if (bestA >= 0)
        for (int o = 0; o < O; o++) {
            const double p = mass[bestA * O + o];
            if (!(p > 0.0)) continue;
            const double diff = __dadd_rn(s_upper[bestA * O + o], -lower[bestA * O + o]);
            const double ov = __dmul_rn(p, diff);
            if (ov > maxOVal) { maxOVal = ov; bestDiff = diff; bestO = o; }
        }
    int added = 0;
    if (mayContinue && !(bestDiff < convTerm) && !s_found && nStored < storedCap) {
        storedKeys[(size_t)nStored * 2] = b0;
        storedKeys[(size_t)nStored * 2 + 1] = b1;
        storedVals[nStored] = maxQ;
        added = 1;
    }
    out->best_a = bestA; out->best_o = bestO; out->max_qv = maxQ; out->best_v_diff = bestDiff;
    out->added = added; out->key0 = (long long)b0; out->key1 = (long long)b1; out->n_possible = nPossible;
}

// next[s] = the chosen successor (b itself when no observation is possible), picked with the indices the choose kernel left on the device
__global__ void __launch_bounds__(256) hsvi_take_next_kernel(const double* __restrict__ succ, const double* __restrict__ b, const HsviOut* __restrict__ res,
                                                             int O, int S, double* __restrict__ next) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= S) return;
    const int a = (int)res->best_a, o = (int)res->best_o;
    next[s] = (a >= 0 && o >= 0) ? succ[((size_t)a * O + o) * S + s] : b[s];
}

}  // namespace pbvi

using namespace pbvi;

extern "C" int pbvi_support_lists(pbvi_model* m, const double* d_rows, int n, const double* d_corner, int32_t* d_idx, double* d_val,
                                  int32_t* d_count, double* d_dot, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n >= 0, "n must be non-negative");
    if (n == 0) return PBVI_OK;
    PBVI_REQUIRE(d_rows && d_corner && d_idx && d_val && d_count && d_dot, "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    support_lists_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(d_rows, d_corner, m->S, d_idx, d_val, d_count, d_dot);
    m->last_launches = 1;
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}

namespace pbvi {
// upper bounds of n_q query rows over the stored beliefs given as support lists; scratch from the arena (no reset here)
static int sawtooth_lists_impl(pbvi_model* m, const double* d_corner, const int32_t* d_idx, const double* d_val, const int32_t* d_count,
                               const double* d_dot, const double* d_ub_values, int n_ub, const double* d_queries, int n_q,
                               const double* d_query_mass, double* d_out, cudaStream_t st) {
    PBVI_TAKE(terms, double, (size_t)n_q * (n_ub + 1));
    PBVI_TAKE(v0, double, (size_t)n_q);
    sawtooth_v0_terms_kernel<<<n_q, 256, 0, st>>>(d_corner, d_queries, m->S, n_ub, v0, terms);
    if (n_ub > 0)
        sawtooth_lists_kernel<<<n_ub, 256, 0, st>>>(d_idx, d_val, d_count, d_dot, d_ub_values, n_ub, d_queries, n_q, m->S, v0, d_query_mass, terms);
    row_min_lists_kernel<<<ceil_div(n_q, 128), 128, 0, st>>>(terms, n_q, n_ub + 1, d_out);
    m->last_launches += 3;
    PBVI_CUDA(cudaGetLastError());
    return PBVI_OK;
}
}  // namespace pbvi

extern "C" int pbvi_sawtooth_lists(pbvi_model* m, const double* d_corner, const int32_t* d_idx, const double* d_val, const int32_t* d_count,
                                   const double* d_dot, const double* d_ub_values, int n_ub, const double* d_queries, int n_q, double* d_out,
                                   void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(n_ub >= 0 && n_q >= 0, "counts must be non-negative");
    if (n_q == 0) return PBVI_OK;
    PBVI_REQUIRE(d_corner && d_queries && d_out && (n_ub == 0 || (d_idx && d_val && d_count && d_dot && d_ub_values)), "NULL pointer argument");
    PBVI_CUDA(cudaSetDevice(m->device));
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    m->last_launches = 0;
    return sawtooth_lists_impl(m, d_corner, d_idx, d_val, d_count, d_dot, d_ub_values, n_ub, d_queries, n_q, nullptr, d_out, (cudaStream_t)stream);
}

extern "C" int pbvi_hsvi_level(pbvi_model* m, const double* d_b, const double* d_alphas, int nV, double gamma, const double* d_corner,
                               const int32_t* d_idx, const double* d_val, const int32_t* d_count, const double* d_dot,
                               const double* d_ub_values, int n_ub, uint64_t* d_stored_keys, double* d_stored_vals, int n_stored,
                               int stored_capacity, double conv_term, int may_continue, double* d_next, double* d_succ, double* d_mass,
                               double* h_out8, void* stream) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_REQUIRE(d_b && d_alphas && d_corner && h_out8, "NULL pointer argument");
    PBVI_REQUIRE(nV > 0 && n_ub >= 0 && n_stored >= 0 && stored_capacity >= n_stored, "bad counts");
    PBVI_REQUIRE(n_ub == 0 || (d_idx && d_val && d_count && d_dot && d_ub_values), "support lists are required when n_ub > 0");
    PBVI_REQUIRE(n_stored == 0 || stored_capacity == 0 || (d_stored_keys && d_stored_vals), "stored key / value arrays are required");
    PBVI_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    PBVI_TRY(enter_call(m, (cudaStream_t)stream));
    m->last_launches = 0;
    const int nZ = m->nZ, S = m->S;
    PBVI_TAKE(keys, unsigned long long, (size_t)(nZ + 1) * 2);          // successors, then b itself
    PBVI_TAKE(upper, double, (size_t)nZ);
    PBVI_TAKE(lower, double, (size_t)nZ);
    PBVI_TAKE(rb, double, (size_t)m->A);
    PBVI_TAKE(outDev, HsviOut, 1);
    if (!d_succ) { d_succ = m->arena.take<double>((size_t)nZ * S); if (!d_succ) return PBVI_ERR_OOM; }     // the caller does not want the block
    if (!d_mass) { d_mass = m->arena.take<double>((size_t)nZ); if (!d_mass) return PBVI_ERR_OOM; }
    PBVI_TRY(belief_successors_impl(m, d_b, 1, 1, d_succ, d_mass, st));
    PBVI_TRY(row_hash_launch(m, d_succ, nZ, S, reinterpret_cast<uint64_t*>(keys), st));
    PBVI_TRY(row_hash_launch(m, d_b, 1, S, reinterpret_cast<uint64_t*>(keys + (size_t)nZ * 2), st));
    PBVI_TRY(sawtooth_lists_impl(m, d_corner, d_idx, d_val, d_count, d_dot, d_ub_values, n_ub, d_succ, nZ, d_mass, upper, st));
    reward_dot_kernel<<<ceil_div(m->A * 32, 256), 256, 0, st>>>(d_b, m->rbarNzPtr, m->rbarNzIdx, m->rbarNzVal, m->A, rb);
    m->last_launches++;
    PBVI_TAKE(scores, double, (size_t)nZ * nV);
    few_rows_scores_kernel<<<dim3(ceil_div(nV, FR_A), ceil_div(nZ, FR_B)), 256, 0, st>>>(d_succ, nZ, d_alphas, nV, S, scores);
    few_rows_max_kernel<<<nZ, 256, 0, st>>>(scores, nV, lower);
    m->last_launches += 2;
    hsvi_choose_kernel<<<1, 256, (size_t)nZ * sizeof(double), st>>>(d_mass, upper, lower, keys, keys + (size_t)nZ * 2, rb,
                                                                    reinterpret_cast<unsigned long long*>(d_stored_keys), d_stored_vals, n_stored,
                                                                    stored_capacity, m->A, m->O, gamma, conv_term, may_continue, outDev);
    m->last_launches++;
    if (d_next) {
        hsvi_take_next_kernel<<<ceil_div(S, 256), 256, 0, st>>>(d_succ, d_b, outDev, m->O, S, d_next);
        m->last_launches++;
    }
    PBVI_CUDA(cudaGetLastError());
    PBVI_CUDA(cudaMemcpyAsync(h_out8, outDev, sizeof(HsviOut), cudaMemcpyDeviceToHost, st));
    PBVI_CUDA(cudaStreamSynchronize(st));
    return PBVI_OK;
}
