// Model handle of the PBVI B200 engine: uploads the reference's model tensors re-laid out action-major
// (replaces Model.gpu_model, reference src/mdp.py:533-560), owns the scratch arena and the error string.
#include <algorithm>
#include <cstdarg>
#include <numeric>

#include "pbvi_common.cuh"

namespace pbvi {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

void* Arena::take_bytes(size_t bytes) {
    bytes = padded(bytes == 0 ? 1 : bytes);
    if (!chunks.empty()) {
        Chunk& c = chunks.back();
        if (c.off + bytes <= c.cap) {
            void* p = c.base + c.off;
            c.off += bytes;
            return p;
        }
    }
    // geometric growth: a new chunk is at least TWICE the arena so far, so that it alone can hold the whole next call of a solve whose
    // value function grows a little every iteration (alphaT, partial argmax buffers): the arena re-allocates O(log) times, and the
    // consolidation below usually only frees the outgrown chunks instead of allocating again
    size_t total = 0;
    for (auto& c : chunks) total += c.cap;
    size_t cap = std::max(bytes, std::max(2 * total, size_t(64) << 20));
    char* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        cudaGetLastError();
        cap = std::max(bytes, total);
        if (cudaMalloc(&p, cap) != cudaSuccess) {
            cudaGetLastError();
            if (cap == bytes || cudaMalloc(&p, bytes) != cudaSuccess) {
                cudaGetLastError();
                set_error("out of device memory: scratch arena could not grow by %zu bytes", bytes);
                return nullptr;
            }
            cap = bytes;
        }
    }
    chunks.push_back({p, cap, bytes});
    return p;
}

// Start of an API call: rewind; if the previous call had to grow, consolidate into one chunk (cudaFree synchronises
// with the work that may still be using the old chunks).
void Arena::reset() {
    if (chunks.size() > 1) {
        size_t total = 0, used = 0, best = 0;
        for (size_t i = 0; i < chunks.size(); i++) {
            total += chunks[i].cap;
            used += padded(chunks[i].off);
            if (chunks[i].cap > chunks[best].cap) best = i;
        }
        if (chunks[best].cap >= used + (used >> 3)) {
            // the largest chunk alone holds what the previous call needed (with 12 % to spare): free the others, allocate nothing
            Chunk keep = chunks[best];
            for (size_t i = 0; i < chunks.size(); i++)
                if (i != best) cudaFree(chunks[i].base);
            chunks.clear();
            chunks.push_back(keep);
        } else {
            release();
            char* p = nullptr;
            if (cudaMalloc(&p, total) == cudaSuccess) chunks.push_back({p, total, 0});
            else cudaGetLastError();
        }
    }
    for (auto& c : chunks) c.off = 0;
}

void Arena::release() {
    for (auto& c : chunks) cudaFree(c.base);
    chunks.clear();
}

// The combine tree of NumPy's float64 pairwise summation over n contiguous elements
// (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum): blocks of <= 128 elements are leaves,
// larger ranges split at n/2 rounded down to a multiple of 8.  Returns the node/leaf reference of the range.
static int build_pairwise(int off, int n, std::vector<int2>& leaves, std::vector<int2>& nodes) {
    if (n <= 128) {
        leaves.push_back(make_int2(off, n));
        return ~(int)(leaves.size() - 1);
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int l = build_pairwise(off, n2, leaves, nodes);
    const int r = build_pairwise(off + n2, n - n2, leaves, nodes);
    nodes.push_back(make_int2(l, r));
    return (int)nodes.size() - 1;
}

int enter_call(pbvi_model* m, cudaStream_t st) {
    if (m->has_last_stream && m->last_stream != st) {
        if (!m->evGuard) PBVI_CUDA(cudaEventCreateWithFlags(&m->evGuard, cudaEventDisableTiming));
        PBVI_CUDA(cudaEventRecord(m->evGuard, m->last_stream));
        PBVI_CUDA(cudaStreamWaitEvent(st, m->evGuard, 0));
    }
    m->last_stream = st;
    m->has_last_stream = true;
    m->arena.reset();
    return PBVI_OK;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
    PBVI_CUDA(cudaMalloc(dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
    if (!src.empty()) PBVI_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return PBVI_OK;
}

}  // namespace pbvi

using namespace pbvi;

extern "C" int pbvi_version(void) { return 100; }

extern "C" const char* pbvi_last_error(void) { return g_last_error.c_str(); }

extern "C" int pbvi_model_create(int S, int A, int O, int R, const int64_t* h_reach, const double* h_probs,
                                 const double* h_rto, const double* h_rbar, int device, pbvi_model** out) {
    PBVI_REQUIRE(out != nullptr, "out handle pointer is NULL");
    *out = nullptr;
    PBVI_REQUIRE(S > 0 && A > 0 && O > 0 && R > 0, "S, A, O, R must be positive");
    PBVI_REQUIRE(h_reach && h_rto && h_rbar, "reach / rto / rbar tables are required");
    PBVI_REQUIRE((long long)S * R < (1ll << 31) && (long long)S < (1ll << 24) * KC, "model too large for int32 indexing");
    int ndev = 0;
    PBVI_CUDA(cudaGetDeviceCount(&ndev));
    PBVI_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    PBVI_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PBVI_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return PBVI_ERR_UNSUPPORTED;
    }

    PBVI_TRY(configure_backup_kernels());
    PBVI_TRY(configure_belief_kernels());
    PBVI_TRY(configure_misc_kernels());

    pbvi_model* m = new pbvi_model();
    m->S = S; m->A = A; m->O = O; m->R = R;
    m->K = S * R;
    m->Sp = ceil_div(S, 16) * 16;          // whole pipeline stages (SUB * KC <= 16 states); pad states land on the zero row S with RTO 0
    m->nChunks = m->Sp / KC;
    m->nZ = A * O;
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->has_probs = h_probs != nullptr;
    const int K = m->K, Sp = m->Sp, nC = m->nChunks, nZ = m->nZ;

    for (size_t i = 0; i < (size_t)S * A * R; i++) {
        if (h_reach[i] < 0 || h_reach[i] >= S) {
            delete m;
            set_error("bad argument: reachable_states[%zu] = %lld is outside [0, %d)", i, (long long)h_reach[i], S);
            return PBVI_ERR_BAD_ARG;
        }
    }

    // ---- action-major re-layout
    std::vector<int32_t> reachK((size_t)A * K);
    std::vector<double> rtoK((size_t)A * O * K), probK, rbarT((size_t)A * S);
    if (h_probs) probK.resize((size_t)A * K);
    for (int s = 0; s < S; s++)
        for (int a = 0; a < A; a++) {
            rbarT[(size_t)a * S + s] = h_rbar[(size_t)s * A + a];
            for (int r = 0; r < R; r++) {
                const size_t src = ((size_t)s * A + a) * R + r, k = (size_t)s * R + r;
                reachK[(size_t)a * K + k] = (int32_t)h_reach[src];
                if (h_probs) probK[(size_t)a * K + k] = h_probs[src];
                for (int o = 0; o < O; o++)
                    rtoK[((size_t)a * O + o) * K + k] = h_rto[(((size_t)s * A + a) * O + o) * R + r];
            }
        }

    m->model_nonneg = true;
    for (double v : rtoK) if (!(v >= 0.0)) { m->model_nonneg = false; break; }
    for (double v : rbarT) if (!(v >= 0.0)) { m->model_nonneg = false; break; }

    // ---- CSR of the non-zero expected rewards
    std::vector<int32_t> rbarNzPtr(A + 1, 0), rbarNzIdx;
    std::vector<double> rbarNzVal;
    for (int a = 0; a < A; a++) {
        for (int s = 0; s < S; s++) {
            const double v = rbarT[(size_t)a * S + s];
            if (v != 0.0) { rbarNzIdx.push_back(s); rbarNzVal.push_back(v); }
        }
        rbarNzPtr[a + 1] = (int32_t)rbarNzIdx.size();
    }

    // ---- live-chunk masks per z = (a,o) and the heavy-first order
    std::vector<uint8_t> zMask((size_t)nZ * nC, 0);
    std::vector<long long> live(nZ, 0);
    for (int z = 0; z < nZ; z++)
        for (int s = 0; s < S; s++) {
            bool nz = false;
            for (int r = 0; r < R && !nz; r++) nz = rtoK[(size_t)z * K + (size_t)s * R + r] != 0.0;
            if (nz && !zMask[(size_t)z * nC + s / KC]) { zMask[(size_t)z * nC + s / KC] = 1; live[z]++; }
        }
    std::vector<int32_t> zOrder(nZ);
    std::iota(zOrder.begin(), zOrder.end(), 0);
    std::stable_sort(zOrder.begin(), zOrder.end(), [&](int x, int y) { return live[x] > live[y]; });

    // ---- chunk-padded tables of the R == 1 gather path
    std::vector<int32_t> reachP;
    std::vector<double> rtoP;
    if (R == 1) {
        reachP.assign((size_t)A * Sp, S);           // pad states gather row S of alphaT, which the select call keeps all-zero
        rtoP.assign((size_t)nZ * Sp, 0.0);
        for (int a = 0; a < A; a++) std::copy_n(&reachK[(size_t)a * K], S, &reachP[(size_t)a * Sp]);
        for (int z = 0; z < nZ; z++) std::copy_n(&rtoK[(size_t)z * K], S, &rtoP[(size_t)z * Sp]);
    }

    // ---- CSR over landing states, sources in ascending k (the accumulation order of np.bincount)
    std::vector<int32_t> predPtr((size_t)A * (S + 1), 0), predK((size_t)A * K);
    for (int a = 0; a < A; a++) {
        int32_t* ptr = &predPtr[(size_t)a * (S + 1)];
        for (int k = 0; k < K; k++) ptr[reachK[(size_t)a * K + k] + 1]++;
        for (int s = 0; s < S; s++) ptr[s + 1] += ptr[s];
        std::vector<int32_t> fill(ptr, ptr + S);
        for (int k = 0; k < K; k++) predK[(size_t)a * K + fill[reachK[(size_t)a * K + k]]++] = k;
    }

    std::vector<int2> leaves, nodes;
    build_pairwise(0, S, leaves, nodes);
    m->nLeaves = (int)leaves.size();
    m->nNodes = (int)nodes.size();

    int rc = PBVI_OK;
    auto up = [&](auto** dst, const auto& v) { if (rc == PBVI_OK) rc = upload(dst, v); };
    up(&m->reachK, reachK); up(&m->rtoK, rtoK); up(&m->rbarT, rbarT);
    up(&m->rbarNzPtr, rbarNzPtr); up(&m->rbarNzIdx, rbarNzIdx); up(&m->rbarNzVal, rbarNzVal);
    if (h_probs) up(&m->probK, probK);
    if (R == 1) { up(&m->reachP, reachP); up(&m->rtoP, rtoP); }
    up(&m->zMask, zMask); up(&m->zOrder, zOrder);
    up(&m->predPtr, predPtr); up(&m->predK, predK);
    up(&m->pwLeaves, leaves); up(&m->pwNodes, nodes);
    {   // the combine tree by levels (a node's level = 1 + the deeper child's; leaves are level 0): nodes of one level are independent
        std::vector<int> level(nodes.size(), 0);
        int maxLevel = 0;
        for (size_t j = 0; j < nodes.size(); j++) {      // post-order: children precede their parent
            const int l = nodes[j].x < 0 ? 0 : level[nodes[j].x], r = nodes[j].y < 0 ? 0 : level[nodes[j].y];
            level[j] = 1 + std::max(l, r);
            maxLevel = std::max(maxLevel, level[j]);
        }
        std::vector<int32_t> levelNodes, levelPtr(1, 0);
        for (int lev = 1; lev <= maxLevel; lev++) {
            for (size_t j = 0; j < nodes.size(); j++) if (level[j] == lev) levelNodes.push_back((int32_t)j);
            levelPtr.push_back((int32_t)levelNodes.size());
        }
        m->nLevels = maxLevel;
        up(&m->pwLevelNodes, levelNodes); up(&m->pwLevelPtr, levelPtr);
    }
    std::vector<uint4> hashKeys((size_t)S);
    for (int s = 0; s < S; s++) hashKeys[s] = row_key_words(s);
    up(&m->hashKeys, hashKeys);
    if (rc == PBVI_OK && cudaMalloc(&m->d_signs, 8 * sizeof(int)) != cudaSuccess) {
        set_error("cudaMalloc(signs) failed");
        rc = PBVI_ERR_CUDA;
    }
    if (rc == PBVI_OK && cudaMalloc(&m->d_stats, sizeof(unsigned long long)) != cudaSuccess) {
        set_error("cudaMalloc(stats) failed");
        rc = PBVI_ERR_CUDA;
    }
    if (rc != PBVI_OK) { pbvi_model_destroy(m); return rc; }
    *out = m;
    return PBVI_OK;
}

extern "C" int pbvi_model_destroy(pbvi_model* m) {
    if (!m) return PBVI_OK;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    cudaFree(m->reachK); cudaFree(m->rtoK); cudaFree(m->probK); cudaFree(m->rbarT);
    cudaFree(m->rbarNzPtr); cudaFree(m->rbarNzIdx); cudaFree(m->rbarNzVal);
    cudaFree(m->reachP); cudaFree(m->rtoP); cudaFree(m->zMask); cudaFree(m->zOrder);
    cudaFree(m->predPtr); cudaFree(m->predK); cudaFree(m->pwLeaves); cudaFree(m->pwNodes); cudaFree(m->hashKeys);
    cudaFree(m->pwLevelNodes); cudaFree(m->pwLevelPtr);
    cudaFree(m->d_stats);
    cudaFree(m->d_signs);
    if (m->evScore0) { cudaEventDestroy(m->evScore0); cudaEventDestroy(m->evScore1); }
    if (m->evGuard) cudaEventDestroy(m->evGuard);
    if (m->hostIn) cudaStreamDestroy(m->hostIn);
    if (m->hostOut) cudaStreamDestroy(m->hostOut);
    for (int i = 0; i < 2; i++) {
        if (m->evIn[i]) cudaEventDestroy(m->evIn[i]);
        if (m->evDone[i]) cudaEventDestroy(m->evDone[i]);
        if (m->evOut[i]) cudaEventDestroy(m->evOut[i]);
    }
    m->arena.release();
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_pack) cudaFreeHost(m->h_pack);
    if (m->h_io) cudaFreeHost(m->h_io);
    delete m;
    return PBVI_OK;
}

extern "C" int pbvi_model_dims(const pbvi_model* m, int* S, int* A, int* O, int* R) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    if (S) *S = m->S;
    if (A) *A = m->A;
    if (O) *O = m->O;
    if (R) *R = m->R;
    return PBVI_OK;
}

extern "C" int pbvi_last_launches(const pbvi_model* m) { return m ? m->last_launches : 0; }

extern "C" int pbvi_set_option(pbvi_model* m, const char* name, int value) {
    PBVI_REQUIRE(m != nullptr && name != nullptr, "NULL argument");
    if (std::strcmp(name, "chain_kernel") == 0) { m->no_chain_kernel = value == 0; m->chain_mode = value == 1 ? 1 : 2; return PBVI_OK; }
    set_error("bad argument: unknown option '%s'", name);
    return PBVI_ERR_BAD_ARG;
}

extern "C" int pbvi_set_profiling(pbvi_model* m, int enable) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    PBVI_CUDA(cudaSetDevice(m->device));
    if (enable && !m->evScore0) {
        PBVI_CUDA(cudaEventCreate(&m->evScore0));
        PBVI_CUDA(cudaEventCreate(&m->evScore1));
    }
    m->profile = enable != 0;
    m->score_timed = false;
    return PBVI_OK;
}

extern "C" int pbvi_last_score_ms(pbvi_model* m, float* ms) {
    PBVI_REQUIRE(m != nullptr && ms != nullptr, "NULL argument");
    PBVI_REQUIRE(m->profile && m->score_timed, "profiling is off or no score kernel has run since it was enabled");
    PBVI_CUDA(cudaEventSynchronize(m->evScore1));
    PBVI_CUDA(cudaEventElapsedTime(ms, m->evScore0, m->evScore1));
    return PBVI_OK;
}

extern "C" int pbvi_last_stats(const pbvi_model* m, double* executed_flops, double* dense_flops, int* launches) {
    PBVI_REQUIRE(m != nullptr, "model handle is NULL");
    unsigned long long visited = 0;
    PBVI_CUDA(cudaMemcpy(&visited, m->d_stats, sizeof(visited), cudaMemcpyDeviceToHost));
    if (executed_flops) *executed_flops = (double)visited * m->last_exec_scale;
    if (dense_flops) *dense_flops = m->last_dense_flops;
    if (launches) *launches = m->last_launches;
    return PBVI_OK;
}
