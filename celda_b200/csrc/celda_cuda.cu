// libcelda_cuda: C ABI (include/celda_cuda.h) over the sm_100a kernels.  Host side of the drop-in boundary.
#include "../../include/celda_cuda.h"

#include <algorithm>
#include <cmath>
#include <memory>

#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a profiler injects itself

#include "common.cuh"
#include "kernels.cuh"
#include "sweeps.cuh"

using namespace celda;

namespace {

int g_device = 0;

inline int grid_for(long long n, int threads, int sms, int per_sm = 8) {
    long long b = (n + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(b, (long long)sms * per_sm));
}

int check_labels_host(const int* g, long long n, int K, const char* what, const char* kname) {
    for (long long i = 0; i < n; ++i)
        if (g[i] < 1 || g[i] > K) return fail("The entries in '%s' need to be between 1 and '%s'.", what, kname);
    return 0;
}

}  // namespace

// NCCL is resolved at run time (dlopen by SONAME): a host process that already carries an NCCL (e.g. a Python
// process that imported torch) shares that copy, a plain R process loads the system one.  No link-time dependency.
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int load() {
        if (so) return 0;
        so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!so) return fail("cannot load libnccl.so.2: %s", dlerror());
        GetUniqueId = (decltype(GetUniqueId))dlsym(so, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(so, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(so, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(so, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(so, "ncclGetErrorString");
        GroupStart = (decltype(GroupStart))dlsym(so, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(so, "ncclGroupEnd");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString || !GroupStart || !GroupEnd)
            return fail("libnccl.so.2 lacks an expected symbol");
        return 0;
    }
};
static NcclApi g_nccl;
#define CELDA_NCCL_OK(expr)                                                                              \
    do {                                                                                                 \
        ncclResult_t r__ = (expr);                                                                       \
        if (r__ != ncclSuccess) return celda::fail("NCCL error %s at %s:%d", g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

struct celda_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
};

// Work space of the sweep kernels (ring buffers, speculation flags, barrier, status, statistics).
struct SweepBufs {
    DevBuf<double> bufP, bufA, colsum, lgnCP, drawMx;
    DevBuf<int> spec, status, drawNz;
    DevBuf<unsigned> barrier;
    DevBuf<long long> stats;  // [which][8]
    // value tables of the z sweep (k_sweep_z): row histogram, per-row offsets, per-CTA global tables
    DevBuf<int> ztHist, ztOffg, ztGrp, ztTmp, ztXT, ztZT, ztNT;
    DevBuf<double> ztVal;  // [K][ZT_MAXGRP * EsCap] the round's value tables, shared by the grid
    int Wmax = 32768, Wmin = 32, lgT = 2048;  // Wmax: ring capacity, a power of two (ring_slot)
    int alloc(int ncand, cudaStream_t st) {
        CELDA_TRY(bufP.alloc((size_t)Wmax * ncand));
        CELDA_TRY(bufA.alloc((size_t)Wmax * ncand));
        CELDA_TRY(spec.alloc(Wmax));
        CELDA_TRY(drawMx.alloc(Wmax));
        CELDA_TRY(drawNz.alloc((size_t)4 * Wmax));
        CELDA_TRY(colsum.alloc((size_t)ncand + 4 * MCMAX));
        CELDA_TRY(lgnCP.alloc((size_t)ncand + 4 * MCMAX));
        CELDA_TRY(status.alloc(2));  // [0] sampler status of the sweeps, [1] visiting order is not a permutation
        CELDA_TRY(status.zero(st));
        CELDA_TRY(barrier.alloc(4));  // [0] grid barrier, [1] task counter of the celda_G sweep
        CELDA_TRY(stats.alloc(40));
        CELDA_TRY(stats.zero(st));
        return 0;
    }
};

// ================================================================================================ chain
struct celda_chain {
    int device = 0, model = 0, nG = 0, nM = 0, nS = 0, K = 0, L = 0, sms = 148;
    long long Z = 0;
    double alpha = 1, beta = 1, delta = 1, gamma = 1;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // celda_chain_iterate (celda_CG, injected draws): the z sweep's order and uniforms are copied on a second stream
    // while the y sweep runs; the iteration's results come back through one pinned block
    cudaStream_t stCopy = nullptr;
    cudaEvent_t evCopyGo = nullptr, evCopyDone = nullptr;
    double* hostRes = nullptr;  // pinned: [0] log-likelihood, [1] the two status words
    bool has_state = false;
    DevBuf<int> cp, ci, cx, s, z, y, zprev, yprev;
    // gene-major twin of the counts (rows sorted by column), built on the device at create for celda_CG / celda_G
    DevBuf<long long> trp;
    DevBuf<int> tcol, tx, changed, nchanged;
    DevBuf<int> nTSByC, nGByCP, nTSByCP, nByC, nGByTS, mCPByS;
    DevBuf<long long> nCP, nByG, nByTS;
    DevBuf<double> probsZ, probsY;
    SweepBufs wb;  // sweep work space
    // cell-sharded mode (EM z step + aggregation): this handle holds a column range of the counts;
    // *_loc are this shard's partial tables, the plain names hold the all-reduced (global) tables.
    celda_comm* comm = nullptr;
    DevBuf<int> nGByCP_loc, mCPByS_loc;
    DevBuf<int> ixy, ixz;
    DevBuf<double> uy, uz;
    DevBuf<double> llPartial, llOut;
    DevBuf<double> phiBuf, thetaBuf;  // EM step
    DevBuf<double> lgt;               // celda_G: lgamma(nTSByC + beta), maintained by the y sweep
    int scanMode = 0;                 // celda_G: 0 exact full-column scan, 1 count columns only (celda_chain_set_scan_mode)
    DevBuf<unsigned long long> marginBuf;  // [3] bits of: min margin, max |probs|, max |lgt| of the last mode-1 sweep
    bool marginValid = false;
    DevBuf<double> lgG;               // celda_CG: lgamma(k + beta), k < 2^22, read through L2 by the y sweep
    // Philox mode
    unsigned long long ph_seed = 0, ph_stream = 0;
    unsigned ph_sweep = 0;
    bool ph_seeded = false;
    int last_draw_n[2] = {0, 0};  // for get_draws: y, z
    DevBuf<int> altLab, altT, altM, altG;  // celda_chain_loglik_alt: candidate labels and their small tables
    DevBuf<long long> altTS;
    DevBuf<int> emFlag, ordSeen, ordSeenZ, resCum, resX;  // resX / resCum: multinomial resample of the counts (perplexity)
    // staged draws
    int staged_iters = 0;
    DevBuf<int> st_ixy, st_ixz;
    DevBuf<double> st_uy, st_uz;
    // profiling
    bool profile = false;
    double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;

    ~celda_chain() {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);  // the buffers go back to the cache (DevCache): no kernel may still use them
        for (auto& pe : pending) {
            cudaEventDestroy(pe.second.first);
            cudaEventDestroy(pe.second.second);
        }
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stCopy) {
            cudaStreamSynchronize(stCopy);
            cudaStreamDestroy(stCopy);
        }
        if (evCopyGo) cudaEventDestroy(evCopyGo);
        if (evCopyDone) cudaEventDestroy(evCopyDone);
        if (hostRes) cudaFreeHost(hostRes);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {

struct ProfScope {
    celda_chain* h;
    int which;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(celda_chain* h_, int w) : h(h_), which(w) {
        static const char* const names[] = {"celda:sweep_y", "celda:rowSumByGroupChange", "celda:sweep_z",
                                            "celda:colSumByGroupChange", "celda:logLik", "celda:rowSumByGroup",
                                            "celda:colSumByGroup"};
        nvtxRangePushA(names[w >= 0 && w < 7 ? w : 0]);  // the chain's phases as NVTX ranges (nsys / ncu --nvtx)
        if (h->profile) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, h->st);
        }
    }
    ~ProfScope() {
        if (h->profile) {
            cudaEventRecord(b, h->st);
            h->pending.push_back({which, {a, b}});
        }
        nvtxRangePop();
    }
};

void prof_collect(celda_chain* h) {
    for (auto& pe : h->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.second.first, pe.second.second) == cudaSuccess) {
            h->prof_ms[pe.first] += ms;
            h->prof_launches[pe.first] += 1;
        }
        cudaEventDestroy(pe.second.first);
        cudaEventDestroy(pe.second.second);
    }
    h->pending.clear();
}

int chain_sync(celda_chain* h) {
    CELDA_CUDA_OK(cudaStreamSynchronize(h->st));
    prof_collect(h);
    return 0;
}

int sweep_status(int stv);

// Status of the sweeps launched since the last check.  A failure is reported once: the words are cleared so that the
// handle can be used again (after set_state -- the chain state behind a failed sweep is unspecified).
int chain_check_status(celda_chain* h) {
    int stv[2] = {0, 0};
    CELDA_CUDA_OK(cudaMemcpyAsync(stv, h->wb.status.p, sizeof(stv), cudaMemcpyDeviceToHost, h->st));
    CELDA_TRY(chain_sync(h));
    if (stv[0] || stv[1]) {
        CELDA_TRY(h->wb.status.zero(h->st));
        CELDA_TRY(chain_sync(h));
    }
    if (stv[1]) return fail("the visiting order must be a permutation of 1..n");
    return sweep_status(stv[0]);
}

// Injected visiting order: copied to the device and checked there (every value in 1..n exactly once).  A bad order
// raises status word 1, which makes the sweep kernels return at once; the error surfaces at the end-of-call check.
int chain_upload_order(celda_chain* h, int* d_ix, const int* ix, long long n) {
    CELDA_CUDA_OK(cudaMemcpyAsync(d_ix, ix, sizeof(int) * n, cudaMemcpyHostToDevice, h->st));
    if (h->ordSeen.n < (size_t)n) CELDA_TRY(h->ordSeen.alloc((size_t)n));
    CELDA_CUDA_OK(cudaMemsetAsync(h->ordSeen.p, 0, sizeof(int) * n, h->st));
    k_validate_order<<<grid_for(n, 256, h->sms), 256, 0, h->st>>>(d_ix, n, h->ordSeen.p, h->wb.status.p + 1);
    count_launch();
    return 0;
}

// The z sweep's injected draws of a celda_CG iteration (12 bytes per cell), issued after the y sweep has been launched:
// the copies run on the copy engine behind the y sweep instead of in front of it.  The copy stream starts behind
// everything already queued on the chain's stream (evCopyGo) and the chain's stream resumes behind it (evCopyDone).
int chain_copy_stream(celda_chain* h) {
    if (h->stCopy) return 0;
    CELDA_CUDA_OK(cudaEventCreateWithFlags(&h->evCopyGo, cudaEventDisableTiming));
    CELDA_CUDA_OK(cudaEventCreateWithFlags(&h->evCopyDone, cudaEventDisableTiming));
    CELDA_CUDA_OK(cudaStreamCreateWithFlags(&h->stCopy, cudaStreamNonBlocking));
    return 0;
}

int chain_upload_z_draws_behind(celda_chain* h, const int* ix_z, const double* u_z) {
    const long long n = h->nM;
    if (h->ordSeenZ.n < (size_t)n) CELDA_TRY(h->ordSeenZ.alloc((size_t)n));
    CELDA_CUDA_OK(cudaStreamWaitEvent(h->stCopy, h->evCopyGo, 0));
    CELDA_CUDA_OK(cudaMemcpyAsync(h->ixz.p, ix_z, sizeof(int) * n, cudaMemcpyHostToDevice, h->stCopy));
    CELDA_CUDA_OK(cudaMemcpyAsync(h->uz.p, u_z, sizeof(double) * n, cudaMemcpyHostToDevice, h->stCopy));
    CELDA_CUDA_OK(cudaMemsetAsync(h->ordSeenZ.p, 0, sizeof(int) * n, h->stCopy));
    k_validate_order<<<grid_for(n, 256, h->sms), 256, 0, h->stCopy>>>(h->ixz.p, n, h->ordSeenZ.p, h->wb.status.p + 1);
    count_launch();
    CELDA_CUDA_OK(cudaEventRecord(h->evCopyDone, h->stCopy));
    CELDA_CUDA_OK(cudaStreamWaitEvent(h->st, h->evCopyDone, 0));
    return 0;
}

// End of an iteration: log-likelihood and status words through one pinned block and one synchronisation.
int chain_finish_iteration(celda_chain* h, double* ll) {
    if (!h->hostRes) CELDA_CUDA_OK(cudaMallocHost(&h->hostRes, 2 * sizeof(double)));
    int* stv = reinterpret_cast<int*>(h->hostRes + 1);
    CELDA_CUDA_OK(cudaMemcpyAsync(h->hostRes, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CELDA_CUDA_OK(cudaMemcpyAsync(stv, h->wb.status.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CELDA_TRY(chain_sync(h));
    *ll = h->hostRes[0];
    const int s0 = stv[0], s1 = stv[1];
    if (s0 || s1) {
        CELDA_TRY(h->wb.status.zero(h->st));
        CELDA_TRY(chain_sync(h));
    }
    if (s1) return fail("the visiting order must be a permutation of 1..n");
    return sweep_status(s0);
}

// Cell-sharded mode: combine the shards' partial G x K and K x S tables with one all-reduce each (int32 sums are
// exact and order-independent, so every rank ends up with the bit-identical global tables).
int chain_reduce_tables(celda_chain* h) {
    if (!h->comm) return 0;
    // one NCCL group: the G x K and the K x S table are reduced by a single fused launch
    CELDA_NCCL_OK(g_nccl.GroupStart());
    CELDA_NCCL_OK(g_nccl.AllReduce(h->nGByCP_loc.p, h->nGByCP.p, (size_t)h->nG * h->K, ncclInt32, ncclSum, h->comm->comm, h->st));
    CELDA_NCCL_OK(g_nccl.AllReduce(h->mCPByS_loc.p, h->mCPByS.p, (size_t)h->K * h->nS, ncclInt32, ncclSum, h->comm->comm, h->st));
    CELDA_NCCL_OK(g_nccl.GroupEnd());
    count_launch(1);
    return 0;
}

unsigned env_u32(const char* env, unsigned dflt);

// ---- decompose (R/celda_CG.R:902-934, R/celda_C.R:800-822, R/celda_G.R:712-733)
int chain_decompose(celda_chain* h) {
    cudaStream_t st = h->st;
    const int sms = h->sms;
    if (h->model != CELDA_MODEL_C) {  // nTSByC = rowSumByGroup(counts, y, L)
        const int threads = 256, nwarp = threads / 32;
        {
            ProfScope ps(h, 5);
            // labels packed into shared memory + 16-byte loads when they fit (k_rowsum_csc_v4), else the generic kernel
            const int t4 = 512, w4 = t4 / 32;
            const size_t smem4 = sizeof(int) * (size_t)w4 * h->L + 2 * (size_t)h->nG;
            if (h->L <= 65535 && smem4 <= 100 * 1024 && env_u32("CELDA_CUDA_AGG_V1", 0) != 1) {
                CELDA_CUDA_OK(cudaFuncSetAttribute(k_rowsum_csc_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
                k_rowsum_csc_v4<<<std::min((h->nM + w4 - 1) / w4, sms * 2), t4, smem4, st>>>(h->nM, h->nG, h->cp.p, h->ci.p, h->cx.p,
                                                                                          h->y.p, h->L, h->nTSByC.p);
            } else {
                const size_t smem1 = sizeof(int) * nwarp * h->L;
                if (smem1 > 200 * 1024) return fail("rowSumByGroup: L=%d needs %zu bytes of shared memory", h->L, smem1);
                CELDA_CUDA_OK(cudaFuncSetAttribute(k_rowsum_csc<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
                k_rowsum_csc<int><<<std::min((h->nM + nwarp - 1) / nwarp, sms * 8), threads, smem1, st>>>(
                    h->nM, h->cp.p, h->ci.p, h->cx.p, h->y.p, h->L, h->nTSByC.p, nullptr);
            }
        }
        count_launch();
        if (h->model == CELDA_MODEL_G) {
            k_lgt_fill<<<grid_for((long long)h->L * h->nM, 256, sms), 256, 0, st>>>(h->nTSByC.p, (long long)h->L * h->nM, h->beta,
                                                                                 h->lgt.p);
            count_launch();
        }
        CELDA_TRY(h->nByTS.zero(st));
        k_fill_int<<<1, 256, 0, st>>>(h->nGByTS.p, h->L, 1);
        k_module_totals<<<grid_for(h->nG, 256, sms), 256, 0, st>>>(h->y.p, h->nByG.p, h->nG,
                                                                    (unsigned long long*)h->nByTS.p, h->nGByTS.p);
        count_launch(2);
    }
    if (h->model != CELDA_MODEL_G) {  // nGByCP = colSumByGroup(counts, z, K); mCPByS = table(z, s)
        const int threads = 256, nwarp = threads / 32;
        DevBuf<int>& ng = h->comm ? h->nGByCP_loc : h->nGByCP;
        DevBuf<int>& mm = h->comm ? h->mCPByS_loc : h->mCPByS;
        if (h->tcol.n) {  // gene-major twin available (celda_CG): no global atomics
            ProfScope ps(h, 6);
            const size_t smem4 = sizeof(int) * 8 * CS_REP * (size_t)h->K;
            const size_t smem5 = sizeof(int) * 32 * CS_REP * (size_t)h->K + (size_t)h->nM;
            const unsigned aggv = env_u32("CELDA_CUDA_AGG_V1", 0);  // measurement aid: 1 = generic kernels, 4 = v4 colsum
            if (h->K <= 256 && smem5 <= 200 * 1024 && h->nG >= 1024 && !aggv) {
                // labels as bytes in shared memory, one warp per gene handed out through a counter
                CELDA_TRY(h->nchanged.zero(st));
                CELDA_CUDA_OK(cudaFuncSetAttribute(k_colsum_twin_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
                k_colsum_twin_v5<<<sms, 1024, smem5, st>>>(h->nG, h->nM, h->trp.p, h->tcol.p, h->tx.p, h->z.p, h->K, ng.p,
                                                           h->nchanged.p);
            } else if (smem4 <= 48 * 1024 && aggv != 1) {
                k_colsum_twin_v4<<<std::min(h->nG, sms * 8), 256, smem4, st>>>(h->nG, h->trp.p, h->tcol.p, h->tx.p, h->z.p, h->K, ng.p);
            } else {
                const size_t smem1 = sizeof(int) * 8 * h->K;
                if (smem1 > 200 * 1024) return fail("colSumByGroup: K=%d needs %zu bytes of shared memory", h->K, smem1);
                CELDA_CUDA_OK(cudaFuncSetAttribute(k_colsum_twin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
                k_colsum_twin<<<std::min(h->nG, sms * 8), 256, smem1, st>>>(h->nG, h->trp.p, h->tcol.p, h->tx.p, h->z.p, h->K, ng.p);
            }
        } else {
            CELDA_TRY(ng.zero(st));
            ProfScope ps(h, 6);
            k_colsum_csc<int><<<std::min((h->nM + nwarp - 1) / nwarp, sms * 16), threads, 0, st>>>(
                h->nM, h->cp.p, h->ci.p, h->cx.p, h->z.p, h->K, ng.p);
        }
        CELDA_TRY(mm.zero(st));
        k_table_zs<<<grid_for(h->nM, 256, sms), 256, 0, st>>>(h->z.p, h->s.p, h->nM, h->K, mm.p);
        count_launch(2);
        CELDA_TRY(chain_reduce_tables(h));
    }
    if (h->model == CELDA_MODEL_CG) {  // nTSByCP = colSumByGroup(nTSByC, z, K); nCP = colSums(nTSByCP)
        CELDA_TRY(h->nTSByCP.zero(st));
        if (sizeof(int) * h->L * h->K > 200 * 1024) return fail("L x K = %d x %d tables exceed the shared-memory budget", h->L, h->K);
        CELDA_CUDA_OK(cudaFuncSetAttribute(k_colsum_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * h->L * h->K)));
        k_colsum_dense<<<std::min(sms * 2, std::max(1, h->nM / 64)), 512, sizeof(int) * h->L * h->K, st>>>(
            h->nTSByC.p, h->L, h->nM, h->z.p, h->K, h->nTSByCP.p);
        k_colsums_small<<<1, 256, 0, st>>>(h->nTSByCP.p, h->L, h->K, h->nCP.p);
        count_launch(2);
    } else if (h->model == CELDA_MODEL_C) {
        CELDA_TRY(h->nCP.zero(st));
        k_colsums_genemajor<<<grid_for((long long)h->nG * h->K, 256, sms), 256, 0, st>>>(
            h->nGByCP.p, h->nG, h->K, (unsigned long long*)h->nCP.p);
        count_launch();
    }
    CELDA_CUDA_OK(cudaGetLastError());
    return 0;
}

// Log-likelihood from device-resident integer tables (shared by the chain handle and the stateless entry points).
struct LLIn {
    int model, K, L, nS, nG;
    long long nM;
    const int *mCPByS, *nTSByCP, *nGByCP /* gene-major */, *nTSByC, *nGByTS;
    const long long *nByG, *nByTS;
    double alpha, beta, delta, gamma;
};

int run_loglik(const LLIn& in, double* partial /* >= 1024 */, double* out, cudaStream_t st) {
    LLArgs a{};
    a.model = in.model;
    a.K = in.K;
    a.L = in.L;
    a.nS = in.nS;
    a.nG = in.nG;
    a.nM = in.nM;
    a.mCPByS = in.mCPByS;
    a.phiTab = in.nTSByCP;
    a.nByTS = in.nByTS;
    a.nGByTS = in.nGByTS;
    a.partial = partial;
    a.alpha = in.alpha;
    a.beta = in.beta;
    a.delta = in.delta;
    a.gamma = in.gamma;
    a.out = out;
    const int NB = 64;  // fixed partial-block counts => deterministic summation order
    int off = 0;
    if (in.model != CELDA_MODEL_C) {
        k_sum_lgamma<long long><<<NB, 256, 0, st>>>(in.nByG, in.nG, in.delta, partial);
        count_launch();
        a.npG = NB;
        off = NB;
    }
    if (in.model == CELDA_MODEL_C) {
        // phi over nGByCP (gene-major [nG][K]): numerator in fixed blocks, denominator one serial column sum per k
        k_sum_lgamma<int><<<NB, 256, 0, st>>>(in.nGByCP, (long long)in.nG * in.K, in.beta, partial + off);
        k_ll_phi_den_genemajor<<<1, 256, 0, st>>>(in.nGByCP, in.nG, in.K, in.beta, partial + off + NB);
        a.npPhiB = NB;
        a.npPhiD = 1;
        count_launch(2);
    }
    if (in.model == CELDA_MODEL_G) {
        const int NBG = 256;
        k_sum_lgamma<int><<<NBG, 256, 0, st>>>(in.nTSByC, (long long)in.L * in.nM, in.beta, partial + off);
        k_sum_lgamma_colsums<<<NBG, 256, 0, st>>>(in.nTSByC, in.L, in.nM, in.beta, partial + off + NBG);
        a.npPhiB = NBG;
        a.npPhiD = NBG;
        count_launch(2);
    }
    k_loglik_final<<<1, 1024, 0, st>>>(a);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    return 0;
}

int chain_loglik_async(celda_chain* h) {
    ProfScope ps(h, 4);
    LLIn in{h->model, h->K, h->L, h->nS, h->nG, h->nM, h->mCPByS.p, h->nTSByCP.p, h->nGByCP.p, h->nTSByC.p,
            h->nGByTS.p, h->nByG.p, h->nByTS.p, h->alpha, h->beta, h->delta, h->gamma};
    return run_loglik(in, h->llPartial.p, h->llOut.p, h->st);
}

// Steady-state window caps (items); CELDA_CUDA_YCAP / CELDA_CUDA_ZCAP override them for tuning runs.
int window_cap(const char* env, int dflt) {
    const char* v = getenv(env);
    const int c = v ? atoi(v) : 0;
    return c > 0 ? c : dflt;
}

unsigned env_u32(const char* env, unsigned dflt) {
    const char* v = getenv(env);
    return v ? (unsigned)strtoul(v, nullptr, 0) : dflt;
}

// Window sizes (items).  unit: the item count whose units (items x candidates) fill one grid-wide lane pass;
// full: the largest multiple of it that fits the ring buffer, so lane-per-unit rounds leave no thread idle.
void window_sizes(const SweepBufs& b, int sms, int threads, int ncand, int cap, int* full, int* unit) {
    const long long lanes = (long long)sms * threads;
    long long u = std::max<long long>(lanes / ncand, b.Wmin);
    if (u > cap) u = cap;
    *unit = (int)u;
    *full = (int)(std::max<long long>(1, cap / u) * u);
}

// Tentative commits per round of k_sweep_z / k_sweep_y (sweeps.cuh, "multi-commit rounds"); CELDA_CUDA_MC overrides
// (1 reproduces the one-mover-per-round schedule).
int sweep_mc() {
    const char* v = getenv("CELDA_CUDA_MC");
    const int m = v ? atoi(v) : MCMAX - 1;
    return std::max(1, std::min(m, MCMAX - 1));
}

// cap: largest steady-state window.  Items behind a mover are patched and redrawn every round, so a sweep whose
// items keep moving (genes of low count in the y sweep) wants a smaller window than one that is quiet.
SweepWork make_work(SweepBufs& b, int sms, int threads, int ncand, int which, int cap) {
    int wfull, wunit;
    window_sizes(b, sms, threads, ncand, std::min(cap, b.Wmax), &wfull, &wunit);
    return SweepWork{b.bufP.p, b.bufA.p, b.spec.p, b.drawMx.p, b.drawNz.p, b.colsum.p, b.lgnCP.p, b.barrier.p, b.status.p,
                     b.stats.p + 8 * which, b.Wmax, wfull, wunit, b.Wmin, b.lgT, sweep_mc()};
}

// a: data pointers and sizes filled by the caller; work space, window policy and launch geometry set here.
int run_sweep_z(ZArgs a, SweepBufs& b, int sms, int which, cudaStream_t st) {
    int threads = 1024;
    int MC = sweep_mc();  // fewer tentative commits per round when large tables leave no room for their column slots
    while (MC > 1 && z_sweep_smem(a.R, a.K, a.nS, b.lgT, threads / 32, 0, MC) > 160 * 1024) MC /= 2;
    if (z_sweep_smem(a.R, a.K, a.nS, b.lgT, threads / 32, 0, MC) > 220 * 1024) threads = 512;
    a.w = make_work(b, sms, threads, a.K, which, window_cap("CELDA_CUDA_ZCAP", 32768));
    a.w.MC = MC;
    // value tables: the shared memory left over holds the heads; sized from this sweep's count distribution
    a.EsCap = 0;
    {
        // the table head shares its buffer with the commit slots beyond the second commit (z_sweep_vtcap)
        const size_t base = z_sweep_smem(a.R, a.K, a.nS, b.lgT, threads / 32, 0, MC) -
                            sizeof(double) * z_sweep_vtcap(a.R, a.K, 0, MC);
        // shared-memory budget of the kernel in KB (the rest of the 228 KB stays L1); CELDA_CUDA_ZSMEM_KB overrides
        const long long budget = (long long)window_cap("CELDA_CUDA_ZSMEM_KB", 207) * 1024;
        const long long room = (budget - (long long)base) / (long long)sizeof(double);
        if (room >= 8LL * a.R && a.nM >= 4096) {
            a.EsCap = (int)std::min<long long>(room, 1 << 20);
            if (b.ztOffg.n < (size_t)a.R + 1) {
                CELDA_TRY(b.ztHist.alloc((size_t)a.R));  // per-row maximum count
                CELDA_TRY(b.ztOffg.alloc((size_t)a.R + 1));
                CELDA_TRY(b.ztGrp.alloc(ZT_MAXGRP + 3));
                CELDA_TRY(b.ztTmp.alloc((size_t)a.R));
            }
            if (b.ztXT.n < (size_t)a.R * a.nM) CELDA_TRY(b.ztXT.alloc((size_t)a.R * a.nM));
            if (b.ztZT.n < (size_t)a.nM) {
                CELDA_TRY(b.ztZT.alloc((size_t)a.nM));
                CELDA_TRY(b.ztNT.alloc((size_t)a.nM));
            }
            a.vtStride = (long long)ZT_MAXGRP * a.EsCap;  // k_pick_rowV never lets the rows' tables exceed this
            if (b.ztVal.n < (size_t)a.K * a.vtStride) CELDA_TRY(b.ztVal.alloc((size_t)a.K * a.vtStride));
            a.vtg = b.ztVal.p;
            CELDA_TRY(b.ztHist.zero(st));
            const int RC = std::min(a.R, 1024);
            const size_t ptSmem = sizeof(int) * ((size_t)32 * (RC | 1) + RC);
            CELDA_CUDA_OK(cudaFuncSetAttribute(k_permute_transpose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ptSmem));
            k_permute_transpose<<<(int)std::min<long long>((a.nM + 31) / 32, (long long)sms * 8), 256, ptSmem, st>>>(
                a.counts, a.R, RC, a.nM, a.ix, b.ztXT.p, b.ztHist.p, a.z, a.nByC, b.ztZT.p, b.ztNT.p, b.status.p);
            k_pick_rowV<<<1, 256, 0, st>>>(b.ztHist.p, a.R, a.EsCap, b.ztOffg.p, b.ztGrp.p, b.ztTmp.p);
            count_launch(2);
            a.offg = b.ztOffg.p;
            a.grp = b.ztGrp.p;
            a.xT = b.ztXT.p;
            a.zT = b.ztZT.p;
            a.nT = b.ztNT.p;
            // Steady-state window for table rounds: the sweep is cut into nr equal windows such that the (candidate,
            // 32-cell chunk) tasks and the draws fill whole grid-wide warp passes (a 32768-cell cap on 100k cells would
            // run 3 windows of 6.5 passes, each rounded up to 7, plus a remainder)
            const long long tw = (long long)sms * (threads / 32), cap = a.w.Wfull;
            long long bestW = cap;
            double bestCost = 1e300;
            for (long long nr = (a.nM + cap - 1) / cap, q = 0; q < 4; ++q, ++nr) {
                const long long W = std::min<long long>(cap, (((a.nM + nr - 1) / nr + 31) / 32) * 32);
                const long long tablePasses = ((long long)a.K * (W / 32) + tw - 1) / tw, drawPasses = (W + tw - 1) / tw;
                const double cost = (double)((a.nM + W - 1) / W) * (6.5 * (double)tablePasses + (double)drawPasses + 1.0);
                if (cost < bestCost) {
                    bestCost = cost;
                    bestW = W;
                }
            }
            a.w.Wfull = (int)bestW;
            a.w.Wunit = std::min(a.w.Wunit, a.w.Wfull);
        }
    }
    a.vtCap = (int)z_sweep_vtcap(a.R, a.K, a.EsCap, MC);
    const size_t smem = z_sweep_smem(a.R, a.K, a.nS, b.lgT, threads / 32, a.EsCap, MC);
    if (smem > 220 * 1024) return fail("z sweep: K=%d x R=%d tables need %zu bytes of shared memory (> 220 KB)", a.K, a.R, smem);
    CELDA_CUDA_OK(cudaFuncSetAttribute(k_sweep_z, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CELDA_CUDA_OK(cudaMemsetAsync(b.barrier.p, 0, sizeof(unsigned), st));
    void* args[] = {&a};
    CELDA_CUDA_OK(cudaLaunchCooperativeKernel((void*)k_sweep_z, dim3(sms), dim3(threads), args, smem, st));
    count_launch();
    return 0;
}

int run_sweep_y(YArgs a, SweepBufs& b, int sms, int which, cudaStream_t st) {
    int threads = 1024;
    int MC = sweep_mc();  // fewer tentative commits per round when large tables leave no room for their row slots
    while (MC > 1 && y_sweep_smem(a.L, a.ncol, b.lgT, threads / 32, MC) > 220 * 1024) MC /= 2;
    if (y_sweep_smem(a.L, a.ncol, b.lgT, threads / 32, MC) > 220 * 1024) threads = 512;
    a.w = make_work(b, sms, threads, a.L, which, window_cap("CELDA_CUDA_YCAP", 8192));
    a.w.MC = MC;
    const size_t smem = y_sweep_smem(a.L, a.ncol, b.lgT, threads / 32, MC);
    if (smem > 220 * 1024) return fail("y sweep: L=%d x K=%d tables need %zu bytes of shared memory (> 220 KB)", a.L, a.ncol, smem);
    CELDA_CUDA_OK(cudaFuncSetAttribute(k_sweep_y, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CELDA_CUDA_OK(cudaMemsetAsync(b.barrier.p, 0, sizeof(unsigned), st));
    void* args[] = {&a};
    CELDA_CUDA_OK(cudaLaunchCooperativeKernel((void*)k_sweep_y, dim3(sms), dim3(threads), args, smem, st));
    count_launch();
    return 0;
}

int run_sweep_yg(YGArgs a, SweepBufs& b, int sms, int which, cudaStream_t st) {
    int threads = YG_THREADS;
    const size_t budget = 220 * 1024;
    // the draw phase's scratch (per warp) and the unit phase's tile ring share the same memory
    while (threads > 128 && yg_sweep_smem(a.L, b.lgT, threads / 32, 4) > budget) threads /= 2;
    if (yg_sweep_smem(a.L, b.lgT, threads / 32, 4) > budget)
        return fail("celda_G y sweep: L=%d needs %zu bytes of shared memory (> 220 KB)", a.L, yg_sweep_smem(a.L, b.lgT, threads / 32, 4));
    // columns per tile: as many as the budget allows (a multiple of 4, at most 64)
    int TC = 64;
    while (TC > 4 && yg_sweep_smem(a.L, b.lgT, threads / 32, TC) > budget) TC -= 4;
    a.TC = TC;
    const size_t smem = yg_sweep_smem(a.L, b.lgT, threads / 32, TC);
    // window: whole waves of (gene group) x (module group) tasks over the grid; a task takes one gene per warp
    const int nwarp = threads / 32, ngroup = (a.L + 32 * YG_MAXMC - 1) / (32 * YG_MAXMC);
    long long unit = (long long)nwarp * std::max(1, sms / ngroup);
    const long long cap = std::min<long long>(b.Wmax, ((a.flags & 8) ? 2 : 3) * unit);
    if (unit > cap) unit = std::max<long long>(1, cap / nwarp) * nwarp;
    const long long full = std::max<long long>(1, cap / unit) * unit;
    a.w = SweepWork{b.bufP.p, b.bufA.p, b.spec.p, b.drawMx.p, b.drawNz.p, b.colsum.p, b.lgnCP.p, b.barrier.p, b.status.p,
                    b.stats.p + 8 * which, b.Wmax, (int)full, (int)unit, std::max(b.Wmin, nwarp), b.lgT, 1};
    CELDA_CUDA_OK(cudaFuncSetAttribute(k_sweep_yg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CELDA_CUDA_OK(cudaMemsetAsync(b.barrier.p, 0, 4 * sizeof(unsigned), st));
    void* args[] = {&a};
    CELDA_CUDA_OK(cudaLaunchCooperativeKernel((void*)k_sweep_yg, dim3(sms), dim3(threads), args, smem, st));
    count_launch();
    return 0;
}

// celda_C Gibbs z sweep on the raw counts (k_sweep_zc): windows are whole waves of (cell group) x (population chunk)
int run_sweep_zc(ZCArgs a, SweepBufs& b, int sms, int which, cudaStream_t st) {
    int threads = 1024;
    if (zc_sweep_smem(a.K, a.nS, b.lgT, threads / 32) > 220 * 1024) threads = 512;
    if (zc_sweep_smem(a.K, a.nS, b.lgT, threads / 32) > 220 * 1024) threads = 256;
    const size_t smem = zc_sweep_smem(a.K, a.nS, b.lgT, threads / 32);
    if (smem > 220 * 1024)
        return fail("celda_C z sweep: K=%d, %d samples need %zu bytes of shared memory (> 220 KB)", a.K, a.nS, smem);
    const int nwarp = threads / 32, nchunk = (a.K + 31) / 32;
    long long unit = std::max<long long>(nwarp, (long long)nwarp * sms / nchunk / nwarp * nwarp);
    const long long cap = std::min<long long>(b.Wmax, 8192);
    if (unit > cap) unit = cap / nwarp * nwarp;
    const long long full = std::max<long long>(1, cap / unit) * unit;
    a.w = SweepWork{b.bufP.p, b.bufA.p, b.spec.p, b.drawMx.p, b.drawNz.p, b.colsum.p, b.lgnCP.p, b.barrier.p, b.status.p,
                    b.stats.p + 8 * which, b.Wmax, (int)full, (int)unit, std::max(b.Wmin, nwarp), b.lgT, 1};
    CELDA_CUDA_OK(cudaFuncSetAttribute(k_sweep_zc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CELDA_CUDA_OK(cudaMemsetAsync(b.barrier.p, 0, sizeof(unsigned), st));
    void* args[] = {&a};
    CELDA_CUDA_OK(cudaLaunchCooperativeKernel((void*)k_sweep_zc, dim3(sms), dim3(threads), args, smem, st));
    count_launch();
    return 0;
}

int sweep_status(int stv) {
    if (stv == -1) return fail("sample.int would take the Walker alias branch (> 200 likely categories): not supported");
    if (stv == -2) return fail("NA in probability vector");
    if (stv == -3) return fail("too few positive probabilities");
    if (stv) return fail("sweep failed with status %d", stv);
    return 0;
}

int launch_sweep_z(celda_chain* h, const int* d_ix, const double* d_u, int doSample, int which) {
    ProfScope ps(h, which);
    ZArgs a{};
    a.R = h->L;
    a.K = h->K;
    a.nS = h->nS;
    a.nM = h->nM;
    a.counts = h->nTSByC.p;
    a.nByC = h->nByC.p;
    a.s = h->s.p;
    a.z = h->z.p;
    a.nTab = h->nTSByCP.p;
    a.nCP = h->nCP.p;
    a.mCPByS = h->mCPByS.p;
    a.alpha = h->alpha;
    a.beta = h->beta;
    a.ix = d_ix;
    a.u = d_u;
    a.doSample = doSample;
    a.probs = h->probsZ.p;
    a.lgG = h->lgG.p;
    a.TG = (int)h->lgG.n;
    return run_sweep_z(a, h->wb, h->sms, which, h->st);
}

int launch_sweep_y(celda_chain* h, const int* d_ix, const double* d_u, int doSample, int which) {
    ProfScope ps(h, which);
    YArgs a{};
    a.nG = h->nG;
    a.ncol = h->K;
    a.L = h->L;
    a.counts = h->nGByCP.p;
    a.nByG = h->nByG.p;
    a.y = h->y.p;
    a.tTab = h->nTSByCP.p;
    a.nByTS = h->nByTS.p;
    a.nGByTS = h->nGByTS.p;
    a.beta = h->beta;
    a.gamma = h->gamma;
    a.delta = h->delta;
    a.ix = d_ix;
    a.u = d_u;
    a.doSample = doSample;
    a.probs = h->probsY.p;
    a.lgG = h->lgG.p;
    a.TG = (int)h->lgG.n;
    a.lgMask = env_u32("CELDA_CUDA_LGMASK", 0x77777777u);
    return run_sweep_y(a, h->wb, h->sms, which, h->st);
}

// y sweep + rowSumByGroupChange (R/celda_CG.R:544-564)
int chain_step_y(celda_chain* h, const int* d_ix, const double* d_u, int doSample) {
    if (h->model == CELDA_MODEL_C) return fail("sweep_y: celda_C has no feature modules");
    if (h->model == CELDA_MODEL_G) {  // raw counts against nTSByC (L x nM); the kernel keeps nTSByC itself current
        ProfScope ps(h, 0);
        YGArgs a{};
        a.nG = h->nG;
        a.L = h->L;
        a.nM = h->nM;
        a.rp = h->trp.p;
        a.tcol = h->tcol.p;
        a.tx = h->tx.p;
        a.nByG = h->nByG.p;
        a.y = h->y.p;
        a.t = h->nTSByC.p;
        a.lgt = h->lgt.p;
        a.nByTS = h->nByTS.p;
        a.nGByTS = h->nGByTS.p;
        a.beta = h->beta;
        a.gamma = h->gamma;
        a.delta = h->delta;
        a.ix = d_ix;
        a.u = d_u;
        a.doSample = doSample;
        a.probs = h->probsY.p;
        a.noTma = (int)env_u32("CELDA_CUDA_YG_NOTMA", 0);
        a.flags = (int)env_u32("CELDA_CUDA_YG_FLAGS", 3) | (h->scanMode ? 32 : 0);
        CELDA_TRY(run_sweep_yg(a, h->wb, h->sms, 0, h->st));
        h->marginValid = false;
        if (h->scanMode && doSample) {  // the report of scan mode 1 (see celda_chain_scan_margin)
            if (!h->marginBuf.n) CELDA_TRY(h->marginBuf.alloc(3));
            const unsigned long long init[3] = {0x7ff0000000000000ULL, 0ULL, 0ULL};
            CELDA_CUDA_OK(cudaMemcpyAsync(h->marginBuf.p, init, sizeof(init), cudaMemcpyHostToDevice, h->st));
            std::vector<long long> tot((size_t)h->L);
            CELDA_CUDA_OK(cudaMemcpyAsync(tot.data(), h->nByTS.p, sizeof(long long) * h->L, cudaMemcpyDeviceToHost, h->st));
            CELDA_CUDA_OK(cudaStreamSynchronize(h->st));  // `init` and `tot` live on this stack frame
            double ntot = 0;
            for (long long v : tot) ntot += (double)v;
            k_max_abs<<<h->sms * 8, 256, 0, h->st>>>(h->lgt.p, (long long)h->L * h->nM, h->marginBuf.p + 2);
            k_yg_margin<<<h->sms * 4, 256, 0, h->st>>>(h->nG, h->L, (long long)h->nM, h->probsY.p, h->y.p, d_ix, d_u, h->nByG.p,
                                                      std::log(ntot + h->beta + 1.0), h->marginBuf.p + 2, h->marginBuf.p);
            count_launch(2);
            CELDA_CUDA_OK(cudaGetLastError());
            h->marginValid = true;
        }
        return 0;
    }
    CELDA_CUDA_OK(cudaMemcpyAsync(h->yprev.p, h->y.p, sizeof(int) * h->nG, cudaMemcpyDeviceToDevice, h->st));
    CELDA_TRY(launch_sweep_y(h, d_ix, d_u, doSample, 0));
    if (doSample) {
        ProfScope ps(h, 1);
        CELDA_TRY(h->nchanged.zero(h->st));
        k_list_changed<<<grid_for(h->nG, 256, h->sms), 256, 0, h->st>>>(h->y.p, h->yprev.p, h->nG, h->changed.p, h->nchanged.p);
        k_rowsum_change_twin<<<h->sms * 4, 256, 0, h->st>>>(h->changed.p, h->nchanged.p, h->trp.p, h->tcol.p, h->tx.p, h->y.p,
                                                          h->yprev.p, h->L, h->nTSByC.p);
        count_launch(2);
    }
    CELDA_CUDA_OK(cudaGetLastError());
    return 0;
}

// z sweep + colSumByGroupChange (R/celda_CG.R:567-585)
int chain_step_z(celda_chain* h, const int* d_ix, const double* d_u, int doSample) {
    if (h->model == CELDA_MODEL_G) return fail("sweep_z_gibbs: celda_G has no cell populations");
    if (h->model == CELDA_MODEL_C) {  // raw counts against nGByCP (R/celda_C.R:427-440); the kernel keeps the tables current
        if (h->comm) return fail("sweep_z_gibbs: a cell-sharded celda_C chain runs the EM z step only");
        ProfScope ps(h, 2);
        const long long nt = (long long)h->nG * h->K;
        k_lgt_fill<<<grid_for(nt, 256, h->sms), 256, 0, h->st>>>(h->nGByCP.p, nt, h->beta, h->lgt.p);
        count_launch();
        ZCArgs a{};
        a.nG = h->nG;
        a.K = h->K;
        a.nS = h->nS;
        a.nM = h->nM;
        a.cp = h->cp.p;
        a.ci = h->ci.p;
        a.cx = h->cx.p;
        a.nByC = h->nByC.p;
        a.s = h->s.p;
        a.z = h->z.p;
        a.n = h->nGByCP.p;
        a.lgn = h->lgt.p;
        a.nCP = h->nCP.p;
        a.mCPByS = h->mCPByS.p;
        a.alpha = h->alpha;
        a.beta = h->beta;
        a.ix = d_ix;
        a.u = d_u;
        a.doSample = doSample;
        a.probs = h->probsZ.p;
        return run_sweep_zc(a, h->wb, h->sms, 2, h->st);
    }
    CELDA_CUDA_OK(cudaMemcpyAsync(h->zprev.p, h->z.p, sizeof(int) * h->nM, cudaMemcpyDeviceToDevice, h->st));
    CELDA_TRY(launch_sweep_z(h, d_ix, d_u, doSample, 2));
    if (doSample) {
        ProfScope ps(h, 3);
        const int threads = 256, nwarp = threads / 32;
        k_colsum_change_csc<int><<<std::min((h->nM + nwarp - 1) / nwarp, h->sms * 16), threads, 0, h->st>>>(
            h->nM, h->cp.p, h->ci.p, h->cx.p, h->z.p, h->zprev.p, h->K, h->nGByCP.p);
        count_launch();
    }
    CELDA_CUDA_OK(cudaGetLastError());
    return 0;
}

// EM z step (.cCCalcEMProbZ, R/celda_C.R:717-756) + re-decomposition (.cCReDecomposeCounts :825-839; in celda_CG
// additionally nGByCP <- colSumByGroupChange, R/celda_CG.R:584).  celda_C: counts = the CSC; celda_CG: counts =
// nTSByC (dense L x nM) against nTSByCP.
int chain_step_z_em(celda_chain* h, int doSample) {
    if (h->model == CELDA_MODEL_G) return fail("sweep_z_em: celda_G has no cell populations");
    ProfScope ps(h, 2);
    cudaStream_t st = h->st;
    const bool isC = h->model == CELDA_MODEL_C;
    const int R = isC ? h->nG : h->L, K = h->K;
    CELDA_TRY(h->emFlag.zero(st));
    k_em_phi<<<grid_for((long long)R * K, 256, h->sms), 256, 0, st>>>(isC ? h->nGByCP.p : h->nTSByCP.p, isC, R, K, h->nCP.p,
                                                                      h->beta, h->phiBuf.p, h->emFlag.p);
    k_em_theta<<<std::min(h->nS, 1024), 64, 0, st>>>(h->mCPByS.p, K, h->nS, h->alpha, h->thetaBuf.p, h->emFlag.p);
    CELDA_CUDA_OK(cudaMemcpyAsync(h->zprev.p, h->z.p, sizeof(int) * h->nM, cudaMemcpyDeviceToDevice, st));
    const int threads = 256, nwarp = threads / 32;
    const int grid = (int)std::min<long long>(((long long)h->nM + nwarp - 1) / nwarp, (long long)h->sms * 32);
    if (isC && K <= 64 && (long long)h->nM >= 4096) {
        // large celda_C problems: phi streamed through shared memory in gene tiles (TMA bulk copies, double-buffered)
        // shared by a block of cells
        const int GT = std::min((R + 1) & ~1, (int)(200 * 1024 / (sizeof(double) * K)) & ~1);  // even: 16-byte aligned tiles
        if (GT > 65536) return fail("EM step: gene tile of %d rows exceeds the packed row field", GT);
        const size_t smem = 128 + sizeof(double) * ((size_t)GT * K + 2);
        CELDA_CUDA_OK(cudaFuncSetAttribute(k_em_probs_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long cpb = 32LL * EM_CPW;
        const int tgrid = (int)std::min<long long>((h->nM + cpb - 1) / cpb, (long long)h->sms);
        k_em_probs_tma<<<tgrid, 1024, smem, st>>>(R, h->nM, K, GT, h->cp.p, h->ci.p, h->cx.p, h->phiBuf.p, h->thetaBuf.p,
                                                 h->s.p, h->probsZ.p, doSample ? h->z.p : nullptr);
    } else if (isC)
        k_em_probs<true><<<grid, threads, 0, st>>>(R, h->nM, K, h->cp.p, h->ci.p, h->cx.p, h->phiBuf.p, h->thetaBuf.p, h->s.p,
                                                  h->probsZ.p, doSample ? h->z.p : nullptr);
    else
        k_em_probs<false><<<grid, threads, 0, st>>>(R, h->nM, K, nullptr, nullptr, h->nTSByC.p, h->phiBuf.p, h->thetaBuf.p,
                                                   h->s.p, h->probsZ.p, doSample ? h->z.p : nullptr);
    count_launch(3);
    if (doSample) {
        if (!isC) {
            k_colsum_change_dense<<<grid_for((long long)h->L * h->nM, 256, h->sms), 256, 0, st>>>(h->nTSByC.p, h->L, h->nM, h->z.p,
                                                                                               h->zprev.p, h->nTSByCP.p);
            k_colsums_small<<<1, 256, 0, st>>>(h->nTSByCP.p, h->L, K, h->nCP.p);
            count_launch(2);
        }
        k_colsum_change_csc<int><<<std::min((h->nM + nwarp - 1) / nwarp, h->sms * 16), threads, 0, st>>>(
            h->nM, h->cp.p, h->ci.p, h->cx.p, h->z.p, h->zprev.p, K, h->comm ? h->nGByCP_loc.p : h->nGByCP.p);
        {
            DevBuf<int>& mm = h->comm ? h->mCPByS_loc : h->mCPByS;
            CELDA_TRY(mm.zero(st));
            k_table_zs<<<grid_for(h->nM, 256, h->sms), 256, 0, st>>>(h->z.p, h->s.p, h->nM, K, mm.p);
            count_launch(2);
        }
        CELDA_TRY(chain_reduce_tables(h));
        if (isC) {
            CELDA_TRY(h->nCP.zero(st));
            k_colsums_genemajor<<<grid_for((long long)h->nG * K, 256, h->sms), 256, 0, st>>>(h->nGByCP.p, h->nG, K,
                                                                                          (unsigned long long*)h->nCP.p);
            count_launch();
        }
    }
    CELDA_CUDA_OK(cudaGetLastError());
    return 0;
}

int chain_check_em_flag(celda_chain* h) {
    int f = 0;
    CELDA_CUDA_OK(cudaMemcpyAsync(&f, h->emFlag.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CELDA_TRY(chain_sync(h));
    if (f) return fail("Division by 0. Make sure colSums of counts does not contain 0 after rounding counts to integers.");
    return 0;
}

// Philox mode: visiting order (philox_order, a keyed permutation evaluated item by item: philox.cuh) and one uniform
// per visited item, generated by one kernel into the chain's ix / u staging buffers.
int chain_philox_draws(celda_chain* h, long long n, int* d_ix, double* d_u) {
    if (!h->ph_seeded) return fail("philox mode: call celda_chain_seed first");
    k_philox_draws<<<grid_for(n, 256, h->sms), 256, 0, h->st>>>(n, philox_order_bits((unsigned long long)n), (unsigned)h->ph_seed,
                                                                (unsigned)(h->ph_seed >> 32), h->ph_sweep,
                                                                (unsigned)h->ph_stream, d_ix, d_u);
    count_launch();
    h->ph_sweep += 1;
    return 0;
}

int validate_order(const int* ix, long long n, const char* what) {
    std::vector<unsigned char> seen((size_t)n, 0);
    for (long long i = 0; i < n; ++i) {
        const int v = ix[i];
        if (v < 1 || v > n || seen[v - 1]) return fail("%s must be a permutation of 1..%lld", what, n);
        seen[v - 1] = 1;
    }
    return 0;
}

// Every buffer whose size depends on K or L (tables, probabilities, sweep work space).  Called at create and by
// celda_chain_reconfigure: the CSC copy, its gene-major twin and the row / column totals are (K, L)-independent.
int chain_alloc_tables(celda_chain* h) {
    const int model = h->model, nG = h->nG, nM = h->nM;
    cudaStream_t st = h->st;
    if (model != CELDA_MODEL_C) {
        CELDA_TRY(h->nTSByC.alloc((size_t)h->L * nM));
        if (model == CELDA_MODEL_G) CELDA_TRY(h->lgt.alloc((size_t)h->L * nM));
        CELDA_TRY(h->nByTS.alloc(h->L));
        CELDA_TRY(h->nGByTS.alloc(h->L));
        CELDA_TRY(h->probsY.alloc((size_t)h->L * nG));
    }
    if (model != CELDA_MODEL_G) {
        CELDA_TRY(h->nGByCP.alloc((size_t)nG * h->K));
        CELDA_TRY(h->mCPByS.alloc((size_t)h->K * h->nS));
        CELDA_TRY(h->nCP.alloc(h->K));
        CELDA_TRY(h->probsZ.alloc((size_t)h->K * nM));
        if (h->comm) {
            CELDA_TRY(h->nGByCP_loc.alloc((size_t)nG * h->K));
            CELDA_TRY(h->mCPByS_loc.alloc((size_t)h->K * h->nS));
        }
    }
    if (model == CELDA_MODEL_CG) CELDA_TRY(h->nTSByCP.alloc((size_t)h->L * h->K));
    if (model == CELDA_MODEL_C) CELDA_TRY(h->lgt.alloc((size_t)nG * h->K));  // lgamma(nGByCP + beta), Gibbs z sweep
    if (model == CELDA_MODEL_CG && h->lgG.n == 0) {
        const long long TG = 1LL << 22;
        CELDA_TRY(h->lgG.alloc((size_t)TG));
        k_lgG_fill<<<grid_for(TG, 256, h->sms), 256, 0, st>>>(h->beta, TG, h->lgG.p);
        count_launch();
    }
    CELDA_TRY(h->wb.alloc(std::max(h->K, h->L), st));
    CELDA_TRY(h->phiBuf.alloc((size_t)std::max(nG, h->L) * h->K + 2));  // + slack for 16-byte bulk copies
    CELDA_TRY(h->thetaBuf.alloc((size_t)h->K * h->nS));
    h->has_state = false;
    return 0;
}

}  // namespace

extern "C" {

const char* celda_cuda_last_error(void) { return g_err; }

int celda_cuda_device_count(int* n) {
    CELDA_CUDA_OK(cudaGetDeviceCount(n));
    return 0;
}

int celda_cuda_set_device(int device) {
    CELDA_CUDA_OK(cudaSetDevice(device));
    g_device = device;
    return 0;
}

long long celda_cuda_launch_count(void) { return g_launches.load(); }

int celda_cuda_fp64_peak(double* tflops) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<double> out;
    CELDA_TRY(out.alloc(1));
    const int sms = num_sms(g_device), blocks = sms * 4, threads = 512, iters = 20000;
    cudaEvent_t e0, e1;
    CELDA_CUDA_OK(cudaEventCreate(&e0));
    CELDA_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_fp64_peak<<<blocks, threads>>>(out.p, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        count_launch();
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CELDA_CUDA_OK(cudaGetLastError());
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}

// ------------------------------------------------------------------------------------------------ math
static int math_vec(int op, const double* x, long long n, double* out) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<double> dx, dout;
    CELDA_TRY(dx.upload(x, (size_t)n));
    CELDA_TRY(dout.alloc((size_t)n));
    const int grid = grid_for(n, 256, num_sms(g_device));
    if (op == 0) k_math_vec<0><<<grid, 256>>>(dx.p, n, dout.p);
    else if (op == 1) k_math_vec<1><<<grid, 256>>>(dx.p, n, dout.p);
    else k_math_vec<2><<<grid, 256>>>(dx.p, n, dout.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dout.download(out, (size_t)n));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
int celda_cuda_lgamma(const double* x, long long n, double* out) { return math_vec(0, x, n, out); }
int celda_cuda_log(const double* x, long long n, double* out) { return math_vec(1, x, n, out); }
int celda_cuda_exp(const double* x, long long n, double* out) { return math_vec(2, x, n, out); }

// ------------------------------------------------------------------------------------------------ sparse aggregation
// The stateless entry points compute in double like the reference (NumericMatrix): sums of integer-valued
// doubles are exact and order-independent, so atomics reproduce the reference bit for bit.
struct CscDev {
    DevBuf<int> p, i;
    DevBuf<double> x;
    int upload(int ncol, const int* hp, const int* hi, const double* hx) {
        const long long nnz = hp[ncol];
        CELDA_TRY(p.upload(hp, (size_t)ncol + 1));
        CELDA_TRY(i.upload(hi, (size_t)nnz));
        CELDA_TRY(x.upload(hx, (size_t)nnz));
        return 0;
    }
};

int celda_colSumByGroupSparse(int nrow, int ncol, const int* p, const int* i, const double* x, const int* group,
                              long long ngroup, int K, double* out) {
    if (ncol != ngroup) return fail("Length of 'group' must be equal to the number of columns in 'counts'.");
    for (long long c = 0; c < ngroup; ++c)  // src/matrixSumsSparse.cpp:21: this one message of the file has no full stop
        if (group[c] < 1 || group[c] > K) return fail("The entries in 'group' need to be between 1 and 'K'");
    if (K > ncol) return fail("'K' cannot be bigger than the number of columns in 'counts'.");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    CscDev m;
    DevBuf<int> g;
    DevBuf<double> acc, res;
    CELDA_TRY(m.upload(ncol, p, i, x));
    CELDA_TRY(g.upload(group, (size_t)ngroup));
    CELDA_TRY(acc.alloc((size_t)nrow * K));
    CELDA_TRY(acc.zero());
    CELDA_TRY(res.alloc((size_t)nrow * K));
    k_colsum_csc<double><<<std::min((ncol + 7) / 8, sms * 16), 256>>>(ncol, m.p.p, m.i.p, m.x.p, g.p, K, acc.p);
    k_transpose_convert<double, double><<<grid_for((long long)nrow * K, 256, sms), 256>>>(acc.p, nrow, K, res.p, true);
    count_launch(2);
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(res.download(out, (size_t)nrow * K));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_rowSumByGroupSparse(int nrow, int ncol, const int* p, const int* i, const double* x, const int* group,
                              long long ngroup, int L, double* out) {
    if (nrow != ngroup) return fail("Length of 'group' must be equal to the number of rows in 'counts'.");
    CELDA_TRY(check_labels_host(group, ngroup, L, "group", "L"));
    if (L > nrow) return fail("'L' cannot be bigger than the number of rows in 'counts'.");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    CscDev m;
    DevBuf<int> g;
    DevBuf<double> res;
    CELDA_TRY(m.upload(ncol, p, i, x));
    CELDA_TRY(g.upload(group, (size_t)ngroup));
    CELDA_TRY(res.alloc((size_t)L * ncol));
    const int threads = (sizeof(double) * 8 * L <= 96 * 1024) ? 256 : 32, nwarp = threads / 32;
    const size_t smem = sizeof(double) * nwarp * L;
    if (smem > 200 * 1024) return fail("'L' = %d is too large for the aggregation kernel", L);
    CELDA_CUDA_OK(cudaFuncSetAttribute(k_rowsum_csc<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rowsum_csc<double><<<std::min((ncol + nwarp - 1) / nwarp, sms * 8), threads, smem>>>(ncol, m.p.p, m.i.p, m.x.p, g.p, L,
                                                                                          res.p, nullptr);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(res.download(out, (size_t)L * ncol));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_colSumByGroupChangeSparse(int nrow, int ncol, const int* p, const int* i, const double* x, double* px,
                                    int px_nrow, const int* group, long long ngroup, const int* pgroup,
                                    long long npgroup, int K) {
    if (ncol != ngroup) return fail("Length of 'group' must be equal to the number of columns in 'counts'.");
    if (ngroup != npgroup) return fail("Length of 'group' must equal 'pgroup'.");
    CELDA_TRY(check_labels_host(group, ngroup, K, "group", "K"));
    CELDA_TRY(check_labels_host(pgroup, npgroup, K, "pgroup", "K"));
    if (px_nrow != nrow) return fail("'px' and 'counts' must have the same number of rows.");
    if (K > ncol) return fail("'K' cannot be bigger than the number of columns in 'counts'.");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    CscDev m;
    DevBuf<int> g, pg;
    DevBuf<double> dpx, acc;
    CELDA_TRY(m.upload(ncol, p, i, x));
    CELDA_TRY(g.upload(group, (size_t)ngroup));
    CELDA_TRY(pg.upload(pgroup, (size_t)npgroup));
    CELDA_TRY(dpx.upload(px, (size_t)nrow * K));
    CELDA_TRY(acc.alloc((size_t)nrow * K));
    const int grid = grid_for((long long)nrow * K, 256, sms);
    k_transpose_convert<double, double><<<grid, 256>>>(dpx.p, nrow, K, acc.p, false);
    k_colsum_change_csc<double><<<std::min((ncol + 7) / 8, sms * 16), 256>>>(ncol, m.p.p, m.i.p, m.x.p, g.p, pg.p, K, acc.p);
    k_transpose_convert<double, double><<<grid, 256>>>(acc.p, nrow, K, dpx.p, true);
    count_launch(3);
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dpx.download(px, (size_t)nrow * K));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_rowSumByGroupChangeSparse(int nrow, int ncol, const int* p, const int* i, const double* x, double* px,
                                    int px_ncol, const int* group, long long ngroup, const int* pgroup,
                                    long long npgroup, int L) {
    if (nrow != ngroup) return fail("Length of 'group' must be equal to the number of rows in 'counts'.");
    if (ngroup != npgroup) return fail("Length of 'group' must equal 'pgroup'.");
    CELDA_TRY(check_labels_host(group, ngroup, L, "group", "L"));
    CELDA_TRY(check_labels_host(pgroup, npgroup, L, "pgroup", "L"));
    if (px_ncol != ncol) return fail("'px' and 'counts' must have the same number of rows.");
    if (L > nrow) return fail("'L' cannot be bigger than the number of rows in 'counts'.");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    CscDev m;
    DevBuf<int> g, pg;
    DevBuf<double> dpx;
    CELDA_TRY(m.upload(ncol, p, i, x));
    CELDA_TRY(g.upload(group, (size_t)ngroup));
    CELDA_TRY(pg.upload(pgroup, (size_t)npgroup));
    CELDA_TRY(dpx.upload(px, (size_t)L * ncol));
    k_rowsum_change_csc<double><<<std::min((ncol + 7) / 8, sms * 16), 256>>>(ncol, m.p.p, m.i.p, m.x.p, g.p, pg.p, L, dpx.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dpx.download(px, (size_t)L * ncol));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------ dense twins
#define NA_INT_ (-2147483647 - 1)
static int has_na(const int* g, long long n) {
    for (long long i = 0; i < n; ++i)
        if (g[i] == NA_INT_) return 1;
    return 0;
}
static int in_levels(const int* g, long long n, int nl) {
    for (long long i = 0; i < n; ++i)
        if (g[i] < 1 || g[i] > nl) return 0;
    return 1;
}

}  // extern "C"
template <typename T>
static int dense_sum(bool byRow, const T* x, int nr, int nc, const int* group, long long ngroup, int nl, T* ans) {
    if (nl < 0) return fail("The grouping argument must be a factor");
    if (byRow && ngroup != nr)
        return fail("The length of the grouping argument must match the number of rows in the matrix");
    if (!byRow && ngroup != nc)
        return fail("The length of the grouping argument must match the number of columns in the matrix");
    if (has_na(group, ngroup)) return fail("Labels in group and pgroup must not be NA.");
    if (!in_levels(group, ngroup, nl)) return fail("factor codes of the grouping argument exceed its number of levels");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    const size_t nout = byRow ? (size_t)nl * nc : (size_t)nr * nl;
    DevBuf<T> dx, dout;
    DevBuf<int> dg;
    CELDA_TRY(dx.upload(x, (size_t)nr * nc));
    CELDA_TRY(dg.upload(group, (size_t)ngroup));
    CELDA_TRY(dout.alloc(nout));
    if (byRow) k_dense_rowsum<T><<<grid_for((long long)nout, 128, sms, 16), 128>>>(dx.p, nr, nc, dg.p, nl, dout.p);
    else k_dense_colsum<T><<<grid_for((long long)nout, 128, sms, 16), 128>>>(dx.p, nr, nc, dg.p, nl, dout.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dout.download(ans, nout));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
extern "C" {

}  // extern "C"
template <typename T>
static int dense_change(bool byRow, const T* x, int nr, int nc, T* px, int px_nr, int px_nc, const int* group,
                        long long ngroup, int nl, const int* pgroup, long long npgroup, int nlp) {
    if (nl < 0 || nlp < 0) return fail("The grouping arguments must be factors");
    if (byRow) {
        if (nl != nlp || nl != px_nr)
            return fail("group and pgroup must have the same number of levels equal to row number of px");
        if (nc != px_nc) return fail("x and the previously summed matrix, px, must have the same number of columns.");
        if (ngroup != npgroup || ngroup != nr)
            return fail("group label and previous group label must be the same length as the number of rows in x.");
    } else {
        if (nl != nlp || nl != px_nc)
            return fail("group and pgroup must have the same number of levels equal to column number of px");
        if (nr != px_nr) return fail("x and the previously summed matrix, pxc must have the same number of rows");
        if (ngroup != npgroup || ngroup != nc)
            return fail("group label and previous group label must be the same length as the number of columns in x.");
    }
    if (has_na(group, ngroup) || has_na(pgroup, npgroup)) return fail("Labels in group and pgroup must not be NA.");
    if (!in_levels(group, ngroup, nl) || !in_levels(pgroup, npgroup, nl))
        return fail("factor codes of the grouping arguments exceed their number of levels");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    const size_t nout = (size_t)px_nr * px_nc;
    DevBuf<T> dx, dpx;
    DevBuf<int> dg, dpg;
    CELDA_TRY(dx.upload(x, (size_t)nr * nc));
    CELDA_TRY(dpx.upload(px, nout));
    CELDA_TRY(dg.upload(group, (size_t)ngroup));
    CELDA_TRY(dpg.upload(pgroup, (size_t)npgroup));
    if (byRow) k_dense_rowsum_change<T><<<grid_for((long long)nout, 128, sms, 16), 128>>>(dx.p, nr, nc, dpx.p, dg.p, dpg.p, nl);
    else k_dense_colsum_change<T><<<grid_for((long long)nout, 128, sms, 16), 128>>>(dx.p, nr, nc, dpx.p, dg.p, dpg.p, nl);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dpx.download(px, nout));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
extern "C" {

int celda_rowSumByGroup_int(const int* x, int nr, int nc, const int* group, long long ngroup, int nlevels, int* ans) {
    return dense_sum<int>(true, x, nr, nc, group, ngroup, nlevels, ans);
}
int celda_rowSumByGroup_numeric(const double* x, int nr, int nc, const int* group, long long ngroup, int nlevels,
                                double* ans) {
    return dense_sum<double>(true, x, nr, nc, group, ngroup, nlevels, ans);
}
int celda_colSumByGroup_int(const int* x, int nr, int nc, const int* group, long long ngroup, int nlevels, int* ans) {
    return dense_sum<int>(false, x, nr, nc, group, ngroup, nlevels, ans);
}
int celda_colSumByGroup_numeric(const double* x, int nr, int nc, const int* group, long long ngroup, int nlevels,
                                double* ans) {
    return dense_sum<double>(false, x, nr, nc, group, ngroup, nlevels, ans);
}
int celda_rowSumByGroupChange_int(const int* x, int nr, int nc, int* px, int px_nr, int px_nc, const int* group,
                                  long long ngroup, int nlevels, const int* pgroup, long long npgroup, int pnlevels) {
    return dense_change<int>(true, x, nr, nc, px, px_nr, px_nc, group, ngroup, nlevels, pgroup, npgroup, pnlevels);
}
int celda_rowSumByGroupChange_numeric(const double* x, int nr, int nc, double* px, int px_nr, int px_nc,
                                      const int* group, long long ngroup, int nlevels, const int* pgroup,
                                      long long npgroup, int pnlevels) {
    return dense_change<double>(true, x, nr, nc, px, px_nr, px_nc, group, ngroup, nlevels, pgroup, npgroup, pnlevels);
}
int celda_colSumByGroupChange_int(const int* x, int nr, int nc, int* px, int px_nr, int px_nc, const int* group,
                                  long long ngroup, int nlevels, const int* pgroup, long long npgroup, int pnlevels) {
    return dense_change<int>(false, x, nr, nc, px, px_nr, px_nc, group, ngroup, nlevels, pgroup, npgroup, pnlevels);
}
int celda_colSumByGroupChange_numeric(const double* x, int nr, int nc, double* px, int px_nr, int px_nc,
                                      const int* group, long long ngroup, int nlevels, const int* pgroup,
                                      long long npgroup, int pnlevels) {
    return dense_change<double>(false, x, nr, nc, px, px_nr, px_nc, group, ngroup, nlevels, pgroup, npgroup, pnlevels);
}

// ------------------------------------------------------------------------------------------------ a10 / a9 helpers
int celda_cG_CalcGibbsProbY(int index, const double* counts, int ncol, const double* nTSbyC, const double* nbyTS,
                            const int* nGbyTS, const double* nbyG, const int* y, int L, int nG, double beta,
                            double gamma, double delta, double* probs) {
    if (index < 1 || index > nG) return fail("'index' must be between 1 and nG");
    if (y[index - 1] < 1 || y[index - 1] > L) return fail("module label of the indexed feature is out of range");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<double> dc, dt, dts, dg, dp;
    DevBuf<int> dgt, dy;
    CELDA_TRY(dc.upload(counts, ncol));
    CELDA_TRY(dt.upload(nTSbyC, (size_t)L * ncol));
    CELDA_TRY(dts.upload(nbyTS, L));
    CELDA_TRY(dgt.upload(nGbyTS, L));
    CELDA_TRY(dg.upload(nbyG, nG));
    CELDA_TRY(dy.upload(y, nG));
    CELDA_TRY(dp.alloc(L));
    k_cG_CalcGibbsProbY<<<(L + 63) / 64, 64>>>(index - 1, dc.p, ncol, dt.p, dts.p, dgt.p, dg.p, dy.p, L, beta, gamma,
                                                delta, dp.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dp.download(probs, L));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_fastNormPropLog(const double* counts, int nr, int nc, double alpha, double* res) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<double> dx, dr;
    DevBuf<int> flag;
    CELDA_TRY(dx.upload(counts, (size_t)nr * nc));
    CELDA_TRY(dr.alloc((size_t)nr * nc));
    CELDA_TRY(flag.alloc(1));
    CELDA_TRY(flag.zero());
    k_fastNormPropLog<<<std::max(1, std::min(nc, num_sms(g_device) * 8)), 256>>>(dx.p, nr, nc, alpha, dr.p, flag.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    int f = 0;
    CELDA_CUDA_OK(cudaMemcpy(&f, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (f)
        return fail("Division by 0. Make sure colSums of counts does not contain 0 after rounding counts to integers.");
    CELDA_TRY(dr.download(res, (size_t)nr * nc));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
template <typename TB>
static int mat_mult_t(const double* A, int n, int k, const TB* B, int m, double* Cm) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<double> dA, dC;
    DevBuf<TB> dB;
    CELDA_TRY(dA.upload(A, (size_t)n * k));
    CELDA_TRY(dB.upload(B, (size_t)n * m));
    CELDA_TRY(dC.alloc((size_t)k * m));
    k_matMultT<TB><<<grid_for((long long)k * m, 128, num_sms(g_device), 16), 128>>>(dA.p, n, k, dB.p, m, dC.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dC.download(Cm, (size_t)k * m));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
extern "C" {
int celda_eigenMatMultInt(const double* A, int n, int k, const int* B, int m, double* Cm) {
    return mat_mult_t<int>(A, n, k, B, m, Cm);
}
int celda_eigenMatMultNumeric(const double* A, int n, int k, const double* B, int m, double* Cm) {
    return mat_mult_t<double>(A, n, k, B, m, Cm);
}

int celda_perplexityG(const int* x, int nr, int nc, const double* phi, int phi_nr, int phi_nc, const double* psi,
                      int psi_nr, int psi_nc, const int* group, long long ngroup, int nlevels, double* ans) {
    const int nl = nlevels;
    if (nl < 0) return fail("The grouping argument must be a factor");
    if (ngroup != nr) return fail("The length of the grouping argument must match the number of rows in the matrix.");
    if (phi_nc != nc) return fail("The R_phi and R_x must have the same number of colums.");
    if (phi_nr != nl) return fail("R_phi must have the same number of rows as the number of levels in R_group.");
    if (psi_nr != nr) return fail("The R_psi and R_x must have the same number of rows.");
    if (psi_nc != nl) return fail("R_phi must have the same number of columns as the number of levels in R_group.");
    for (long long i = 0; i < ngroup; ++i)
        if (group[i] == NA_INT_ || group[i] < 1 || group[i] > nl)
            return fail("Labels in group and pgroup must not be NA and must less than or equal to the number of rows in the matrix.");
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<int> dx, dg;
    DevBuf<double> dphi, dpsi, part, out;
    CELDA_TRY(dx.upload(x, (size_t)nr * nc));
    CELDA_TRY(dg.upload(group, (size_t)ngroup));
    CELDA_TRY(dphi.upload(phi, (size_t)phi_nr * phi_nc));
    CELDA_TRY(dpsi.upload(psi, (size_t)psi_nr * psi_nc));
    const int NB = 128;
    CELDA_TRY(part.alloc(NB));
    CELDA_TRY(out.alloc(1));
    k_perplexityG<<<NB, 256>>>(dx.p, nr, nc, dphi.p, dpsi.p, dg.p, nl, part.p);
    k_sum_partials<<<1, 32>>>(part.p, NB, out.p);
    count_launch(2);
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(out.download(ans, 1));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------ stateless LL / sweeps
}  // extern "C"
namespace {
// host double array (integer-valued) -> device int32 / int64, with an integrality check
int up_int(const double* h, size_t n, DevBuf<int>& out, const char* what) {
    DevBuf<double> d;
    DevBuf<int> flag;
    CELDA_TRY(d.upload(h, n));
    CELDA_TRY(out.alloc(n));
    CELDA_TRY(flag.alloc(1));
    CELDA_TRY(flag.zero());
    k_narrow<<<grid_for((long long)n, 256, num_sms(g_device)), 256>>>(d.p, (long long)n, out.p, flag.p);
    count_launch();
    int f = 0;
    CELDA_CUDA_OK(cudaMemcpy(&f, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (f) return fail("'%s' must hold integer-valued counts below 2^31", what);
    return 0;
}
int up_ll(const double* h, size_t n, DevBuf<long long>& out, const char* what) {
    std::vector<long long> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i] = (long long)h[i];
        if ((double)v[i] != h[i]) return fail("'%s' must hold integer-valued counts", what);
    }
    return out.upload(v.data(), n);
}
template <typename T>
int down_d(const T* dev, size_t n, double* host) {
    DevBuf<double> tmp;
    CELDA_TRY(tmp.alloc(n));
    k_convert<T, double><<<grid_for((long long)n, 256, num_sms(g_device)), 256>>>(dev, (long long)n, tmp.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(tmp.download(host, n));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
int ll_finish(const LLIn& in, double* ll) {
    DevBuf<double> partial, out;
    CELDA_TRY(partial.alloc(1024));
    CELDA_TRY(out.alloc(1));
    CELDA_TRY(run_loglik(in, partial.p, out.p, 0));
    CELDA_TRY(out.download(ll, 1));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}
}  // namespace
extern "C" {

int celda_cCGCalcLL(int K, int L, const int* mCPByS, const double* nTSByCP, const double* nByG, int nGenes,
                    const double* nByTS, const int* nGByTS, int nS, double alpha, double beta, double delta,
                    double gamma, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<int> m, t, gt;
    DevBuf<long long> g, ts;
    CELDA_TRY(m.upload(mCPByS, (size_t)K * nS));
    CELDA_TRY(up_int(nTSByCP, (size_t)L * K, t, "nTSByCP"));
    CELDA_TRY(up_ll(nByG, nGenes, g, "nByG"));
    CELDA_TRY(up_ll(nByTS, L, ts, "nByTS"));
    CELDA_TRY(gt.upload(nGByTS, L));
    LLIn in{CELDA_MODEL_CG, K, L, nS, nGenes, 0, m.p, t.p, nullptr, nullptr, gt.p, g.p, ts.p, alpha, beta, delta, gamma};
    return ll_finish(in, ll);
}

int celda_cCCalcLL(const int* mCPByS, const double* nGByCP, int K, int nS, int nG, double alpha, double beta,
                   double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<int> m, ncm, n;
    CELDA_TRY(m.upload(mCPByS, (size_t)K * nS));
    CELDA_TRY(up_int(nGByCP, (size_t)nG * K, ncm, "nGByCP"));
    CELDA_TRY(n.alloc((size_t)nG * K));
    k_transpose_convert<int, int><<<grid_for((long long)nG * K, 256, num_sms(g_device)), 256>>>(ncm.p, nG, K, n.p, false);
    count_launch();
    LLIn in{CELDA_MODEL_C, K, 1, nS, nG, 0, m.p, nullptr, n.p, nullptr, nullptr, nullptr, nullptr, alpha, beta, 1.0, 1.0};
    return ll_finish(in, ll);
}

int celda_cGCalcLL(const double* nTSByC, const double* nByTS, const double* nByG, int nGenes, const int* nGByTS,
                   long long nM, int L, double beta, double delta, double gamma, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<int> t, gt;
    DevBuf<long long> g, ts;
    CELDA_TRY(up_int(nTSByC, (size_t)L * nM, t, "nTSByC"));
    CELDA_TRY(up_ll(nByG, nGenes, g, "nByG"));
    CELDA_TRY(up_ll(nByTS, L, ts, "nByTS"));
    CELDA_TRY(gt.upload(nGByTS, L));
    LLIn in{CELDA_MODEL_G, 1, L, 1, nGenes, nM, nullptr, nullptr, nullptr, t.p, gt.p, g.p, ts.p, 1.0, beta, delta, gamma};
    return ll_finish(in, ll);
}

int celda_cCCalcGibbsProbZ(const double* counts, int* mCPByS, int nS, double* nGByCP, const double* nByC, double* nCP,
                           int* z, const int* s, int K, int nG, long long nM, double alpha, double beta, int doSample,
                           const int* ix, const double* u, double* probs) {
    if (K < 1 || K > 32767) return fail("'K' must be between 1 and 32767");
    CELDA_TRY(check_labels_host(z, nM, K, "z", "K"));
    CELDA_TRY(check_labels_host(s, nM, nS, "s", "nS"));
    CELDA_TRY(validate_order(ix, nM, "ix"));
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    DevBuf<int> dc, dm, dn, dbc, dz, ds, dix;
    DevBuf<long long> dcp;
    DevBuf<double> du, dprobs;
    SweepBufs wb;
    CELDA_TRY(up_int(counts, (size_t)nG * nM, dc, "counts"));
    CELDA_TRY(dm.upload(mCPByS, (size_t)K * nS));
    CELDA_TRY(up_int(nGByCP, (size_t)nG * K, dn, "nGByCP"));
    CELDA_TRY(up_int(nByC, (size_t)nM, dbc, "nByC"));
    CELDA_TRY(up_ll(nCP, K, dcp, "nCP"));
    CELDA_TRY(dz.upload(z, (size_t)nM));
    CELDA_TRY(ds.upload(s, (size_t)nM));
    CELDA_TRY(dix.upload(ix, (size_t)nM));
    CELDA_TRY(du.upload(u, (size_t)nM));
    if (probs) CELDA_TRY(dprobs.alloc((size_t)K * nM));
    CELDA_TRY(wb.alloc(K, 0));
    if (z_sweep_smem(nG, K, nS, wb.lgT, 16, 0, 1) > 220 * 1024) {
        // celda_C shapes (counts = the raw G x C matrix, R/celda_C.R:427-440): the G x K table does not fit shared memory,
        // so the sweep runs on the raw-count kernel (k_sweep_zc) over a CSC copy of the dense argument; same arithmetic
        // (the reference sums all G genes per candidate, zeros included)
        std::vector<int> hp((size_t)nM + 1, 0), hi, hx;
        for (long long c = 0; c < nM; ++c) {
            for (int g = 0; g < nG; ++g) {
                const double v = counts[(size_t)c * nG + g];
                if (v != 0.0) {
                    hi.push_back(g);
                    hx.push_back((int)v);
                }
            }
            if (hi.size() > 0x7fffffffULL) return fail("cCCalcGibbsProbZ: more than 2^31 - 1 stored counts");
            hp[(size_t)c + 1] = (int)hi.size();
        }
        DevBuf<int> dp_, di_, dx_, dng;
        DevBuf<double> dlg;
        CELDA_TRY(dp_.upload(hp.data(), hp.size()));
        CELDA_TRY(di_.upload(hi.data(), std::max<size_t>(hi.size(), 1)));
        CELDA_TRY(dx_.upload(hx.data(), std::max<size_t>(hx.size(), 1)));
        CELDA_TRY(dng.alloc((size_t)nG * K));
        CELDA_TRY(dlg.alloc((size_t)nG * K));
        k_transpose_convert<int, int><<<grid_for((long long)nG * K, 256, sms), 256>>>(dn.p, nG, K, dng.p, false);
        k_lgt_fill<<<grid_for((long long)nG * K, 256, sms), 256>>>(dng.p, (long long)nG * K, beta, dlg.p);
        count_launch(2);
        ZCArgs c{};
        c.nG = nG;
        c.K = K;
        c.nS = nS;
        c.nM = nM;
        c.cp = dp_.p;
        c.ci = di_.p;
        c.cx = dx_.p;
        c.nByC = dbc.p;
        c.s = ds.p;
        c.z = dz.p;
        c.n = dng.p;
        c.lgn = dlg.p;
        c.nCP = dcp.p;
        c.mCPByS = dm.p;
        c.alpha = alpha;
        c.beta = beta;
        c.ix = dix.p;
        c.u = du.p;
        c.doSample = doSample;
        c.probs = probs ? dprobs.p : nullptr;
        CELDA_TRY(run_sweep_zc(c, wb, sms, 0, 0));
        k_transpose_convert<int, int><<<grid_for((long long)nG * K, 256, sms), 256>>>(dng.p, nG, K, dn.p, true);
        count_launch();
        int stv = 0;
        CELDA_CUDA_OK(cudaMemcpy(&stv, wb.status.p, sizeof(int), cudaMemcpyDeviceToHost));
        CELDA_TRY(sweep_status(stv));
        CELDA_TRY(dz.download(z, (size_t)nM));
        CELDA_TRY(dm.download(mCPByS, (size_t)K * nS));
        CELDA_TRY(down_d<int>(dn.p, (size_t)nG * K, nGByCP));
        CELDA_TRY(down_d<long long>(dcp.p, K, nCP));
        if (probs) CELDA_TRY(dprobs.download(probs, (size_t)K * nM));
        CELDA_CUDA_OK(cudaDeviceSynchronize());
        return 0;
    }
    ZArgs a{};
    a.R = nG;
    a.K = K;
    a.nS = nS;
    a.nM = nM;
    a.counts = dc.p;
    a.nByC = dbc.p;
    a.s = ds.p;
    a.z = dz.p;
    a.nTab = dn.p;
    a.nCP = dcp.p;
    a.mCPByS = dm.p;
    a.alpha = alpha;
    a.beta = beta;
    a.ix = dix.p;
    a.u = du.p;
    a.doSample = doSample;
    a.probs = probs ? dprobs.p : nullptr;
    CELDA_TRY(run_sweep_z(a, wb, sms, 0, 0));
    int stv = 0;
    CELDA_CUDA_OK(cudaMemcpy(&stv, wb.status.p, sizeof(int), cudaMemcpyDeviceToHost));
    CELDA_TRY(sweep_status(stv));
    CELDA_TRY(dz.download(z, (size_t)nM));
    CELDA_TRY(dm.download(mCPByS, (size_t)K * nS));
    CELDA_TRY(down_d<int>(dn.p, (size_t)nG * K, nGByCP));
    CELDA_TRY(down_d<long long>(dcp.p, K, nCP));
    if (probs) CELDA_TRY(dprobs.download(probs, (size_t)K * nM));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_cGCalcGibbsProbY(const double* counts, int ncol, double* nTSByC, double* nByTS, int* nGByTS,
                           const double* nByG, int* y, int L, int nG, double beta, double delta, double gamma,
                           int doSample, const int* ix, const double* u, double* probs) {
    if (L < 1 || L > 32767) return fail("'L' must be between 1 and 32767");
    CELDA_TRY(check_labels_host(y, nG, L, "y", "L"));
    CELDA_TRY(validate_order(ix, nG, "ix"));
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    DevBuf<int> dcm, dc, dt, dgt, dy, dix;
    DevBuf<long long> dts, dg;
    DevBuf<double> du, dprobs;
    SweepBufs wb;
    CELDA_TRY(up_int(counts, (size_t)nG * ncol, dcm, "counts"));
    CELDA_TRY(dc.alloc((size_t)nG * ncol));
    k_transpose_convert<int, int><<<grid_for((long long)nG * ncol, 256, sms), 256>>>(dcm.p, nG, ncol, dc.p, false);
    count_launch();
    CELDA_TRY(up_int(nTSByC, (size_t)L * ncol, dt, "nTSByC"));
    CELDA_TRY(up_ll(nByTS, L, dts, "nByTS"));
    CELDA_TRY(dgt.upload(nGByTS, L));
    CELDA_TRY(up_ll(nByG, nG, dg, "nByG"));
    CELDA_TRY(dy.upload(y, nG));
    CELDA_TRY(dix.upload(ix, nG));
    CELDA_TRY(du.upload(u, nG));
    if (probs) CELDA_TRY(dprobs.alloc((size_t)L * nG));
    CELDA_TRY(wb.alloc(L, 0));
    YArgs a{};
    a.nG = nG;
    a.ncol = ncol;
    a.L = L;
    a.counts = dc.p;
    a.nByG = dg.p;
    a.y = dy.p;
    a.tTab = dt.p;
    a.nByTS = dts.p;
    a.nGByTS = dgt.p;
    a.beta = beta;
    a.gamma = gamma;
    a.delta = delta;
    a.ix = dix.p;
    a.u = du.p;
    a.doSample = doSample;
    a.probs = probs ? dprobs.p : nullptr;
    CELDA_TRY(run_sweep_y(a, wb, sms, 0, 0));
    int stv = 0;
    CELDA_CUDA_OK(cudaMemcpy(&stv, wb.status.p, sizeof(int), cudaMemcpyDeviceToHost));
    CELDA_TRY(sweep_status(stv));
    CELDA_TRY(dy.download(y, nG));
    CELDA_TRY(dgt.download(nGByTS, L));
    CELDA_TRY(down_d<int>(dt.p, (size_t)L * ncol, nTSByC));
    CELDA_TRY(down_d<long long>(dts.p, L, nByTS));
    if (probs) CELDA_TRY(dprobs.download(probs, (size_t)L * nG));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_cCCalcEMProbZ(int nG, long long nM, const int* p, const int* i, const double* x, int* mCPByS, int nS,
                        double* nGByCP, double* nCP, int* z, const int* s, int K, double alpha, double beta,
                        int doSample, double* probs) {
    if (K < 1 || K > 256) return fail("'K' must be between 1 and 256 for the EM step");
    CELDA_TRY(check_labels_host(z, nM, K, "z", "K"));
    CELDA_TRY(check_labels_host(s, nM, nS, "s", "nS"));
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    const int sms = num_sms(g_device);
    const bool sparse = p != nullptr;
    DevBuf<int> dp, di, dx, dm, dncm, dn, dz, dzp, ds, flag;
    DevBuf<long long> dcp;
    DevBuf<double> phi, theta, dprobs;
    if (sparse) {
        CELDA_TRY(dp.upload(p, (size_t)nM + 1));
        CELDA_TRY(di.upload(i, (size_t)p[nM]));
        CELDA_TRY(up_int(x, (size_t)p[nM], dx, "counts"));
    } else {
        CELDA_TRY(up_int(x, (size_t)nG * nM, dx, "counts"));
    }
    CELDA_TRY(dm.upload(mCPByS, (size_t)K * nS));
    CELDA_TRY(up_int(nGByCP, (size_t)nG * K, dncm, "nGByCP"));
    CELDA_TRY(up_ll(nCP, K, dcp, "nCP"));
    CELDA_TRY(dz.upload(z, (size_t)nM));
    CELDA_TRY(dzp.upload(z, (size_t)nM));
    CELDA_TRY(ds.upload(s, (size_t)nM));
    CELDA_TRY(phi.alloc((size_t)nG * K));
    CELDA_TRY(theta.alloc((size_t)K * nS));
    CELDA_TRY(dprobs.alloc((size_t)K * nM));
    CELDA_TRY(flag.alloc(1));
    CELDA_TRY(flag.zero());
    k_em_phi<<<grid_for((long long)nG * K, 256, sms), 256>>>(dncm.p, false, nG, K, dcp.p, beta, phi.p, flag.p);
    k_em_theta<<<std::min(nS, 1024), 64>>>(dm.p, K, nS, alpha, theta.p, flag.p);
    const int threads = 256, nwarp = threads / 32;
    const int grid = (int)std::min<long long>((nM + nwarp - 1) / nwarp, (long long)sms * 32);
    if (sparse)
        k_em_probs<true><<<grid, threads>>>(nG, nM, K, dp.p, di.p, dx.p, phi.p, theta.p, ds.p, dprobs.p,
                                            doSample ? dz.p : nullptr);
    else
        k_em_probs<false><<<grid, threads>>>(nG, nM, K, nullptr, nullptr, dx.p, phi.p, theta.p, ds.p, dprobs.p,
                                             doSample ? dz.p : nullptr);
    count_launch(3);
    int f = 0;
    CELDA_CUDA_OK(cudaMemcpy(&f, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (f) return fail("Division by 0. Make sure colSums of counts does not contain 0 after rounding counts to integers.");
    if (doSample) {
        if (sparse) {
            // gene-major working copy for the sparse change kernel
            CELDA_TRY(dn.alloc((size_t)nG * K));
            k_transpose_convert<int, int><<<grid_for((long long)nG * K, 256, sms), 256>>>(dncm.p, nG, K, dn.p, false);
            k_colsum_change_csc<int><<<(int)std::min<long long>((nM + nwarp - 1) / nwarp, (long long)sms * 16), threads>>>(
                (int)nM, dp.p, di.p, dx.p, dz.p, dzp.p, K, dn.p);
            k_transpose_convert<int, int><<<grid_for((long long)nG * K, 256, sms), 256>>>(dn.p, nG, K, dncm.p, true);
            count_launch(3);
        } else {
            k_colsum_change_dense<<<grid_for((long long)nG * nM, 256, sms), 256>>>(dx.p, nG, nM, dz.p, dzp.p, dncm.p);
            count_launch();
        }
        k_colsums_small<<<1, 256>>>(dncm.p, nG, K, dcp.p);
        CELDA_TRY(dm.zero());
        k_table_zs<<<grid_for(nM, 256, sms), 256>>>(dz.p, ds.p, nM, K, dm.p);
        count_launch(2);
        CELDA_CUDA_OK(cudaGetLastError());
        CELDA_TRY(dz.download(z, (size_t)nM));
        CELDA_TRY(dm.download(mCPByS, (size_t)K * nS));
        CELDA_TRY(down_d<int>(dncm.p, (size_t)nG * K, nGByCP));
        CELDA_TRY(down_d<long long>(dcp.p, K, nCP));
    }
    CELDA_CUDA_OK(cudaGetLastError());
    if (probs) CELDA_TRY(dprobs.download(probs, (size_t)K * nM));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------------------------------------ chain
int celda_chain_create(celda_chain** out, int device, int model, int nG, int nM, const int* p, const int* i,
                       const double* x, const int* s, int nS, int K, int L, double alpha, double beta, double delta,
                       double gamma) {
    *out = nullptr;
    if (model < 0 || model > 2) return fail("unknown model %d", model);
    if (nG < 1 || nM < 1) return fail("counts must have at least one row and one column");
    if (model != CELDA_MODEL_G && (K < 1 || K > nM)) return fail("'K' must be between 1 and the number of cells");
    if (model != CELDA_MODEL_C && (L < 1 || L > nG)) return fail("'L' must be between 1 and the number of features");
    // the sampler keeps one bit per candidate and lane in a 32-bit word (warp_sample_ll): at most 1024 candidates
    if (K > 1024 || L > 1024) return fail("K and L must not exceed 1024");
    if (p[0] != 0) return fail("column pointers of 'counts' must start at 0");
    if (model != CELDA_MODEL_G) {
        if (!s) return fail("sample labels are required");
        CELDA_TRY(check_labels_host(s, nM, nS, "s", "nS"));
    }
    const long long Z = p[nM];
    for (int c = 0; c < nM; ++c)
        if (p[c + 1] < p[c]) return fail("column pointers of 'counts' must be non-decreasing");
    CELDA_CUDA_OK(cudaSetDevice(device));
    std::unique_ptr<celda_chain> h(new celda_chain());
    h->device = device;
    h->model = model;
    h->nG = nG;
    h->nM = nM;
    h->nS = (model == CELDA_MODEL_G) ? 1 : nS;
    h->K = (model == CELDA_MODEL_G) ? 1 : K;
    h->L = (model == CELDA_MODEL_C) ? 1 : L;
    h->Z = Z;
    h->alpha = alpha;
    h->beta = beta;
    h->delta = delta;
    h->gamma = gamma;
    h->sms = num_sms(device);
    CELDA_CUDA_OK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CELDA_CUDA_OK(cudaEventCreate(&h->ev0));
    CELDA_CUDA_OK(cudaEventCreate(&h->ev1));
    cudaStream_t st = h->st;
    // ---- a1: GPU-resident CSC copy, values narrowed to int32 after an integrality check
    CELDA_TRY(h->cp.upload(p, (size_t)nM + 1, st));
    CELDA_TRY(h->ci.upload(i, (size_t)Z, st));
    CELDA_TRY(h->cx.alloc((size_t)Z));
    {
        DevBuf<double> dx;
        DevBuf<int> flag;
        CELDA_TRY(dx.upload(x, (size_t)Z, st));
        CELDA_TRY(flag.alloc(1));
        CELDA_TRY(flag.zero(st));
        k_narrow<<<grid_for(Z, 256, h->sms), 256, 0, st>>>(dx.p, Z, h->cx.p, flag.p);
        count_launch();
        int f = 0;
        CELDA_CUDA_OK(cudaMemcpyAsync(&f, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
        if (f) return fail("counts must be integer-valued and below 2^31 (found a non-integer entry)");
    }
    for (long long t = 0; t < Z; ++t)
        if (i[t] < 0 || i[t] >= nG) return fail("row index out of range in 'counts'");
    if (model != CELDA_MODEL_G) CELDA_TRY(h->s.upload(s, (size_t)nM, st));
    CELDA_TRY(h->z.alloc(nM));
    CELDA_TRY(h->zprev.alloc(nM));
    CELDA_TRY(h->y.alloc(nG));
    CELDA_TRY(h->yprev.alloc(nG));
    CELDA_TRY(h->nByC.alloc(nM));
    CELDA_TRY(h->nByG.alloc(nG));
    CELDA_TRY(h->nByG.zero(st));
    k_rowtotals_csc<<<grid_for(Z, 256, h->sms), 256, 0, st>>>(Z, h->ci.p, h->cx.p, (unsigned long long*)h->nByG.p);
    {
        // nByC (Matrix::colSums): reuse the row-sum kernel with a single group
        DevBuf<int> ones, tmp;
        CELDA_TRY(ones.alloc(nG));
        CELDA_TRY(tmp.alloc(nM));
        k_fill_int<<<grid_for(nG, 256, h->sms), 256, 0, st>>>(ones.p, nG, 1);
        k_rowsum_csc<int><<<std::min((nM + 7) / 8, h->sms * 8), 256, sizeof(int) * 8, st>>>(nM, h->cp.p, h->ci.p, h->cx.p,
                                                                                         ones.p, 1, tmp.p, h->nByC.p);
        count_launch(3);
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (model != CELDA_MODEL_C) {
        // gene-major twin: stable radix sort of (row index, position) pairs keeps each gene's entries in ascending
        // column order; row pointers from a histogram scanned on the host (nG values).
        DevBuf<int> colOf, iota, keysOut, perm, cnt;
        DevBuf<unsigned char> tmp;
        CELDA_TRY(colOf.alloc((size_t)Z));
        CELDA_TRY(iota.alloc((size_t)Z));
        CELDA_TRY(keysOut.alloc((size_t)Z));
        CELDA_TRY(perm.alloc((size_t)Z));
        CELDA_TRY(cnt.alloc(nG));
        CELDA_TRY(cnt.zero(st));
        k_csc_cols<<<std::min((nM + 7) / 8, h->sms * 16), 256, 0, st>>>(nM, h->cp.p, colOf.p, iota.p);
        k_row_counts<<<grid_for(Z, 256, h->sms), 256, 0, st>>>(Z, h->ci.p, cnt.p);
        int bits = 1;
        while ((1LL << bits) < nG) ++bits;
        size_t bytes = 0;
        CELDA_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->ci.p, keysOut.p, iota.p, perm.p, (int)Z, 0, bits, st));
        CELDA_TRY(tmp.alloc(bytes));
        CELDA_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, h->ci.p, keysOut.p, iota.p, perm.p, (int)Z, 0, bits, st));
        CELDA_TRY(h->tcol.alloc((size_t)Z));
        CELDA_TRY(h->tx.alloc((size_t)Z));
        k_gather_twin<<<grid_for(Z, 256, h->sms), 256, 0, st>>>(Z, perm.p, colOf.p, h->cx.p, h->tcol.p, h->tx.p);
        count_launch(7);
        std::vector<int> hc(nG);
        CELDA_TRY(cnt.download(hc.data(), nG, st));
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
        std::vector<long long> rp((size_t)nG + 1, 0);
        for (int g = 0; g < nG; ++g) rp[g + 1] = rp[g] + hc[g];
        CELDA_TRY(h->trp.upload(rp.data(), (size_t)nG + 1, st));
        CELDA_TRY(h->changed.alloc(nG));
        CELDA_TRY(h->nchanged.alloc(1));
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
    }
    CELDA_TRY(chain_alloc_tables(h.get()));
    CELDA_TRY(h->ixy.alloc(nG));
    CELDA_TRY(h->uy.alloc(nG));
    CELDA_TRY(h->ixz.alloc(nM));
    CELDA_TRY(h->uz.alloc(nM));
    CELDA_TRY(h->llPartial.alloc(1024));
    CELDA_TRY(h->emFlag.alloc(1));
    CELDA_TRY(h->llOut.alloc(1));
    CELDA_CUDA_OK(cudaStreamSynchronize(st));
    CELDA_CUDA_OK(cudaGetLastError());
    *out = h.release();
    return 0;
}

void celda_chain_destroy(celda_chain* h) { delete h; }

// celdaGridSearch runs every (K, L, chain) combination on the same counts (R/celdaGridSearch.R:290-391): the resident
// CSC copy, its gene-major twin and the totals are kept, only the (K, L)-sized tables are re-allocated.  The handle
// needs set_state before the next sweep.
int celda_chain_reconfigure(celda_chain* h, int K, int L) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_TRY(chain_sync(h));
    if (h->comm) return fail("reconfigure: not available on a cell-sharded chain");
    if (h->model != CELDA_MODEL_G && (K < 1 || K > h->nM)) return fail("'K' must be between 1 and the number of cells");
    if (h->model != CELDA_MODEL_C && (L < 1 || L > h->nG)) return fail("'L' must be between 1 and the number of features");
    if (K > 1024 || L > 1024) return fail("K and L must not exceed 1024");
    h->K = (h->model == CELDA_MODEL_G) ? 1 : K;
    h->L = (h->model == CELDA_MODEL_C) ? 1 : L;
    CELDA_TRY(chain_alloc_tables(h));
    CELDA_CUDA_OK(cudaGetLastError());
    return chain_sync(h);
}

int celda_chain_set_state(celda_chain* h, const int* z, const int* y) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (h->model != CELDA_MODEL_G) {
        if (!z) return fail("z labels are required");
        for (int c = 0; c < h->nM; ++c)
            if (z[c] < 1 || z[c] > h->K)
                return fail("Assigned value of cell cluster greater than the total number of cell clusters!");
        CELDA_CUDA_OK(cudaMemcpyAsync(h->z.p, z, sizeof(int) * h->nM, cudaMemcpyHostToDevice, h->st));
    }
    if (h->model != CELDA_MODEL_C) {
        if (!y) return fail("y labels are required");
        for (int g = 0; g < h->nG; ++g)
            if (y[g] < 1 || y[g] > h->L)
                return fail("Assigned value of feature module greater than the total number of feature modules!");
        CELDA_CUDA_OK(cudaMemcpyAsync(h->y.p, y, sizeof(int) * h->nG, cudaMemcpyHostToDevice, h->st));
    }
    CELDA_TRY(chain_decompose(h));
    CELDA_TRY(chain_sync(h));
    h->has_state = true;
    return 0;
}

int celda_chain_get_state(celda_chain* h, int* z, int* y) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (z && h->model != CELDA_MODEL_G) CELDA_TRY(h->z.download(z, h->nM, h->st));
    if (y && h->model != CELDA_MODEL_C) CELDA_TRY(h->y.download(y, h->nG, h->st));
    return chain_sync(h);
}

int celda_chain_sweep_y(celda_chain* h, const int* ix, const double* u, int doSample) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_upload_order(h, h->ixy.p, ix, h->nG));
    CELDA_CUDA_OK(cudaMemcpyAsync(h->uy.p, u, sizeof(double) * h->nG, cudaMemcpyHostToDevice, h->st));
    CELDA_TRY(chain_step_y(h, h->ixy.p, h->uy.p, doSample));
    return chain_check_status(h);
}

int celda_chain_sweep_z_gibbs(celda_chain* h, const int* ix, const double* u, int doSample) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_upload_order(h, h->ixz.p, ix, h->nM));
    CELDA_CUDA_OK(cudaMemcpyAsync(h->uz.p, u, sizeof(double) * h->nM, cudaMemcpyHostToDevice, h->st));
    CELDA_TRY(chain_step_z(h, h->ixz.p, h->uz.p, doSample));
    return chain_check_status(h);
}

// ---- cell-sharded EM: one process per GPU, one NCCL communicator
int celda_comm_unique_id(char* id128) {
    CELDA_TRY(g_nccl.load());
    ncclUniqueId id;
    CELDA_NCCL_OK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

int celda_comm_create(celda_comm** out, const char* id128, int nranks, int rank, int device) {
    *out = nullptr;
    CELDA_TRY(g_nccl.load());
    CELDA_CUDA_OK(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    std::unique_ptr<celda_comm> c(new celda_comm());
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    CELDA_NCCL_OK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    *out = c.release();
    return 0;
}

void celda_comm_destroy(celda_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

// Turn a celda_C chain into one shard of a cell-sharded chain: its counts are a column range of the full matrix,
// its partial nGByCP / mCPByS tables are all-reduced over `comm` after every (re-)decomposition.  Call before
// set_state; every rank must make the same sequence of calls.
int celda_chain_attach_comm(celda_chain* h, celda_comm* comm) {
    if (h->model != CELDA_MODEL_C) return fail("attach_comm: only the celda_C (EM) chain shards cells");
    if (comm->device != h->device) return fail("attach_comm: communicator and chain live on different devices");
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_TRY(h->nGByCP_loc.alloc((size_t)h->nG * h->K));
    CELDA_TRY(h->mCPByS_loc.alloc((size_t)h->K * h->nS));
    h->comm = comm;
    h->has_state = false;
    return 0;
}

int celda_chain_set_scan_mode(celda_chain* h, int mode) {
    if (h->model != CELDA_MODEL_G) return fail("scan mode: only the celda_G y sweep scans the raw counts");
    if (mode != 0 && mode != 1) return fail("scan mode must be 0 (exact) or 1 (count columns only)");
    h->scanMode = mode;
    h->marginValid = false;
    return 0;
}

int celda_chain_scan_margin(celda_chain* h, double* margin) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->marginValid) return fail("scan margin: no sweep in scan mode 1 (with doSample) since the mode was set");
    unsigned long long bits[3];
    CELDA_TRY(chain_sync(h));
    CELDA_CUDA_OK(cudaMemcpy(bits, h->marginBuf.p, sizeof(bits), cudaMemcpyDeviceToHost));
    double ratio;
    memcpy(&ratio, bits, sizeof(ratio));
    margin[0] = ratio;
    margin[1] = (double)bits[1];
    return 0;
}

int celda_chain_seed(celda_chain* h, unsigned long long seed, unsigned long long stream) {
    h->ph_seed = seed;
    h->ph_stream = stream;
    h->ph_sweep = 0;
    h->ph_seeded = true;
    return 0;
}

int celda_chain_sweep_y_philox(celda_chain* h, int doSample) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_philox_draws(h, h->nG, h->ixy.p, h->uy.p));
    CELDA_TRY(chain_step_y(h, h->ixy.p, h->uy.p, doSample));
    return chain_check_status(h);
}

int celda_chain_sweep_z_gibbs_philox(celda_chain* h, int doSample) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_philox_draws(h, h->nM, h->ixz.p, h->uz.p));
    CELDA_TRY(chain_step_z(h, h->ixz.p, h->uz.p, doSample));
    return chain_check_status(h);
}

// The order and uniforms the last sweep of each kind consumed (injected or generated): which = 0 y, 1 z.
int celda_chain_get_draws(celda_chain* h, int which, int* ix, double* u) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (which == 0) {
        CELDA_TRY(h->ixy.download(ix, h->nG, h->st));
        CELDA_TRY(h->uy.download(u, h->nG, h->st));
    } else {
        CELDA_TRY(h->ixz.download(ix, h->nM, h->st));
        CELDA_TRY(h->uz.download(u, h->nM, h->st));
    }
    return chain_sync(h);
}

int celda_cuda_philox4x32(const int* counters, const int* key, int n, int* out) {
    CELDA_CUDA_OK(cudaSetDevice(g_device));
    DevBuf<unsigned> dc, dk, dout;
    CELDA_TRY(dc.upload(reinterpret_cast<const unsigned*>(counters), (size_t)4 * n));
    CELDA_TRY(dk.upload(reinterpret_cast<const unsigned*>(key), 2));
    CELDA_TRY(dout.alloc((size_t)4 * n));
    k_philox_raw<<<grid_for(n, 128, num_sms(g_device)), 128>>>(dc.p, dk.p, n, dout.p);
    count_launch();
    CELDA_CUDA_OK(cudaGetLastError());
    CELDA_TRY(dout.download(reinterpret_cast<unsigned*>(out), (size_t)4 * n));
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int celda_chain_loglik(celda_chain* h, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_loglik_async(h));
    CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    return chain_sync(h);
}

int celda_chain_iterate(celda_chain* h, const int* ix_y, const double* u_y, const int* ix_z, const double* u_z,
                        double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    cudaStream_t st = h->st;
    if (h->model == CELDA_MODEL_C) {  // .celda_C loop body, algorithm = "Gibbs" (R/celda_C.R:427-470): z sweep, log-likelihood
        if (ix_z && u_z) {
            CELDA_TRY(chain_upload_order(h, h->ixz.p, ix_z, h->nM));
            CELDA_CUDA_OK(cudaMemcpyAsync(h->uz.p, u_z, sizeof(double) * h->nM, cudaMemcpyHostToDevice, st));
        } else {
            CELDA_TRY(chain_philox_draws(h, h->nM, h->ixz.p, h->uz.p));
        }
        CELDA_TRY(chain_step_z(h, h->ixz.p, h->uz.p, 1));
        CELDA_TRY(chain_loglik_async(h));
        CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        return chain_check_status(h);
    }
    if (h->model == CELDA_MODEL_G) {  // .celda_G loop body (R/celda_G.R:400-424): y sweep, log-likelihood
        if (ix_y && u_y) {
            CELDA_TRY(chain_upload_order(h, h->ixy.p, ix_y, h->nG));
            CELDA_CUDA_OK(cudaMemcpyAsync(h->uy.p, u_y, sizeof(double) * h->nG, cudaMemcpyHostToDevice, st));
        } else {
            CELDA_TRY(chain_philox_draws(h, h->nG, h->ixy.p, h->uy.p));
        }
        CELDA_TRY(chain_step_y(h, h->ixy.p, h->uy.p, 1));
        CELDA_TRY(chain_loglik_async(h));
        CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        return chain_check_status(h);
    }
    if (!ix_y && !u_y && !ix_z && !u_z) {  // Philox mode
        CELDA_TRY(chain_philox_draws(h, h->nG, h->ixy.p, h->uy.p));
        CELDA_TRY(chain_philox_draws(h, h->nM, h->ixz.p, h->uz.p));
    } else {
        if (!ix_y || !u_y || !ix_z || !u_z) return fail("iterate: pass all four of ix_y, u_y, ix_z, u_z or none (Philox)");
        CELDA_TRY(chain_upload_order(h, h->ixy.p, ix_y, h->nG));
        CELDA_CUDA_OK(cudaMemcpyAsync(h->uy.p, u_y, sizeof(double) * h->nG, cudaMemcpyHostToDevice, st));
        CELDA_TRY(chain_copy_stream(h));
        CELDA_CUDA_OK(cudaEventRecord(h->evCopyGo, st));  // ixz / uz are free once everything queued so far has run
    }
    CELDA_TRY(chain_step_y(h, h->ixy.p, h->uy.p, 1));
    if (ix_z) CELDA_TRY(chain_upload_z_draws_behind(h, ix_z, u_z));
    CELDA_TRY(chain_step_z(h, h->ixz.p, h->uz.p, 1));
    CELDA_TRY(chain_loglik_async(h));
    return chain_finish_iteration(h, ll);
}

int celda_chain_sweep_z_em(celda_chain* h, int doSample) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    CELDA_TRY(chain_step_z_em(h, doSample));
    return chain_check_em_flag(h);
}

// One iteration with the EM z step (algorithm = "EM"): celda_C -- .cCCalcEMProbZ then .cCCalcLL (R/celda_C.R:427-449);
// celda_CG -- y Gibbs sweep, rowSumByGroupChange, EM z step, colSumByGroupChange, .cCGCalcLL (R/celda_CG.R:543-602).
int celda_chain_iterate_em(celda_chain* h, const int* ix_y, const double* u_y, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    if (h->model == CELDA_MODEL_G) return fail("iterate_em: celda_G has no cell populations");
    if (h->model == CELDA_MODEL_CG) {
        if (!ix_y || !u_y) return fail("iterate_em: injected order and uniforms are required for the y sweep");
        CELDA_TRY(chain_upload_order(h, h->ixy.p, ix_y, h->nG));
        CELDA_CUDA_OK(cudaMemcpyAsync(h->uy.p, u_y, sizeof(double) * h->nG, cudaMemcpyHostToDevice, h->st));
        CELDA_TRY(chain_step_y(h, h->ixy.p, h->uy.p, 1));
    }
    CELDA_TRY(chain_step_z_em(h, 1));
    CELDA_TRY(chain_loglik_async(h));
    CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CELDA_TRY(chain_check_em_flag(h));
    return chain_check_status(h);
}

// perplexity(x) on the model's own counts (R/perplexity.R:233-325, newCounts = NULL).
}  // extern "C"
namespace {

// perplexity of the chain's model on a count matrix with the chain's dimensions (R/perplexity.R:233-325): the chain's
// own counts, another matrix (newCounts) or a resample.  colTot: the column totals of that matrix.
int chain_perplexity_on(celda_chain* h, const int* dp, const int* di, const int* dx, const int* colTot, double* perplexity) {
    if (!h->has_state) return fail("set_state must be called first");
    cudaStream_t st = h->st;
    const int sms = h->sms, K = h->K, L = h->L, nG = h->nG;
    DevBuf<double> psiDen, colDen, M, theta, percell, partial, out;
    CELDA_TRY(percell.alloc(h->nM));
    const int NB = 256;
    CELDA_TRY(partial.alloc(NB));
    CELDA_TRY(out.alloc(3));
    const int threads = 256, nwarp = threads / 32;
    const int grid = (int)std::min<long long>(((long long)h->nM + nwarp - 1) / nwarp, (long long)sms * 32);
    if (h->model != CELDA_MODEL_C) {
        CELDA_TRY(psiDen.alloc(L));
        k_perp_psi_den<<<(L + 63) / 64, 64, 0, st>>>(h->nByG.p, h->y.p, nG, L, h->delta, psiDen.p);
        count_launch();
    }
    if (h->model == CELDA_MODEL_G) {
        k_perp_cells_G<<<grid, threads, 0, st>>>(h->nM, L, dp, di, dx, h->nTSByC.p, h->y.p, h->nByG.p, psiDen.p, h->beta,
                                                h->delta, percell.p);
        count_launch();
    } else {
        if (K > 256) return fail("perplexity: K must be <= 256");
        CELDA_TRY(colDen.alloc(K));
        CELDA_TRY(M.alloc((size_t)nG * K));
        CELDA_TRY(theta.alloc((size_t)K * h->nS));
        if (h->model == CELDA_MODEL_CG) k_perp_col_den<<<1, 256, 0, st>>>(h->nTSByCP.p, false, L, K, h->beta, colDen.p);
        else k_perp_col_den<<<1, 256, 0, st>>>(h->nGByCP.p, true, nG, K, h->beta, colDen.p);
        k_perp_M<<<grid_for((long long)nG * K, 256, sms), 256, 0, st>>>(h->model, nG, K, L, h->y.p, h->nByG.p, psiDen.p,
                                                                       h->nTSByCP.p, h->nGByCP.p, colDen.p, h->beta,
                                                                       h->delta, M.p);
        k_perp_theta<<<(h->nS + 63) / 64, 64, 0, st>>>(h->mCPByS.p, K, h->nS, h->alpha, theta.p);
        k_perp_cells<<<grid, threads, 0, st>>>(h->nM, K, dp, di, dx, M.p, theta.p, h->s.p, percell.p);
        count_launch(4);
    }
    k_sum_array<<<NB, 256, 0, st>>>(percell.p, h->nM, partial.p);
    k_perp_final<<<1, 256, 0, st>>>(partial.p, NB, colTot, h->nM, out.p);
    count_launch(2);
    CELDA_CUDA_OK(cudaGetLastError());
    if (h->comm) {  // shards: all-reduce (logPx, total counts); the sum order differs from one GPU: 1e-9, not bit-exact
        CELDA_NCCL_OK(g_nccl.AllReduce(out.p + 1, out.p + 1, 2, ncclDouble, ncclSum, h->comm->comm, st));
        double v[2];
        CELDA_CUDA_OK(cudaMemcpyAsync(v, out.p + 1, sizeof(v), cudaMemcpyDeviceToHost, st));
        CELDA_TRY(chain_sync(h));
        *perplexity = std::exp(-(v[0] / v[1]));
        return 0;
    }
    CELDA_CUDA_OK(cudaMemcpyAsync(perplexity, out.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    return chain_sync(h);
}

}  // namespace
extern "C" {

// Log-likelihood of the chain's model under OTHER labels, the chain left untouched: what the split heuristics and the
// cluster-removal loops of the 'split' initialisation evaluate hundreds of times per call (.cCGSplitZ / .cCGSplitY
// R/split_clusters.R:236-330,483-560; .initializeSplit* R/initialize_clusters.R:150-216).  Exactly one of z_alt, y_alt
// replaces the chain's labels.  celda_CG needs no pass over the counts: other z -> nTSByCP = colSumByGroup(nTSByC, z),
// mCPByS = table(z, s); other y -> nTSByCP = rowSumByGroup(nGByCP, y), nByTS, nGByTS from nByG.
int celda_chain_loglik_alt(celda_chain* h, const int* z_alt, const int* y_alt, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    if ((z_alt != nullptr) == (y_alt != nullptr)) return fail("loglik_alt: pass exactly one of z_alt, y_alt");
    if (h->model != CELDA_MODEL_CG) return fail("loglik_alt: celda_CG chains only (use set_state + loglik)");
    if (h->comm) return fail("loglik_alt: not available on a cell-sharded chain");
    cudaStream_t st = h->st;
    const int K = h->K, L = h->L, nG = h->nG, nM = h->nM;
    const int n = z_alt ? nM : nG, top = z_alt ? K : L;
    for (int q = 0; q < n; ++q)
        if ((z_alt ? z_alt[q] : y_alt[q]) < 1 || (z_alt ? z_alt[q] : y_alt[q]) > top)
            return fail(z_alt ? "Assigned value of cell cluster greater than the total number of cell clusters!"
                              : "Assigned value of feature module greater than the total number of feature modules!");
    if (h->altLab.n < (size_t)std::max(nG, nM)) CELDA_TRY(h->altLab.alloc((size_t)std::max(nG, nM)));
    if (h->altT.n < (size_t)L * K) CELDA_TRY(h->altT.alloc((size_t)L * K));
    if (h->altM.n < (size_t)K * h->nS) CELDA_TRY(h->altM.alloc((size_t)K * h->nS));
    if (h->altG.n < (size_t)L) {
        CELDA_TRY(h->altG.alloc((size_t)L));
        CELDA_TRY(h->altTS.alloc((size_t)L));
    }
    CELDA_CUDA_OK(cudaMemcpyAsync(h->altLab.p, z_alt ? z_alt : y_alt, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CELDA_CUDA_OK(cudaMemsetAsync(h->altT.p, 0, sizeof(int) * (size_t)L * K, st));
    LLIn in{h->model, K, L, h->nS, nG, nM, h->mCPByS.p, h->altT.p, h->nGByCP.p, h->nTSByC.p, h->nGByTS.p, h->nByG.p,
            h->nByTS.p, h->alpha, h->beta, h->delta, h->gamma};
    if (z_alt) {
        if (sizeof(int) * L * K > 200 * 1024) return fail("L x K = %d x %d tables exceed the shared-memory budget", L, K);
        CELDA_CUDA_OK(cudaFuncSetAttribute(k_colsum_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * L * K)));
        k_colsum_dense<<<std::min(h->sms * 2, std::max(1, nM / 64)), 512, sizeof(int) * L * K, st>>>(h->nTSByC.p, L, nM,
                                                                                                  h->altLab.p, K, h->altT.p);
        CELDA_CUDA_OK(cudaMemsetAsync(h->altM.p, 0, sizeof(int) * (size_t)K * h->nS, st));
        k_table_zs<<<grid_for(nM, 256, h->sms), 256, 0, st>>>(h->altLab.p, h->s.p, nM, K, h->altM.p);
        count_launch(2);
        in.mCPByS = h->altM.p;
    } else {
        k_rowsum_genemajor<<<grid_for((long long)nG * K, 256, h->sms), 256, 0, st>>>(h->nGByCP.p, nG, K, h->altLab.p, L, h->altT.p);
        CELDA_CUDA_OK(cudaMemsetAsync(h->altTS.p, 0, sizeof(long long) * L, st));
        k_fill_int<<<1, 256, 0, st>>>(h->altG.p, L, 1);
        k_module_totals<<<grid_for(nG, 256, h->sms), 256, 0, st>>>(h->altLab.p, h->nByG.p, nG,
                                                                  (unsigned long long*)h->altTS.p, h->altG.p);
        count_launch(3);
        in.nByTS = h->altTS.p;
        in.nGByTS = h->altG.p;
    }
    CELDA_TRY(run_loglik(in, h->llPartial.p, h->llOut.p, st));
    CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    return chain_sync(h);
}

int celda_chain_perplexity(celda_chain* h, double* perplexity) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    return chain_perplexity_on(h, h->cp.p, h->ci.p, h->cx.p, h->nByC.p, perplexity);
}

// perplexity(x, celdaMod, newCounts = ...) (R/perplexity.R:233-325): the model of the chain evaluated on another count
// matrix of the same dimensions (any sparsity pattern).
int celda_chain_perplexity_new(celda_chain* h, const int* p, const int* i, const double* x, double* perplexity) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!p || !i || !x) return fail("perplexity_new: the new count matrix is required");
    if (p[0] != 0) return fail("column pointers of 'newCounts' must start at 0");
    for (int c = 0; c < h->nM; ++c)
        if (p[c + 1] < p[c]) return fail("column pointers of 'newCounts' must be non-decreasing");
    const long long Z = p[h->nM];
    for (long long t = 0; t < Z; ++t)
        if (i[t] < 0 || i[t] >= h->nG) return fail("row index out of range in 'newCounts'");
    cudaStream_t st = h->st;
    DevBuf<int> dp, di, dxi, flag, tot, ones, tmp;
    DevBuf<double> dx;
    CELDA_TRY(dp.upload(p, (size_t)h->nM + 1, st));
    CELDA_TRY(di.upload(i, (size_t)Z, st));
    CELDA_TRY(dx.upload(x, (size_t)Z, st));
    CELDA_TRY(dxi.alloc((size_t)Z));
    CELDA_TRY(flag.alloc(1));
    CELDA_TRY(flag.zero(st));
    k_narrow<<<grid_for(Z, 256, h->sms), 256, 0, st>>>(dx.p, Z, dxi.p, flag.p);
    int f = 0;
    CELDA_CUDA_OK(cudaMemcpyAsync(&f, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CELDA_CUDA_OK(cudaStreamSynchronize(st));
    if (f) return fail("newCounts must be integer-valued and below 2^31 (found a non-integer entry)");
    CELDA_TRY(tot.alloc(h->nM));
    CELDA_TRY(ones.alloc(h->nG));
    CELDA_TRY(tmp.alloc(h->nM));
    k_fill_int<<<grid_for(h->nG, 256, h->sms), 256, 0, st>>>(ones.p, h->nG, 1);
    k_rowsum_csc<int><<<std::min((h->nM + 7) / 8, h->sms * 8), 256, sizeof(int) * 8, st>>>(h->nM, dp.p, di.p, dxi.p, ones.p, 1,
                                                                                        tmp.p, tot.p);
    count_launch(3);
    return chain_perplexity_on(h, dp.p, di.p, dxi.p, tot.p, perplexity);
}

// .resamplePerplexity(doResampling = TRUE) (R/perplexity.R:449-476) for the chain's model: n multinomial resamples of
// the counts (.resampleCountMatrix, :764-775) drawn on the device with Philox (key = seed), perplexity on each.  The
// resample shares the counts' (p, i) pattern and column totals.
int celda_chain_perplexity_resampled(celda_chain* h, unsigned long long seed, int n_resample, double* perplexity) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (n_resample < 1) return fail("The 'numResample' parameter needs to be an integer greater than 0.");
    cudaStream_t st = h->st;
    if (h->resCum.n < (size_t)h->Z) {
        CELDA_TRY(h->resCum.alloc((size_t)h->Z));
        CELDA_TRY(h->resX.alloc((size_t)h->Z));
    }
    const int threads = 256, nwarp = threads / 32;
    const int grid = (int)std::min<long long>(((long long)h->nM + nwarp - 1) / nwarp, (long long)h->sms * 16);
    for (int j = 0; j < n_resample; ++j) {
        k_resample_prefix<<<grid, threads, 0, st>>>(h->nM, h->cp.p, h->cx.p, h->resCum.p, h->resX.p);
        k_resample_draw<<<grid, threads, 0, st>>>(h->nM, h->cp.p, h->resCum.p, (unsigned)seed, (unsigned)(seed >> 32), (unsigned)j,
                                                 h->resX.p);
        count_launch(2);
        CELDA_TRY(chain_perplexity_on(h, h->cp.p, h->ci.p, h->resX.p, h->nByC.p, perplexity + j));
    }
    return 0;
}

// The resampled values of the last celda_chain_perplexity_resampled call, aligned with the stored entries of the counts.
int celda_chain_get_resampled(celda_chain* h, int* x) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (h->resX.n < (size_t)h->Z) return fail("get_resampled: no resample has been drawn");
    CELDA_TRY(h->resX.download(x, (size_t)h->Z, h->st));
    return chain_sync(h);
}

int celda_chain_stage_draws(celda_chain* h, int n_iter, const int* ix_y, const double* u_y, const int* ix_z,
                            const double* u_z) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (h->model != CELDA_MODEL_CG) return fail("stage_draws: staged iterations run on a celda_CG chain");
    if (n_iter < 1 || !ix_y || !u_y || !ix_z || !u_z) return fail("stage_draws: n_iter >= 1 and all four arrays are required");
    for (int it = 0; it < n_iter; ++it) {
        CELDA_TRY(validate_order(ix_y + (size_t)it * h->nG, h->nG, "ix_y"));
        CELDA_TRY(validate_order(ix_z + (size_t)it * h->nM, h->nM, "ix_z"));
    }
    CELDA_TRY(h->st_ixy.upload(ix_y, (size_t)n_iter * h->nG, h->st));
    CELDA_TRY(h->st_uy.upload(u_y, (size_t)n_iter * h->nG, h->st));
    CELDA_TRY(h->st_ixz.upload(ix_z, (size_t)n_iter * h->nM, h->st));
    CELDA_TRY(h->st_uz.upload(u_z, (size_t)n_iter * h->nM, h->st));
    h->staged_iters = n_iter;
    return chain_sync(h);
}

int celda_chain_iterate_staged(celda_chain* h, int iter, double* ll) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (!h->has_state) return fail("set_state must be called first");
    if (iter < 0 || iter >= h->staged_iters) return fail("iterate_staged: iteration %d was not staged", iter);
    CELDA_TRY(chain_step_y(h, h->st_ixy.p + (size_t)iter * h->nG, h->st_uy.p + (size_t)iter * h->nG, 1));
    CELDA_TRY(chain_step_z(h, h->st_ixz.p + (size_t)iter * h->nM, h->st_uz.p + (size_t)iter * h->nM, 1));
    CELDA_TRY(chain_loglik_async(h));
    CELDA_CUDA_OK(cudaMemcpyAsync(ll, h->llOut.p, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    return chain_check_status(h);
}

int celda_chain_get_tables(celda_chain* h, int* mCPByS, double* nTSByC, double* nTSByCP, double* nGByCP, double* nCP,
                           double* nByC, double* nByG, double* nByTS, int* nGByTS) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = h->st;
    const int sms = h->sms;
    DevBuf<double> tmp;
    auto out_d = [&](auto* src, long long n, double* dst) -> int {
        CELDA_TRY(tmp.alloc((size_t)n));
        using T = std::remove_pointer_t<decltype(src)>;
        k_convert<T, double><<<grid_for(n, 256, sms), 256, 0, st>>>(src, n, tmp.p);
        count_launch();
        CELDA_TRY(tmp.download(dst, (size_t)n, st));
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    };
    if (mCPByS && h->model != CELDA_MODEL_G) CELDA_TRY(h->mCPByS.download(mCPByS, (size_t)h->K * h->nS, st));
    if (nTSByC && h->model != CELDA_MODEL_C) CELDA_TRY(out_d(h->nTSByC.p, (long long)h->L * h->nM, nTSByC));
    if (nTSByCP && h->model == CELDA_MODEL_CG) CELDA_TRY(out_d(h->nTSByCP.p, (long long)h->L * h->K, nTSByCP));
    if (nGByCP && h->model != CELDA_MODEL_G) {
        const long long n = (long long)h->nG * h->K;
        CELDA_TRY(tmp.alloc((size_t)n));
        k_transpose_convert<int, double><<<grid_for(n, 256, sms), 256, 0, st>>>(h->nGByCP.p, h->nG, h->K, tmp.p, true);
        count_launch();
        CELDA_TRY(tmp.download(nGByCP, (size_t)n, st));
        CELDA_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (nCP && h->model != CELDA_MODEL_G) CELDA_TRY(out_d(h->nCP.p, h->K, nCP));
    if (nByC) CELDA_TRY(out_d(h->nByC.p, h->nM, nByC));
    if (nByG) CELDA_TRY(out_d(h->nByG.p, h->nG, nByG));
    if (nByTS && h->model != CELDA_MODEL_C) CELDA_TRY(out_d(h->nByTS.p, h->L, nByTS));
    if (nGByTS && h->model != CELDA_MODEL_C) CELDA_TRY(h->nGByTS.download(nGByTS, h->L, st));
    CELDA_CUDA_OK(cudaGetLastError());
    return chain_sync(h);
}

int celda_chain_get_probs(celda_chain* h, double* probsZ, double* probsY) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    if (probsZ && h->model != CELDA_MODEL_G) CELDA_TRY(h->probsZ.download(probsZ, (size_t)h->K * h->nM, h->st));
    if (probsY && h->model != CELDA_MODEL_C) CELDA_TRY(h->probsY.download(probsY, (size_t)h->L * h->nG, h->st));
    return chain_sync(h);
}

int celda_chain_timer_start(celda_chain* h) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_CUDA_OK(cudaEventRecord(h->ev0, h->st));
    return 0;
}

int celda_chain_timer_stop(celda_chain* h, float* ms) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_CUDA_OK(cudaEventRecord(h->ev1, h->st));
    CELDA_CUDA_OK(cudaEventSynchronize(h->ev1));
    CELDA_CUDA_OK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

int celda_chain_profile(celda_chain* h, int enable) {
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_TRY(chain_sync(h));
    h->profile = enable != 0;
    for (int q = 0; q < 8; ++q) {
        h->prof_ms[q] = 0;
        h->prof_launches[q] = 0;
    }
    CELDA_TRY(h->wb.stats.zero(h->st));
    return chain_sync(h);
}

#ifdef CELDA_DEBUG_STATS
int celda_cuda_debug_counters(long long* out8, int reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    CELDA_CUDA_OK(cudaMemcpyFromSymbol(out8, g_dbg, sizeof(z)));
    if (reset) CELDA_CUDA_OK(cudaMemcpyToSymbol(g_dbg, z, sizeof(z)));
    return 0;
}
int celda_cuda_debug_segments(long long* out8, int reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CELDA_CUDA_OK(cudaDeviceSynchronize());
    CELDA_CUDA_OK(cudaMemcpyFromSymbol(out8, g_dbg2, sizeof(z)));
    if (reset) CELDA_CUDA_OK(cudaMemcpyToSymbol(g_dbg2, z, sizeof(z)));
    return 0;
}
#endif

int celda_chain_profile_get(celda_chain* h, int which, double* ms, long long* launches, long long* rounds) {
    if (which < 0 || which > 6) return fail("profile_get: which must be 0..6");
    CELDA_CUDA_OK(cudaSetDevice(h->device));
    CELDA_TRY(chain_sync(h));
    long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (which < 5) CELDA_CUDA_OK(cudaMemcpy(stats, h->wb.stats.p + 8 * which, sizeof(stats), cudaMemcpyDeviceToHost));
    if (ms) *ms = h->prof_ms[which];
    if (launches) *launches = h->prof_launches[which];
    if (rounds) {
        for (int q = 0; q < 8; ++q) rounds[q] = stats[q];
    }
    return 0;
}

}  // extern "C"
