// Aggregation, contingency-table, log-likelihood and element-wise kernels (sm_100a).
// All count data on the device is int32 (values) / int64 (totals); sums of integer-valued doubles in the
// reference (src/matrixSumsSparse.cpp, src/matrixSums.c) are exact, so integer arithmetic reproduces them bit
// for bit and is order-independent (atomics are safe).
#pragma once
#include "celda_math.cuh"
#include "philox.cuh"

namespace celda {

// ---------------------------------------------------------------------------------------------- element-wise
template <int OP>
__global__ void k_math_vec(const double* __restrict__ x, long long n, double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        out[i] = OP == 0 ? dlgamma_hot(v) : (OP == 1 ? dlog(v) : dexp(v));  // hot lgamma falls back to the generic one off range
    }
}

// FP64 FMA peak probe: 8 independent chains per thread, fma() is explicit so -fmad=false does not split it.
__global__ void k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

// double -> int32 with an integrality / range check (flag[0] |= 1 when a value is not an int32 integer)
__global__ void k_narrow(const double* __restrict__ x, long long n, int* __restrict__ out, int* flag) {
    int bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        const int iv = (int)v;
        bad |= !((double)iv == v);
        out[i] = iv;
    }
    if (bad) atomicOr(flag, 1);
}

template <typename TI, typename TO>
__global__ void k_convert(const TI* __restrict__ x, long long n, TO* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (TO)x[i];
}

// out[r + c*nr] = in[c + r*nc]  (gene-major [nr][nc] <-> column-major nr x nc)
template <typename TI, typename TO>
__global__ void k_transpose_convert(const TI* __restrict__ in, int nr, int nc, TO* __restrict__ out, bool in_is_rowmajor) {
    const long long n = (long long)nr * nc;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(q % nr), c = (int)(q / nr);  // q indexes the column-major side
        if (in_is_rowmajor) out[q] = (TO)in[(size_t)r * nc + c];
        else out[(size_t)r * nc + c] = (TO)in[q];
    }
}

// ---------------------------------------------------------------------------------------------- CSC aggregation
// rowSumByGroup (src/matrixSumsSparse.cpp:44-73): out[L x nc] column-major; one warp per column, per-warp shared
// accumulator so every HBM byte of (i, x) is read once, coalesced, and the output is written once.
template <typename T>
__global__ void k_rowsum_csc(int nc, const int* __restrict__ p, const int* __restrict__ ri, const T* __restrict__ xv,
                             const int* __restrict__ group, int L, T* __restrict__ out, T* __restrict__ nByC) {
    extern __shared__ __align__(8) unsigned char acc_raw[];
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* acc = reinterpret_cast<T*>(acc_raw) + warp * L;
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        for (int l = lane; l < L; l += 32) acc[l] = 0;
        __syncwarp();
        T tot = 0;
        const int e = p[c + 1];
        int t = p[c] + lane;
        for (; t + 224 < e; t += 256) {  // eight independent index/value/label loads in flight per lane
            int r[8];
            T v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                r[q] = ri[t + 32 * q];
                v[q] = xv[t + 32 * q];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = group[r[q]];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                atomicAdd(&acc[r[q] - 1], v[q]);
                tot += v[q];
            }
        }
        for (; t + 96 < e; t += 128) {
            const int r0 = ri[t], r1 = ri[t + 32], r2 = ri[t + 64], r3 = ri[t + 96];
            const T v0 = xv[t], v1 = xv[t + 32], v2 = xv[t + 64], v3 = xv[t + 96];
            const int g0 = group[r0], g1 = group[r1], g2 = group[r2], g3 = group[r3];
            atomicAdd(&acc[g0 - 1], v0);
            atomicAdd(&acc[g1 - 1], v1);
            atomicAdd(&acc[g2 - 1], v2);
            atomicAdd(&acc[g3 - 1], v3);
            tot += v0 + v1 + v2 + v3;
        }
        for (; t < e; t += 32) {
            const T v = xv[t];
            atomicAdd(&acc[group[ri[t]] - 1], v);
            tot += v;
        }
        __syncwarp();
        for (int l = lane; l < L; l += 32) out[(size_t)c * L + l] = acc[l];
        if (nByC) {  // integer instantiation only (order of a shuffle tree is irrelevant for integers)
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) nByC[c] = tot;
        }
        __syncwarp();
    }
}

// rowSumByGroup for the chain handle (int32 counts): the same result as k_rowsum_csc<int>, shaped for HBM bandwidth.
// The per-entry label look-up group[row] is the expensive part of the generic kernel (a 32-way gather through L1 per
// warp instruction), so the labels are first packed to 16 bits in SHARED memory (nG * 2 bytes, shared by the CTA's
// warps) where a gather costs a few bank conflicts; indices and values are read with 16-byte loads (the column's
// unaligned head and tail with scalar loads), four vectors in flight per lane.  Requires L <= 65535 and
// 2 * nG + 4 * nwarp * L bytes of shared memory; the launcher falls back to k_rowsum_csc otherwise.
__global__ void __launch_bounds__(512) k_rowsum_csc_v4(int nc, int nG, const int* __restrict__ p, const int* __restrict__ ri,
                                                       const int* __restrict__ xv, const int* __restrict__ group, int L,
                                                       int* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* acc = reinterpret_cast<int*>(rs_smem) + warp * L;
    unsigned short* lab = reinterpret_cast<unsigned short*>(rs_smem + sizeof(int) * (size_t)nwarp * L);
    for (int g = threadIdx.x; g < nG; g += blockDim.x) lab[g] = (unsigned short)(group[g] - 1);
    __syncthreads();
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        for (int l = lane; l < L; l += 32) acc[l] = 0;
        __syncwarp();
        const int b = p[c], e = p[c + 1];
        const int b4 = min(e, (b + 3) & ~3), e4 = max(b4, e & ~3);  // [b4, e4) is 16-byte aligned in both arrays
        if (b + lane < b4) atomicAdd(&acc[lab[ri[b + lane]]], xv[b + lane]);
        if (e4 + lane < e) atomicAdd(&acc[lab[ri[e4 + lane]]], xv[e4 + lane]);
        const int4* r4 = reinterpret_cast<const int4*>(ri + b4);
        const int4* x4 = reinterpret_cast<const int4*>(xv + b4);
        const int n4 = (e4 - b4) >> 2;
        int t = lane;
        for (; t + 96 < n4; t += 128) {  // four index and four value vectors in flight per lane
            int4 r[4], v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                r[q] = __ldg(r4 + t + 32 * q);
                v[q] = __ldg(x4 + t + 32 * q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                atomicAdd(&acc[lab[r[q].x]], v[q].x);
                atomicAdd(&acc[lab[r[q].y]], v[q].y);
                atomicAdd(&acc[lab[r[q].z]], v[q].z);
                atomicAdd(&acc[lab[r[q].w]], v[q].w);
            }
        }
        for (; t < n4; t += 32) {
            const int4 r = __ldg(r4 + t), v = __ldg(x4 + t);
            atomicAdd(&acc[lab[r.x]], v.x);
            atomicAdd(&acc[lab[r.y]], v.y);
            atomicAdd(&acc[lab[r.z]], v.z);
            atomicAdd(&acc[lab[r.w]], v.w);
        }
        __syncwarp();
        for (int l = lane; l < L; l += 32) out[(size_t)c * L + l] = acc[l];
        __syncwarp();
    }
}

// colSumByGroup (src/matrixSumsSparse.cpp:9-38): out gene-major [nr][K] (pre-zeroed), integer atomics into L2.
template <typename T>
__global__ void k_colsum_csc(int nc, const int* __restrict__ p, const int* __restrict__ ri, const T* __restrict__ xv,
                             const int* __restrict__ group, int K, T* __restrict__ out) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        const int g = group[c] - 1, e = p[c + 1];
        int t = p[c] + lane;
        for (; t + 96 < e; t += 128) {
            const int r0 = ri[t], r1 = ri[t + 32], r2 = ri[t + 64], r3 = ri[t + 96];
            const T v0 = xv[t], v1 = xv[t + 32], v2 = xv[t + 64], v3 = xv[t + 96];
            atomicAdd(&out[(size_t)r0 * K + g], v0);
            atomicAdd(&out[(size_t)r1 * K + g], v1);
            atomicAdd(&out[(size_t)r2 * K + g], v2);
            atomicAdd(&out[(size_t)r3 * K + g], v3);
        }
        for (; t < e; t += 32) atomicAdd(&out[(size_t)ri[t] * K + g], xv[t]);
    }
}

// colSumByGroupChange (src/matrixSumsSparse.cpp:85-131): only columns whose label changed are walked.
template <typename T>
__global__ void k_colsum_change_csc(int nc, const int* __restrict__ p, const int* __restrict__ ri,
                                    const T* __restrict__ xv, const int* __restrict__ group,
                                    const int* __restrict__ pgroup, int K, T* __restrict__ out) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        const int g = group[c] - 1, pg = pgroup[c] - 1;
        if (g == pg) continue;
        const int e = p[c + 1];
        for (int t = p[c] + lane; t < e; t += 32) {
            const T v = xv[t];
            atomicAdd(&out[(size_t)ri[t] * K + g], v);
            atomicAdd(&out[(size_t)ri[t] * K + pg], -v);
        }
    }
}

// rowSumByGroupChange (src/matrixSumsSparse.cpp:141-187): like the reference, every stored entry is inspected
// (the row index stream is read once, coalesced; values only for rows whose label changed).  px: L x nc.
template <typename T>
__global__ void k_rowsum_change_csc(int nc, const int* __restrict__ p, const int* __restrict__ ri,
                                    const T* __restrict__ xv, const int* __restrict__ group,
                                    const int* __restrict__ pgroup, int L, T* __restrict__ px) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        const int e = p[c + 1];
        for (int t = p[c] + lane; t < e; t += 32) {
            const int r = ri[t];
            const int g = group[r], pg = pgroup[r];
            if (g != pg) {
                const T v = xv[t];
                atomicAdd(&px[(size_t)c * L + g - 1], v);
                atomicAdd(&px[(size_t)c * L + pg - 1], -v);
            }
        }
    }
}

// per-row totals of a CSC matrix (Matrix::rowSums, R/celda_CG.R:913): int64 atomics
__global__ void k_rowtotals_csc(long long nnz, const int* __restrict__ ri, const int* __restrict__ xv,
                                unsigned long long* __restrict__ nByG) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nnz; t += (long long)gridDim.x * blockDim.x)
        atomicAdd(&nByG[ri[t]], (unsigned long long)(long long)xv[t]);
}

// ---------------------------------------------------------------------------------------------- dense aggregation
// colSumByGroup on a dense int32 R x nc column-major matrix (nTSByC -> nTSByCP, R/celda_CG.R:910): out R x K.
__global__ void k_colsum_dense(const int* __restrict__ x, int R, long long nc, const int* __restrict__ group, int K,
                               int* __restrict__ out) {
    extern __shared__ int acc[];  // R x K
    for (int q = threadIdx.x; q < R * K; q += blockDim.x) acc[q] = 0;
    __syncthreads();
    const long long per = (nc + gridDim.x - 1) / gridDim.x;
    const long long c0 = blockIdx.x * per, c1 = min(nc, c0 + per);
    const long long n = (c1 > c0) ? (c1 - c0) * R : 0;
    for (long long q = threadIdx.x; q < n; q += blockDim.x) {
        const long long c = c0 + q / R;
        const int r = (int)(q % R);
        const int v = x[c * R + r];
        if (v) atomicAdd(&acc[(group[c] - 1) * R + r], v);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < R * K; q += blockDim.x)
        if (acc[q]) atomicAdd(&out[q], acc[q]);
}

// colSumByGroupChange on the same dense layout, in place on out (R x K)
__global__ void k_colsum_change_dense(const int* __restrict__ x, int R, long long nc, const int* __restrict__ group,
                                      const int* __restrict__ pgroup, int* __restrict__ out) {
    const long long n = nc * R;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const long long c = q / R;
        const int g = group[c], pg = pgroup[c];
        if (g == pg) continue;
        const int r = (int)(q % R);
        const int v = x[q];
        if (v) {
            atomicAdd(&out[(size_t)(g - 1) * R + r], v);
            atomicSub(&out[(size_t)(pg - 1) * R + r], v);
        }
    }
}

// table(factor(z, 1:K), s) (R/celda_CG.R:904): K x nS
__global__ void k_table_zs(const int* __restrict__ z, const int* __restrict__ s, long long n, int K, int* __restrict__ m) {
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x)
        atomicAdd(&m[(size_t)(s[c] - 1) * K + z[c] - 1], 1);
}

// nByTS = rowSumByGroup(nByG, y), nGByTS = tabulate(y, L) + 1  (R/celda_CG.R:914-916); nGByTS pre-set to 1
__global__ void k_module_totals(const int* __restrict__ y, const long long* __restrict__ nByG, int nG,
                                unsigned long long* __restrict__ nByTS, int* __restrict__ nGByTS) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nG; g += gridDim.x * blockDim.x) {
        atomicAdd(&nByTS[y[g] - 1], (unsigned long long)nByG[g]);
        atomicAdd(&nGByTS[y[g] - 1], 1);
    }
}

__global__ void k_fill_int(int* p, long long n, int v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// nCP = colSums(table) for an int32 R x K column-major table (one thread per column; K is small)
__global__ void k_colsums_small(const int* __restrict__ t, int R, int K, long long* __restrict__ out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        long long s = 0;
        for (int r = 0; r < R; ++r) s += t[(size_t)k * R + r];
        out[k] = s;
    }
}

// column totals of a gene-major [nG][K] table (nCP for celda_C)
__global__ void k_colsums_genemajor(const int* __restrict__ t, int nG, int K, unsigned long long* __restrict__ out) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < (long long)nG * K;
         q += (long long)gridDim.x * blockDim.x)
        if (t[q]) atomicAdd(&out[q % K], (unsigned long long)(long long)t[q]);
}

// ---------------------------------------------------------------------------------------------- label checks
// flag bit 0: label outside [1, K]
__global__ void k_check_labels(const int* __restrict__ g, long long n, int K, int* flag) {
    int bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        bad |= (g[i] < 1 || g[i] > K);
    if (bad) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------------------------- log-likelihood
// Deterministic block reduction (fixed tree), result on thread 0.
__device__ double block_sum(double v, double* red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// Partial sums of lgamma(x + add) over an integer array, one partial per block (grid fixed => deterministic).
template <typename T>
__global__ void k_sum_lgamma(const T* __restrict__ x, long long n, double add, double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += dlgamma((double)x[i] + add);
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Partial sums of lgamma(colSum(x[, c] + add)) for an int32 R x nc column-major matrix (one thread per column,
// the column is added serially like R's colSums of (x + add)).
__global__ void k_sum_lgamma_colsums(const int* __restrict__ x, int R, long long nc, double add,
                                     double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < nc; c += (long long)gridDim.x * blockDim.x) {
        double cs = 0.0;
        for (int r = 0; r < R; ++r) cs += (double)x[c * R + r] + add;
        acc += dlgamma(cs);
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct LLArgs {
    int model;  // 0 CG, 1 C, 2 G
    int K, L, nS, nG;
    long long nM;
    const int* mCPByS;        // K x nS
    const int* phiTab;        // CG: nTSByCP (L x K); C/G: unused (partials)
    const long long* nByTS;   // L
    const int* nGByTS;        // L
    const double* partial;    // [0..np): sum lgamma(nByG + delta); then model-specific partial blocks
    int npG, npPhiB, npPhiD;  // partial counts: genes, phi numerator, phi denominator (C and G models)
    double alpha, beta, delta, gamma;
    double* out;
};

// Final assembly in the reference's order: a + b + c + d per component, then theta + phi + psi + eta
// (R/celda_CG.R:854-887, R/celda_C.R:771-787, R/celda_G.R:677-701).  One block; every thread first accumulates its
// share of all seven table sums, then ONE block reduction combines them (a chain of seven would cost seven
// barrier round trips on a kernel that is pure latency).
__global__ void k_loglik_final(LLArgs a) {
    __shared__ double red[7][32];
    const int K = a.K, L = a.L, nS = a.nS;
    const double* part = a.partial;
    const int tid = threadIdx.x, nt = blockDim.x, rev = nt - 1 - tid;  // column sums start from the last thread
    double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (a.model != 2) {  // theta (CG, C)
        for (int q = tid; q < K * nS; q += nt) acc[0] += dlgamma((double)a.mCPByS[q] + a.alpha);
        for (int s = rev; s < nS; s += nt) {
            double cs = 0.0;
            for (int k = 0; k < K; ++k) cs += (double)a.mCPByS[(size_t)s * K + k] + a.alpha;
            acc[1] += dlgamma(cs);
        }
    }
    if (a.model == 0) {  // phi over nTSByCP (L x K)
        for (int q = tid; q < L * K; q += nt) acc[2] += dlgamma((double)a.phiTab[q] + a.beta);
        for (int k = rev; k < K; k += nt) {
            double cs = 0.0;
#pragma unroll 8
            for (int l = 0; l < L; ++l) cs += (double)a.phiTab[(size_t)k * L + l] + a.beta;
            acc[3] += dlgamma(cs);
        }
    }
    if (a.model != 1) {  // psi, eta (CG, G)
        for (int l = tid; l < L; l += nt) {
            const double g = (double)a.nGByTS[l];
            acc[4] += dlgamma(g * a.delta);
            acc[5] += dlgamma((double)a.nByTS[l] + (g * a.delta));
            acc[6] += dlgamma(g + a.gamma);
        }
    }
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    // what thread 0 needs afterwards, fetched / evaluated in parallel: the per-block partials, the module sizes and
    // the seven scalar lgamma terms (lane 0 of warps 1..7; dlgamma() is a few hundred dependent instructions)
    __shared__ double sp[640];
    __shared__ int sG[1024];
    __shared__ double sc[8];
    const int np = min(640, a.npG + a.npPhiB + a.npPhiD);
    for (int q = tid; q < np; q += nt) sp[q] = part[q];
    const bool gS = a.model != 1 && L <= 1024;
    if (gS)
        for (int l = tid; l < L; l += nt) sG[l] = a.nGByTS[l];
    if (lane == 0 && warp >= 1 && warp <= 7) {
        const int Rr = (a.model == 0) ? L : ((a.model == 1) ? a.nG : L);
        double x;
        switch (warp) {
            case 1: x = (double)K * a.alpha; break;
            case 2: x = a.alpha; break;
            case 3: x = (double)Rr * a.beta; break;
            case 4: x = a.beta; break;
            case 5: x = a.delta; break;
            case 6: x = (double)L * a.gamma; break;
            default: x = a.gamma; break;
        }
        sc[warp] = dlgamma(x);
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        double v = acc[c];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[c][warp] = v;
    }
    __syncthreads();
    if (tid < 7) {
        double sres = 0.0;
        for (int w = 0; w < nw; ++w) sres += red[tid][w];
        red[tid][0] = sres;
    }
    __syncthreads();
    if (tid != 0) return;
    double comp[4] = {0.0, 0.0, 0.0, 0.0};
    if (a.model != 2) {
        const double aa = (double)nS * sc[1];
        const double c = (double)(-nS * K) * sc[2];
        comp[0] = aa + red[0][0] + c + -red[1][0];
    }
    {
        double b = red[2][0], d = red[3][0];
        int R = L;
        long long ncols = K;
        if (a.model != 0) {
            R = (a.model == 1) ? a.nG : L;
            ncols = (a.model == 1) ? K : a.nM;
            b = 0.0;
            d = 0.0;
            for (int q = 0; q < a.npPhiB; ++q) b += sp[a.npG + q];
            for (int q = 0; q < a.npPhiD; ++q) d += sp[a.npG + a.npPhiB + q];
        }
        const double aa = (double)ncols * sc[3];
        const double c = -(double)ncols * (double)R * sc[4];
        comp[1] = aa + b + c + -d;
    }
    if (a.model != 1) {
        double pb = 0.0;
        for (int q = 0; q < a.npG; ++q) pb += sp[q];
        double nGs = 0.0, es = 0.0;
        for (int l = 0; l < L; ++l) {
            const double g = (double)(gS ? sG[l] : a.nGByTS[l]);
            nGs += g;
            es += g + a.gamma;
        }
        const double pc = -nGs * sc[5];
        comp[2] = red[4][0] + pb + pc + -red[5][0];
        const double ea = sc[6];
        const double ec = (double)(-L) * sc[7];
        comp[3] = ea + red[6][0] + ec + -dlgamma(es);
    }
    if (a.model == 0) *a.out = comp[0] + comp[1] + comp[2] + comp[3];
    else if (a.model == 1) *a.out = comp[0] + comp[1];
    else *a.out = comp[1] + comp[2] + comp[3];
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- dense twins
// src/matrixSums.c.  One thread per OUTPUT element, gathering its contributions in the reference's loop order
// (ascending row / column index), so results are bit-identical for arbitrary doubles, not only integer-valued
// ones, and the int32 variants wrap like the reference's int arithmetic.
template <typename T> struct AccOf { using type = T; };
template <> struct AccOf<int> { using type = unsigned; };

// _rowSumByGroup[_numeric] (:5-48, :198-236): ans[nl x nc], ans[g(i), j] += x[i, j]
template <typename T>
__global__ void k_dense_rowsum(const T* __restrict__ x, int nr, int nc, const int* __restrict__ group, int nl,
                               T* __restrict__ ans) {
    using A = typename AccOf<T>::type;
    const long long n = (long long)nl * nc;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(q % nl) + 1;
        const long long j = q / nl;
        A acc = 0;
        for (int i = 0; i < nr; ++i)
            if (group[i] == l) acc += (A)x[j * nr + i];
        ans[q] = (T)acc;
    }
}

// _colSumByGroup[_numeric] (:50-93, :240-278): ans[nr x nl], ans[i, g(j)] += x[i, j]
template <typename T>
__global__ void k_dense_colsum(const T* __restrict__ x, int nr, int nc, const int* __restrict__ group, int nl,
                               T* __restrict__ ans) {
    using A = typename AccOf<T>::type;
    const long long n = (long long)nr * nl;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nr), l = (int)(q / nr) + 1;
        A acc = 0;
        for (int j = 0; j < nc; ++j)
            if (group[j] == l) acc += (A)x[(long long)j * nr + i];
        ans[q] = (T)acc;
    }
}

// _rowSumByGroupChange[_numeric] (:97-143, :282-327): px[nl x nc] in place
template <typename T>
__global__ void k_dense_rowsum_change(const T* __restrict__ x, int nr, int nc, T* __restrict__ px,
                                      const int* __restrict__ group, const int* __restrict__ pgroup, int nl) {
    using A = typename AccOf<T>::type;
    const long long n = (long long)nl * nc;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(q % nl) + 1;
        const long long j = q / nl;
        A acc = (A)px[q];
        for (int i = 0; i < nr; ++i) {
            const int g = group[i], pg = pgroup[i];
            if (g != pg) {
                if (pg == l) acc -= (A)x[j * nr + i];
                if (g == l) acc += (A)x[j * nr + i];
            }
        }
        px[q] = (T)acc;
    }
}

// _colSumByGroupChange[_numeric] (:147-194, :331-379): px[nr x nl] in place
template <typename T>
__global__ void k_dense_colsum_change(const T* __restrict__ x, int nr, int nc, T* __restrict__ px,
                                      const int* __restrict__ group, const int* __restrict__ pgroup, int nl) {
    using A = typename AccOf<T>::type;
    const long long n = (long long)nr * nl;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nr), l = (int)(q / nr) + 1;
        A acc = (A)px[q];
        for (int j = 0; j < nc; ++j) {
            const int g = group[j], pg = pgroup[j];
            if (g != pg) {
                if (g == l) acc += (A)x[(long long)j * nr + i];
                if (pg == l) acc -= (A)x[(long long)j * nr + i];
            }
        }
        px[q] = (T)acc;
    }
}

// ---------------------------------------------------------------------------------------------- a10 single gene
// cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:187-249) for one gene: one thread per module, the column loop
// and the six trailing terms in the reference's order.  Table look-ups lg_beta[k] etc. are evaluated directly.
__global__ void k_cG_CalcGibbsProbY(int index0, const double* __restrict__ counts, int ncol,
                                    const double* __restrict__ nTSbyC, const double* __restrict__ nbyTS,
                                    const int* __restrict__ nGbyTS, const double* __restrict__ nbyG,
                                    const int* __restrict__ y, int L, double beta, double gamma, double delta,
                                    double* __restrict__ probs) {
    const int cur = y[index0] - 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
        double p = 0.0;
        for (int col = 0; col < ncol; ++col) {
            const double t = nTSbyC[(size_t)col * L + i], c = counts[col];
            if (i == cur) {
                p += dlgamma_hot((double)(long long)t + beta);
                p -= dlgamma_hot((double)(long long)(t - c) + beta);
            } else {
                p += dlgamma_hot((double)(long long)(t + c) + beta);
                p -= dlgamma_hot((double)(long long)t + beta);
            }
        }
        const int g = nGbyTS[i];
        if (i == cur) {
            p += dlgamma_hot((double)g + gamma);
            p -= dlgamma_hot((double)(g - 1) + gamma);
            p += dlgdelta(g, delta);
            p -= dlgdelta(g - 1, delta);
            p += dlgamma_hot(nbyTS[i] - nbyG[index0] + (double)(g - 1) * delta);
            p -= dlgamma_hot(nbyTS[i] + (double)g * delta);
        } else {
            p += dlgamma_hot((double)(g + 1) + gamma);
            p -= dlgamma_hot((double)g + gamma);
            p += dlgdelta(g + 1, delta);
            p -= dlgdelta(g, delta);
            p += dlgamma_hot(nbyTS[i] + (double)g * delta);
            p -= dlgamma_hot(nbyTS[i] + nbyG[index0] + (double)(g + 1) * delta);
        }
        probs[i] = p;
    }
}

// ---------------------------------------------------------------------------------------------- a9 helpers
// fastNormPropLog (src/matrixNorm.cpp:36-52): res = log((x + a) / (colSum + nrow a)); one block per column,
// the column sum added serially by thread 0 (Rcpp sugar colSums order).  flag |= 1 on a zero denominator.
__global__ void k_fastNormPropLog(const double* __restrict__ x, int nr, int nc, double a, double* __restrict__ res,
                                  int* flag) {
    __shared__ double s_den;
    for (int j = blockIdx.x; j < nc; j += gridDim.x) {
        if (threadIdx.x == 0) {
            double cs = 0.0;
            for (int i = 0; i < nr; ++i) cs += x[(size_t)j * nr + i];
            s_den = cs + (double)nr * a;
            if (s_den == 0.0) atomicOr(flag, 1);
        }
        __syncthreads();
        const double den = s_den;
        for (int i = threadIdx.x; i < nr; i += blockDim.x)
            res[(size_t)j * nr + i] = dlog((x[(size_t)j * nr + i] + a) / den);
        __syncthreads();
    }
}

// eigenMatMultInt / eigenMatMultNumeric (src/eigenMatMultInt.cpp:11-26): C = t(A) %*% B, A n x k, B n x m, C k x m.
// One thread per output element, serial dot product over n (unfused multiply-add).
template <typename TB>
__global__ void k_matMultT(const double* __restrict__ A, int n, int k, const TB* __restrict__ B, int m,
                           double* __restrict__ Cm) {
    const long long tot = (long long)k * m;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(q % k);
        const long long c = q / k;
        double acc = 0.0;
        for (int r = 0; r < n; ++r) acc += A[(size_t)a * n + r] * (double)B[c * n + r];
        Cm[q] = acc;
    }
}

// _perplexityG (src/perplexity.c:51-55): sum_ij x[i,j] log(phi[g(i), j] psi[i, g(i)]); per-block partials
__global__ void k_perplexityG(const int* __restrict__ x, int nr, int nc, const double* __restrict__ phi,
                              const double* __restrict__ psi, const int* __restrict__ group, int nl,
                              double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    const long long n = (long long)nr * nc;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nr);
        const long long j = q / nr;
        const int g = group[i] - 1;
        acc += (double)x[q] * dlog(phi[j * nl + g] * psi[(size_t)nr * g + i]);
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void k_sum_partials(const double* __restrict__ partial, int n, double* out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partial[i];
        *out = s;
    }
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- EM z step (a9)
// .cCCalcEMProbZ (R/celda_C.R:717-756).  phi is laid out [R][K] (row r's K log-probabilities contiguous) so that
// the lanes of a warp (one lane per population) read it coalesced.

// phi[r][k] = log((n[r,k] + beta) / (nCP[k] + R beta)); n is gene-major [R][K] (rowMajor) or column-major R x K.
// flag |= 1 on a zero denominator (fastNormPropLog's "Division by 0" error, src/matrixNorm.cpp:44-46).
__global__ void k_em_phi(const int* __restrict__ n, bool rowMajor, int R, int K, const long long* __restrict__ nCP,
                         double beta, double* __restrict__ phi, int* flag) {
    const long long tot = (long long)R * K;
    const double atot = (double)R * beta;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(q / K), k = (int)(q % K);
        const double den = (double)nCP[k] + atot;
        if (den == 0.0) atomicOr(flag, 1);
        const int v = rowMajor ? n[q] : n[(size_t)k * R + r];
        phi[q] = dlog(((double)v + beta) / den);
    }
}

// theta[k, s] = log((m[k,s] + alpha) / (colSum_s + K alpha)), K x nS column-major
__global__ void k_em_theta(const int* __restrict__ m, int K, int nS, double alpha, double* __restrict__ theta, int* flag) {
    for (int s = blockIdx.x; s < nS; s += gridDim.x) {
        double cs = 0.0;
        for (int k = 0; k < K; ++k) cs += (double)m[(size_t)s * K + k];
        const double den = cs + (double)K * alpha;
        if (den == 0.0 && threadIdx.x == 0) atomicOr(flag, 1);
        for (int k = threadIdx.x; k < K; k += blockDim.x) theta[(size_t)s * K + k] = dlog(((double)m[(size_t)s * K + k] + alpha) / den);
    }
}

// probs[, c] = t(phi) %*% counts[, c] + theta[, s_c] and which.max (first maximum), one warp per cell, lane k (+32q)
// per population, the stored entries of the column added serially in storage order (unfused multiply, add).
// SPARSE: counts is the CSC (ri, xv); else a dense int32 R x nM column-major matrix.
template <bool SPARSE>
__global__ void k_em_probs(int R, long long nM, int K, const int* __restrict__ p, const int* __restrict__ ri,
                           const int* __restrict__ xv, const double* __restrict__ phi, const double* __restrict__ theta,
                           const int* __restrict__ s, double* __restrict__ probs, int* __restrict__ znew) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KPL = 8;  // populations per lane: K <= 256
    for (long long c = blockIdx.x * (long long)nwarp + warp; c < nM; c += (long long)gridDim.x * nwarp) {
        double acc[KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) acc[q] = 0.0;
        if (SPARSE) {
            // The column's (row, value) pairs are fetched 32 at a time, coalesced, one per lane, and handed round by
            // shuffles; four phi rows are requested before they are consumed so several L2 requests are in flight.
            // The products are still added in storage order.
            const int e = p[c + 1];
            for (int t0 = p[c]; t0 < e; t0 += 32) {
                const int cnt = min(32, e - t0);
                const int myr = (lane < cnt) ? ri[t0 + lane] : 0;
                const double myx = (lane < cnt) ? (double)xv[t0 + lane] : 0.0;
                for (int u = 0; u < cnt; u += 4) {
                    const double* row[4];
                    double x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        row[j] = phi + (size_t)__shfl_sync(0xffffffffu, myr, (u + j) & 31) * K;
                        x[j] = __shfl_sync(0xffffffffu, myx, (u + j) & 31);
                    }
#pragma unroll
                    for (int q = 0; q < KPL; ++q) {
                        const int k = lane + 32 * q;
                        if (k < K) {
                            double v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) v[j] = row[j][k];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (u + j < cnt) acc[q] += v[j] * x[j];
                        }
                    }
                }
            }
        } else {
            const int* col = xv + c * R;
            for (int r = 0; r < R; ++r) {
                const double x = (double)col[r];
                const double* row = phi + (size_t)r * K;
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int k = lane + 32 * q;
                    if (k < K) acc[q] += row[k] * x;
                }
            }
        }
        const double* th = theta + (size_t)(s[c] - 1) * K;
        double bv = -__longlong_as_double(0x7ff0000000000000LL);
        int bi = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int k = lane + 32 * q;
            if (k < K) {
                const double v = acc[q] + th[k];
                if (probs) probs[c * K + k] = v;
                if (v > bv) {  // ascending k within the lane: strict > keeps the first maximum
                    bv = v;
                    bi = k;
                }
            }
        }
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (znew && lane == 0) znew[c] = (bi == 0x7fffffff ? 0 : bi) + 1;
    }
}

// sum over k of lgamma(colSum_k(n + beta)) for a gene-major [R][K] table, columns added serially (R/celda_C.R:783)
__global__ void k_ll_phi_den_genemajor(const int* __restrict__ n, int R, int K, double beta, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double cs = 0.0;
        for (int r = 0; r < R; ++r) cs += (double)n[(size_t)r * K + k] + beta;
        acc += dlgamma(cs);
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = s;
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- perplexity (a14)
// R/perplexity.R:233-325 with the posterior factors of R/factorizeMatrix.R:280-296.

// den[l] = sum over genes of module l of (nByG + delta), genes added in ascending order (one thread per module)
__global__ void k_perp_psi_den(const long long* __restrict__ nByG, const int* __restrict__ y, int nG, int L, double delta,
                               double* __restrict__ den) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < nG; ++g)
            if (y[g] == l + 1) s += (double)nByG[g] + delta;
        den[l] = s;
    }
}

// column denominators of (table + add): tab is R x K column-major (rowMajor = false) or [R][K] (rowMajor = true)
__global__ void k_perp_col_den(const int* __restrict__ tab, bool rowMajor, int R, int K, double add, double* __restrict__ den) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += (double)(rowMajor ? tab[(size_t)r * K + k] : tab[(size_t)k * R + r]) + add;
        den[k] = s;
    }
}

// M[g][k] (gene-major): celda_CG log(psi[g] phi[y_g, k]) ; celda_C log((nGByCP[g,k] + beta) / colden[k])
__global__ void k_perp_M(int model, int nG, int K, int L, const int* __restrict__ y, const long long* __restrict__ nByG,
                         const double* __restrict__ psiDen, const int* __restrict__ nTSByCP /* L x K col-major */,
                         const int* __restrict__ nGByCP /* [nG][K] */, const double* __restrict__ colDen, double beta,
                         double delta, double* __restrict__ M) {
    const long long tot = (long long)nG * K;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(q / K), k = (int)(q % K);
        if (model == 0) {
            const int l = y[g] - 1;
            const double psi = ((double)nByG[g] + delta) / psiDen[l];
            const double phi = ((double)nTSByCP[(size_t)k * L + l] + beta) / colDen[k];
            M[q] = dlog(psi * phi);
        } else {
            M[q] = dlog(((double)nGByCP[q] + beta) / colDen[k]);
        }
    }
}

// theta[k, s] = log((m + alpha) / colSum(m + alpha))
__global__ void k_perp_theta(const int* __restrict__ m, int K, int nS, double alpha, double* __restrict__ theta) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nS; s += gridDim.x * blockDim.x) {
        double cs = 0.0;
        for (int k = 0; k < K; ++k) cs += (double)m[(size_t)s * K + k] + alpha;
        for (int k = 0; k < K; ++k) theta[(size_t)s * K + k] = dlog(((double)m[(size_t)s * K + k] + alpha) / cs);
    }
}

// per cell: logSumExp_k( sum_nz M[i][k] x + theta[k, s_c] ) -> lse[c]; column total -> tot[c].  One warp per cell.
__global__ void k_perp_cells(long long nM, int K, const int* __restrict__ p, const int* __restrict__ ri,
                             const int* __restrict__ xv, const double* __restrict__ M, const double* __restrict__ theta,
                             const int* __restrict__ s, double* __restrict__ lse) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KPL = 8;
    for (long long c = blockIdx.x * (long long)nwarp + warp; c < nM; c += (long long)gridDim.x * nwarp) {
        double acc[KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) acc[q] = 0.0;
        const int e = p[c + 1];
        for (int t = p[c]; t < e; ++t) {
            const double x = (double)xv[t];
            const double* row = M + (size_t)ri[t] * K;
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const int k = lane + 32 * q;
                if (k < K) acc[q] += row[k] * x;
            }
        }
        const double* th = theta + (size_t)(s[c] - 1) * K;
        double mx = -__longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int k = lane + 32 * q;
            if (k < K) {
                acc[q] += th[k];
                mx = fmax(mx, acc[q]);
            }
        }
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double se = 0.0;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int k = lane + 32 * q;
            if (k < K) se += dexp(acc[q] - mx);
        }
        for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (lane == 0) lse[c] = mx + dlog(se);  // the maximum contributes exp(0) = 1, so se >= 1
    }
}

// celda_G: per cell sum_nz x log(psi[g] phi_cell[y_g, c]), phi_cell = (nTSByC[, c] + beta) / colSum
__global__ void k_perp_cells_G(long long nM, int L, const int* __restrict__ p, const int* __restrict__ ri,
                               const int* __restrict__ xv, const int* __restrict__ nTSByC, const int* __restrict__ y,
                               const long long* __restrict__ nByG, const double* __restrict__ psiDen, double beta,
                               double delta, double* __restrict__ out) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long c = blockIdx.x * (long long)nwarp + warp; c < nM; c += (long long)gridDim.x * nwarp) {
        const int* col = nTSByC + c * L;
        double den = 0.0;
        for (int l = 0; l < L; ++l) den += (double)col[l] + beta;  // every lane, same order
        double acc = 0.0;
        const int e = p[c + 1];
        for (int t = p[c] + lane; t < e; t += 32) {
            const int g = ri[t], l = y[g] - 1;
            const double psi = ((double)nByG[g] + delta) / psiDen[l];
            acc += (double)xv[t] * dlog(psi * (((double)col[l] + beta) / den));
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[c] = acc;
    }
}

// Deterministic sum of a double array and of an int32 array (fixed blocks) -> partial[blockIdx], totals[blockIdx]
__global__ void k_sum_array(const double* __restrict__ v, long long n, double* __restrict__ partial) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += v[i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void k_perp_final(const double* __restrict__ partial, int np, const int* __restrict__ nByC, long long nM,
                             double* out) {
    __shared__ unsigned long long tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    unsigned long long t = 0;
    for (long long c = threadIdx.x; c < nM; c += blockDim.x) t += (unsigned long long)nByC[c];
    atomicAdd(&tot, t);  // nByC: the column totals of the matrix the per-cell terms were computed on
    __syncthreads();
    if (threadIdx.x == 0) {
        double logPx = 0.0;
        for (int i = 0; i < np; ++i) logPx += partial[i];
        out[0] = dexp(-(logPx / (double)tot));
        out[1] = logPx;
        out[2] = (double)tot;
    }
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__global__ void k_philox_raw(const unsigned* __restrict__ ctr, const unsigned* __restrict__ key, int n,
                             unsigned* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Philox4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[0], key[1]);
        for (int q = 0; q < 4; ++q) out[4 * i + q] = r.x[q];
    }
}

// Draws of one sweep in Philox mode (philox.cuh): ix[t] = the t-th visited item (1-based), u[t] = its uniform.
__global__ void k_philox_draws(long long n, unsigned bits, unsigned k0, unsigned k1, unsigned sweep, unsigned stream,
                               int* __restrict__ ix, double* __restrict__ u) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        ix[t] = (int)philox_order((unsigned long long)t, (unsigned long long)n, bits, k0, k1, sweep, stream) + 1;
        u[t] = philox_uniform((unsigned long long)t, k0, k1, sweep, stream);
    }
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- gene-major twin
// column index of every stored entry of a CSC matrix (input of the stable sort that builds the gene-major twin)
__global__ void k_csc_cols(int nc, const int* __restrict__ p, int* __restrict__ colOf, int* __restrict__ iota) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * nwarp + warp; c < nc; c += gridDim.x * nwarp) {
        const int e = p[c + 1];
        for (int t = p[c] + lane; t < e; t += 32) {
            colOf[t] = c;
            iota[t] = t;
        }
    }
}

__global__ void k_gather_twin(long long nnz, const int* __restrict__ perm, const int* __restrict__ colOf,
                              const int* __restrict__ xv, int* __restrict__ tcol, int* __restrict__ tx) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nnz; q += (long long)gridDim.x * blockDim.x) {
        const int t = perm[q];
        tcol[q] = colOf[t];
        tx[q] = xv[t];
    }
}

__global__ void k_row_counts(long long nnz, const int* __restrict__ ri, int* __restrict__ cnt) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nnz; t += (long long)gridDim.x * blockDim.x)
        atomicAdd(&cnt[ri[t]], 1);
}

// list[0 .. *count) = indices g with group[g] != pgroup[g]
__global__ void k_list_changed(const int* __restrict__ group, const int* __restrict__ pgroup, int n, int* __restrict__ list,
                               int* __restrict__ count) {
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        if (group[g] != pgroup[g]) list[atomicAdd(count, 1)] = g;
}

// rowSumByGroupChange (src/matrixSumsSparse.cpp:141-187) through the gene-major twin: only the stored entries of
// genes whose label changed are visited (the reference scans all of them).  One block per changed gene (strided).
__global__ void k_rowsum_change_twin(const int* __restrict__ list, const int* __restrict__ count,
                                     const long long* __restrict__ rp, const int* __restrict__ tcol,
                                     const int* __restrict__ tx, const int* __restrict__ group,
                                     const int* __restrict__ pgroup, int L, int* __restrict__ px) {
    const int n = *count;
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int g = list[q];
        const int a = pgroup[g] - 1, b = group[g] - 1;
        const long long e = rp[g + 1];
        for (long long t = rp[g] + threadIdx.x; t < e; t += blockDim.x) {
            const int v = tx[t];
            const size_t base = (size_t)tcol[t] * L;
            atomicAdd(&px[base + b], v);
            atomicSub(&px[base + a], v);
        }
    }
}

}  // namespace celda

namespace celda {

// ---------------------------------------------------------------------------------------------- EM z step, tiled
// cells per warp of the tiled EM kernel (accumulators in registers, one cell after the other)
constexpr int EM_CPW = 4;

}  // namespace celda

namespace celda {

// colSumByGroup (src/matrixSumsSparse.cpp:9-38) over the gene-major twin: one block per gene (strided), per-warp
// shared accumulators (K counters), no global atomics; every HBM byte of the twin is read once, coalesced, the
// G x K table is written once.  out: gene-major [nG][K].
__global__ void k_colsum_twin(int nG, const long long* __restrict__ rp, const int* __restrict__ tcol,
                              const int* __restrict__ tx, const int* __restrict__ group, int K, int* __restrict__ out) {
    extern __shared__ int acc_tw[];  // [nwarp][K]
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* acc = acc_tw + warp * K;
    for (int g = blockIdx.x; g < nG; g += gridDim.x) {
        for (int k = lane; k < K; k += 32) acc[k] = 0;
        __syncwarp();
        const long long e = rp[g + 1];
        long long t = rp[g] + threadIdx.x;
        const long long step = blockDim.x;
        for (; t + 3 * step < e; t += 4 * step) {
            const int c0 = tcol[t], c1 = tcol[t + step], c2 = tcol[t + 2 * step], c3 = tcol[t + 3 * step];
            const int v0 = tx[t], v1 = tx[t + step], v2 = tx[t + 2 * step], v3 = tx[t + 3 * step];
            const int g0 = group[c0], g1 = group[c1], g2 = group[c2], g3 = group[c3];
            atomicAdd(&acc[g0 - 1], v0);
            atomicAdd(&acc[g1 - 1], v1);
            atomicAdd(&acc[g2 - 1], v2);
            atomicAdd(&acc[g3 - 1], v3);
        }
        for (; t < e; t += step) atomicAdd(&acc[group[tcol[t]] - 1], tx[t]);
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            int sum = 0;
            for (int w = 0; w < nwarp; ++w) sum += acc_tw[w * K + k];
            out[(size_t)g * K + k] = sum;
        }
        __syncthreads();
    }
}

// colSumByGroup over the gene-major twin with 16-byte loads and lane-replicated accumulators: a gene's entries all
// fall into K counters, so plain per-warp counters serialise on same-address conflicts; here each warp keeps CS_REP
// copies of the K counters (copy = lane % CS_REP) and the block folds them once per gene.  out: gene-major [nG][K].
constexpr int CS_REP = 8;
__global__ void __launch_bounds__(256) k_colsum_twin_v4(int nG, const long long* __restrict__ rp, const int* __restrict__ tcol,
                                                        const int* __restrict__ tx, const int* __restrict__ group, int K,
                                                        int* __restrict__ out) {
    extern __shared__ int acc_tw4[];  // [nwarp][CS_REP][K]
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* acc = acc_tw4 + (warp * CS_REP + (lane & (CS_REP - 1))) * K - 1;  // labels are 1-based
    const int nacc = nwarp * CS_REP * K;
    for (int g = blockIdx.x; g < nG; g += gridDim.x) {
        for (int k = threadIdx.x; k < nacc; k += blockDim.x) acc_tw4[k] = 0;
        __syncthreads();
        const long long b = rp[g], e = rp[g + 1];
        const long long b4 = min(e, (b + 3) & ~3LL), e4 = max(b4, e & ~3LL);
        if (b + threadIdx.x < b4) atomicAdd(&acc[group[tcol[b + threadIdx.x]]], tx[b + threadIdx.x]);
        if (e4 + threadIdx.x < e) atomicAdd(&acc[group[tcol[e4 + threadIdx.x]]], tx[e4 + threadIdx.x]);
        const int4* c4 = reinterpret_cast<const int4*>(tcol + b4);
        const int4* x4 = reinterpret_cast<const int4*>(tx + b4);
        const long long n4 = (e4 - b4) >> 2, step = blockDim.x;
        long long t = threadIdx.x;
        for (; t + step < n4; t += 2 * step) {
            const int4 c0 = __ldg(c4 + t), c1 = __ldg(c4 + t + step), v0 = __ldg(x4 + t), v1 = __ldg(x4 + t + step);
            const int g0 = group[c0.x], g1 = group[c0.y], g2 = group[c0.z], g3 = group[c0.w];
            const int g4 = group[c1.x], g5 = group[c1.y], g6 = group[c1.z], g7 = group[c1.w];
            atomicAdd(&acc[g0], v0.x);
            atomicAdd(&acc[g1], v0.y);
            atomicAdd(&acc[g2], v0.z);
            atomicAdd(&acc[g3], v0.w);
            atomicAdd(&acc[g4], v1.x);
            atomicAdd(&acc[g5], v1.y);
            atomicAdd(&acc[g6], v1.z);
            atomicAdd(&acc[g7], v1.w);
        }
        for (; t < n4; t += step) {
            const int4 c0 = __ldg(c4 + t), v0 = __ldg(x4 + t);
            atomicAdd(&acc[group[c0.x]], v0.x);
            atomicAdd(&acc[group[c0.y]], v0.y);
            atomicAdd(&acc[group[c0.z]], v0.z);
            atomicAdd(&acc[group[c0.w]], v0.w);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            int sum = 0;
            for (int w = 0; w < nwarp * CS_REP; ++w) sum += acc_tw4[w * K + k];
            out[(size_t)g * K + k] = sum;
        }
        __syncthreads();
    }
}

// colSumByGroup over the gene-major twin, one WARP per gene (genes handed out through an atomic counter), with the
// cell labels packed to bytes in SHARED memory: the per-entry look-up group[column] is what holds k_colsum_twin_v4
// back (ncu: L1 at 80 %, DRAM at 60 %), a byte gather in shared memory costs a few bank conflicts instead.
// Requires K <= 256 and nM + 4 * nwarp * CS_REP * K bytes of shared memory.  out: gene-major [nG][K].
__global__ void __launch_bounds__(1024, 1) k_colsum_twin_v5(int nG, int nM, const long long* __restrict__ rp,
                                                            const int* __restrict__ tcol, const int* __restrict__ tx,
                                                            const int* __restrict__ group, int K, int* __restrict__ out,
                                                            int* __restrict__ next /* zeroed */) {
    extern __shared__ __align__(16) unsigned char cs5_smem[];
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* accw = reinterpret_cast<int*>(cs5_smem) + warp * CS_REP * K;    // [CS_REP][K] of this warp
    int* acc = accw + (lane & (CS_REP - 1)) * K;
    unsigned char* lab = cs5_smem + sizeof(int) * (size_t)nwarp * CS_REP * K;  // [nM] 0-based labels
    for (int c = threadIdx.x; c < nM; c += blockDim.x) lab[c] = (unsigned char)(group[c] - 1);
    for (int k = lane; k < CS_REP * K; k += 32) accw[k] = 0;
    __syncthreads();
    while (true) {
        int g = 0;
        if (lane == 0) g = atomicAdd(next, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= nG) break;
        const long long b = rp[g], e = rp[g + 1];
        const long long b4 = min(e, (b + 3) & ~3LL), e4 = max(b4, e & ~3LL);
        if (b + lane < b4) atomicAdd(&acc[lab[tcol[b + lane]]], tx[b + lane]);
        if (e4 + lane < e) atomicAdd(&acc[lab[tcol[e4 + lane]]], tx[e4 + lane]);
        const int4* c4 = reinterpret_cast<const int4*>(tcol + b4);
        const int4* x4 = reinterpret_cast<const int4*>(tx + b4);
        const long long n4 = (e4 - b4) >> 2;
        // software pipeline: the next two column / value vectors are in flight while the current two are accumulated
        long long t = lane;
        int4 c0, c1, v0, v1;
        const int4 zero4 = make_int4(0, 0, 0, 0);
        c0 = (t < n4) ? __ldg(c4 + t) : zero4;
        v0 = (t < n4) ? __ldg(x4 + t) : zero4;
        c1 = (t + 32 < n4) ? __ldg(c4 + t + 32) : zero4;
        v1 = (t + 32 < n4) ? __ldg(x4 + t + 32) : zero4;
        for (; t < n4; t += 64) {
            const long long tn = t + 64;
            const int4 nc0 = (tn < n4) ? __ldg(c4 + tn) : zero4, nv0 = (tn < n4) ? __ldg(x4 + tn) : zero4;
            const int4 nc1 = (tn + 32 < n4) ? __ldg(c4 + tn + 32) : zero4, nv1 = (tn + 32 < n4) ? __ldg(x4 + tn + 32) : zero4;
            // a vector beyond the row carries zero counts at column 0: adding 0 is harmless
            atomicAdd(&acc[lab[c0.x]], v0.x);
            atomicAdd(&acc[lab[c0.y]], v0.y);
            atomicAdd(&acc[lab[c0.z]], v0.z);
            atomicAdd(&acc[lab[c0.w]], v0.w);
            atomicAdd(&acc[lab[c1.x]], v1.x);
            atomicAdd(&acc[lab[c1.y]], v1.y);
            atomicAdd(&acc[lab[c1.z]], v1.z);
            atomicAdd(&acc[lab[c1.w]], v1.w);
            c0 = nc0;
            v0 = nv0;
            c1 = nc1;
            v1 = nv1;
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) {
            int sum = 0;
#pragma unroll
            for (int w = 0; w < CS_REP; ++w) {
                sum += accw[w * K + k];
                accw[w * K + k] = 0;
            }
            out[(size_t)g * K + k] = sum;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ TMA helpers (sm_100a)
// One-dimensional bulk copies global -> shared (cp.async.bulk, SASS UBLKCP) completing on an mbarrier.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// generic-proxy writes (this thread's, and those made visible to it) are ordered before its later async-proxy accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned opaque_u32(unsigned x) {
    unsigned y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
// non-blocking variant: polls with test_wait instead of the (possibly suspending) try_wait
__device__ __forceinline__ void mbar_wait_poll(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// .countsTimesProbs + which.max (R/celda_C.R:717-756, 967-977) for large celda_C problems.  Same arithmetic as
// k_em_probs<true> (the stored entries of a column are added in storage order, product and sum unfused).  The kernel is
// bound by the shared-memory / shuffle pipe: every stored entry needs its K phi values (K * 8 bytes of shared-memory
// traffic, 128 bytes per clock and SM) plus the broadcast of (row, count) to the 32 lanes.  So: phi is streamed through
// shared memory in gene tiles by one TMA bulk copy per tile (mbarrier completion), lane l holds populations 2l and
// 2l + 1 and reads them with ONE 16-byte load, and (row - tile start, count) travel in ONE 32-bit shuffle (16 bits each;
// a chunk with a count >= 65536 takes two).  A warp owns EM_CPW cells, one after the other.
__global__ void __launch_bounds__(1024, 1)
k_em_probs_tma(int R, long long nM, int K, int GT, const int* __restrict__ p, const int* __restrict__ ri,
               const int* __restrict__ xv, const double* __restrict__ phi, const double* __restrict__ theta,
               const int* __restrict__ s, double* __restrict__ probs, int* __restrict__ znew) {
    extern __shared__ __align__(128) unsigned char em_smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(em_smem);
    double* tile = reinterpret_cast<double*>(em_smem + 128);  // [GT][K]
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cpb = (long long)nwarp * EM_CPW;  // cells per block pass
    const int k0 = 2 * lane, k1 = 2 * lane + 1;       // K <= 64
    const bool pair = k1 < K, single = k0 < K && !pair;
    const bool vec = (K & 1) == 0;  // rows are 16-byte aligned only for even K
    // lanes beyond K / 2 re-read the row's LAST pair: the same address as the lane before them in their quarter-warp (a
    // broadcast), where the row's first pair would sit in the same banks at another address (one more wavefront per entry)
    const unsigned rowBytes = (unsigned)K * 8u, tileLane = smem_addr(tile) + (pair ? (unsigned)k0 : (unsigned)max(K - 2, 0)) * 8u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    unsigned phase = 0;
    for (long long blk = blockIdx.x; blk * cpb < nM; blk += gridDim.x) {
        long long cell[EM_CPW];
        int cur[EM_CPW], end[EM_CPW];
        double acc0[EM_CPW], acc1[EM_CPW];
#pragma unroll
        for (int j = 0; j < EM_CPW; ++j) {
            cell[j] = blk * cpb + (long long)warp * EM_CPW + j;
            const bool ok = cell[j] < nM;
            cur[j] = ok ? p[cell[j]] : 0;
            end[j] = ok ? p[cell[j] + 1] : 0;
            acc0[j] = 0.0;
            acc1[j] = 0.0;
        }
        for (int g0 = 0; g0 < R; g0 += GT) {
            const int g1 = min(R, g0 + GT);
            __syncthreads();  // the previous tile is no longer read
            if (threadIdx.x == 0) {
                // bulk copies move multiples of 16 bytes: GT is even, an odd last tile of odd K reads 8 bytes of slack
                const unsigned bytes = (unsigned)(((size_t)(g1 - g0) * K * sizeof(double) + 15) & ~(size_t)15);
                mbar_arrive_expect_tx(bar, bytes);
                bulk_copy_g2s(tile, phi + (size_t)g0 * K, bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int j = 0; j < EM_CPW; ++j) {
                while (cur[j] < end[j]) {
                    const int cnt = min(32, end[j] - cur[j]);
                    const int myr = (lane < cnt) ? ri[cur[j] + lane] : 0x7fffffff;
                    const int myxi = (lane < cnt) ? xv[cur[j] + lane] : 0;
                    const int inTile = __popc(__ballot_sync(0xffffffffu, myr < g1));  // rows ascend: a prefix
                    const bool wide = __any_sync(0xffffffffu, lane < inTile && (unsigned)myxi >= 65536u);
                    double a0 = acc0[j], a1 = acc1[j];
                    if (!wide && vec) {
                        // branch-free body: every lane loads 16 bytes (lanes beyond K / 2 re-read the row's last pair and
                        // accumulate into registers nobody reads), 32-bit shared-memory addressing
                        const unsigned pk = ((unsigned)(myr - g0) << 16) | (unsigned)myxi;
#pragma unroll 4
                        for (int u = 0; u < inTile; ++u) {
                            const unsigned q = __shfl_sync(0xffffffffu, pk, u);
                            const double2 v = lds_f64x2(tileLane + (q >> 16) * rowBytes);
                            const unsigned xi = q & 0xffffu;
                            if (xi == 1u) {  // warp-uniform; phi * 1.0 is phi exactly, so the product is skipped (most counts are 1)
                                a0 += v.x;
                                a1 += v.y;
                            } else {
                                const double x = (double)(int)xi;
                                a0 += v.x * x;
                                a1 += v.y * x;
                            }
                        }
                    } else {
                        for (int u = 0; u < inTile; ++u) {
                            const int r = __shfl_sync(0xffffffffu, myr, u) - g0;
                            const double x = (double)__shfl_sync(0xffffffffu, myxi, u);
                            const double* row = tile + (size_t)r * K;
                            if (k0 < K) a0 += row[k0] * x;
                            if (k1 < K) a1 += row[k1] * x;
                        }
                    }
                    acc0[j] = a0;
                    acc1[j] = a1;
                    cur[j] += inTile;
                    if (inTile < cnt || inTile == 0) break;  // the rest of the column belongs to later tiles
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EM_CPW; ++j) {
            if (cell[j] >= nM) continue;
            const long long c = cell[j];
            const double* th = theta + (size_t)(s[c] - 1) * K;
            double bv = -__longlong_as_double(0x7ff0000000000000LL);
            int bi = 0x7fffffff;
            if (k0 < K) {
                const double v = acc0[j] + th[k0];
                if (probs) probs[c * K + k0] = v;
                bv = v;
                bi = k0;
            }
            if (k1 < K) {
                const double v = acc1[j] + th[k1];
                if (probs) probs[c * K + k1] = v;
                if (v > bv) {  // which.max: the lowest index wins a tie
                    bv = v;
                    bi = k1;
                }
            }
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) {
                    bv = ov;
                    bi = oi;
                }
            }
            if (znew && lane == 0) znew[c] = (bi == 0x7fffffff ? 0 : bi) + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ resampled counts
// .resampleCountMatrix (R/perplexity.R:764-775): every cell's counts are redrawn as rmultinom(1, size = colSum,
// prob = column / colSum), i.e. colSum categorical draws over the column's stored entries.  The resample keeps the
// matrix's (p, i) pattern; only the values change (possibly to 0).  One warp per cell: inclusive prefix sums of the
// column in `cum`, then draw t of cell c takes the Philox4x32-10 block (counter = (t / 4, cell, resample, 0x5253),
// key = seed), scales 32 bits to [0, colSum) and finds its entry by binary search.  R's own generator cannot be
// replayed on the device, so this path is pinned statistically.
__global__ void k_resample_prefix(long long nM, const int* __restrict__ p, const int* __restrict__ xv, int* __restrict__ cum,
                                  int* __restrict__ xres) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long c = blockIdx.x * (long long)nwarp + warp; c < nM; c += (long long)gridDim.x * nwarp) {
        const int b = p[c], e = p[c + 1];
        int run = 0;
        for (int t0 = b; t0 < e; t0 += 32) {
            const int t = t0 + lane;
            int v = (t < e) ? xv[t] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += nb;
            }
            if (t < e) {
                cum[t] = run + v;
                xres[t] = 0;
            }
            run += __shfl_sync(0xffffffffu, v, 31);
        }
    }
}

__global__ void k_resample_draw(long long nM, const int* __restrict__ p, const int* __restrict__ cum, unsigned k0, unsigned k1,
                                unsigned resample, int* __restrict__ xres) {
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long c = blockIdx.x * (long long)nwarp + warp; c < nM; c += (long long)gridDim.x * nwarp) {
        const int b = p[c], e = p[c + 1];
        if (e == b) continue;
        const unsigned N = (unsigned)cum[e - 1];
        for (unsigned t4 = lane; 4ull * t4 < N; t4 += 32) {  // one Philox block = four draws
            const Philox4 r = philox4x32_10(t4, (unsigned)c, resample, 0x5253u, k0, k1);
            const unsigned w[4] = {r.x[0], r.x[1], r.x[2], r.x[3]};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (4ull * t4 + q >= N) break;
                const unsigned target = (unsigned)(((unsigned long long)w[q] * N) >> 32);  // uniform on [0, N)
                int lo = b, hi = e - 1;  // first entry with cum > target
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((unsigned)cum[mid] > target) hi = mid; else lo = mid + 1;
                }
                atomicAdd(&xres[lo], 1);
            }
        }
    }
}

// nTSByCP (L x K column-major) of other module labels from the gene-major nGByCP: .cGReDecomposeCounts applied to
// nGByCP as .cCGSplitY does (R/split_clusters.R:483-486).  out pre-zeroed.
__global__ void k_rowsum_genemajor(const int* __restrict__ nGByCP, int nG, int K, const int* __restrict__ y, int L,
                                   int* __restrict__ out) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < (long long)nG * K;
         q += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(q / K), k = (int)(q - (long long)g * K);
        const int v = nGByCP[q];
        if (v) atomicAdd(&out[(size_t)k * L + (y[g] - 1)], v);
    }
}

// An injected visiting order must hold every value of 1..n exactly once (seen: n ints, zeroed): flag is raised otherwise.
__global__ void k_validate_order(const int* __restrict__ ix, long long n, int* __restrict__ seen, int* __restrict__ flag) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int v = ix[t];
        if (v < 1 || v > n || atomicExch(&seen[v - 1], 1) != 0) *flag = 1;
    }
}

}  // namespace celda
