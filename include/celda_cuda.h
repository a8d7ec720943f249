/*
 * celda_cuda.h -- C ABI of libcelda_cuda.so (B200 / sm_100a), the drop-in boundary for celda's
 * collapsed-Gibbs hot path.  Plain pointers and sizes only; every call is synchronous; every
 * function returns 0 on success or a non-zero status whose text is celda_cuda_last_error().
 *
 * Conventions follow R: matrices are column-major, labels are 1-based int32, a dgCMatrix is
 * (p int32[ncol+1], i int32[nnz] 0-based, x double[nnz]).  All pointers are HOST pointers; the
 * library owns device memory.  `*Change` routines update `px` IN PLACE (reference quirk Q3).
 *
 * Each entry point names the reference interface it replaces (paths relative to the celda
 * source tree, campbio/celda 1.18.2).  The R-side binding (.Call shim) is shown in INTEGRATION.md.
 */
#ifndef CELDA_CUDA_H
#define CELDA_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
const char *celda_cuda_last_error(void);
int celda_cuda_device_count(int *n);
/* Select the device used by the stateless entry points below (chain handles carry their own). */
int celda_cuda_set_device(int device);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long celda_cuda_launch_count(void);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (2 flop per FMA; independent register chains on
 * every SM).  The roofline denominator for the FP64-bound sweep kernels: MEASURED_PEAKS.json has no FP64 entry. */
int celda_cuda_fp64_peak(double *tflops);

/* ---- a12: lgamma tables  (R/celda_CG.R:428-429,543; R/celda_G.R:329-332) --------------------
 * out[k] = lgamma(x[k]) / log / exp with the library's device arithmetic (DESIGN.md "arithmetic"). */
int celda_cuda_lgamma(const double *x, long long n, double *out);
int celda_cuda_log(const double *x, long long n, double *out);
int celda_cuda_exp(const double *x, long long n, double *out);

/* ---- a2-a5: by-group aggregation on a dgCMatrix  (src/matrixSumsSparse.cpp) ------------------
 * replaces colSumByGroupSparse(counts, group, K)                  src/matrixSumsSparse.cpp:9-38   */
int celda_colSumByGroupSparse(int nrow, int ncol, const int *p, const int *i, const double *x,
                              const int *group, long long ngroup, int K, double *out /* nrow x K */);
/* replaces rowSumByGroupSparse(counts, group, L)                  src/matrixSumsSparse.cpp:44-73  */
int celda_rowSumByGroupSparse(int nrow, int ncol, const int *p, const int *i, const double *x,
                              const int *group, long long ngroup, int L, double *out /* L x ncol */);
/* replaces colSumByGroupChangeSparse(counts, px, group, pgroup, K) src/matrixSumsSparse.cpp:85-131 */
int celda_colSumByGroupChangeSparse(int nrow, int ncol, const int *p, const int *i, const double *x,
                                    double *px, int px_nrow, const int *group, long long ngroup,
                                    const int *pgroup, long long npgroup, int K);
/* replaces rowSumByGroupChangeSparse(counts, px, group, pgroup, L) src/matrixSumsSparse.cpp:141-187 */
int celda_rowSumByGroupChangeSparse(int nrow, int ncol, const int *p, const int *i, const double *x,
                                    double *px, int px_ncol, const int *group, long long ngroup,
                                    const int *pgroup, long long npgroup, int L);

/* ---- a2-a5 dense twins  (src/matrixSums.c) ---------------------------------------------------
 * `nlevels` stands for nlevels(group); pass a negative value for "group is not a factor".
 * replaces _rowSumByGroup / _rowSumByGroup_numeric                 src/matrixSums.c:5-48,198-236   */
int celda_rowSumByGroup_int(const int *x, int nr, int nc, const int *group, long long ngroup, int nlevels, int *ans);
int celda_rowSumByGroup_numeric(const double *x, int nr, int nc, const int *group, long long ngroup, int nlevels,
                                double *ans);
/* replaces _colSumByGroup / _colSumByGroup_numeric                 src/matrixSums.c:50-93,240-278  */
int celda_colSumByGroup_int(const int *x, int nr, int nc, const int *group, long long ngroup, int nlevels, int *ans);
int celda_colSumByGroup_numeric(const double *x, int nr, int nc, const int *group, long long ngroup, int nlevels,
                                double *ans);
/* replaces _rowSumByGroupChange[_numeric]                          src/matrixSums.c:97-143,282-327 */
int celda_rowSumByGroupChange_int(const int *x, int nr, int nc, int *px, int px_nr, int px_nc, const int *group,
                                  long long ngroup, int nlevels, const int *pgroup, long long npgroup, int pnlevels);
int celda_rowSumByGroupChange_numeric(const double *x, int nr, int nc, double *px, int px_nr, int px_nc,
                                      const int *group, long long ngroup, int nlevels, const int *pgroup,
                                      long long npgroup, int pnlevels);
/* replaces _colSumByGroupChange[_numeric]                          src/matrixSums.c:147-194,331-379 */
int celda_colSumByGroupChange_int(const int *x, int nr, int nc, int *px, int px_nr, int px_nc, const int *group,
                                  long long ngroup, int nlevels, const int *pgroup, long long npgroup, int pnlevels);
int celda_colSumByGroupChange_numeric(const double *x, int nr, int nc, double *px, int px_nr, int px_nc,
                                      const int *group, long long ngroup, int nlevels, const int *pgroup,
                                      long long npgroup, int pnlevels);

/* ---- a10: one gene's module log-probabilities -----------------------------------------------
 * replaces cG_CalcGibbsProbY(index, counts, nTSbyC, nbyTS, nGbyTS, nbyG, y, L, nG, lg_beta, lg_gamma,
 * lg_delta, delta)                                                src/cG_calcGibbsProbY.cpp:187-249
 * The three lgamma tables of the reference are value-identical to direct evaluation and are not taken. */
int celda_cG_CalcGibbsProbY(int index, const double *counts, int ncol, const double *nTSbyC, const double *nbyTS,
                            const int *nGbyTS, const double *nbyG, const int *y, int L, int nG, double beta,
                            double gamma, double delta, double *probs /* L */);

/* ---- a9 helpers -----------------------------------------------------------------------------
 * replaces fastNormPropLog(R_counts, R_alpha)                     src/matrixNorm.cpp:36-52        */
int celda_fastNormPropLog(const double *counts, int nr, int nc, double alpha, double *res);
/* replaces eigenMatMultInt(A, B) / eigenMatMultNumeric(A, B): C = t(A) %*% B   src/eigenMatMultInt.cpp:11-26 */
int celda_eigenMatMultInt(const double *A, int n, int k, const int *B, int m, double *C /* k x m */);
int celda_eigenMatMultNumeric(const double *A, int n, int k, const double *B, int m, double *C);
/* replaces _perplexityG(x, phi, psi, group)                       src/perplexity.c:5-62           */
int celda_perplexityG(const int *x, int nr, int nc, const double *phi, int phi_nr, int phi_nc, const double *psi,
                      int psi_nr, int psi_nc, const int *group, long long ngroup, int nlevels, double *ans);

/* ---- a13: collapsed log-likelihoods -----------------------------------------------------------
 * replaces .cCGCalcLL  R/celda_CG.R:839-888 ; .cCCalcLL  R/celda_C.R:760-788 ; .cGCalcLL  R/celda_G.R:664-702 */
int celda_cCGCalcLL(int K, int L, const int *mCPByS, const double *nTSByCP, const double *nByG, int nGenes,
                    const double *nByTS, const int *nGByTS, int nS, double alpha, double beta, double delta,
                    double gamma, double *ll);
int celda_cCCalcLL(const int *mCPByS, const double *nGByCP, int K, int nS, int nG, double alpha, double beta,
                   double *ll);
int celda_cGCalcLL(const double *nTSByC, const double *nByTS, const double *nByG, int nGenes, const int *nGByTS,
                   long long nM, int L, double beta, double delta, double gamma, double *ll);

/* ---- a8/a9/a10 at sweep granularity, R-closure signatures ------------------------------------
 * replaces .cCCalcGibbsProbZ(counts, mCPByS, nGByCP, nByC, nCP, z, s, K, nG, nM, alpha, beta, doSample)
 *                                                                 R/celda_C.R:620-714
 * counts: dense nG x nM (nTSByC in celda_CG; the raw count matrix in celda_C -- when its nG x K table does not fit
 * shared memory the sweep runs on the raw-count kernel over a CSC copy, same results).  ix/u: injected visiting order
 * (1-based) and uniforms.  mCPByS, nGByCP, nCP, z are updated in place; probs (K x nM) may be NULL. */
int celda_cCCalcGibbsProbZ(const double *counts, int *mCPByS, int nS, double *nGByCP, const double *nByC, double *nCP,
                           int *z, const int *s, int K, int nG, long long nM, double alpha, double beta, int doSample,
                           const int *ix, const double *u, double *probs);
/* replaces .cGCalcGibbsProbY(counts, nTSByC, nByTS, nGByTS, nByG, y, L, nG, beta, delta, gamma, lgbeta,
 * lggamma, lgdelta, doSample)                                     R/celda_G.R:602-660
 * counts: dense nG x ncol (nGByCP in celda_CG).  nTSByC (L x ncol), nByTS, nGByTS, y updated in place. */
int celda_cGCalcGibbsProbY(const double *counts, int ncol, double *nTSByC, double *nByTS, int *nGByTS,
                           const double *nByG, int *y, int L, int nG, double beta, double delta, double gamma,
                           int doSample, const int *ix, const double *u, double *probs);
/* replaces .cCCalcEMProbZ (same arguments as the Gibbs closure)   R/celda_C.R:717-756
 * counts as a dgCMatrix (p != NULL) or dense nG x nM double (p == NULL, x = the matrix). */
int celda_cCCalcEMProbZ(int nG, long long nM, const int *p, const int *i, const double *x, int *mCPByS, int nS,
                        double *nGByCP, double *nCP, int *z, const int *s, int K, double alpha, double beta,
                        int doSample, double *probs);


/* ---- device-resident chain (a1, a6, a7, a15): counts and tables stay in HBM -------------------
 * One handle = one chain of .celda_CG / .celda_C / .celda_G (R/celda_CG.R:367-835, R/celda_C.R:306-616,
 * R/celda_G.R:286-583) on one device.  */
typedef struct celda_chain celda_chain;
enum { CELDA_MODEL_CG = 0, CELDA_MODEL_C = 1, CELDA_MODEL_G = 2 };

/* Upload the count matrix (integrality-checked, narrowed to int32) and build its gene-major twin.
 * Range: the per-entry tables (nTSByC, nGByCP, nTSByCP, nByC) are int32 on the device, the totals (nCP, nByG, nByTS)
 * int64: a cell's total count and every (gene or module, population) total must stay below 2^31 -- the matrix total
 * may exceed it (configs[2]: 5.5e9 counts over 50 populations).  The reference's dgCMatrix path holds these tables
 * as doubles; its integer-matrix path has the same 2^31 bound (R integer overflow). */
int celda_chain_create(celda_chain **out, int device, int model, int nG, int nM, const int *p, const int *i,
                       const double *x, const int *s /* 1-based sample labels, NULL for celda_G */, int nS, int K,
                       int L, double alpha, double beta, double delta, double gamma);
void celda_chain_destroy(celda_chain *h);
/* celdaGridSearch runs every (K, L, chain) of paramsTest on the same counts (R/celdaGridSearch.R:290-391, one
 * `counts` shared by all .celda_CG calls): keep the resident counts, re-size the (K, L)-dependent tables.  The handle
 * needs celda_chain_set_state before the next sweep.  K, L <= 1024. */
int celda_chain_reconfigure(celda_chain *h, int K, int L);
/* .cCGDecomposeCounts / .cCDecomposeCounts / .cGDecomposeCounts from labels (z or y may be NULL per model). */
int celda_chain_set_state(celda_chain *h, const int *z, const int *y);
int celda_chain_get_state(celda_chain *h, int *z, int *y);
/* Sweeps with injected order/uniforms (host pointers).  sweep_y: celda_CG and celda_G handles (the latter on the raw
 * counts, every cell scanned); sweep_z_gibbs: celda_CG and celda_C handles (the latter on the raw counts). */
int celda_chain_sweep_y(celda_chain *h, const int *ix, const double *u, int doSample);
int celda_chain_sweep_z_gibbs(celda_chain *h, const int *ix, const double *u, int doSample);
/* EM z step (.cCCalcEMProbZ, R/celda_C.R:717-756) and one EM-mode iteration of the chain loop. */
int celda_chain_sweep_z_em(celda_chain *h, int doSample);
int celda_chain_iterate_em(celda_chain *h, const int *ix_y, const double *u_y, double *ll);
/* ... or with the device Philox4x32-10 generator (key = seed; counter = item, sweep number, stream): the visiting
 * order is a keyed pseudo-random permutation computed item by item (a Feistel network over Philox with cycle walking,
 * celda_b200/csrc/philox.cuh: no sort), one uniform in (0,1) per visited item.  get_draws returns the order
 * and uniforms the last sweep consumed (which = 0 y, 1 z), so a Philox run can be replayed in injection mode. */
int celda_chain_seed(celda_chain *h, unsigned long long seed, unsigned long long stream);
int celda_chain_sweep_y_philox(celda_chain *h, int doSample);
int celda_chain_sweep_z_gibbs_philox(celda_chain *h, int doSample);
int celda_chain_get_draws(celda_chain *h, int which, int *ix, double *u);
/* Raw generator (known-answer tests): out[4 i ..] = philox4x32_10(counters[4 i ..], key[0..1]) as 32-bit words. */
int celda_cuda_philox4x32(const int *counters, const int *key, int n, int *out);
/* One iteration of the chain loop without split heuristics (R/celda_CG.R:543-602): y sweep, rowSumByGroupChange,
 * z sweep (Gibbs), colSumByGroupChange, log-likelihood.  All four of ix/u NULL => Philox mode.
 * On a celda_C handle: the Gibbs z sweep on the raw counts + .cCCalcLL (R/celda_C.R:427-470; ix_y/u_y ignored);
 * on a celda_G handle: the y sweep on the raw counts + .cGCalcLL (R/celda_G.R:400-424; ix_z/u_z ignored). */
int celda_chain_iterate(celda_chain *h, const int *ix_y, const double *u_y, const int *ix_z, const double *u_z,
                        double *ll);
/* Pre-stage injected draws for n_iter iterations in HBM, then run them without host traffic. */
int celda_chain_stage_draws(celda_chain *h, int n_iter, const int *ix_y, const double *u_y, const int *ix_z,
                            const double *u_z);
int celda_chain_iterate_staged(celda_chain *h, int iter, double *ll);
int celda_chain_loglik(celda_chain *h, double *ll);
/* .cCGCalcLL of the chain's model with ONE label vector replaced (exactly one of z_alt, y_alt non-NULL), chain state
 * untouched: the candidate evaluations of .cCGSplitZ / .cCGSplitY (R/split_clusters.R:236-330, 483-560), which
 * re-aggregate nTSByC by z resp. nGByCP by y (.cCReDecomposeCounts / .cGReDecomposeCounts) and never touch the counts. */
int celda_chain_loglik_alt(celda_chain *h, const int *z_alt, const int *y_alt, double *ll);
/* perplexity(x) of the current state on the model's own counts: .perplexityCelda_CG / _C / _G
 * (R/perplexity.R:233-325) over the posterior factors of R/factorizeMatrix.R:280-296. */
int celda_chain_perplexity(celda_chain *h, double *perplexity);
/* perplexity(x, celdaMod, newCounts) (R/perplexity.R:233-325): the chain's model evaluated on another count matrix of
 * the same dimensions (dgCMatrix slots p, i, x; any sparsity pattern). */
int celda_chain_perplexity_new(celda_chain *h, const int *p, const int *i, const double *x, double *perplexity);
/* .resamplePerplexity(doResampling = TRUE) (R/perplexity.R:449-476) + .resampleCountMatrix (:764-775) for the chain's
 * model: n_resample multinomial resamples of the counts, drawn on the device with Philox4x32-10 (key = seed), and the
 * perplexity of each; perplexity: [n_resample]. */
int celda_chain_perplexity_resampled(celda_chain *h, unsigned long long seed, int n_resample, double *perplexity);
/* The values of the last resample, aligned with the stored entries of the chain's counts (x: [nnz]). */
int celda_chain_get_resampled(celda_chain *h, int *x);
/* Tables as R holds them (double, column-major); any pointer may be NULL. */
int celda_chain_get_tables(celda_chain *h, int *mCPByS, double *nTSByC, double *nTSByCP, double *nGByCP, double *nCP,
                           double *nByC, double *nByG, double *nByTS, int *nGByTS);
int celda_chain_get_probs(celda_chain *h, double *probsZ /* K x nM */, double *probsY /* L x nG */);
/* ---- a9 / a2-a7 across GPUs: cells sharded over ranks (one process per GPU), the partial G x K and K x S count
 * tables combined with an NCCL all-reduce.  id128: a 128-byte ncclUniqueId made on one rank and
 * distributed by the host (R: a file or socket; bench.py: torch.distributed).  NCCL is dlopen'ed at run time. */
typedef struct celda_comm celda_comm;
int celda_comm_unique_id(char *id128);
int celda_comm_create(celda_comm **out, const char *id128, int nranks, int rank, int device);
void celda_comm_destroy(celda_comm *c);
int celda_chain_attach_comm(celda_chain *h, celda_comm *comm);

/* celda_G y sweep (celda_chain_sweep_y / iterate on a celda_G handle): how a gene's row is scanned.
 * mode 0 (default): every cell, exactly like cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:212-225), which adds and
 *   subtracts the same table entry for the cells where the gene has no count -- bit-identical probabilities.
 * mode 1 (opt-in): only the cells where the gene has a count.  The skipped pairs change a log-probability by a few
 *   units of rounding at most, so the draws agree with mode 0 unless a uniform falls within that distance of a
 *   boundary of its cumulative distribution.  celda_chain_scan_margin reports, for the last sweep: per draw, the
 *   distance m between its uniform and the nearest boundary and a bound b on how far the skipped operations can move
 *   that boundary (b = 4 nM 2^-53 (2 N_g log(total count) + max |lgamma table entry|), N_g the gene's total count);
 *   margin[0] = the smallest m / b over the sweep's draws (0 for a draw on the Walker alias branch or with an exact
 *   tie), margin[1] = the number of draws with m <= b.  margin[1] == 0 proves that every label of the sweep is the one
 *   the exact scan draws; the bound is a worst case (it grows with nM), so large matrices leave draws unproven. */
int celda_chain_set_scan_mode(celda_chain *h, int mode);
int celda_chain_scan_margin(celda_chain *h, double *margin /* [2] */);
/* CUDA-event timing on the chain's stream (bench.py): start, then stop returns elapsed ms. */
int celda_chain_timer_start(celda_chain *h);
int celda_chain_timer_stop(celda_chain *h, float *ms);
/* Per-kernel accumulated device time (ms) and launch counts since the last reset, for roofline accounting:
 * which = 0 y sweep, 1 rowSumByGroupChange, 2 z sweep (Gibbs or EM), 3 colSumByGroupChange, 4 log-likelihood,
 * 5 rowSumByGroup (decompose), 6 colSumByGroup (decompose). */
int celda_chain_profile(celda_chain *h, int enable);
int celda_chain_profile_get(celda_chain *h, int which, double *ms, long long *launches,
                            long long *stats /* [8]: rounds, units, movers, CTA-0 clocks in commit / units / draw / barriers,
                                                  carried items drawn again in full */);

#ifdef __cplusplus
}
#endif
#endif
