/*
 * Synthetic-count generator for bench.py and the large parity tests: the two-stage multinomial of celda's
 * .simulateCellscelda_CG (R/simulateCells.R:302-435) -- per cell N_c transcripts, each assigned a module from
 * phi[z_c, ] and then a gene from psi[module, ] -- written for speed (alias tables, one RNG stream per cell,
 * pthreads over cells).  Benchmark support code: not part of libcelda_cuda and not an oracle.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double *prob; int *alias; int n; } alias_t;

static void alias_build(alias_t *t, const double *p, int n)
{
    t->n = n; t->prob = malloc(sizeof(double) * n); t->alias = malloc(sizeof(int) * n);
    int *small = malloc(sizeof(int) * n), *large = malloc(sizeof(int) * n), ns = 0, nl = 0;
    double *q = malloc(sizeof(double) * n), sum = 0;
    for (int i = 0; i < n; i++) sum += p[i];
    for (int i = 0; i < n; i++) { q[i] = p[i] / sum * n; if (q[i] < 1.0) small[ns++] = i; else large[nl++] = i; }
    while (ns && nl) {
        int s = small[--ns], l = large[--nl];
        t->prob[s] = q[s]; t->alias[s] = l;
        q[l] = (q[l] + q[s]) - 1.0;
        if (q[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    while (nl) { int l = large[--nl]; t->prob[l] = 1.0; t->alias[l] = l; }
    while (ns) { int s = small[--ns]; t->prob[s] = 1.0; t->alias[s] = s; }
    free(small); free(large); free(q);
}

static inline uint64_t splitmix(uint64_t *x) { uint64_t z = (*x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoshiro(uint64_t *s) { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
static inline int alias_draw(const alias_t *t, uint64_t *s)
{
    uint64_t r = xoshiro(s);
    int i = (int)(((r >> 32) * (uint64_t)t->n) >> 32);
    double u = (double)(r & 0xffffffffu) * (1.0 / 4294967296.0);
    return u < t->prob[i] ? i : t->alias[i];
}

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

typedef struct {
    int nG, tid, nthreads; long long nM, cell_offset; const int *z, *N, *mod_off, *mod_genes; const alias_t *aphi, *apsi;
    uint64_t seed; int **crow, **cval, *cnnz;
} job_t;

static void *worker(void *arg)
{
    job_t *j = arg;
    int nG = j->nG;
    int *dense = calloc(nG, sizeof(int)), *touched = malloc(sizeof(int) * nG);
    for (long long c = j->tid; c < j->nM; c += j->nthreads) {
        uint64_t sm = j->seed ^ (0xd1342543de82ef95ULL * (uint64_t)(c + j->cell_offset + 1)), s[4];
        for (int q = 0; q < 4; q++) s[q] = splitmix(&sm);
        const alias_t *ap = &j->aphi[j->z[c] - 1];
        int nt = 0;
        for (int t = 0; t < j->N[c]; t++) {
            int l = alias_draw(ap, s);
            int g = j->mod_genes[j->mod_off[l] + alias_draw(&j->apsi[l], s)];
            if (dense[g]++ == 0) touched[nt++] = g;
        }
        /* rows ascending within the column, like a dgCMatrix */
        qsort(touched, nt, sizeof(int), cmp_int);
        j->crow[c] = malloc(sizeof(int) * (nt ? nt : 1)); j->cval[c] = malloc(sizeof(int) * (nt ? nt : 1)); j->cnnz[c] = nt;
        for (int a = 0; a < nt; a++) { j->crow[c][a] = touched[a]; j->cval[c][a] = dense[touched[a]]; dense[touched[a]] = 0; }
    }
    free(dense); free(touched);
    return NULL;
}

/* phi: K x L row-major (cell population -> module distribution); psi: concatenated per-module gene
 * distributions, genes of module l are mod_genes[mod_off[l] .. mod_off[l+1]) with probabilities psi[same range].
 * Output: CSC. colptr int64[nM+1]; *rows_out / *vals_out malloc'd (free with sim_free). Returns nnz.
 * The result depends only on (seed, inputs), not on nthreads: one RNG stream per cell, keyed by the global cell
 * index cell_offset + c, so a rank can generate just its own cell range of a larger matrix. */
long long sim_counts(int nG, long long nM, int K, int L, const int *z /* 1-based */, const int *N, const double *phi,
                     const int *mod_off, const int *mod_genes, const double *psi, uint64_t seed, int nthreads,
                     long long cell_offset, long long *colptr, int **rows_out, int **vals_out)
{
    alias_t *aphi = malloc(sizeof(alias_t) * K), *apsi = malloc(sizeof(alias_t) * L);
    for (int k = 0; k < K; k++) alias_build(&aphi[k], phi + (size_t)k * L, L);
    for (int l = 0; l < L; l++) alias_build(&apsi[l], psi + mod_off[l], mod_off[l + 1] - mod_off[l]);
    int **crow = malloc(sizeof(int *) * nM), **cval = malloc(sizeof(int *) * nM);
    int *cnnz = malloc(sizeof(int) * nM);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64]; job_t jobs[64];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){nG, t, nthreads, nM, cell_offset, z, N, mod_off, mod_genes, aphi, apsi, seed, crow, cval, cnnz};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    colptr[0] = 0;
    for (long long c = 0; c < nM; c++) colptr[c + 1] = colptr[c] + cnnz[c];
    long long nnz = colptr[nM];
    int *rows = malloc(sizeof(int) * (nnz ? nnz : 1)), *vals = malloc(sizeof(int) * (nnz ? nnz : 1));
    for (long long c = 0; c < nM; c++) {
        memcpy(rows + colptr[c], crow[c], sizeof(int) * cnnz[c]);
        memcpy(vals + colptr[c], cval[c], sizeof(int) * cnnz[c]);
        free(crow[c]); free(cval[c]);
    }
    free(crow); free(cval); free(cnnz);
    for (int k = 0; k < K; k++) { free(aphi[k].prob); free(aphi[k].alias); }
    for (int l = 0; l < L; l++) { free(apsi[l].prob); free(apsi[l].alias); }
    free(aphi); free(apsi);
    *rows_out = rows; *vals_out = vals;
    return nnz;
}

void sim_free(int *p) { free(p); }
