// Sequential Gibbs sweeps as persistent cooperative kernels (sm_100a).
//
// The reference visits cells (R/celda_C.R:637-705) or genes (R/celda_G.R:620-651) strictly in order; step i
// reads the count tables step i-1 may have changed.  Two facts make an exact parallel schedule possible:
//   (1) the log-probability of candidate j for an item depends only on column j of the table (plus the
//       per-candidate scalars), and on whether j is the item's current label;
//   (2) an item that does not move changes nothing.
// So a window of upcoming items is evaluated speculatively by the whole grid against the current tables, the
// first item (in visiting order) whose draw differs from its label is committed, and for the items behind it
// only the two candidates the move touched are re-evaluated ("patch units") before they are drawn again.
// Every floating-point value is produced by the same operations in the same order as a sequential sweep, so
// labels, tables and probabilities are bit-identical to it.
//
// Work decomposition: one warp per (item, candidate) "unit": lanes evaluate the lgamma terms in parallel into
// shared memory, lane 0 adds them in the reference's serial order.  One warp per item draws.
#pragma once
#include "celda_math.cuh"

namespace celda {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned*)counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// R sort.c revsort semantics (descending heapsort carrying the index permutation; not stable), 1 thread.
__device__ __noinline__ void revsort_serial(double* a, int* ib, int n) {
    if (n <= 1) return;
    a--;
    ib--;
    int l = (n >> 1) + 1, ir = n, i, j, ii;
    double ra;
    for (;;) {
        if (l > 1) {
            l = l - 1;
            ra = a[l];
            ii = ib[l];
        } else {
            ra = a[ir];
            ii = ib[ir];
            a[ir] = a[1];
            ib[ir] = ib[1];
            if (--ir == 1) {
                a[1] = ra;
                ib[1] = ii;
                return;
            }
        }
        i = l;
        j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) {
                a[i] = a[j];
                ib[i] = ib[j];
                j += (i = j);
            } else
                j = ir + 1;
        }
        a[i] = ra;
        ib[i] = ii;
    }
}

// Ordered (ascending index) sum of r[0..n), identical on every lane: the serial order of R's sum().  Entries are
// >= +0 and adding +0.0 to a non-negative partial sum is the identity, so zeros are skipped: a chunk of 32 with few
// positive entries is added through shuffles, a dense chunk by a serial loop over shared memory (broadcast reads).
__device__ __forceinline__ double warp_ordered_sum_pos(const double* r, int n, int lane) {
    // few positive entries overall: chunk by chunk through shuffles
    int npos = 0;
    for (int base = 0; base < n; base += 32) npos += __popc(__ballot_sync(0xffffffffu, base + lane < n && r[base + lane] > 0.0));
    double s = 0.0;
    if (npos <= 12) {
        for (int base = 0; base < n; base += 32) {
            const double v = (base + lane < n) ? r[base + lane] : 0.0;
            unsigned m = __ballot_sync(0xffffffffu, v > 0.0);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                s += shfl_d(v, b);
            }
        }
        return s;
    }
    // dense vector: one serial pass, loads (broadcast) running ahead of the dependent adds
#pragma unroll 8
    for (int i = 0; i < n; ++i) s += r[i];
    return s;
}

#ifdef CELDA_DEBUG_STATS
__device__ unsigned long long g_dbg2[8];  // clocks by sampler segment (general path), lane 0
#define CELDA_DBG_SEG(k) do { if (lane == 0) { const long long c_ = clock64(); atomicAdd(&g_dbg2[k], (unsigned long long)(c_ - dbg_c0)); dbg_c0 = c_; } } while (0)
__device__ unsigned long long g_dbg[8];  // 0 draws, 1 compact hits, 2 general path, 3 rank sort, 4 exact path, 5 max clocks of one draw
#define CELDA_DBG_ADD(k) do { if (lane == 0) atomicAdd(&g_dbg[k], 1ull); } while (0)
#else
#define CELDA_DBG_ADD(k) do { } while (0)
#define CELDA_DBG_SEG(k) do { } while (0)
#endif

// .sampleLl (R/celda_functions.R:1-11) + sample.int(n, 1, TRUE, prob) with the uniform supplied.
// pl: n log-probabilities in shared memory (overwritten on the rare exact path).  r: n doubles of warp scratch;
// hs may alias pl; perm: n ints.  n <= 1024.
// Returns the 0-based label, or <0: -1 Walker-alias branch (>200 likely categories), -2 non-finite
// probability, -3 no positive probability.  Warp-cooperative; all lanes return the same value.
// mx_out / nz_out describe the vector the draw was made from: the maximum log-probability and the (up to four)
// entries whose exp(pl - max) is positive, x = -2 when there are more (see draw_unchanged).
__device__ int warp_sample_ll(double* pl, int n, double u, double* r, int* perm, int lane, double& mx_out, int4& nz_out) {
    double* hs = pl;
    const unsigned FULL = 0xffffffffu;
    nz_out = make_int4(-2, -1, -1, -1);
    CELDA_DBG_ADD(0);
#ifdef CELDA_DEBUG_STATS
    long long dbg_c0 = clock64();
#endif
    double mx = -__longlong_as_double(0x7ff0000000000000LL);
    for (int i = lane; i < n; i += 32) mx = fmax(mx, pl[i]);
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    int bad = 0;
    for (int i = lane; i < n; i += 32) {
        const double e = dexp(pl[i] - mx);
        r[i] = e;
        bad |= !(e >= 0.0) || (e > 1.7976931348623157e308);
    }
    __syncwarp();
    mx_out = mx;
    if (__any_sync(FULL, bad)) return -2;
    {
        int np = 0, i0 = -1, i1 = -1, i2 = -1, i3 = -1;
        for (int base = 0; base < n && np <= 4; base += 32) {
            unsigned m = __ballot_sync(FULL, base + lane < n && r[base + lane] > 0.0);
            np += __popc(m);
            if (np > 4) break;
            for (int q = np - __popc(m); m; m &= m - 1, ++q) {
                const int b = base + __ffs(m) - 1;
                i0 = (q == 0) ? b : i0;
                i1 = (q == 1) ? b : i1;
                i2 = (q == 2) ? b : i2;
                i3 = (q == 3) ? b : i3;
            }
        }
        if (np <= 4) nz_out = make_int4(i0, i1, i2, i3);
    }
    // Compact path: at most 32 entries have exp() > 0 (the usual case once a chain has settled).  Lane k takes the k-th
    // of them in index order and the whole pipeline -- the two ordered sums, the two normalisations, the descending
    // order and the cumulative walk -- runs on one entry per lane.  Zero entries contribute +0.0 to the sums, fail
    // n * q > 0.1 and sort behind every positive entry, so the arithmetic is the one the general path below performs.
    // Anything unusual (exact ties, u beyond the total mass, an all-zero vector) falls through to the general path.
    if (n > 1) {
        int P = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const double v = (i < n) ? r[i] : 0.0;
            const unsigned m = __ballot_sync(FULL, v > 0.0);
            if (v > 0.0) {
                const int at = P + __popc(m & ((1u << lane) - 1u));
                if (at < 32) {
                    perm[at] = i;
                    hs[at] = v;
                }
            }
            P += __popc(m);
        }
        __syncwarp();
        if (P == 1 && u <= 1.0) {  // a single positive entry: e = 1 exactly, q = 1 / 1 / 1, and u <= 1 selects it at the first position
            const int res = perm[0];
            __syncwarp();
            CELDA_DBG_ADD(1);
            return res;
        }
        if (P >= 1 && P <= 32) {
            const int myIdx = (lane < P) ? perm[lane] : -1;
            const double myE = (lane < P) ? hs[lane] : 0.0;
            __syncwarp();
            double c1 = 0.0;
            for (int k = 0; k < P; ++k) c1 += shfl_d(myE, k);
            const double r1 = myE / c1;
            double c2 = 0.0;
            for (int k = 0; k < P; ++k) {
                const double v = shfl_d(r1, k);
                if (v > 0.0) c2 += v;
            }
            if (c2 > 0.0) {
                const double q = r1 / c2;
                const int Ppos = __popc(__ballot_sync(FULL, q > 0.0));
                int gt = 0, eq = 0;
                for (int k = 0; k < P; ++k) {
                    const double v = shfl_d(q, k);
                    gt += (v > q);
                    eq += (v == q);
                }
                const bool tie = __any_sync(FULL, q > 0.0 && eq > 1);
                if (!tie) {
                    if (q > 0.0) {
                        hs[gt] = q;
                        perm[gt] = myIdx;
                    }
                    __syncwarp();
                    double cumc = 0.0;
                    int res = -1;
                    for (int k = 0; k < Ppos; ++k) {  // every lane, same order
                        cumc = (k == 0) ? hs[k] : hs[k] + cumc;
                        if (k == n - 1 || u <= cumc) {
                            res = perm[k];
                            break;
                        }
                    }
                    __syncwarp();
                    if (res >= 0) {
                        CELDA_DBG_ADD(1);
                        return res;
                    }
                }
            }
        }
        __syncwarp();
    }
    CELDA_DBG_ADD(2);
    CELDA_DBG_SEG(0);  // max, exp, list, compact attempt
    const double s1 = warp_ordered_sum_pos(r, n, lane);
    CELDA_DBG_SEG(1);
    for (int i = lane; i < n; i += 32) r[i] = r[i] / s1;
    __syncwarp();
    CELDA_DBG_SEG(2);
    const double s2 = warp_ordered_sum_pos(r, n, lane);  // FixupProb: plain-double sum of the positive entries
    CELDA_DBG_SEG(3);
    if (!(s2 > 0.0)) return -3;
    int nc = 0, nzero = 0;
    for (int i = lane; i < n; i += 32) {
        const double q = r[i] / s2;
        r[i] = q;
        nc += ((double)n * q > 0.1);
        nzero += (q == 0.0);
    }
    for (int o = 16; o; o >>= 1) {
        nc += __shfl_xor_sync(FULL, nc, o);
        nzero += __shfl_xor_sync(FULL, nzero, o);
    }
    __syncwarp();
    CELDA_DBG_SEG(4);
    if (nc > 200) {
        // sample.int switches to Walker's alias method when more than 200 categories are likely (R random.c,
        // walker_ProbSampleReplace): alias tables built serially from q = n p, then ONE uniform picks a cell and either
        // its own category or its alias.  Rare (K or L > 200 with a diffuse vector): one lane, hs (the log-probabilities,
        // no longer needed) holds the small/large work list, perm the aliases.
        int res = 0;
        __syncwarp();
        if (lane == 0) {
            int* HL = reinterpret_cast<int*>(hs);
            int H = -1, Lp = n;
            for (int i = 0; i < n; ++i) {
                r[i] = r[i] * (double)n;
                perm[i] = i;  // an entry that is never given an alias keeps itself (R leaves it unset; reachable only through rounding)
                if (r[i] < 1.0) HL[++H] = i; else HL[--Lp] = i;
            }
            if (H >= 0 && Lp < n) {
                for (int k = 0; k < n - 1; ++k) {
                    const int i = HL[k], j = HL[Lp];
                    perm[i] = j;
                    r[j] += r[i] - 1.0;
                    if (r[j] < 1.0) ++Lp;
                    if (Lp >= n) break;
                }
            }
            const double rU = u * (double)n;
            const int k = min((int)rU, n - 1);
            res = (rU < r[k] + (double)k) ? k : perm[k];
        }
        res = __shfl_sync(FULL, res, 0);
        __syncwarp();
        return res;
    }
    if (n == 1) return 0;
    // Scan the values in descending order (the order revsort produces is unique up to ties).  Peaked
    // distributions end within a few warp arg-max steps; a diffuse one (more than 16 positive entries) switches to a rank sort after two.
    double cum = 0.0;
    int result = -1;
    bool slow = false;
    unsigned taken = 0;  // bit q: entry lane + 32 q of this lane has been consumed (n <= 1024)
    const int P = n - nzero;
    const int klimit = (P > 16) ? 2 : n;
    int k = 0;
    for (; k < klimit; ++k) {
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int i = lane, q = 0; i < n; i += 32, ++q) {
            const double v = ((taken >> q) & 1u) ? -1.0 : r[i];
            if (v > bv) {
                bv = v;
                bi = i;
            }
        }
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        const int rem = n - k;
        bool select;
        if (bv == 0.0) {
            // only zero-probability entries remain: the cumulative sum no longer moves, so no earlier position
            // can satisfy u <= cum and R falls through to the last sorted position.
            if (nzero == 1) {
                result = bi;
            } else
                slow = true;
            break;
        }
        cum = (k == 0) ? bv : bv + cum;
        select = (rem == 1) || (u <= cum);
        if (select) {
            int mult = 0;
            for (int i = lane; i < n; i += 32) mult += (r[i] == bv);
            for (int o = 16; o; o >>= 1) mult += __shfl_xor_sync(FULL, mult, o);
            if (mult == 1)
                result = bi;
            else
                slow = true;  // the label inside a group of exactly equal probabilities depends on heapsort order
            break;
        }
        if ((bi & 31) == lane) taken |= 1u << (bi >> 5);
    }
    CELDA_DBG_SEG(5);  // arg-max steps
    if (!slow && result < 0) {
        CELDA_DBG_ADD(3);
        // Rank sort of the positive entries (hs aliases pl, which is no longer needed): position = number of larger
        // values.  Entries are >= +0, so their bit patterns order like the values and integer compares leave the FP64
        // pipe alone.  A lane keeps up to five of its entries in registers and passes over the vector once per group.
        // Any exact tie among positive entries leaves the order to heapsort => exact path.
        bool tie = false;
        const long long* rb = reinterpret_cast<const long long*>(r);
        for (int q = lane; q < n; q += 32) perm[q] = -1;
        __syncwarp();
        for (int i0 = lane; i0 < n; i0 += 160) {
            long long b[5];
            int gt[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int idx = i0 + 32 * q;
                b[q] = (idx < n) ? rb[idx] : 0;
                gt[q] = 0;
            }
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const long long v = rb[j];
#pragma unroll
                for (int q = 0; q < 5; ++q) gt[q] += (v > b[q]);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int idx = i0 + 32 * q;
                if (idx < n && b[q] > 0) {
                    hs[gt[q]] = __longlong_as_double(b[q]);
                    perm[gt[q]] = idx;
                }
            }
        }
        __syncwarp();
        for (int q = lane; q < P; q += 32) tie |= (perm[q] < 0);  // equal entries share a position and leave a hole behind it
        __syncwarp();
        if (__any_sync(FULL, tie)) {
            slow = true;
        } else {
            for (; k < P; ++k) {  // continue the scan where the arg-max steps stopped; every lane, same order
                cum = (k == 0) ? hs[k] : hs[k] + cum;
                if (k == n - 1 || u <= cum) {
                    result = perm[k];
                    break;
                }
            }
            if (result < 0) {  // every positive entry passed: R falls through to the last (zero) position
                if (nzero == 1) {
                    int zi = 0x7fffffff;
                    for (int i = lane; i < n; i += 32)
                        if (r[i] == 0.0) zi = i;
                    for (int o = 16; o; o >>= 1) zi = min(zi, __shfl_xor_sync(FULL, zi, o));
                    result = zi;
                } else
                    slow = true;
            }
        }
        __syncwarp();
        CELDA_DBG_SEG(6);  // rank sort + walk
    }
    if (!slow) return result;
    // Exact emulation of ProbSampleReplace on one lane (rare: exact ties, or u beyond the total mass).
    CELDA_DBG_ADD(4);
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < n; ++i) {
            hs[i] = r[i];
            perm[i] = i;
        }
        revsort_serial(hs, perm, n);
        for (int i = 1; i < n; ++i) hs[i] += hs[i - 1];
        int j;
        for (j = 0; j < n - 1; ++j)
            if (u <= hs[j]) break;
        result = perm[j];
    }
    result = __shfl_sync(FULL, result, 0);
    __syncwarp();
    return result;
}

// Block-wide: smallest c in [0, nwin) whose ring slot spec[ring_slot(pos + c)] holds a mover, or -1.  Result valid on
// all threads.
__device__ int block_first_mover(const int* spec, long long pos, int Wmax, int nwin, int* s_first) {
    if (threadIdx.x == 0) *s_first = 0x7fffffff;
    __syncthreads();
    int best = 0x7fffffff;
    for (int c = threadIdx.x; c < nwin; c += blockDim.x) {
        if (__ldcg(spec + ((pos + c) & (long long)(Wmax - 1))) >= 0) {
            best = c;
            break;  // ascending per thread
        }
    }
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != 0x7fffffff) atomicMin(s_first, best);
    __syncthreads();
    const int f = *s_first;
    __syncthreads();
    return f == 0x7fffffff ? -1 : f;
}

// Ring slot of item position t; the ring capacity is a power of two (SweepBufs).
__device__ __forceinline__ size_t ring_slot(long long t, int Wmax) { return (size_t)(t & (long long)(Wmax - 1)); }

struct SweepWork {
    double* bufP;       // [Wmax * ncand] per-(item, candidate) partial results
    double* bufA;       // [Wmax * ncand] second per-candidate scalar (z sweep only)
    int* spec;          // [Wmax] ring: packed (old << 16 | new) for movers, -1 otherwise
    double* drawMx;     // [Wmax] ring: max log-probability of the item's last draw
    int* drawNz;        // [Wmax][4] ring: its positive entries (see warp_sample_ll)
    double* colsum;     // [K + 4 MCMAX]  (z sweep) cached column sums: base columns, then two banks of commit slots
    double* lgnCP;      // [K + 4 MCMAX]  (z sweep)
    unsigned* barrier;  // zeroed before launch
    int* status;        // 0 ok, else sampler error code
    long long* stats;   // [0] rounds, [1] units, [2] movers, [3..6] CTA-0 clocks in commit / units / draw / barriers,
                        // [7] low 40 bits: carried items drawn again in full (all CTAs); high bits: discarded tentative commits
    int Wmax;           // ring capacity (items)
    int Wfull;          // steady-state window: units per round = a whole number of grid-wide lane passes
    int Wunit;          // items whose units fill one grid-wide lane pass
    int Wmin;
    int lgT;            // entries of the shared-memory lgamma(k + beta) table
    int MC;             // tentative commits per round (k_sweep_z / k_sweep_y), 1..MCMAX
};

// A carried item after the commit (da, db): its draw is a function of e = exp(pl - max) and u alone, and only the
// entries da and db of pl changed.  If both had e == 0 at the item's last draw (neither is among its recorded positive
// entries) and both still have, the vector e is bit-identical and so is the outcome: the recorded draw stands.
// The test is tried only while movers are sparse (gap estimate gavg >= 32): in a churning sweep nearly every item has
// many positive entries and the failed test would only add latency to the round.
// Two steps so that an item that fails the list test costs one 16-byte load: draw_recorded_zero() -- neither da nor
// db was a positive entry at the last draw (mx = that draw's maximum); then the caller evaluates the two new
// log-probabilities and asks draw_still_zero().
__device__ __forceinline__ bool draw_recorded_zero(const SweepWork& w, size_t slot, int da, int db, double& mx) {
    const int4 nz = __ldcg(reinterpret_cast<const int4*>(w.drawNz) + slot);
    if (nz.x == -2) return false;
    if (nz.x == da || nz.y == da || nz.z == da || nz.w == da) return false;
    if (nz.x == db || nz.y == db || nz.z == db || nz.w == db) return false;
    mx = __ldcg(w.drawMx + slot);
    return true;
}
__device__ __forceinline__ bool draw_still_zero(double mx, double pa, double pb) {
    return dexp(pa - mx) == 0.0 && dexp(pb - mx) == 0.0;
}

// ------------------------------------------------------------------------------------------------ multi-commit rounds
// k_sweep_z / k_sweep_y commit SEVERAL movers per grid-wide round.  After the draws of a window, the first M movers
// (visiting order) become *tentative* commits.  Commit i yields two new table columns ("slots" K + 2i, K + 2i + 1:
// the old label's column without the item, the new label's column with it); the base columns stay untouched.  An item
// with q tentative commits in front of it sees, for column j, the slot of the last of those q commits that touched j,
// else the base column: exactly the table state a sequential sweep would show it if those commits are real.  Its
// "stale" columns -- the ones touched by the q commits, plus the columns of commits that were discarded in the round
// before (Dcarry) -- are re-evaluated against those slots and its draw is repeated (or proven unchanged).  An item
// whose outcome differs from the one recorded in the ring is marked INVALID: every commit in front of the first
// invalid item stands (all items before it, movers included, reproduced their outcome under exactly the commits in
// front of them), every commit behind it is discarded, and its columns become Dcarry for the items still in the
// window.  The first invalid item's new outcome was computed from standing commits only, so it is the true one; the
// next round starts there.  With M = 1 this is the single-commit schedule; the arithmetic of every unit and every
// draw is unchanged, so results stay bit-identical to the sequential sweep (scripts/proto_multicommit.py checks the
// bookkeeping against a sequential sweep on random problems).
constexpr int MCMAX = 32;           // array bound of the commit records; SweepWork::MC <= MCMAX - 1 (one lane per commit + 1)
constexpr int SPEC_INV = 1 << 30;   // spec ring, movers: the last validation changed this item's outcome

// spec ring encoding: -1 non-mover; (old << 16 | new) mover; invalid marks carry the 3-bit tag of the round that set
// them (non-mover: -2 - tag; mover: | SPEC_INV | tag << 27), so a mark is honoured by the next round only.
__device__ __forceinline__ int spec_mark(int v, int tag) { return v < 0 ? -2 - tag : (v | SPEC_INV | (tag << 27)); }
__device__ __forceinline__ int spec_clean(int v) { return v < 0 ? -1 : (v & 0x03ffffff); }
__device__ __forceinline__ bool spec_invalid(int v, int tag) {
    return v < -1 ? (-2 - v) == tag : ((v & SPEC_INV) && ((v >> 27) & 7) == tag);
}

// Block-wide: smallest c in [0, nwin) whose ring slot carries an invalid mark of round `tag`, or -1; anyMover tells
// whether the slots examined hold a mover at all (a quiet window needs no selection pass).  s_first: 2 ints.
__device__ int block_first_invalid(const int* spec, long long pos, int Wmax, int nwin, int tag, int* s_first, bool& anyMover) {
    if (threadIdx.x == 0) {
        s_first[0] = 0x7fffffff;
        s_first[1] = 0;
    }
    __syncthreads();
    int best = 0x7fffffff, mv = 0;
    for (int c0 = threadIdx.x; c0 < nwin && best == 0x7fffffff; c0 += 4 * blockDim.x) {
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + k * blockDim.x;
            v[k] = (c < nwin) ? __ldcg(spec + ((pos + c) & (long long)(Wmax - 1))) : -1;
        }
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            if (spec_invalid(v[k], tag)) best = c0 + k * blockDim.x;
            mv |= (v[k] >= 0);
        }
    }
    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    mv = __any_sync(0xffffffffu, mv);
    if ((threadIdx.x & 31) == 0) {
        if (best != 0x7fffffff) atomicMin(&s_first[0], best);
        if (mv) s_first[1] = 1;
    }
    __syncthreads();
    const int f = s_first[0];
    // a thread that stopped at an invalid item has not seen the rest of the window: assume movers there
    anyMover = s_first[1] != 0 || f != 0x7fffffff;
    __syncthreads();
    return f == 0x7fffffff ? -1 : f;
}

// Block-wide ordered compaction: the first (visiting order) <= cap movers among the ring slots of [pos, pos + nwin).
// out_t[k] = position, out_pk[k] = clean packed (old << 16 | new).  s_tmp: 130 ints.  Returns the number selected.
// Four chunks of blockDim items per pass (four loads in flight per thread, two barriers per pass).
__device__ int block_select_movers(const int* spec, long long pos, int Wmax, int nwin, int cap, long long* out_t, int* out_pk,
                                   int* s_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int total = 0;
    for (int base = 0; base < nwin && total < cap; base += 4 * blockDim.x) {
        int v[4];
        unsigned bal[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = base + k * blockDim.x + threadIdx.x;
            v[k] = (c < nwin) ? __ldcg(spec + ((pos + c) & (long long)(Wmax - 1))) : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bal[k] = __ballot_sync(0xffffffffu, v[k] >= 0);
            if (lane == 0) s_tmp[k * 32 + warp] = __popc(bal[k]);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive prefix over the 4 x nwarp counts in (chunk, warp) order: lane <-> 4 consecutive counts
            int c4[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = lane * 4 + q, k = idx / nwarp, w = idx - k * nwarp;
                c4[q] = (idx < 4 * nwarp) ? s_tmp[k * 32 + w] : 0;
                sum += c4[q];
            }
            int inc = sum;
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += nb;
            }
            int run = inc - sum;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = lane * 4 + q, k = idx / nwarp, w = idx - k * nwarp;
                if (idx < 4 * nwarp) s_tmp[k * 32 + w] = run;
                run += c4[q];
            }
            if (lane == 31) s_tmp[128] = inc;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rank = total + s_tmp[k * 32 + warp] + __popc(bal[k] & ((1u << lane) - 1u));
            if (v[k] >= 0 && rank < cap) {
                out_t[rank] = pos + base + k * blockDim.x + threadIdx.x;
                out_pk[rank] = spec_clean(v[k]);
            }
        }
        total += s_tmp[128];
        __syncthreads();
    }
    return min(total, cap);
}

// Tentative commits wanted per round: a round costs a fixed latency (two grid barriers, the scans, one draw) plus
// patch work that grows with (window x commits); the window itself has to span the commits, so with g items between
// movers the cost per mover is about F / M + c g (M + 4), smallest near M = sqrt(F / (c g)).
__device__ __forceinline__ int mc_target(int gavg, int MC) {
    const int m = (int)sqrtf(48000.0f / (float)max(gavg, 1));
    return max(1, min(m, MC));
}

// Round bookkeeping done by ONE warp (lane <-> tentative commit) while the other warps copy table columns: the
// Dcarry list of the discarded commits, the next tentative commits' records, their source slots, the stale-column
// list CL with its per-commit prefix lengths nq, and the patch-unit segment offsets.  Everything here is a few
// shared-memory words per lane with __syncwarp between steps; one __syncthreads publishes the result.
struct CommitBank {
    long long t[MCMAX];   // visiting position of the mover
    long long N[MCMAX];   // its total count
    int a[MCMAX], b[MCMAX], item[MCMAX], aux[MCMAX], srcA[MCMAX], srcB[MCMAX], pk[MCMAX];
};

// slot of candidate j for an item with q commits of bank cb in front of it (ncand base slots first)
__device__ __forceinline__ int bank_slot_of(const CommitBank* cb, int ncand, int q, int j) {
    for (int i = q - 1; i >= 0; --i) {
        if (cb->b[i] == j) return ncand + 2 * i + 1;
        if (cb->a[i] == j) return ncand + 2 * i;
    }
    return j;
}

// Warp-level ordered selection of the first <= cap movers of [pos, pos + nsel) (nsel <= a few thousand).
__device__ __forceinline__ int warp_select_movers(const int* spec, long long pos, int Wmax, int nsel, int cap, CommitBank* nb,
                                                  int lane) {
    int Mn = 0;
    for (int base = 0; base < nsel && Mn < cap; base += 256) {
        int v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = base + k * 32 + lane;
            v[k] = (c < nsel) ? __ldcg(spec + ((pos + c) & (long long)(Wmax - 1))) : -1;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool mover = v[k] >= 0;
            const unsigned bal = __ballot_sync(0xffffffffu, mover);
            const int rank = Mn + __popc(bal & ((1u << lane) - 1u));
            if (mover && rank < cap) {
                nb->t[rank] = pos + base + k * 32 + lane;
                nb->pk[rank] = spec_clean(v[k]);
            }
            Mn += __popc(bal);
        }
    }
    return min(Mn, cap);
}

// Dcarry of the commits [qf, M) of the old bank, in commit order without repeats; returns its length (all lanes).
__device__ __forceinline__ int warp_build_dcarry(const CommitBank* ob, int qf, int M, int* dc_s, int lane) {
    bool newA = false, newB = false;
    int ca = -1, cb = -1;
    if (lane >= qf && lane < M) {
        ca = ob->a[lane];
        cb = ob->b[lane];
        newA = newB = true;
        for (int k = qf; k < lane; ++k) {
            const int ak = ob->a[k], bk = ob->b[k];
            newA &= (ak != ca) & (bk != ca);
            newB &= (ak != cb) & (bk != cb);
        }
    }
    const int cnt = (int)newA + (int)newB;
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
    }
    const int at = inc - cnt;
    if (newA) dc_s[at] = ca;
    if (newB) dc_s[at + (int)newA] = cb;
    return __shfl_sync(0xffffffffu, inc, 31);
}

// Source slots, stale list CL (Dcarry first, then each commit's columns where new), nq, segment offsets.
// The records a/b/t of the new bank are in place.  hiC = end of the carried items.
__device__ __forceinline__ void warp_build_lists(CommitBank* nb, int ncand, int Mn, int ndc, const int* dc_s, int* CL_s, int* nq_s,
                                                 int* seg_s, long long pos, long long hiC, int lane) {
    bool newA = false, newB = false;
    int ca = -1, cb = -1;
    if (lane < Mn) {
        ca = nb->a[lane];
        cb = nb->b[lane];
        nb->srcA[lane] = bank_slot_of(nb, ncand, lane, ca);
        nb->srcB[lane] = bank_slot_of(nb, ncand, lane, cb);
        newA = newB = true;
        for (int d = 0; d < ndc; ++d) {
            const int c = dc_s[d];
            newA &= (c != ca);
            newB &= (c != cb);
        }
        for (int k = 0; k < lane; ++k) {
            const int ak = nb->a[k], bk = nb->b[k];
            newA &= (ak != ca) & (bk != ca);
            newB &= (ak != cb) & (bk != cb);
        }
    }
    for (int d = lane; d < ndc; d += 32) CL_s[d] = dc_s[d];
    const int cnt = (int)newA + (int)newB;
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int at = ndc + inc - cnt;
    if (newA) CL_s[at] = ca;
    if (newB) CL_s[at + (int)newA] = cb;
    if (lane == 0) nq_s[0] = ndc;
    if (lane < Mn) nq_s[lane + 1] = ndc + inc;
    __syncwarp();
    // segment q = the carried items with exactly q commits in front: (m_{q-1}, m_q], the last one up to hiC - 1
    int units = 0;
    if (lane <= Mn) {
        const long long last = (lane < Mn) ? nb->t[lane] : hiC - 1;
        const long long prevLast = (lane == 0) ? pos - 1 : nb->t[lane - 1];
        units = (int)max(0LL, last - prevLast) * nq_s[lane];
    }
    int incU = units;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incU, o);
        if (lane >= o) incU += v;
    }
    if (lane == 0) seg_s[0] = 0;
    seg_s[lane + 1] = incU;  // lanes beyond Mn repeat the total: seg_s[Mn + 1] is what the kernels read
    __syncwarp();
}


// ------------------------------------------------------------------------------------------------ z sweep
// .cCCalcGibbsProbZ (R/celda_C.R:620-714) on a dense R x nM int32 count matrix (nTSByC in celda_CG).
constexpr int ZT_MAXGRP = 30;  // row groups of the value tables

// columns of the z sweep's table with memory of their own: the K base columns and the slots of the first two tentative
// commits, rounded up so that the value-table buffer behind them stays 8-byte aligned (the row stride is odd)
__host__ __device__ inline int z_sweep_kb(int K, int MC) {
    const int kb = K + 2 * (MC < 2 ? MC : 2);
    return kb + (kb & 1);
}

struct ZArgs {
    int R, K, nS;
    long long nM;
    const int* counts;  // R x nM column-major
    const int* nByC;    // [nM]
    const int* s;       // [nM] 1-based
    int* z;             // [nM] 1-based, updated
    int* nTab;          // R x K column-major (nGByCP argument), updated
    long long* nCP;     // [K], updated
    int* mCPByS;        // K x nS, updated
    double alpha, beta;
    const int* ix;      // visiting order, 1-based
    const double* u;    // one uniform per visited item
    int doSample;
    double* probs;      // K x nM or null
    // value tables of the lane-per-unit rounds (see k_sweep_z)
    const int* offg;    // [R + 1] row r tabulates counts v in [0, offg[r+1] - offg[r])
    const int* grp;     // [ZT_MAXGRP + 3]: grp[0] = number of row groups, then the group boundaries (rows); last entry = 1 when
                        // every row's table covers all its counts (no look-up can miss)
    int EsCap;          // shared-memory table capacity in entries; 0 disables the tables
    int vtCap;          // doubles reserved for the table head / the commit slots beyond the second commit (z_sweep_vtcap)
    const int* xT;      // [R][nM]: xT[r][t] = counts[r, ix[t] - 1], the counts in visiting order, cell-minor
    const int* zT;      // [nM] zT[t] = z[ix[t] - 1] at the start of the sweep (a cell's label changes only at its own visit)
    const int* nT;      // [nM] nT[t] = nByC[ix[t] - 1]
    const double* lgG;  // lgamma(k + beta), k < TG, L2-resident (see lg_tab2); null / 0 without a chain
    int TG;
    double* vtg;        // [K][vtStride] the value tables of a round, built once by the whole grid, copied by the users
    long long vtStride; // entries per candidate (>= offg[R])
    SweepWork w;
};

__device__ __forceinline__ double lg_tab(const double* lgtab, int T, long long k, double beta) {
    return (k < T) ? lgtab[k] : dlgamma_hot((double)k + beta);
}
// The same value with a second, L2-resident level lgG[k] = lgamma(k + beta), k < TG (built once per chain by
// k_lgG_fill with the same routine, so the bits agree); TG = 0 without a chain.
__device__ __forceinline__ double lg_tab2(const double* lgtab, int T, const double* __restrict__ lgG, int TG, long long k,
                                          double beta) {
    if (k < T) return lgtab[k];
    if (k < TG) return __ldg(lgG + k);
    return dlgamma_hot((double)k + beta);
}

// One (candidate, cell) partial sum over the rows [r0, r1) of a row group: S += table[row][count], rows in order.
// xp points at the cell's count in row r0 (stride nM per row); a row's table starts at vt_s[offg_s[row] - e0].
// CHECK = false when every table covers all the counts of its row.
template <bool CHECK>
__device__ __forceinline__ double z_table_rows(const int* xp, long long nM, int r0, int r1, const int* offg_s, int e0,
                                               const double* vt_s, const int* ncol, const double* lgtab, int T,
                                               double beta, double S) {
    int r = r0;
    for (; r + 8 <= r1; r += 8) {  // eight row loads in flight, then eight look-ups in row order
        int v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = __ldg(xp);
            xp += nM;
        }
        int og = offg_s[r];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ogn = offg_s[r + k + 1];
            double val;
            if (CHECK)
                val = (v[k] < ogn - og) ? vt_s[og - e0 + v[k]] : lg_tab(lgtab, T, (long long)ncol[r + k] + v[k], beta);
            else
                val = vt_s[og - e0 + v[k]];
            S += val;
            og = ogn;
        }
    }
    for (; r < r1; ++r) {
        const int v = __ldg(xp);
        xp += nM;
        const int og = offg_s[r];
        double val;
        if (CHECK)
            val = (v < offg_s[r + 1] - og) ? vt_s[og - e0 + v] : lg_tab(lgtab, T, (long long)ncol[r] + v, beta);
        else
            val = vt_s[og - e0 + v];
        S += val;
    }
    return S;
}

__global__ void __launch_bounds__(1024, 1) k_sweep_z(ZArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.w.status[1]) return;  // the injected visiting order failed its check: nothing may be indexed through it
    const int R = a.R, K = a.K, nS = a.nS, T = a.w.lgT, MC = a.w.MC;
    // Column slots: K base columns + two per tentative commit.  The first KB columns have memory of their own; the
    // slots beyond (commits 2, 3, ...) lie over the head of the value-table buffer vt_s, which only table rounds use --
    // and a round that may become a table round selects at most two commits (see the window policy below).
    const int KS = K + 2 * MC, KB = z_sweep_kb(K, MC);
    const int Rp = R | 1;       // odd stride: candidate-major rows, conflict-free across lanes of r
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scratchD = max(R, 2 * K + (K + 1) / 2) + 2;
    double* lgtab = reinterpret_cast<double*>(smem_raw);
    double* scratch_all = lgtab + T;
    long long* nCP_s = reinterpret_cast<long long*>(scratch_all + (size_t)nwarp * scratchD);  // [KS]
    CommitBank* banks = reinterpret_cast<CommitBank*>(nCP_s + KS);  // [2] tentative commits of this round / the last
    long long* st_s = reinterpret_cast<long long*>(banks + 2);  // [12] statistics and CTA-0 clocks (thread 0)
    int* n_s = reinterpret_cast<int*>(st_s + 12);            // [KB][Rp] (+ the slots KB.. over vt_s); KB * Rp is even
    double* vt_s = reinterpret_cast<double*>(n_s + (size_t)KB * Rp);  // [vtCap] value table head of this CTA's candidate
    int* m_s = reinterpret_cast<int*>(vt_s + a.vtCap);
    int* offg_s = m_s + K * nS;           // [R + 1]
    int* grp_s = offg_s + R + 1;          // [ZT_MAXGRP + 3]
    int* fin_s = grp_s + ZT_MAXGRP + 3;   // [K] slot of column j with every tentative commit applied
    int* CL_s = fin_s + K;                // [K] stale-column list: Dcarry first, then the commits' columns in order
    int* dc_s = CL_s + K;                 // [K] Dcarry
    int* nq_s = dc_s + K;                 // [MCMAX + 1] stale columns of an item with q commits in front
    int* seg_s = nq_s + MCMAX + 1;        // [MCMAX + 2] patch-unit offsets per segment
    int* s_misc = seg_s + MCMAX + 2;      // [144]
    double* scr = scratch_all + (size_t)warp * scratchD;
    // Lane-per-unit rounds: a CTA tabulates, for the candidate j it is serving,
    // lgamma(n[r][j] + v + beta) over the counts v that occur in row r, one group of rows at a time (as many rows
    // as fit the shared-memory table).  A unit then is R look-ups added in the reference's order instead of R
    // lgamma evaluations; its partial sum waits in the ring buffer between row groups.  The tables hold exactly
    // the values lg_tab() returns, and a count beyond a row's table is evaluated directly.
    const bool useTab = a.EsCap > 0;
    (void)KS;

    const double Rbeta = (double)R * a.beta;
    const int gwarp = blockIdx.x * nwarp + warp, tw = gridDim.x * nwarp;
    // warp-granular work goes round the CTAs first (item c -> CTA c mod grid), so a short list of items lands on
    // many SMs with one warp each instead of filling a few SMs
    const int gwarpT = warp * gridDim.x + blockIdx.x;

    for (int q = threadIdx.x; q < R * K; q += blockDim.x) n_s[(q / R) * Rp + (q % R)] = a.nTab[q];
    for (int q = threadIdx.x; q < K; q += blockDim.x) {
        nCP_s[q] = a.nCP[q];
        fin_s[q] = q;
    }
    for (int q = threadIdx.x; q < K * nS; q += blockDim.x) m_s[q] = a.mCPByS[q];
    for (int q = threadIdx.x; q < T; q += blockDim.x) lgtab[q] = dlgamma((double)q + a.beta);
    if (useTab) {
        for (int q = threadIdx.x; q <= R; q += blockDim.x) offg_s[q] = a.offg[q];
        for (int q = threadIdx.x; q < ZT_MAXGRP + 3; q += blockDim.x) grp_s[q] = a.grp[q];
    }
    if (threadIdx.x <= MCMAX) nq_s[threadIdx.x] = 0;
    if (threadIdx.x <= MCMAX + 1) seg_s[threadIdx.x] = 0;
    CommitBank* cb = banks;
    __syncthreads();

    long long pos = 0, hi = 0, prev_end = 0;
    int W = a.w.Wfull, gavg = a.w.Wfull, M = 0, mcAimd = MC, par = 0, tagPrev = 0;
    bool have_spec = false, first = true;
    unsigned bar_target = 0;
    int roundNo = 0;
    int cheapScore = 4, cheapTick = 0;
    if (threadIdx.x < 12) st_s[threadIdx.x] = (threadIdx.x == 9) ? clock64() : 0;  // 0 rounds 1 units 2 movers 3 discards 4 redraws 5..8 clocks
    __syncthreads();
#define CELDA_TICKR(acc) do { ck1 = clock64(); acc += ck1 - ck0; ck0 = ck1; } while (0)
#define CELDA_TICK(k) do { if (threadIdx.x == 0) { const long long c_ = clock64(); st_s[k] += c_ - st_s[9]; st_s[9] = c_; } } while (0)
#ifdef CELDA_COMMIT_PROF
#define CELDA_CSUB(k) do { const long long c_ = clock64(); csub[k] += c_ - csub0; csub0 = c_; } while (0)
#define CELDA_CSUB_BEGIN() csub0 = clock64()
#else
#define CELDA_CSUB(k) do { } while (0)
#define CELDA_CSUB_BEGIN() do { } while (0)
#endif

    // the per-slot cached sums live in global memory (written by whichever warp evaluated them); the commit slots
    // alternate between two banks so that a round never overwrites what the next round's first phase still reads
    auto gslot = [&](int slot, int bank) { return slot < K ? slot : K + bank * 2 * MCMAX + (slot - K); };
    auto window_for = [&](int g, int mc) -> int {
        const long long target = (long long)g * (mc + 4);
        if (target >= a.w.Wunit) return (int)min((target / a.w.Wunit) * a.w.Wunit, (long long)a.w.Wfull);
        return (int)max((long long)a.w.Wmin, target);
    };

    long long csub[6] = {0, 0, 0, 0, 0, 0}, csub0 = 0;
    (void)csub; (void)csub0;
    while (true) {
        // ---- settle the previous round: which tentative commits stand, where the window restarts, the next commits
        CELDA_CSUB_BEGIN();
        CommitBank* ob = cb;              // the finished round's commits
        cb = banks + par;                 // this round's
        bool hasDc = false;
        if (have_spec) {
            const int nwinPrev = (int)(prev_end - pos);
            bool anyMover;
            const int f = block_first_invalid(a.w.spec, pos, a.w.Wmax, nwinPrev, tagPrev, s_misc, anyMover);
            CELDA_CSUB(0);
            const long long tf = pos + f;
            const int Mold = M;
            int qf = Mold;
            if (f >= 0) {
                qf = 0;
                for (int i = 0; i < Mold; ++i) qf += (ob->t[i] < tf);
            }
            hasDc = qf < Mold;
            const long long oldpos = pos;
            if (qf < Mold) pos = tf;  // no unselected mover can lie in front of a discarded selected one
            else if (Mold > 0) pos = ob->t[Mold - 1] + 1;
            if (threadIdx.x == 0) {
                st_s[2] += qf;
                st_s[3] += Mold - qf;
            }
            if (Mold > 0) mcAimd = (qf == Mold) ? min(MC, 2 * mcAimd) : max(1, qf + 1);
            if (qf > 0) gavg = max(1, (int)((3LL * gavg + max(1LL, (pos - oldpos) / qf)) / 4));
            int mcCur = min(mcAimd, mc_target(gavg, MC));
            W = window_for(gavg, mcCur);
            if (useTab && 2LL * W * K + 64 >= (long long)tw * 16) {
                // this window could make a table round (total >= tw * 16), whose value tables lie over the commit slots
                // beyond the second commit: take the most commits whose window keeps every round small, else two
                const long long fit = ((long long)tw * 16 - 64) / (2LL * K * gavg) - 4;
                mcCur = (fit >= 3) ? (int)min((long long)mcCur, fit) : min(mcCur, 2);
                W = window_for(gavg, mcCur);
            }
            hi = min(prev_end, pos + W);
            const int nsel = (int)(hi - pos);
            const bool wideSel = nsel > 1024;  // a large window is scanned by the whole block, a small one by warp 0
            int Mwide = 0;
            if (wideSel && anyMover) Mwide = block_select_movers(a.w.spec, pos, a.w.Wmax, nsel, mcCur, cb->t, cb->pk, s_misc + 2);
            if (warp == 0) {
                // one warp: Dcarry, the next commits and their records, source slots, stale list, segments
                const int ndc = warp_build_dcarry(ob, qf, Mold, dc_s, lane);
                const int Mn = wideSel ? Mwide : (anyMover ? warp_select_movers(a.w.spec, pos, a.w.Wmax, nsel, mcCur, cb, lane) : 0);
                __syncwarp();
                if (lane < Mn) {
                    const int cell = a.ix[cb->t[lane]] - 1, pk = cb->pk[lane];
                    cb->item[lane] = cell;
                    cb->a[lane] = pk >> 16;
                    cb->b[lane] = pk & 0xffff;
                    cb->aux[lane] = a.s[cell] - 1;
                    cb->N[lane] = a.nByC[cell];
                }
                __syncwarp();
                warp_build_lists(cb, K, Mn, ndc, dc_s, CL_s, nq_s, seg_s, pos, hi, lane);
                if (lane == 0) s_misc[140] = Mn;
            } else {
                // the other warps: standing commits advance the base columns (in order: a later commit's slot
                // supersedes an earlier one's) and the per-column scalars
                for (int r = threadIdx.x - 32; r < R; r += blockDim.x - 32)
                    for (int i = 0; i < qf; ++i) {
                        n_s[ob->a[i] * Rp + r] = n_s[(K + 2 * i) * Rp + r];
                        n_s[ob->b[i] * Rp + r] = n_s[(K + 2 * i + 1) * Rp + r];
                    }
                for (int j = (int)blockDim.x - 1 - (int)threadIdx.x; j < K; j += blockDim.x - 32) {
                    const int sl = bank_slot_of(ob, K, qf, j);
                    if (sl != j) {
                        nCP_s[j] = nCP_s[sl];
                        if (blockIdx.x == 0) {
                            a.w.colsum[j] = __ldcg(a.w.colsum + gslot(sl, par ^ 1));
                            a.w.lgnCP[j] = __ldcg(a.w.lgnCP + gslot(sl, par ^ 1));
                        }
                    }
                }
                if (warp == 1 && lane < qf) {
                    atomicAdd(&m_s[ob->aux[lane] * K + ob->a[lane]], -1);
                    atomicAdd(&m_s[ob->aux[lane] * K + ob->b[lane]], 1);
                    if (blockIdx.x == 0) a.z[ob->item[lane]] = ob->b[lane] + 1;
                }
            }
            __syncthreads();
            M = s_misc[140];
            CELDA_CSUB(1);
            if (M == 0 && !hasDc) {  // nothing moves and nothing is stale: the carried items are settled
                gavg = (int)min(2LL * gavg + (hi - pos), (long long)(1 << 24));
                pos = hi;
                W = window_for(gavg, 1);
            }
        }
        if (pos >= a.nM) break;
        CELDA_CSUB(2);
        // ---- this round's tentative commits: new column slots
        {
            // the movers' count columns, all loads in flight at once (staged in the warp scratch, idle here)
            int* xs = reinterpret_cast<int*>(scratch_all);
            for (int q = threadIdx.x; q < M * R; q += blockDim.x) {
                const int i = q / R, r = q - i * R;
                xs[q] = __ldg(a.counts + (size_t)cb->item[i] * R + r);
            }
            __syncthreads();
            CELDA_CSUB(3);
            for (int r = threadIdx.x; r < R; r += blockDim.x)
                for (int i = 0; i < M; ++i) {
                    const int xv = xs[i * R + r];
                    n_s[(K + 2 * i) * Rp + r] = n_s[cb->srcA[i] * Rp + r] - xv;
                    n_s[(K + 2 * i + 1) * Rp + r] = n_s[cb->srcB[i] * Rp + r] + xv;
                }
        }
        if (threadIdx.x == blockDim.x - 1)
            for (int i = 0; i < M; ++i) {
                nCP_s[K + 2 * i] = nCP_s[cb->srcA[i]] - cb->N[i];
                nCP_s[K + 2 * i + 1] = nCP_s[cb->srcB[i]] + cb->N[i];
            }
        for (int j = threadIdx.x; j < K; j += blockDim.x) fin_s[j] = bank_slot_of(cb, K, M, j);
        __syncthreads();
        CELDA_CSUB(4);
        CELDA_TICK(5);
        // after commits the window is topped up with fresh items only once it is less than half full: carried items
        // cost a few patch units and a check, fresh ones a full evaluation and draw, cheaper in batches
        const bool topup = first || (M == 0 && !hasDc) || (hi - pos) * 2 < W;
        const long long end = topup ? max(hi, min(pos + (long long)W, a.nM)) : hi;
        const int nwin = (int)(end - pos);
        const int nColU = first ? K : 2 * M;     // cached column sums: every base column once, then the new slots
        const long long nPatch = seg_s[M + 1];   // stale columns of the carried items
        const long long nFresh = (end - hi) * K;
        const long long total = nColU + nPatch + nFresh;
        const int tag = roundNo & 7;
        ++roundNo;
        // ---- A1: units.  Large rounds: one LANE per unit (the lane evaluates and adds its unit's terms serially, no
        // staging, no single-lane phases).  Small rounds (churn): one WARP per unit for latency.  Both orders of
        // evaluation are the reference's, so the results are identical.
        // unit index -> (position t or -1 for a column sum, candidate j, column slot)
        auto decode = [&](long long uidx, long long& t, int& j, int& slot) {
            t = -1;
            if (uidx < nColU) {
                slot = first ? (int)uidx : K + (int)uidx;
                j = -1;
            } else if (uidx < nColU + nPatch) {
                int rel = (int)(uidx - nColU), q = 0;
                while (rel >= seg_s[q + 1]) ++q;
                rel -= seg_s[q];
                const int wq = nq_s[q];
                t = ((q == 0) ? pos : cb->t[q - 1] + 1) + rel / wq;
                j = CL_s[rel % wq];
                slot = bank_slot_of(cb, K, q, j);
            } else {
                const long long q = uidx - nColU - nPatch;
                t = hi + q / K;
                j = (int)(q % K);
                slot = fin_s[j];
            }
        };
        // one warp per unit: lanes evaluate the lgamma terms into shared memory, lane 0 adds them serially
        auto warp_unit = [&](long long t, int j, int slot) {
            const int* ncol = n_s + slot * Rp;
            if (t < 0) {
                for (int r = lane; r < R; r += 32) scr[r] = lg_tab2(lgtab, T, a.lgG, a.TG, ncol[r], a.beta);
                __syncwarp();
                if (lane == 0) {
                    double S = 0.0;
#pragma unroll 8
                    for (int r = 0; r < R; ++r) S += scr[r];  // loads run ahead of the serial adds
                    a.w.colsum[gslot(slot, par)] = S;
                    a.w.lgnCP[gslot(slot, par)] = dlgamma((double)nCP_s[slot] + Rbeta);
                }
                __syncwarp();
            } else {
                const long long i = a.ix[t] - 1;
                const int zi = a.z[i] - 1;
                const int sgn = (j == zi) ? -1 : 1;
                const int* x = a.counts + i * R;
                for (int r = lane; r < R; r += 32) scr[r] = lg_tab2(lgtab, T, a.lgG, a.TG, (long long)ncol[r] + sgn * x[r], a.beta);
                __syncwarp();
                if (lane == 0) {
                    double S = 0.0;
#pragma unroll 8
                    for (int r = 0; r < R; ++r) S += scr[r];  // loads run ahead of the serial adds
                    const size_t o = ring_slot(t, a.w.Wmax) * K + j;
                    a.w.bufP[o] = S;
                    a.w.bufA[o] = dlgamma_hot((double)(nCP_s[slot] + (long long)sgn * a.nByC[i]) + Rbeta);
                }
                __syncwarp();
            }
        };
        // one lane per unit, direct evaluation
        auto lane_unit = [&](long long t, int j, int slot) {
            const int* ncol = n_s + slot * Rp;
            if (t < 0) {
                double S = 0.0;
                for (int r = 0; r < R; ++r) S += lg_tab2(lgtab, T, a.lgG, a.TG, ncol[r], a.beta);
                a.w.colsum[gslot(slot, par)] = S;
                a.w.lgnCP[gslot(slot, par)] = dlgamma((double)nCP_s[slot] + Rbeta);
            } else {
                const long long i = a.ix[t] - 1;
                const int zi = a.z[i] - 1;
                const int sgn = (j == zi) ? -1 : 1;
                const int* x = a.counts + i * R;
                double S = 0.0;
#pragma unroll 2
                for (int r = 0; r < R; ++r) S += lg_tab2(lgtab, T, a.lgG, a.TG, (long long)ncol[r] + sgn * __ldg(x + r), a.beta);
                const size_t o = ring_slot(t, a.w.Wmax) * K + j;
                a.w.bufP[o] = S;
                a.w.bufA[o] = dlgamma_hot((double)(nCP_s[slot] + (long long)sgn * a.nByC[i]) + Rbeta);
            }
        };
        const bool bigRound = total >= (long long)tw * 16;
        if (bigRound && useTab && M <= 2) {
            // (1) cached column sums and patch units: one warp each while they are few, else one lane each
            const long long nSmall = nColU + nPatch;
            if (nSmall < (long long)tw * 8) {
                for (long long uidx = gwarpT; uidx < nSmall; uidx += tw) {
                    int j, slot;
                    long long t;
                    decode(uidx, t, j, slot);
                    warp_unit(t, j, slot);
                }
            } else {
                for (long long uidx = (long long)gwarp * 32 + lane; uidx < nSmall; uidx += (long long)tw * 32) {
                    int j, slot;
                    long long t;
                    decode(uidx, t, j, slot);
                    lane_unit(t, j, slot);
                }
            }
            const long long nFreshCells = end - hi;
            // (1b) this round's value tables, every entry once: vtg[j][e] = lgamma(n[r][j] + v + beta) for the e-th
            // (row, count) pair; the CTAs that serve candidate j copy the row group they need into shared memory
            {
                const int Etot = offg_s[R];
                const long long nEnt = (long long)K * Etot;
                for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nEnt;
                     idx += (long long)gridDim.x * blockDim.x) {
                    const int j = (int)(idx / Etot), e = (int)(idx - (long long)j * Etot);
                    int lo = 0, hiR = R;  // row with offg[row] <= e < offg[row + 1]
                    while (hiR - lo > 1) {
                        const int mid = (lo + hiR) >> 1;
                        if (offg_s[mid] <= e) lo = mid; else hiR = mid;
                    }
                    a.vtg[(size_t)j * a.vtStride + e] =
                        lg_tab2(lgtab, T, a.lgG, a.TG, (long long)n_s[fin_s[j] * Rp + lo] + (e - offg_s[lo]), a.beta);
                }
                grid_sync(a.w.barrier, bar_target);
            }
            // (2) fresh cells x candidates other than the cell's own: table look-ups.  The K x nchunk (candidate,
            // 32-cell chunk) tasks are cut into gridDim equal contiguous runs, candidate-major, so a CTA serves one
            // candidate or the tail of one and the head of the next, building that candidate's table first.
            {
                const int ngrp = grp_s[0];
                const bool complete = grp_s[ZT_MAXGRP + 2] != 0;
                const long long nchunk = (nFreshCells + 31) / 32;
                const long long totalCh = (long long)K * nchunk;
                const long long c0 = totalCh * blockIdx.x / gridDim.x, c1 = totalCh * (blockIdx.x + 1) / gridDim.x;
                for (long long cs = c0; cs < c1;) {
                    const int j = (int)(cs / nchunk);
                    const long long segEnd = min(c1, (long long)(j + 1) * nchunk);
                    const long long chLo = cs - (long long)j * nchunk, chHi = segEnd - (long long)j * nchunk;
                    cs = segEnd;
                    const int* ncol = n_s + fin_s[j] * Rp;
                    const long long nCPj = nCP_s[fin_s[j]];
                    for (int gI = 0; gI < ngrp; ++gI) {
                        const int r0 = grp_s[1 + gI], r1 = grp_s[2 + gI];
                        const int e0 = offg_s[r0], ne = offg_s[r1] - e0;
                        __syncthreads();
                        {
                            const double* src = a.vtg + (size_t)j * a.vtStride + e0;
                            for (int q0 = threadIdx.x; q0 < ne; q0 += 8 * blockDim.x) {  // eight loads in flight per thread
                                double v[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const int q = q0 + k * blockDim.x;
                                    v[k] = (q < ne) ? __ldcg(src + q) : 0.0;
                                }
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const int q = q0 + k * blockDim.x;
                                    if (q < ne) vt_s[q] = v[k];
                                }
                            }
                        }
                        __syncthreads();
                        for (long long ch = chLo + warp; ch < chHi; ch += nwarp) {
                            const long long q = ch * 32 + lane;
                            if (q >= nFreshCells) continue;
                            const long long t = hi + q;
                            if (__ldg(a.zT + t) - 1 == j) continue;  // the cell's own population: step (3)
                            const int* x = a.xT + t;  // lanes hold consecutive t: one coalesced load per row
                            const size_t o = ring_slot(t, a.w.Wmax) * K + j;
                            double S = (gI == 0) ? 0.0 : a.w.bufP[o];
                            const int* xp = x + (size_t)r0 * a.nM;  // walks down the rows: + nM per row
                            S = complete ? z_table_rows<false>(xp, a.nM, r0, r1, offg_s, e0, vt_s, ncol, lgtab, T, a.beta, S)
                                         : z_table_rows<true>(xp, a.nM, r0, r1, offg_s, e0, vt_s, ncol, lgtab, T, a.beta, S);
                            a.w.bufP[o] = S;
                            if (gI == ngrp - 1)
                                a.w.bufA[o] = dlgamma_hot((double)(nCPj + (long long)__ldg(a.nT + t)) + Rbeta);
                        }
                    }
                }
            }
            // (3) each fresh cell against its own population (counts removed): direct evaluation, one warp each
            for (long long q = gwarpT; q < nFreshCells; q += tw) {
                const long long t = hi + q;
                const int zi = a.z[a.ix[t] - 1] - 1;
                warp_unit(t, zi, fin_s[zi]);
            }
        } else if (bigRound) {
            for (long long uidx = (long long)gwarp * 32 + lane; uidx < total; uidx += (long long)tw * 32) {
                int j, slot;
                long long t;
                decode(uidx, t, j, slot);
                lane_unit(t, j, slot);
            }
        } else {
            for (long long uidx = gwarpT; uidx < total; uidx += tw) {
                int j, slot;
                long long t;
                decode(uidx, t, j, slot);
                warp_unit(t, j, slot);
            }
        }
        if (threadIdx.x == 0) {
            st_s[0] += 1;
            st_s[1] += total;
        }
        CELDA_TICK(6);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICK(8);
        // ---- A2: combine in the reference's order, draw / validate
        double* pl = scr;
        double* rr = scr + K;
        int* perm = reinterpret_cast<int*>(scr + 2 * K);
        // carried items in front of the first tentative commit have nothing stale unless commits were discarded
        const int cstart = (hasDc || M == 0) ? 0 : (int)(cb->t[0] + 1 - pos);
        for (int c = cstart + gwarpT; c < nwin; c += tw) {
            const long long t = pos + c;
            const bool fresh = t >= hi;
            int q = M;
            if (!fresh) {
                q = 0;
                for (int i = 0; i < M; ++i) q += (cb->t[i] < t);
            }
            const int ncols = fresh ? K : nq_s[q];
            if (ncols == 0) continue;
            const long long i = a.ix[t] - 1;
            const int zi = a.z[i] - 1, si = a.s[i] - 1;
            const size_t slot = ring_slot(t, a.w.Wmax), o = slot * K;
            const int oldv = fresh ? -1 : __ldcg(a.w.spec + slot);
            auto combine = [&](int j) {
                const int g = gslot(bank_slot_of(cb, K, q, j), par);
                const double S = __ldcg(a.w.bufP + o + j), A = __ldcg(a.w.bufA + o + j);
                const double cs = __ldcg(a.w.colsum + g), lc = __ldcg(a.w.lgnCP + g);
                int mm = m_s[si * K + j] - (j == zi);
                for (int e = 0; e < q; ++e)
                    if (cb->aux[e] == si) mm += (cb->b[e] == j) - (cb->a[e] == j);
                double p = dlog((double)mm + a.alpha);
                if (j != zi) {
                    p = p + S;
                    p = p - A;
                    p = p - cs;
                    p = p + lc;
                } else {
                    p = p + cs;
                    p = p - lc;
                    p = p - S;
                    p = p + A;
                }
                return p;
            };
            // A carried item's draw is a function of e = exp(p - max) and u alone.  If every stale column had e == 0 at the
            // item's last draw (none is among its recorded positive entries) and still has, the vector e is bit-identical
            // and the recorded outcome stands.  Tried while it keeps succeeding (a failed test only adds latency).
            if (!fresh && a.doSample && ncols <= 32 && (cheapScore > 0 || ((++cheapTick) & 15) == 0)) {
                const int4 nz = __ldcg(reinterpret_cast<const int4*>(a.w.drawNz) + slot);
                bool ok = nz.x != -2;
                if (ok) {
                    const double mxOld = __ldcg(a.w.drawMx + slot);
                    const int cj = (lane < ncols) ? CL_s[lane] : -1;
                    bool mine = true;
                    double p = 0.0;
                    if (cj >= 0) {
                        mine = !(nz.x == cj || nz.y == cj || nz.z == cj || nz.w == cj);
                        if (mine) {
                            p = combine(cj);
                            mine = dexp(p - mxOld) == 0.0;
                        }
                    }
                    ok = __all_sync(0xffffffffu, mine);
                    if (ok) {
                        if (a.probs && cj >= 0) a.probs[i * K + cj] = p;
                        if (lane == 0 && oldv != spec_clean(oldv)) a.w.spec[slot] = spec_clean(oldv);
                    }
                }
                cheapScore = ok ? min(8, cheapScore + 1) : max(0, cheapScore - 1);
                if (ok) continue;
            }
            if (lane == 0 && !fresh) atomicAdd(reinterpret_cast<unsigned long long*>(st_s + 4), 1ull);
            for (int j = lane; j < K; j += 32) {
                const double p = combine(j);
                pl[j] = p;
                if (a.probs) a.probs[i * K + j] = p;
            }
            __syncwarp();
            int znew = zi;
            if (a.doSample) {
                double mx;
                int4 nz;
                znew = warp_sample_ll(pl, K, a.u[t], rr, perm, lane, mx, nz);
                if (znew < 0) {
                    if (lane == 0) atomicCAS(a.w.status, 0, znew);
                    znew = zi;
                }
                if (lane == 0) {
                    a.w.drawMx[slot] = mx;
                    reinterpret_cast<int4*>(a.w.drawNz)[slot] = nz;
                }
            }
            if (lane == 0) {
                const int newv = (znew != zi) ? ((zi << 16) | znew) : -1;
                a.w.spec[slot] = (fresh || newv == spec_clean(oldv)) ? newv : spec_mark(newv, tag);
            }
            __syncwarp();
        }
        prev_end = end;
        have_spec = true;
        first = false;
        tagPrev = tag;
        par ^= 1;
        CELDA_TICK(7);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICK(8);
    }
    __syncthreads();
    if (threadIdx.x == 0 && st_s[4]) atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)st_s[4]);
    #ifdef CELDA_COMMIT_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int q = 0; q < 6; ++q) a.w.stats[8 + q] += csub[q];
#endif
// ---- write the tables back (CTA 0's copy; all copies are identical)
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < R * K; q += blockDim.x) a.nTab[q] = n_s[(q / R) * Rp + (q % R)];
        for (int q = threadIdx.x; q < K; q += blockDim.x) a.nCP[q] = nCP_s[q];
        for (int q = threadIdx.x; q < K * nS; q += blockDim.x) a.mCPByS[q] = m_s[q];
        if (threadIdx.x == 0) {
            a.w.stats[0] += st_s[0];
            a.w.stats[1] += st_s[1];
            a.w.stats[2] += st_s[2];
            atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)st_s[3] << 40);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.w.stats[3] += st_s[5];
        a.w.stats[4] += st_s[6];
        a.w.stats[5] += st_s[7];
        a.w.stats[6] += st_s[8];
    }
}

// vtCap: doubles of the buffer shared by the value-table head (EsCap entries) and the commit slots beyond column KB
inline size_t z_sweep_vtcap(int R, int K, int EsCap, int MC) {
    const int Rp = R | 1, KS = K + 2 * MC, KB = z_sweep_kb(K, MC);
    const size_t ext = ((size_t)(KS - KB) * Rp * sizeof(int) + 7) / 8;
    return std::max((size_t)EsCap, ext);
}
inline size_t z_sweep_smem(int R, int K, int nS, int T, int nwarp, int EsCap = 0, int MC = MCMAX - 1) {
    const int Rp = R | 1, KS = K + 2 * MC, KB = z_sweep_kb(K, MC);
    const size_t scratchD = (size_t)std::max(R, 2 * K + (K + 1) / 2) + 2;
    return sizeof(double) * (T + nwarp * scratchD + z_sweep_vtcap(R, K, EsCap, MC)) + sizeof(long long) * KS +
           2 * sizeof(CommitBank) + 12 * sizeof(long long) +
           sizeof(int) * ((size_t)KB * Rp + K * nS + (R + 1) + ZT_MAXGRP + 3 + 3 * (size_t)K + (MCMAX + 1) + (MCMAX + 2) + 144);
}

// xT[r][t] = counts[r, ix[t] - 1]: the count matrix in visiting order, transposed, so that lanes holding
// consecutive visiting positions read one row coalesced.  A block moves 32 cells x RC rows per step through shared
// memory: each warp has a cell's whole column slice in flight (RC / 32 independent 128-byte loads), the rows go out
// as 128-byte stores, and the per-row maxima are kept per block and merged with one atomic per row at the end.
__global__ void k_permute_transpose(const int* __restrict__ counts, int R, int RC, long long nM, const int* __restrict__ ix,
                                    int* __restrict__ xT, int* __restrict__ rowMax /* [R], zeroed */,
                                    const int* __restrict__ z, const int* __restrict__ nByC, int* __restrict__ zT,
                                    int* __restrict__ nT, const int* __restrict__ status) {
    extern __shared__ int pt_smem[];
    if (status[1]) return;  // bad visiting order (see chain_upload_order)
    const int RCp = RC | 1;
    int* tile = pt_smem;             // [32][RCp]
    int* bmax = pt_smem + 32 * RCp;  // [RC]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nM; t += (long long)gridDim.x * blockDim.x) {
        const int i = ix[t] - 1;
        zT[t] = z[i];
        nT[t] = nByC[i];
    }
    for (int rb = 0; rb < R; rb += RC) {
        const int rc = min(RC, R - rb);
        __syncthreads();
        for (int r = threadIdx.x; r < rc; r += blockDim.x) bmax[r] = 0;
        for (long long t0 = (long long)blockIdx.x * 32; t0 < nM; t0 += (long long)gridDim.x * 32) {
            __syncthreads();
            for (int c = w; c < 32; c += nw) {
                const long long t = t0 + c;
                if (t < nM) {
                    const int* src = counts + (size_t)(ix[t] - 1) * R + rb;
#pragma unroll 4
                    for (int r = lane; r < rc; r += 32) tile[c * RCp + r] = __ldg(src + r);
                } else {
                    for (int r = lane; r < rc; r += 32) tile[c * RCp + r] = 0;
                }
            }
            __syncthreads();
            for (int rr = w; rr < rc; rr += nw) {  // row rb + rr always falls to this warp: bmax needs no atomics
                int v = tile[lane * RCp + rr];
                if (t0 + lane < nM) xT[(size_t)(rb + rr) * nM + t0 + lane] = v;
                for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0 && v > bmax[rr]) bmax[rr] = v;
            }
        }
        __syncthreads();
        for (int r = threadIdx.x; r < rc; r += blockDim.x)
            if (bmax[r] > 0) atomicMax(&rowMax[rb + r], bmax[r]);
    }
}

constexpr int ZT_VMAX = 1 << 16;  // longest table of one row

// Table length per row: the largest count seen (+ margin), then rows are packed greedily into groups of at most
// EsCap entries.  Any length is CORRECT (a count beyond a row's table is evaluated directly); the choice only
// decides how often that happens.  One block.
__global__ void k_pick_rowV(const int* __restrict__ rowMax, int R, int EsCap, int* __restrict__ offg, int* __restrict__ grp,
                            int* __restrict__ tmpV) {
    for (int r = threadIdx.x; r < R; r += blockDim.x) tmpV[r] = min(min(ZT_VMAX, rowMax[r] + 1), EsCap);
    __syncthreads();
    if (threadIdx.x == 0) {
        int og = 0, ngrp = 0, inGrp = 0;
        grp[1] = 0;
        for (int r = 0; r < R; ++r) {
            int v = tmpV[r];
            if (inGrp + v > EsCap) {
                if (ngrp + 1 < ZT_MAXGRP) {  // close the group before this row
                    ++ngrp;
                    grp[1 + ngrp] = r;
                    inGrp = 0;
                } else {
                    v = max(0, EsCap - inGrp);  // out of groups: shorten the remaining rows' tables
                }
            }
            offg[r] = og;
            og += v;
            inGrp += v;
        }
        offg[R] = og;
        ++ngrp;
        grp[1 + ngrp] = R;
        grp[0] = ngrp;
        int complete = 1;
        for (int r = 0; r < R; ++r) complete &= (offg[r + 1] - offg[r] == rowMax[r] + 1);
        grp[ZT_MAXGRP + 2] = complete;
    }
}

// ------------------------------------------------------------------------------------------------ y sweep
// .cGCalcGibbsProbY (R/celda_G.R:602-660) + cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:187-249) on a dense
// gene-major int32 count table (nGByCP in celda_CG; ncol = K).
struct YArgs {
    int nG, ncol, L;
    const int* counts;      // [nG][ncol] gene-major
    const long long* nByG;  // [nG]
    int* y;                 // [nG] 1-based, updated
    int* tTab;              // L x ncol column-major (nTSByC argument), updated
    long long* nByTS;       // [L], updated
    int* nGByTS;            // [L] (includes the +1 pseudo-gene), updated
    double beta, gamma, delta;
    const int* ix;
    const double* u;
    int doSample;
    double* probs;  // L x nG or null
    // lgG[k] = lgamma(k + beta) for k < TG, resident in L2 (same bits as lg_tab): lane-per-unit rounds fetch the terms
    // of the columns selected by lgMask from it and evaluate the others, which keeps the FP64 pipe and the L2 busy
    // side by side.  Null / 0 disables.
    const double* lgG;
    int TG;
    unsigned lgMask;  // bit (c & 31): column c reads the table
    SweepWork w;
};

// lgG[k] = lgamma(k + beta) over k < n (built once per chain)
__global__ void k_lgG_fill(double beta, long long n, double* __restrict__ out) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        out[k] = dlgamma_hot((double)k + beta);
}

__global__ void __launch_bounds__(1024, 1) k_sweep_y(YArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.w.status[1]) return;  // the injected visiting order failed its check: nothing may be indexed through it
    const int L = a.L, C = a.ncol, T = a.w.lgT, MC = a.w.MC;
    const int LS = L + 2 * MC;  // row slots: L base rows + two per tentative commit (see "multi-commit rounds")
    const int Cp = C | 1;
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scratchD = max(2 * C + 8, 2 * L + (L + 1) / 2) + 2;
    double* lgtab = reinterpret_cast<double*>(smem_raw);
    double* lgt_s = lgtab + T;                           // [LS][Cp] lgamma(t + beta)
    double* tail_s = lgt_s + (size_t)LS * Cp;            // [LS][7] module-only terms of the second loop
    double* scratch_all = tail_s + (size_t)LS * 7;
    long long* T_s = reinterpret_cast<long long*>(scratch_all + (size_t)nwarp * scratchD);  // [LS]
    CommitBank* banks = reinterpret_cast<CommitBank*>(T_s + LS);  // [2] tentative commits of this round / the last
    long long* st_s = reinterpret_cast<long long*>(banks + 2);  // [12] statistics and CTA-0 clocks (thread 0)
    int* t_s = reinterpret_cast<int*>(st_s + 12);        // [LS][Cp]
    int* g_s = t_s + (size_t)LS * Cp;                    // [LS]
    int* fin_s = g_s + LS;                               // [L] slot of row l with every tentative commit applied
    int* CL_s = fin_s + L;                               // [L] stale-row list: Dcarry first, then the commits' rows in order
    int* dc_s = CL_s + L;                                // [L] Dcarry
    int* nq_s = dc_s + L;                                // [MCMAX + 1]
    int* seg_s = nq_s + MCMAX + 1;                       // [MCMAX + 2]
    int* s_misc = seg_s + MCMAX + 2;                     // [144]
    double* scr = scratch_all + (size_t)warp * scratchD;
    const int gwarp = blockIdx.x * nwarp + warp, tw = gridDim.x * nwarp;
    // warp-granular work goes round the CTAs first (item c -> CTA c mod grid), so a short list of items lands on
    // many SMs with one warp each instead of filling a few SMs
    const int gwarpT = warp * gridDim.x + blockIdx.x;

    for (int q = threadIdx.x; q < T; q += blockDim.x) lgtab[q] = dlgamma((double)q + a.beta);
    for (int q = threadIdx.x; q < L * C; q += blockDim.x) t_s[(q % L) * Cp + (q / L)] = a.tTab[q];
    for (int q = threadIdx.x; q < L; q += blockDim.x) {
        T_s[q] = a.nByTS[q];
        g_s[q] = a.nGByTS[q];
        fin_s[q] = q;
    }
    if (threadIdx.x <= MCMAX) nq_s[threadIdx.x] = 0;
    if (threadIdx.x <= MCMAX + 1) seg_s[threadIdx.x] = 0;
    CommitBank* cb = banks;
    __syncthreads();
    for (int q = threadIdx.x; q < L * C; q += blockDim.x) {
        const int l = q / C, c = q % C;
        lgt_s[l * Cp + c] = lg_tab2(lgtab, T, a.lgG, a.TG, t_s[l * Cp + c], a.beta);
    }
    // Module-only terms of cG_CalcGibbsProbY's second loop (:230-246), cached per row slot:
    // lgamma(g-1+gamma), (g+gamma), (g+1+gamma), ((g-1) delta), (g delta), ((g+1) delta), (T + g delta) -- the same
    // expressions the per-unit code evaluates, so the values are identical.  lg_delta[0] is NA in the reference
    // (R/celda_CG.R:429): dlgdelta() returns NaN there and the draw reports "NA in probability vector".
    auto tail_fill = [&](int l, int k) {
        const int g = g_s[l];
        double v;
        switch (k) {
            case 0: v = dlgamma_hot((double)(g - 1) + a.gamma); break;
            case 1: v = dlgamma_hot((double)g + a.gamma); break;
            case 2: v = dlgamma_hot((double)(g + 1) + a.gamma); break;
            case 3: v = dlgdelta(g - 1, a.delta); break;
            case 4: v = dlgdelta(g, a.delta); break;
            case 5: v = dlgdelta(g + 1, a.delta); break;
            default: v = dlgamma_hot((double)T_s[l] + (double)g * a.delta); break;
        }
        tail_s[l * 7 + k] = v;
    };
    for (int q = threadIdx.x; q < L * 7; q += blockDim.x) tail_fill(q / 7, q % 7);
    __syncthreads();

    long long pos = 0, hi = 0, prev_end = 0;
    int W = a.w.Wfull, gavg = a.w.Wfull, M = 0, mcAimd = MC, par = 0, tagPrev = 0;
    bool have_spec = false, first = true;
    unsigned bar_target = 0;
    int roundNo = 0;
    int cheapScore = 4, cheapTick = 0;
    if (threadIdx.x < 12) st_s[threadIdx.x] = (threadIdx.x == 9) ? clock64() : 0;  // 0 rounds 1 units 2 movers 3 discards 4 redraws 5..8 clocks
    __syncthreads();

    auto window_for = [&](int g, int mc) -> int {
        const long long target = (long long)g * (mc + 4);
        if (target >= a.w.Wunit) return (int)min((target / a.w.Wunit) * a.w.Wunit, (long long)a.w.Wfull);
        return (int)max((long long)a.w.Wmin, target);
    };

    long long csub[6] = {0, 0, 0, 0, 0, 0}, csub0 = 0;
    (void)csub; (void)csub0;
    while (true) {
        // ---- settle the previous round: which tentative commits stand, where the window restarts, the next commits
        CELDA_CSUB_BEGIN();
        CommitBank* ob = cb;              // the finished round's commits
        cb = banks + par;                 // this round's
        bool hasDc = false;
        if (have_spec) {
            const int nwinPrev = (int)(prev_end - pos);
            bool anyMover;
            const int f = block_first_invalid(a.w.spec, pos, a.w.Wmax, nwinPrev, tagPrev, s_misc, anyMover);
            CELDA_CSUB(0);
            const long long tf = pos + f;
            const int Mold = M;
            int qf = Mold;
            if (f >= 0) {
                qf = 0;
                for (int i = 0; i < Mold; ++i) qf += (ob->t[i] < tf);
            }
            hasDc = qf < Mold;
            const long long oldpos = pos;
            if (qf < Mold) pos = tf;  // no unselected mover can lie in front of a discarded selected one
            else if (Mold > 0) pos = ob->t[Mold - 1] + 1;
            if (threadIdx.x == 0) {
                st_s[2] += qf;
                st_s[3] += Mold - qf;
            }
            if (Mold > 0) mcAimd = (qf == Mold) ? min(MC, 2 * mcAimd) : max(1, qf + 1);
            if (qf > 0) gavg = max(1, (int)((3LL * gavg + max(1LL, (pos - oldpos) / qf)) / 4));
            const int mcCur = min(mcAimd, mc_target(gavg, MC));
            W = window_for(gavg, mcCur);
            hi = min(prev_end, pos + W);
            const int nsel = (int)(hi - pos);
            const bool wideSel = nsel > 1024;  // a large window is scanned by the whole block, a small one by warp 0
            int Mwide = 0;
            if (wideSel && anyMover) Mwide = block_select_movers(a.w.spec, pos, a.w.Wmax, nsel, mcCur, cb->t, cb->pk, s_misc + 2);
            if (warp == 0) {
                // one warp: Dcarry, the next commits and their records, source slots, stale list, segments
                const int ndc = warp_build_dcarry(ob, qf, Mold, dc_s, lane);
                const int Mn = wideSel ? Mwide : (anyMover ? warp_select_movers(a.w.spec, pos, a.w.Wmax, nsel, mcCur, cb, lane) : 0);
                __syncwarp();
                if (lane < Mn) {
                    const int gene = a.ix[cb->t[lane]] - 1, pk = cb->pk[lane];
                    cb->item[lane] = gene;
                    cb->a[lane] = pk >> 16;
                    cb->b[lane] = pk & 0xffff;
                    cb->N[lane] = a.nByG[gene];
                }
                __syncwarp();
                warp_build_lists(cb, L, Mn, ndc, dc_s, CL_s, nq_s, seg_s, pos, hi, lane);
                if (lane == 0) s_misc[140] = Mn;
            } else {
                // the other warps: standing commits advance the base rows (in order: a later commit's slot supersedes
                // an earlier one's) and the per-row scalars
                for (int c = threadIdx.x - 32; c < C + 7; c += blockDim.x - 32)
                    for (int i = 0; i < qf; ++i)
                        for (int side = 0; side < 2; ++side) {
                            const int dst = side ? ob->b[i] : ob->a[i], src = L + 2 * i + side;
                            if (c < C) {
                                t_s[dst * Cp + c] = t_s[src * Cp + c];
                                lgt_s[dst * Cp + c] = lgt_s[src * Cp + c];
                            } else {
                                tail_s[dst * 7 + (c - C)] = tail_s[src * 7 + (c - C)];
                            }
                        }
                for (int j = (int)blockDim.x - 1 - (int)threadIdx.x; j < L; j += blockDim.x - 32) {
                    const int sl = bank_slot_of(ob, L, qf, j);
                    if (sl != j) {
                        g_s[j] = g_s[sl];
                        T_s[j] = T_s[sl];
                    }
                }
                if (warp == 1 && lane < qf && blockIdx.x == 0) a.y[ob->item[lane]] = ob->b[lane] + 1;
            }
            __syncthreads();
            M = s_misc[140];
            CELDA_CSUB(1);
            if (M == 0 && !hasDc) {  // nothing moves and nothing is stale: the carried items are settled
                gavg = (int)min(2LL * gavg + (hi - pos), (long long)(1 << 24));
                pos = hi;
                W = window_for(gavg, min(mcAimd, mc_target(gavg, MC)));
            }
        }
        if (pos >= a.nG) break;
        CELDA_CSUB(2);
        // ---- this round's tentative commits: new row slots
        CELDA_CSUB(3);
        {
            // the movers' count rows, all loads in flight at once (staged in the warp scratch, idle here)
            int* xs = reinterpret_cast<int*>(scratch_all);
            for (int q = threadIdx.x; q < M * C; q += blockDim.x) {
                const int i = q / C, c = q - i * C;
                xs[q] = __ldg(a.counts + (size_t)cb->item[i] * C + c);
            }
            __syncthreads();
            for (int c = threadIdx.x; c < C; c += blockDim.x)
                for (int i = 0; i < M; ++i) {
                    const int xv = xs[i * C + c];
                    t_s[(L + 2 * i) * Cp + c] = t_s[cb->srcA[i] * Cp + c] - xv;
                    t_s[(L + 2 * i + 1) * Cp + c] = t_s[cb->srcB[i] * Cp + c] + xv;
                }
        }
        if (threadIdx.x == blockDim.x - 1)
            for (int i = 0; i < M; ++i) {
                g_s[L + 2 * i] = g_s[cb->srcA[i]] - 1;
                g_s[L + 2 * i + 1] = g_s[cb->srcB[i]] + 1;
                T_s[L + 2 * i] = T_s[cb->srcA[i]] - cb->N[i];
                T_s[L + 2 * i + 1] = T_s[cb->srcB[i]] + cb->N[i];
            }
        for (int j = threadIdx.x; j < L; j += blockDim.x) fin_s[j] = bank_slot_of(cb, L, M, j);
        __syncthreads();
        for (int q = threadIdx.x; q < 2 * M * (C + 7); q += blockDim.x) {
            const int sl = L + q / (C + 7), c = q % (C + 7);
            if (c < C)
                lgt_s[sl * Cp + c] = lg_tab2(lgtab, T, a.lgG, a.TG, t_s[sl * Cp + c], a.beta);
            else
                tail_fill(sl, c - C);
        }
        __syncthreads();
        CELDA_CSUB(4);
        CELDA_TICK(5);
        // after commits the window is topped up with fresh items only once it is less than half full: carried items
        // cost a few patch units and a check, fresh ones a full evaluation and draw, cheaper in batches
        const bool topup = first || (M == 0 && !hasDc) || (hi - pos) * 2 < W;
        const long long end = topup ? max(hi, min(pos + (long long)W, (long long)a.nG)) : hi;
        const int nwin = (int)(end - pos);
        const long long nPatch = seg_s[M + 1];
        const long long nFresh = (end - hi) * L;
        const long long total = nPatch + nFresh;
        const int tag = roundNo & 7;
        ++roundNo;
        // unit index -> (position t, module l, row slot)
        auto decode = [&](long long uidx, long long& t, int& l, int& slot) {
            if (uidx < nPatch) {
                int rel = (int)uidx, q = 0;
                while (rel >= seg_s[q + 1]) ++q;
                rel -= seg_s[q];
                const int wq = nq_s[q];
                t = ((q == 0) ? pos : cb->t[q - 1] + 1) + rel / wq;
                l = CL_s[rel % wq];
                slot = bank_slot_of(cb, L, q, l);
            } else {
                const long long q = uidx - nPatch;
                t = hi + q / L;
                l = (int)(q % L);
                slot = fin_s[l];
            }
        };
        if (total >= (long long)tw * 16) {
            // one lane per (gene, module) unit: the whole cG_CalcGibbsProbY term sequence in one thread
            for (long long uidx = (long long)gwarp * 32 + lane; uidx < total; uidx += (long long)tw * 32) {
                long long t;
                int l, slot;
                decode(uidx, t, l, slot);
                const int i = a.ix[t] - 1;
                const int cur = a.y[i] - 1;
                const bool isCur = (l == cur);
                const int* x = a.counts + (size_t)i * C;
                const int* trow = t_s + slot * Cp;
                const double* lrow = lgt_s + slot * Cp;
                double p = 0.0;
                const int sgn = isCur ? -1 : 1;
#pragma unroll 2
                for (int c = 0; c < C; ++c) {
                    const int xc = __ldg(x + c);  // same gene on (almost) every lane of the warp: a uniform branch
                    const double cached = lrow[c];
                    // a zero count leaves the argument at t + beta, whose lgamma is the cached entry (same bits)
                    const long long kk = (long long)trow[c] + sgn * xc;
                    double fresh;
                    if (xc == 0)
                        fresh = cached;
                    else if (kk < T)
                        fresh = lgtab[kk];
                    else if (((a.lgMask >> (c & 31)) & 1u) && kk < a.TG)
                        fresh = __ldg(a.lgG + kk);  // one 32-byte sector per lane: bound by L2 sector throughput
                    else
                        fresh = dlgamma_hot((double)kk + a.beta);
                    p += isCur ? cached : fresh;
                    p -= isCur ? fresh : cached;
                }
                const int g = g_s[slot];
                const double Tl = (double)T_s[slot], Ni = (double)a.nByG[i];
                {
                    // second loop of cG_CalcGibbsProbY (:230-246); (g0, g1) = (g, g-1) for the current module else
                    // (g+1, g): five of the six terms depend on the module only and come from the cache
                    const double* tl = tail_s + slot * 7;
                    const int g1 = isCur ? g - 1 : g, g0 = g1 + 1;
                    p += tl[isCur ? 1 : 2];
                    p -= tl[isCur ? 0 : 1];
                    p += tl[isCur ? 4 : 5];
                    p -= tl[isCur ? 3 : 4];
                    p += isCur ? dlgamma_hot(Tl - Ni + (double)g1 * a.delta) : tl[6];
                    p -= isCur ? tl[6] : dlgamma_hot(Tl + Ni + (double)g0 * a.delta);
                }
                a.w.bufP[ring_slot(t, a.w.Wmax) * L + l] = p;
            }
        } else {
            for (long long uidx = gwarpT; uidx < total; uidx += tw) {
                long long t;
                int l, slot;
                decode(uidx, t, l, slot);
                const int i = a.ix[t] - 1;
                const int cur = a.y[i] - 1;
                const bool isCur = (l == cur);
                const int* x = a.counts + (size_t)i * C;
                const int* trow = t_s + slot * Cp;
                const double* lrow = lgt_s + slot * Cp;
                // first loop of cG_CalcGibbsProbY (:212-225): per column one fresh lgamma and one cached one
                for (int c = lane; c < C; c += 32) {
                    const double cached = lrow[c];
                    const double fresh =
                        (x[c] == 0) ? cached : lg_tab2(lgtab, T, a.lgG, a.TG, (long long)trow[c] + (isCur ? -x[c] : x[c]), a.beta);
                    scr[2 * c] = isCur ? cached : fresh;      // added
                    scr[2 * c + 1] = isCur ? fresh : cached;  // subtracted
                }
                // second loop (:230-246): five module-only terms from the cache, the sixth evaluated
                if (lane < 6) {
                    const int g = g_s[slot];
                    const double Tl = (double)T_s[slot], Ni = (double)a.nByG[i];
                    const double* tl = tail_s + slot * 7;
                    const int g1 = isCur ? g - 1 : g, g0 = g1 + 1;
                    double v;
                    switch (lane) {
                        case 0: v = tl[isCur ? 1 : 2]; break;
                        case 1: v = tl[isCur ? 0 : 1]; break;
                        case 2: v = tl[isCur ? 4 : 5]; break;
                        case 3: v = tl[isCur ? 3 : 4]; break;
                        case 4: v = isCur ? dlgamma_hot(Tl - Ni + (double)g1 * a.delta) : tl[6]; break;
                        default: v = isCur ? tl[6] : dlgamma_hot(Tl + Ni + (double)g0 * a.delta); break;
                    }
                    scr[2 * C + lane] = v;
                }
                __syncwarp();
                if (lane == 0) {
                    double p = 0.0;
#pragma unroll 8
                    for (int c = 0; c < C; ++c) {
                        p += scr[2 * c];
                        p -= scr[2 * c + 1];
                    }
                    p += scr[2 * C + 0];
                    p -= scr[2 * C + 1];
                    p += scr[2 * C + 2];
                    p -= scr[2 * C + 3];
                    p += scr[2 * C + 4];
                    p -= scr[2 * C + 5];
                    a.w.bufP[ring_slot(t, a.w.Wmax) * L + l] = p;
                }
                __syncwarp();
            }
        }
        if (threadIdx.x == 0) {
            st_s[0] += 1;
            st_s[1] += total;
        }
        CELDA_TICK(6);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICK(8);
        double* pl = scr;
        double* rr = scr + L;
        int* perm = reinterpret_cast<int*>(scr + 2 * L);
        // carried items in front of the first tentative commit have nothing stale unless commits were discarded
        const int cstart = (hasDc || M == 0) ? 0 : (int)(cb->t[0] + 1 - pos);
        for (int c = cstart + gwarpT; c < nwin; c += tw) {
            const long long t = pos + c;
            const bool fresh = t >= hi;
            int q = M;
            if (!fresh) {
                q = 0;
                for (int i = 0; i < M; ++i) q += (cb->t[i] < t);
            }
            const int ncols = fresh ? L : nq_s[q];
            if (ncols == 0) continue;
            const int i = a.ix[t] - 1;
            const int cur = a.y[i] - 1;
            const size_t slot = ring_slot(t, a.w.Wmax), o = slot * L;
            const int oldv = fresh ? -1 : __ldcg(a.w.spec + slot);
            // see k_sweep_z: the recorded draw stands if every stale entry had exp(p - max) == 0 and still has
            if (!fresh && a.doSample && ncols <= 32 && (cheapScore > 0 || ((++cheapTick) & 15) == 0)) {
                const int4 nz = __ldcg(reinterpret_cast<const int4*>(a.w.drawNz) + slot);
                bool ok = nz.x != -2;
                if (ok) {
                    const double mxOld = __ldcg(a.w.drawMx + slot);
                    const int cj = (lane < ncols) ? CL_s[lane] : -1;
                    bool mine = true;
                    double p = 0.0;
                    if (cj >= 0) {
                        mine = !(nz.x == cj || nz.y == cj || nz.z == cj || nz.w == cj);
                        if (mine) {
                            p = __ldcg(a.w.bufP + o + cj);
                            mine = dexp(p - mxOld) == 0.0;
                        }
                    }
                    ok = __all_sync(0xffffffffu, mine);
                    if (ok) {
                        if (a.probs && cj >= 0) a.probs[(size_t)i * L + cj] = p;
                        if (lane == 0 && oldv != spec_clean(oldv)) a.w.spec[slot] = spec_clean(oldv);
                    }
                }
                cheapScore = ok ? min(8, cheapScore + 1) : max(0, cheapScore - 1);
                if (ok) continue;
            }
            if (lane == 0 && !fresh) atomicAdd(reinterpret_cast<unsigned long long*>(st_s + 4), 1ull);
            for (int l = lane; l < L; l += 32) {
                const double p = __ldcg(a.w.bufP + o + l);
                pl[l] = p;
                if (a.probs) a.probs[(size_t)i * L + l] = p;
            }
            __syncwarp();
            int ynew = cur;
            if (a.doSample) {
                double mx;
                int4 nz;
                ynew = warp_sample_ll(pl, L, a.u[t], rr, perm, lane, mx, nz);
                if (ynew < 0) {
                    if (lane == 0) atomicCAS(a.w.status, 0, ynew);
                    ynew = cur;
                }
                if (lane == 0) {
                    a.w.drawMx[slot] = mx;
                    reinterpret_cast<int4*>(a.w.drawNz)[slot] = nz;
                }
            }
            if (lane == 0) {
                const int newv = (ynew != cur) ? ((cur << 16) | ynew) : -1;
                a.w.spec[slot] = (fresh || newv == spec_clean(oldv)) ? newv : spec_mark(newv, tag);
            }
            __syncwarp();
        }
        prev_end = end;
        have_spec = true;
        first = false;
        tagPrev = tag;
        par ^= 1;
        CELDA_TICK(7);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICK(8);
    }
    __syncthreads();
    if (threadIdx.x == 0 && st_s[4]) atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)st_s[4]);
#ifdef CELDA_COMMIT_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int q = 0; q < 6; ++q) a.w.stats[8 + q] += csub[q];
#endif
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < L * C; q += blockDim.x) a.tTab[q] = t_s[(q % L) * Cp + (q / L)];
        for (int q = threadIdx.x; q < L; q += blockDim.x) {
            a.nByTS[q] = T_s[q];
            a.nGByTS[q] = g_s[q];
        }
        if (threadIdx.x == 0) {
            a.w.stats[0] += st_s[0];
            a.w.stats[1] += st_s[1];
            a.w.stats[2] += st_s[2];
            atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)st_s[3] << 40);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.w.stats[3] += st_s[5];
        a.w.stats[4] += st_s[6];
        a.w.stats[5] += st_s[7];
        a.w.stats[6] += st_s[8];
    }
}

inline size_t y_sweep_smem(int L, int C, int T, int nwarp, int MC = MCMAX - 1) {
    const int Cp = C | 1, LS = L + 2 * MC;
    const size_t scratchD = (size_t)std::max(2 * C + 8, 2 * L + (L + 1) / 2) + 2;
    return sizeof(double) * (T + (size_t)LS * Cp + (size_t)LS * 7 + nwarp * scratchD) + sizeof(long long) * LS +
           2 * sizeof(CommitBank) + 12 * sizeof(long long) + sizeof(int) * ((size_t)LS * Cp + LS + 3 * (size_t)L + (MCMAX + 1) + (MCMAX + 2) + 144);
}

// ------------------------------------------------------------------------------------------------ y sweep, celda_G
// .cGCalcGibbsProbY on the RAW counts (R/celda_G.R:401-415, 602-660): every gene's row is scanned over ALL cells
// against nTSByC (L x nM), zero counts included -- cG_CalcGibbsProbY adds and subtracts the same table entry for
// them (src/cG_calcGibbsProbY.cpp:218-222), which is not a floating-point no-op, so bit parity needs the full
// column scan: 2 dependent FP64 adds per (gene, module, cell).  Layout:
//   * nTSByC (t) and its cached lgamma (lgt) stay in HBM / L2.  A CTA task = (one gene per warp) x (a group of up to
//     8 x 32 modules): lane l of the warp carries one accumulation chain per 32-module chunk (modules l, l + 32, ...),
//     i.e. up to 8 independent chains per lane that hide each other's FP64 latency, and every chain of the warp
//     belongs to the SAME gene, so the columns where the gene has a count are a warp-uniform event (a cursor over
//     the gene's gene-major twin row: a 32-entry register window, the next window already in flight).
//   * Columns are contiguous in memory (L entries each), so the (t, lgt) tile of TC columns is ONE contiguous slab
//     per array: two TMA bulk copies (cp.async.bulk) per tile into a ring of YG_NBUF buffers, each completing on its
//     own mbarrier; the warp that is last to finish a buffer refills it -- no CTA-wide barrier inside a task.
//   * Measured bound: shared-memory bandwidth (one 8-byte read per 2 FP64 adds), see DESIGN.md.
// Patch units after a commit (da, db): the same code with one chain, lanes 0 and 1 bound to the two modules and a
// tile of just those two rows.
struct YGArgs {
    int nG, L;
    long long nM;
    const long long* rp;   // gene-major twin: row pointers [nG + 1]
    const int* tcol;       //   column of each stored entry, ascending within a gene
    const int* tx;         //   value
    const long long* nByG;
    int* y;                // [nG] 1-based, updated
    int* t;                // nTSByC, L x nM column-major, updated
    double* lgt;           // lgamma(t + beta), same layout, maintained
    long long* nByTS;      // [L], updated
    int* nGByTS;           // [L], updated
    double beta, gamma, delta;
    const int* ix;
    const double* u;
    int doSample;
    double* probs;         // L x nG or null
    int noTma;             // measurement aid (CELDA_CUDA_YG_NOTMA): stage the tiles with plain loads
    int flags;             // policy switches (CELDA_CUDA_YG_FLAGS): 1 wave-sized top-ups, 2 wide patch tiles, 16 poll the tile
                           // barrier with test_wait; 32 = scan mode 1 (set by the library)
    int TC;                // columns per tile (a multiple of 4; set by the launcher from the shared-memory budget)
    SweepWork w;
};

// the table-miss path of the column loop as a call: inlined copies would not fit the instruction cache
__device__ __noinline__ double dlgamma_hot_call(double x) { return dlgamma_hot(x); }

constexpr int YG_THREADS = 512;   // 16 genes per task; 128 registers per thread for up to 8 chains with their operands in flight
constexpr int YG_NBUF = 3;        // tile buffers in flight (two buffers with 1.5x wider tiles measure the same)
constexpr int YG_MAXMC = 7;       // 32-module chunks per task (chains per lane); 8 would spill at 128 registers

__host__ __device__ inline int yg_group_modules(int L) { return L < 32 * YG_MAXMC ? L : 32 * YG_MAXMC; }
// bytes of the tile ring: TC columns x (modules of a group) x (8 + 4) bytes per buffer
__host__ __device__ inline size_t yg_unit_bytes(int L, int TC) {
    // + 1024: idle chains (the chunk count is rounded up to 1, 2, 4 or 7) read up to 95 entries past the last row
    return (size_t)YG_NBUF * TC * yg_group_modules(L) * (sizeof(double) + sizeof(int)) + 1024;
}
__host__ __device__ inline size_t yg_fixed_bytes(int L, int T) {  // tables that live for the whole kernel, padded to 128
    const size_t b = sizeof(double) * T + sizeof(long long) * L + sizeof(unsigned long long) * YG_NBUF +
                     sizeof(int) * (L + YG_NBUF + 8 + 65);
    return (b + 127) & ~(size_t)127;
}

__global__ void __launch_bounds__(YG_THREADS, 1) k_sweep_yg(YGArgs a) {
    extern __shared__ __align__(128) unsigned char yg_smem[];
    if (a.w.status[1]) return;  // the injected visiting order failed its check: nothing may be indexed through it
    const int L = a.L, T = a.w.lgT, TC = a.TC;
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scratchD = 2 * L + (L + 1) / 2 + 2;
    const unsigned FULL = 0xffffffffu;
    double* lgtab = reinterpret_cast<double*>(yg_smem);
    long long* T_s = reinterpret_cast<long long*>(lgtab + T);
    unsigned long long* fullb = reinterpret_cast<unsigned long long*>(T_s + L);  // [YG_NBUF] tile landed
    int* g_s = reinterpret_cast<int*>(fullb + YG_NBUF);
    int* cnt_s = g_s + L;            // [YG_NBUF] warps done with the buffer
    int* s_misc = cnt_s + YG_NBUF;   // [8]
    int* hist_s = s_misc + 8;        // [65] density buckets of the per-round gene order
    unsigned char* region = yg_smem + yg_fixed_bytes(L, T);
    // draw phase view of the region
    double* scr = reinterpret_cast<double*>(region) + (size_t)warp * scratchD;
    // unit phase view: the tile ring, per buffer TC x GM doubles then (after all buffers) TC x GM ints
    const int GM = yg_group_modules(L);                         // modules per task group (= L unless L > 256)
    double* tile = reinterpret_cast<double*>(region);           // [YG_NBUF][TC][GM] lgamma(t + beta)
    int* ttile = reinterpret_cast<int*>(tile + (size_t)YG_NBUF * TC * GM);  // [YG_NBUF][TC][GM] t
    const int tw = gridDim.x * nwarp;
    // warp-granular work goes round the CTAs first (item c -> CTA c mod grid), so a short list of items lands on
    // many SMs with one warp each instead of filling a few SMs
    const int gwarpT = warp * gridDim.x + blockIdx.x;
    const long long nM = a.nM;
    const int ngroup = (L + GM - 1) / GM;
    const long long ntile = (nM + TC - 1) / TC;
    const int TCp = (a.flags & 2) ? max(4, min(1024, (TC * GM / 2) & ~3)) : TC;  // columns per patch tile (two rows) in the same buffers

    for (int q = threadIdx.x; q < T; q += blockDim.x) lgtab[q] = dlgamma((double)q + a.beta);
    for (int q = threadIdx.x; q < L; q += blockDim.x) {
        T_s[q] = a.nByTS[q];
        g_s[q] = a.nGByTS[q];
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < YG_NBUF; ++b) {
            mbar_init(fullb + b, 1);
            cnt_s[b] = 0;
        }
        mbar_init_fence();
    }
    __syncthreads();

    long long pos = 0, hi = 0, prev_end = 0;
    int W = a.w.Wfull, nd = 0, da = -1, db = -1, gavg = a.w.Wfull;
    bool have_spec = false;
    unsigned bar_target = 0;
    unsigned phaseBits = 0;  // bit b: parity of the next completion of fullb[b]
    long long rounds = 0, units_done = 0, movers = 0;
    long long redraws = 0;  // carried items drawn again in full (this thread's share)
    long long ck_commit = 0, ck_units = 0, ck_draw = 0, ck_bar = 0, ck0 = clock64(), ck1;

    while (true) {
        bool committed = false;
        if (have_spec) {
            const int nwin = (int)(prev_end - pos);
            const int first = block_first_mover(a.w.spec, pos, a.w.Wmax, nwin, s_misc);
            if (first < 0) {
                pos = prev_end;
                hi = pos;
                nd = 0;
                gavg = min(2 * gavg + nwin, 1 << 24);
                W = min(2 * W, a.w.Wfull);
            } else {
                const long long tt = pos + first;
                const int i = a.ix[tt] - 1;
                const int packed = __ldcg(a.w.spec + ring_slot(pos + first, a.w.Wmax));
                da = packed >> 16;
                db = packed & 0xffff;
                // the two table rows change only where the gene has counts: the grid shares the gene's entries
                const long long e0 = a.rp[i], e1 = a.rp[i + 1];
                for (long long e = e0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < e1;
                     e += (long long)gridDim.x * blockDim.x) {
                    const size_t base = (size_t)a.tcol[e] * L;
                    const int xv = a.tx[e];
                    const int na = __ldcg(a.t + base + da) - xv, nb = __ldcg(a.t + base + db) + xv;
                    a.t[base + da] = na;
                    a.t[base + db] = nb;
                    a.lgt[base + da] = lg_tab(lgtab, T, na, a.beta);
                    a.lgt[base + db] = lg_tab(lgtab, T, nb, a.beta);
                }
                if (threadIdx.x == 0) {
                    const long long Ni = a.nByG[i];
                    g_s[da] -= 1;
                    g_s[db] += 1;
                    T_s[da] -= Ni;
                    T_s[db] += Ni;
                    if (blockIdx.x == 0) a.y[i] = db + 1;
                }
                ++movers;
                gavg = max(1, (3 * gavg + first + 1) / 4);
                W = (int)max((long long)a.w.Wmin, min(4LL * gavg, (long long)a.w.Wfull));
                W = ((W + 31) / 32) * 32;
                pos = tt + 1;
                hi = min(prev_end, pos + W);
                nd = 2;
                committed = true;
            }
        }
        if (pos >= a.nG) break;
        // after a commit the window is topped up with fresh items only once it is less than half full: carried items
        // cost two patch units and a check (draw_unchanged), fresh ones a full evaluation and draw, cheaper in batches
        // ... and in batches of at least one grid-wide wave of tasks: a fresh task lasts several patch tasks, a round with
        // a handful of them leaves the grid idle
        const bool topup = nd != 2 || (hi - pos) * 2 < W;
        const long long end = topup ? max(hi, min(max(pos + (long long)W, hi + ((a.flags & 1) ? (long long)a.w.Wunit : 0LL)), (long long)a.nG)) : hi;
        const int nwin = (int)(end - pos);
        // A gene's cost grows with its number of stored entries (a column with a count costs several times a column
        // without), and the warps of a task advance tile by tile together: CTA 0 orders the carried and the fresh genes
        // by density (64 buckets, densest first), so that a task's warps carry similar genes and the longest tasks
        // start first.  Which warp evaluates a gene does not change the arithmetic.
        int* permP = reinterpret_cast<int*>(a.w.bufA);  // [Wmax] slot -> carried gene (offset from pos)
        int* permF = permP + a.w.Wmax;                  // [Wmax] slot -> fresh gene (offset from hi)
        if (blockIdx.x == 0) {
            for (int part = 0; part < 2; ++part) {
                const long long lo = part ? hi : pos;
                const int n = (int)(part ? end - hi : ((nd == 2) ? hi - pos : 0));
                int* perm = part ? permF : permP;
                if (n == 0) continue;
                __syncthreads();
                for (int q = threadIdx.x; q < 65; q += blockDim.x) hist_s[q] = 0;
                __syncthreads();
                for (int q = threadIdx.x; q < n; q += blockDim.x) {
                    const int g = a.ix[lo + q] - 1;
                    const long long nz = a.rp[g + 1] - a.rp[g];
                    atomicAdd(&hist_s[63 - (int)min(63LL, nz * 64 / (nM + 1))], 1);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int run = 0;
                    for (int q = 0; q < 64; ++q) {
                        const int c = hist_s[q];
                        hist_s[q] = run;
                        run += c;
                    }
                }
                __syncthreads();
                for (int q = threadIdx.x; q < n; q += blockDim.x) {
                    const int g = a.ix[lo + q] - 1;
                    const long long nz = a.rp[g + 1] - a.rp[g];
                    perm[atomicAdd(&hist_s[63 - (int)min(63LL, nz * 64 / (nM + 1))], 1)] = q;
                }
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) a.w.barrier[1] = 0;  // this round's task counter
        grid_sync(a.w.barrier, bar_target);  // the order, and after a commit table rows a, b, are visible before they are read
        CELDA_TICKR(ck_commit);
        // ---- tasks: patch = (nwarp carried genes) x {da, db}; fresh = (nwarp new genes) x (module group); one gene per
        // warp, warp slot s of a part <-> gene s of the part's density order
        const long long nP = (nd == 2) ? hi - pos : 0, nF = end - hi;
        // (a second gene per patch warp, to save the second wave of patch tasks at c4, measured 17 % slower)
        const int ngP = 1;
        const long long slotsP = (nP + ngP - 1) / ngP, slotsF = nF;
        const long long nPatchGrp = (slotsP + nwarp - 1) / nwarp;
        const long long nFreshGrp = (slotsF + nwarp - 1) / nwarp;
        const long long nFreshTasks = nFreshGrp * ngroup, ntask = nFreshTasks + nPatchGrp;
        ++rounds;
        // One task with MC chains per lane and gene, NG genes per warp.  (Two genes per warp in FRESH tasks -- a dense one
        // paired with a sparse one -- were measured and are slower: 128 registers do not hold 14 chains and their
        // operands)
        auto run_task = [&](auto mc_tag, auto ng_tag, long long task) {
            constexpr int MC = decltype(mc_tag)::value;
            constexpr int NG = decltype(ng_tag)::value;
            static_assert(NG == 1, "the column loop is written for one gene per warp");
            const bool patch = task >= nFreshTasks;  // fresh tasks (up to 7 chains) are handed out before patch tasks (one chain)
            // a patch tile holds two entries per column, so it spans many more columns in the same buffer: the chain of a
            // patch unit is short per column and a tile has to outlast the latency of its refill
            const int TCt = patch ? TCp : TC;
            const long long ntileT = (nM + TCt - 1) / TCt;
            int l0 = 0, gm = 2, stride = 2;   // first module, modules and row stride of the task's tile
            long long slot, nPart, nSlots, lo;
            const int* perm;
            if (patch) {
                slot = (task - nFreshTasks) * nwarp + warp;
                nPart = nP;
                nSlots = slotsP;
                lo = pos;
                perm = permP;
            } else {
                const long long q = task;
                slot = (q / ngroup) * nwarp + warp;
                nPart = nF;
                nSlots = slotsF;
                lo = hi;
                perm = permF;
                l0 = (int)(q % ngroup) * GM;
                gm = min(GM, L - l0);
                stride = gm;
            }
            // this warp's genes: order positions slot, slot + nSlots, ...
            long long tg[NG];
            int gi[NG], cur[NG];
            unsigned validMask = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const long long o = slot + g * nSlots;
                const bool v = slot < nSlots && o < nPart;
                tg[g] = lo + (v ? __ldcg(perm + o) : 0);
                gi[g] = v ? a.ix[tg[g]] - 1 : 0;
                cur[g] = v ? a.y[gi[g]] - 1 : -1;
                if (v) validMask |= 1u << g;
            }
            // chain m of this lane <-> module (patch: lanes 0 and 1 <-> da, db), idle if outside the group
            unsigned actMask = 0, curMask[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) curMask[g] = 0;
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int l = patch ? (lane == 0 ? da : (lane == 1 ? db : -1)) : ((m * 32 + lane < gm) ? l0 + m * 32 + lane : -1);
                if (l >= 0) actMask |= 1u << m;
#pragma unroll
                for (int g = 0; g < NG; ++g)
                    if (l >= 0 && l == cur[g]) curMask[g] |= 1u << m;
            }
            const int laneOff = patch ? min(lane, 1) : lane;  // this lane's entry within a tile row (+ 32 m)
            // where the gene's current module sits among this task's chains: chunk mstar (warp-uniform), lane lstar
            int mstar = -1, lstar = -1;
            if (patch) {
                if (cur[0] == da || cur[0] == db) {
                    mstar = 0;
                    lstar = (cur[0] == da) ? 0 : 1;
                }
            } else if (cur[0] >= l0 && cur[0] < l0 + gm) {
                mstar = (cur[0] - l0) >> 5;
                lstar = (cur[0] - l0) & 31;
            }
            const unsigned lgA = opaque_u32(smem_addr(lgtab));
            const unsigned tileA = opaque_u32(smem_addr(tile + laneOff)), ttileA = opaque_u32(smem_addr(ttile + laneOff));
            const bool useTma = !patch && !a.noTma && gm == L;
            // cursors over the genes' twin rows: lane k holds entry wbase + k, the following window already in flight
            const int NONE = 0x7fffffff;
            long long rb[NG];
            int nEnt[NG], wbase[NG], widx[NG], wc[NG], wx[NG], wc2[NG], wx2[NG], nextcol[NG], nextx[NG];
            double p[NG][MC];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const bool v = (validMask >> g) & 1u;
                rb[g] = v ? a.rp[gi[g]] : 0;
                nEnt[g] = v ? (int)(a.rp[gi[g] + 1] - rb[g]) : 0;
                wbase[g] = 0;
                widx[g] = 0;
                wc[g] = (lane < nEnt[g]) ? __ldg(a.tcol + rb[g] + lane) : NONE;
                wx[g] = (lane < nEnt[g]) ? __ldg(a.tx + rb[g] + lane) : 0;
                wc2[g] = (32 + lane < nEnt[g]) ? __ldg(a.tcol + rb[g] + 32 + lane) : NONE;
                wx2[g] = (32 + lane < nEnt[g]) ? __ldg(a.tx + rb[g] + 32 + lane) : 0;
                nextcol[g] = __shfl_sync(FULL, wc[g], 0);
                nextx[g] = __shfl_sync(FULL, wx[g], 0);
#pragma unroll
                for (int m = 0; m < MC; ++m) p[g][m] = 0.0;
            }
            // one warp loads tile n into buffer b
            auto issue_tile = [&](long long n, int b) {
                const long long c0 = n * TCt;
                const int tc = (int)min((long long)TCt, nM - c0);
                double* dt = tile + (size_t)b * TC * GM;
                int* di = ttile + (size_t)b * TC * GM;
                const unsigned bytes8 = (unsigned)tc * (unsigned)L * 8u, bytes4 = (unsigned)tc * (unsigned)L * 4u;
                if (useTma && (bytes4 & 15u) == 0) {
                    // the tile's columns are one contiguous slab of each array: two bulk copies (one pair of copies per
                    // column issued by as many lanes measured 10 % slower)
                    if (lane == 0) {
                        mbar_arrive_expect_tx(fullb + b, bytes8 + bytes4);
                        fence_proxy_async();  // rows written by the commit phase (generic proxy) are read by the copy engine
                        bulk_copy_g2s(dt, a.lgt + (size_t)c0 * L, bytes8, fullb + b);
                        bulk_copy_g2s(di, a.t + (size_t)c0 * L, bytes4, fullb + b);
                    }
                    __syncwarp();
                } else {
                    if (patch) {
                        for (int cc0 = lane; cc0 < tc; cc0 += 128) {  // sixteen loads in flight per lane
                            double va[4], vb[4];
                            int ta[4], tb[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int cc = cc0 + 32 * k;
                                const size_t g = (size_t)(c0 + min(cc, tc - 1)) * L;
                                va[k] = __ldcg(a.lgt + g + da);
                                vb[k] = __ldcg(a.lgt + g + db);
                                ta[k] = __ldcg(a.t + g + da);
                                tb[k] = __ldcg(a.t + g + db);
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int cc = cc0 + 32 * k;
                                if (cc < tc) {
                                    dt[cc * 2] = va[k];
                                    dt[cc * 2 + 1] = vb[k];
                                    di[cc * 2] = ta[k];
                                    di[cc * 2 + 1] = tb[k];
                                }
                            }
                        }
                    } else {
                        const int tot = tc * gm;  // entry q <-> (column q / gm, module l0 + q % gm)
                        for (int q0 = lane; q0 < tot; q0 += 256) {  // eight loads in flight per lane
                            double v[8];
                            int tv[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int q = q0 + 32 * k;
                                const size_t g = (size_t)(c0 + q / gm) * L + l0 + q % gm;
                                v[k] = (q < tot) ? __ldcg(a.lgt + g) : 0.0;
                                tv[k] = (q < tot) ? __ldcg(a.t + g) : 0;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int q = q0 + 32 * k;
                                if (q < tot) {
                                    dt[q] = v[k];
                                    di[q] = tv[k];
                                }
                            }
                        }
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(fullb + b);
                }
            };
            __syncthreads();  // every warp is done with the previous task's buffers (and the draw phase's scratch)
            if (warp < YG_NBUF && warp < ntileT) issue_tile(warp, warp);
            for (long long n = 0; n < ntileT; ++n) {
                const int b = (int)(n % YG_NBUF);
                const long long c0 = n * TCt;
                const int tc = (int)min((long long)TCt, nM - c0);
                if (a.flags & 16) mbar_wait_poll(fullb + b, (phaseBits >> b) & 1u);
                else mbar_wait(fullb + b, (phaseBits >> b) & 1u);
                phaseBits ^= 1u << b;
                const double* tl = tile + (size_t)b * TC * GM + laneOff;
                const unsigned tlA = tileA + (unsigned)b * (unsigned)(TC * GM) * 8u, ttA = ttileA + (unsigned)b * (unsigned)(TC * GM) * 4u;
                int cc = validMask ? 0 : tc;  // a warp without a gene only keeps the buffer hand-over going
                while (cc < tc) {
                    // columns where neither gene has a count: (p + cached) - cached, a tight loop over the chains
                    int nc = nextcol[0];
#pragma unroll
                    for (int g = 1; g < NG; ++g) nc = min(nc, nextcol[g]);
                    const long long rel = (long long)nc - c0;
                    const int stop = rel < tc ? (int)rel : tc;
                    if (a.flags & 32) cc = stop;  // opt-in scan mode 1: the count-free columns are skipped (celda_chain_set_scan_mode)
#pragma unroll 2
                    for (; cc < stop; ++cc) {
                        double cached[MC];
#pragma unroll
                        for (int m = 0; m < MC; ++m) cached[m] = tl[cc * stride + 32 * m];
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
#pragma unroll
                            for (int m = 0; m < MC; ++m) p[g][m] += cached[m];
                        }
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
#pragma unroll
                            for (int m = 0; m < MC; ++m) p[g][m] -= cached[m];
                        }
                    }
                    if (cc < tc) {  // a column where the gene has a count: a fresh lgamma against the cached one
                        // Shared-memory loads on 32-bit register addresses with immediate offsets (the compiler otherwise
                        // re-derives the tile and table bases for every chain: 350 instructions per such column, ncu).
                        const unsigned rowA = opaque_u32(tlA + (unsigned)(cc * stride) * 8u),
                                       rowT = opaque_u32(ttA + (unsigned)(cc * stride) * 4u);
                        const int xv = nextx[0], xneg = -xv;
                        double cached[MC], fresh[MC];
                        int arg[MC];
                        unsigned over = 0;  // some chain's argument lies beyond the shared-memory table (rare)
#pragma unroll
                        for (int m = 0; m < MC; ++m) cached[m] = lds_f64(rowA + 256u * m);
#pragma unroll
                        for (int m = 0; m < MC; ++m) {
                            // only one chain of one lane is the gene's current module (curMask; counts removed: - x)
                            arg[m] = lds_s32(rowT + 128u * m) + (((curMask[0] >> m) & 1u) ? xneg : xv);
                            if (!((actMask >> m) & 1u)) arg[m] = 0;
                            over |= (unsigned)(arg[m] >= T);
                        }
                        if (!__any_sync(FULL, over)) {
#pragma unroll
                            for (int m = 0; m < MC; ++m) fresh[m] = lds_f64(lgA + (unsigned)arg[m] * 8u);
                        } else {
#pragma unroll
                            for (int m = 0; m < MC; ++m)
                                fresh[m] = (arg[m] < T) ? lgtab[arg[m]] : dlgamma_hot_call((double)arg[m] + a.beta);
                        }
                        if (++widx[0] == 32) {  // window used up: the one in flight takes over, the next is requested
                            wbase[0] += 32;
                            widx[0] = 0;
                            wc[0] = wc2[0];
                            wx[0] = wx2[0];
                            const int e1 = wbase[0] + 32 + lane;
                            wc2[0] = (e1 < nEnt[0]) ? __ldg(a.tcol + rb[0] + e1) : NONE;
                            wx2[0] = (e1 < nEnt[0]) ? __ldg(a.tx + rb[0] + e1) : 0;
                        }
                        nextcol[0] = __shfl_sync(FULL, wc[0], widx[0]);
                        nextx[0] = __shfl_sync(FULL, wx[0], widx[0]);
#pragma unroll
                        for (int m = 0; m < MC; ++m) {
                            if (m == mstar) {  // warp-uniform: the chunk that holds the current module
                                const bool ic = lane == lstar;
                                p[0][m] += ic ? cached[m] : fresh[m];
                                p[0][m] -= ic ? fresh[m] : cached[m];
                            } else {
                                p[0][m] += fresh[m];
                                p[0][m] -= cached[m];
                            }
                        }
                        ++cc;
                    }
                }
                // ---- buffer b is free once every warp has passed; the last one refills it
                __syncwarp();
                int last = 0;
                if (lane == 0) {
                    __threadfence_block();
                    last = atomicAdd(cnt_s + b, 1) == nwarp - 1;
                    if (last) cnt_s[b] = 0;
                }
                last = __shfl_sync(FULL, last, 0);
                if (last && n + YG_NBUF < ntileT) issue_tile(n + YG_NBUF, b);
            }
            // the per-candidate scalar terms of cG_CalcGibbsProbY's second loop
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (!((validMask >> g) & 1u)) continue;
#pragma unroll
                for (int m = 0; m < MC; ++m) {
                    if ((actMask >> m) & 1u) {
                        const int l = patch ? (lane == 0 ? da : db) : l0 + m * 32 + lane;
                        const bool isCur = (curMask[g] >> m) & 1u;
                        const int gt = g_s[l];
                        const double Tl = (double)T_s[l], Ni = (double)a.nByG[gi[g]];
                        const int g1 = isCur ? gt - 1 : gt, g0 = g1 + 1;
                        double pp = p[g][m];
                        pp += dlgamma_hot_call((double)g0 + a.gamma);
                        pp -= dlgamma_hot_call((double)g1 + a.gamma);
                        pp += (g0 == 0) ? dlgdelta(0, a.delta) : dlgamma_hot_call((double)g0 * a.delta);
                        pp -= (g1 == 0) ? dlgdelta(0, a.delta) : dlgamma_hot_call((double)g1 * a.delta);
                        pp += dlgamma_hot_call(isCur ? Tl - Ni + (double)g1 * a.delta : Tl + (double)g1 * a.delta);
                        pp -= dlgamma_hot_call(isCur ? Tl + (double)g0 * a.delta : Tl + Ni + (double)g0 * a.delta);
                        a.w.bufP[ring_slot(tg[g], a.w.Wmax) * L + l] = pp;
                    }
                }
            }
        };
        // Tasks are handed out through a counter, heaviest first (fresh before patch, densest genes first): a round lasts
        // as long as its busiest CTA, and a fixed round-robin gave CTA 0 the heaviest task of every wave.
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) s_misc[0] = (int)atomicAdd(a.w.barrier + 1, 1u);
            __syncthreads();
            const long long task = s_misc[0];
            if (task >= ntask) break;
            const int q = (int)(task % ngroup);
            const int mc = task >= nFreshTasks ? 0 : (min(GM, L - q * GM) + 31) / 32;
            using one = std::integral_constant<int, 1>;
            // chains per lane: 1, 2, 4 or 7 (a group's chunk count rounded up; idle chains cost adds, no results)
            if (mc <= 1) run_task(one{}, one{}, task);  // mc == 0: a patch task
            else if (mc == 2) run_task(std::integral_constant<int, 2>{}, one{}, task);
            else if (mc <= 4) run_task(std::integral_constant<int, 4>{}, one{}, task);
            else run_task(std::integral_constant<int, 7>{}, one{}, task);
        }
        units_done += ntask;
        CELDA_TICKR(ck_units);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICKR(ck_bar);
        double* pl = scr;
        double* rr = scr + L;
        int* perm = reinterpret_cast<int*>(scr + 2 * L);
        for (int c = gwarpT; c < nwin; c += tw) {
            const long long tt = pos + c;
            const int i = a.ix[tt] - 1;
            const int cur = a.y[i] - 1;
            const size_t slot = ring_slot(tt, a.w.Wmax), o = slot * L;
            double mxOld;
            if (a.doSample && nd == 2 && tt < hi && gavg >= 32 && draw_recorded_zero(a.w, slot, da, db, mxOld)) {
                const double pa = __ldcg(a.w.bufP + o + da), pb = __ldcg(a.w.bufP + o + db);
                if (draw_still_zero(mxOld, pa, pb)) {  // carried over the commit (da, db) with an unchanged draw
                    if (a.probs && lane == 0) {
                        a.probs[(size_t)i * L + da] = pa;
                        a.probs[(size_t)i * L + db] = pb;
                    }
                    continue;
                }
            }
            if (lane == 0 && nd == 2 && tt < hi) ++redraws;
            for (int l = lane; l < L; l += 32) {
                const double p = __ldcg(a.w.bufP + o + l);
                pl[l] = p;
                if (a.probs) a.probs[(size_t)i * L + l] = p;
            }
            __syncwarp();
            int ynew = cur;
            if (a.doSample) {
                double mx;
                int4 nz;
                ynew = warp_sample_ll(pl, L, a.u[tt], rr, perm, lane, mx, nz);
                if (ynew < 0) {
                    if (lane == 0) atomicCAS(a.w.status, 0, ynew);
                    ynew = cur;
                }
                if (lane == 0) {
                    a.w.drawMx[slot] = mx;
                    reinterpret_cast<int4*>(a.w.drawNz)[slot] = nz;
                }
            }
            if (lane == 0) a.w.spec[slot] = (ynew != cur) ? ((cur << 16) | ynew) : -1;
            __syncwarp();
        }
        prev_end = end;
        have_spec = true;
        CELDA_TICKR(ck_draw);
        grid_sync(a.w.barrier, bar_target);
        CELDA_TICKR(ck_bar);
    }
    if (redraws) atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)redraws);
    __syncthreads();  // thread 0's last commit (a mover in the final position) precedes the write-back
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < L; q += blockDim.x) {
            a.nByTS[q] = T_s[q];
            a.nGByTS[q] = g_s[q];
        }
        if (threadIdx.x == 0) {
            a.w.stats[0] += rounds;
            a.w.stats[1] += units_done;
            a.w.stats[2] += movers;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.w.stats[3] += ck_commit;
        a.w.stats[4] += ck_units;
        a.w.stats[5] += ck_draw;
        a.w.stats[6] += ck_bar;
    }
    // unit-phase segment clocks of CTA 0's first and last warp (profile slots 2 and 3, unused by a celda_G chain)
}

inline size_t yg_sweep_smem(int L, int T, int nwarp, int TC) {
    const size_t scratchD = (size_t)2 * L + (L + 1) / 2 + 2;
    return yg_fixed_bytes(L, T) + std::max(sizeof(double) * nwarp * scratchD, yg_unit_bytes(L, TC));
}

// Scan mode 1 of the celda_G sweep, the report: per draw of the last sweep, the distance m between its uniform and the
// nearest boundary of its cumulative distribution (the sampler's descending order) against a bound b on how far the
// skipped operations can move that boundary.  A skipped pair p <- fl(fl(p + a) - a) moves p by at most
// 2^-53 (|p + a| + |p|); along a chain |p| stays below N_g log(ntot + beta + 1) (N_g the gene's total count: every term
// lgamma(t + x + beta) - lgamma(t + beta) is at most x log(t + x + beta)), |a| below amax = max |lgt|; nM pairs at most;
// a change of d in every log-probability moves a boundary by less than 2 d: b = 4 nM 2^-53 (2 N_g log(..) + amax).
// out[0]: min over the draws of m / b (bits of a non-negative double, atomicMin), out[1]: draws with m <= b (a count).
// One warp per gene; plain exp / sums (a report, not part of the chain's arithmetic).
__global__ void k_yg_margin(int nG, int L, long long nM, const double* __restrict__ probs, const int* __restrict__ y,
                            const int* __restrict__ ix, const double* __restrict__ u, const long long* __restrict__ nByG,
                            double logTot, const unsigned long long* __restrict__ amaxBits, unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const double amax = __longlong_as_double((long long)*amaxBits);
    for (int t = gw; t < nG; t += nw) {
        const int i = ix[t] - 1, sel = y[i] - 1;
        const double* p = probs + (size_t)i * L;
        double mx = -1e308;
        for (int l = lane; l < L; l += 32) mx = fmax(mx, p[l]);
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double s = 0.0;
        for (int l = lane; l < L; l += 32) s += exp(p[l] - mx);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double qsel = exp(p[sel] - mx) / s;
        double lower = 0.0;
        int ties = 0, likely = 0, above = 0;
        for (int l = lane; l < L; l += 32) {
            const double q = exp(p[l] - mx) / s;
            if (q > qsel) {
                lower += q;
                ++above;
            }
            ties += (l != sel && q == qsel);
            likely += ((double)L * q > 0.1);
        }
        for (int o = 16; o; o >>= 1) {
            lower += __shfl_xor_sync(0xffffffffu, lower, o);
            ties += __shfl_xor_sync(0xffffffffu, ties, o);
            likely += __shfl_xor_sync(0xffffffffu, likely, o);
            above += __shfl_xor_sync(0xffffffffu, above, o);
        }
        if (lane == 0) {
            const double uu = u[t];
            double m;
            if (ties || likely > 200 || !(s > 0.0)) m = 0.0;  // exact ties and the Walker alias branch are not analysed
            else {
                const double up = lower + qsel - uu, dn = uu - lower;
                // the first sorted position has no boundary below it, the last one (everything else above it) none above
                m = (above == 0) ? up : ((above == L - 1) ? dn : fmin(up, dn));
                m = fmax(m, 0.0);
            }
            const double b = 4.0 * (double)nM * 1.1102230246251565e-16 * (2.0 * (double)nByG[i] * logTot + amax);
            atomicMin(out, (unsigned long long)__double_as_longlong(m / b));
            if (!(m > b)) atomicAdd(out + 1, 1ull);
        }
    }
}

// max |v| over an array of doubles (bits, atomicMax)
__global__ void k_max_abs(const double* __restrict__ v, long long n, unsigned long long* __restrict__ out) {
    double m = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) m = fmax(m, fabs(v[q]));
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// lgt = lgamma(t + beta) for the whole L x nM table (at set_state; the sweep keeps it current afterwards)
__global__ void k_lgt_fill(const int* __restrict__ t, long long n, double beta, double* __restrict__ lgt) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int v = t[q];
        lgt[q] = dlgamma_hot((double)v + beta);
    }
}

// ------------------------------------------------------------------------------------------------ z sweep, celda_C
// .cCCalcGibbsProbZ on the RAW counts (celda_C: R/celda_C.R:427-440 with algorithm = "Gibbs"): per (cell,
// population) the reference sums lgamma over ALL nG genes, zeros included (sum(lgamma(nGByCP[, j] + counts[, i] +
// beta)), :664-687) -- a serial chain of nG adds whose order fixes the bits.  Same structure as k_sweep_yg with the
// roles swapped: nGByCP (gene-major [nG][K]) and its cached lgamma stay in HBM; one CTA task = (one cell per warp)
// x (32 populations, one per lane); the cached-lgamma tile of 128 genes is staged in shared memory once for the
// CTA's cells; each warp walks it in gene order with the cell's CSC column as a cursor, adding the cached value
// where the cell has no count and a fresh lgamma where it has one.
constexpr int YG_TILE = 128;  // genes per staged tile of k_sweep_zc

struct ZCArgs {
    int nG, K, nS;
    long long nM;
    const int* cp;      // CSC column pointers [nM + 1]
    const int* ci;      //   row of each stored entry, ascending within a column
    const int* cx;      //   value
    const int* nByC;
    const int* s;
    int* z;             // [nM] 1-based, updated
    int* n;             // nGByCP gene-major [nG][K], updated
    double* lgn;        // lgamma(n + beta), same layout, maintained
    long long* nCP;     // [K], updated
    int* mCPByS;        // K x nS, updated
    double alpha, beta;
    const int* ix;
    const double* u;
    int doSample;
    double* probs;      // K x nM or null
    SweepWork w;
};

__global__ void __launch_bounds__(1024, 1) k_sweep_zc(ZCArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.w.status[1]) return;  // the injected visiting order failed its check: nothing may be indexed through it
    const int K = a.K, nS = a.nS, T = a.w.lgT, R = a.nG;
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scratchD = 2 * K + (K + 1) / 2 + 2;
    double* lgtab = reinterpret_cast<double*>(smem_raw);
    double* tile = lgtab + T;                             // [YG_TILE][32] lgamma(n + beta)
    double* scratch_all = tile + YG_TILE * 32;
    long long* nCP_s = reinterpret_cast<long long*>(scratch_all + (size_t)nwarp * scratchD);
    int* ntile = reinterpret_cast<int*>(nCP_s + K);       // [YG_TILE][32] n
    int* m_s = ntile + YG_TILE * 32;
    int* s_misc = m_s + K * nS;
    double* scr = scratch_all + (size_t)warp * scratchD;
    const int tw = gridDim.x * nwarp;
    const int gwarpT = warp * gridDim.x + blockIdx.x;
    const int nchunk = (K + 31) / 32;
    const double Rbeta = (double)R * a.beta;

    for (int q = threadIdx.x; q < T; q += blockDim.x) lgtab[q] = dlgamma((double)q + a.beta);
    for (int q = threadIdx.x; q < K; q += blockDim.x) nCP_s[q] = a.nCP[q];
    for (int q = threadIdx.x; q < K * nS; q += blockDim.x) m_s[q] = a.mCPByS[q];
    __syncthreads();

    long long pos = 0, hi = 0, prev_end = 0;
    int W = a.w.Wfull, nd = K, da = -1, db = -1, gavg = a.w.Wfull;  // nd == K: every cached column sum is due
    bool have_spec = false;
    unsigned bar_target = 0;
    long long rounds = 0, units_done = 0, movers = 0;
    long long redraws = 0;  // carried items drawn again in full (this thread's share)

    while (true) {
        bool committed = false;
        if (have_spec) {
            const int nwin = (int)(prev_end - pos);
            const int first = block_first_mover(a.w.spec, pos, a.w.Wmax, nwin, s_misc);
            if (first < 0) {
                pos = prev_end;
                hi = pos;
                nd = 0;
                gavg = min(2 * gavg + nwin, 1 << 24);
                W = min(2 * W, a.w.Wfull);
            } else {
                const long long tt = pos + first;
                const long long i = a.ix[tt] - 1;
                const int packed = __ldcg(a.w.spec + ring_slot(pos + first, a.w.Wmax));
                da = packed >> 16;
                db = packed & 0xffff;
                // the two table columns change only where the cell has counts: the grid shares the cell's entries
                const int e0 = a.cp[i], e1 = a.cp[i + 1];
                for (int e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += gridDim.x * blockDim.x) {
                    const size_t base = (size_t)a.ci[e] * K;
                    const int xv = a.cx[e];
                    const int na = __ldcg(a.n + base + da) - xv, nb = __ldcg(a.n + base + db) + xv;
                    a.n[base + da] = na;
                    a.n[base + db] = nb;
                    a.lgn[base + da] = lg_tab(lgtab, T, na, a.beta);
                    a.lgn[base + db] = lg_tab(lgtab, T, nb, a.beta);
                }
                if (threadIdx.x == 0) {
                    const int Ni = a.nByC[i], si = a.s[i] - 1;
                    nCP_s[da] -= Ni;
                    nCP_s[db] += Ni;
                    m_s[si * K + da] -= 1;
                    m_s[si * K + db] += 1;
                    if (blockIdx.x == 0) a.z[i] = db + 1;
                }
                ++movers;
                gavg = max(1, (3 * gavg + first + 1) / 4);
                W = (int)max((long long)a.w.Wmin, min(4LL * gavg, (long long)a.w.Wfull));
                W = ((W + nwarp - 1) / nwarp) * nwarp;
                pos = tt + 1;
                hi = min(prev_end, pos + W);
                nd = 2;
                committed = true;
            }
        }
        if (pos >= a.nM) break;
        if (committed || rounds == 0) grid_sync(a.w.barrier, bar_target);  // table columns visible before they are read
        // after a commit the window is topped up with fresh items only once it is less than half full: carried items
        // cost two patch units and a check (draw_unchanged), fresh ones a full evaluation and draw, cheaper in batches
        const bool topup = nd != 2 || (hi - pos) * 2 < W;
        const long long end = topup ? max(hi, min(pos + (long long)W, a.nM)) : hi;
        const int nwin = (int)(end - pos);
        // ---- tasks: cached column sums of the changed columns (one warp, lane = column); patch = (nwarp carried
        // cells) x {da, db}; fresh = (nwarp new cells) x (32-population chunk)
        const long long nSumTask = (nd == K) ? nchunk : (nd == 2 ? 1 : 0);
        const long long nPatchGrp = (nd == 2) ? (hi - pos + nwarp - 1) / nwarp : 0;
        const long long nFreshGrp = (end - hi + nwarp - 1) / nwarp;
        const long long ntask = nSumTask + nPatchGrp + nFreshGrp * nchunk;
        ++rounds;
        for (long long task = blockIdx.x; task < ntask; task += gridDim.x) {
            if (task < nSumTask) {
                // sum(lgamma(nGByCP[, j] + beta)) added serially over the genes, one lane per column
                if (warp == 0) {
                    int j;
                    if (nd == K) {
                        j = (int)task * 32 + lane;
                        if (j >= K) j = -1;
                    } else {
                        j = (lane == 0) ? da : (lane == 1 ? db : -1);
                    }
                    if (j >= 0) {
                        double S = 0.0;
#pragma unroll 8
                        for (int r = 0; r < R; ++r) S += __ldcg(a.lgn + (size_t)r * K + j);
                        a.w.colsum[j] = S;
                        a.w.lgnCP[j] = dlgamma((double)nCP_s[j] + Rbeta);
                    }
                }
                continue;
            }
            long long tcell;
            int j;
            bool valid;
            if (task < nSumTask + nPatchGrp) {
                tcell = pos + (task - nSumTask) * nwarp + warp;
                valid = tcell < hi;
                j = (lane == 0) ? da : (lane == 1 ? db : -1);
            } else {
                const long long q = task - nSumTask - nPatchGrp;
                tcell = hi + (q / nchunk) * nwarp + warp;
                valid = tcell < end;
                j = (int)(q % nchunk) * 32 + lane;
                if (j >= K) j = -1;
            }
            const long long i = valid ? a.ix[tcell] - 1 : 0;
            const int zi = valid ? a.z[i] - 1 : -1;
            const bool active = valid && j >= 0;
            const int sgn = (active && j == zi) ? -1 : 1;
            const int jsafe = (j >= 0) ? j : 0;
            // cursor into the cell's CSC column: a register window of 32 entries per warp, the next window in flight
            // (see k_sweep_yg)
            const int nzEnd = valid ? a.cp[i + 1] : 0;
            int wbase = valid ? a.cp[i] : 0;
            const int NONE = 0x7fffffff;
            int wc, wx, wc2, wx2, widx = 0;
            {
                const int e0 = wbase + lane, e1 = e0 + 32;
                wc = (e0 < nzEnd) ? __ldg(a.ci + e0) : NONE;
                wx = (e0 < nzEnd) ? __ldg(a.cx + e0) : 0;
                wc2 = (e1 < nzEnd) ? __ldg(a.ci + e1) : NONE;
                wx2 = (e1 < nzEnd) ? __ldg(a.cx + e1) : 0;
            }
            int nextrow = __shfl_sync(0xffffffffu, wc, 0), nextx = __shfl_sync(0xffffffffu, wx, 0);
            double S = 0.0;
            for (int r0 = 0; r0 < R; r0 += YG_TILE) {
                const int tc = min(YG_TILE, R - r0);
                __syncthreads();
                for (int rr = warp; rr < tc; rr += nwarp) {
                    const size_t g = (size_t)(r0 + rr) * K + jsafe;
                    tile[rr * 32 + lane] = (j >= 0) ? __ldcg(a.lgn + g) : 0.0;
                    ntile[rr * 32 + lane] = (j >= 0) ? __ldcg(a.n + g) : 0;
                }
                __syncthreads();
                if (valid) {  // warp-uniform: the whole warp follows one cell, inactive lanes add zeros
                    int rr = 0;
                    while (rr < tc) {
                        const int stop = ((long long)nextrow - r0 < tc) ? nextrow - r0 : tc;
#pragma unroll 4
                        for (; rr < stop; ++rr) S += tile[rr * 32 + lane];
                        if (rr < tc) {
                            // one warp vote for the rare argument beyond the shared-memory table instead of a divergent
                            // inlined lgamma per lane (see k_sweep_yg)
                            const int karg = ntile[rr * 32 + lane] + sgn * nextx;
                            if (!__any_sync(0xffffffffu, karg >= T)) S += lgtab[karg];
                            else S += (karg < T) ? lgtab[karg] : dlgamma_hot_call((double)karg + a.beta);
                            if (++widx == 32) {
                                wbase += 32;
                                widx = 0;
                                wc = wc2;
                                wx = wx2;
                                const int e1 = wbase + 32 + lane;
                                wc2 = (e1 < nzEnd) ? __ldg(a.ci + e1) : NONE;
                                wx2 = (e1 < nzEnd) ? __ldg(a.cx + e1) : 0;
                            }
                            nextrow = __shfl_sync(0xffffffffu, wc, widx);
                            nextx = __shfl_sync(0xffffffffu, wx, widx);
                            ++rr;
                        }
                    }
                }
            }
            if (active) {
                const size_t o = ring_slot(tcell, a.w.Wmax) * K + j;
                a.w.bufP[o] = S;
                a.w.bufA[o] = dlgamma_hot((double)(nCP_s[j] + (long long)sgn * a.nByC[i]) + Rbeta);
            }
        }
        units_done += ntask;
        grid_sync(a.w.barrier, bar_target);
        // ---- combine in the reference's order (R/celda_C.R:664-687), draw
        double* pl = scr;
        double* rr2 = scr + K;
        int* perm = reinterpret_cast<int*>(scr + 2 * K);
        for (int c = gwarpT; c < nwin; c += tw) {
            const long long tt = pos + c;
            const long long i = a.ix[tt] - 1;
            const int zi = a.z[i] - 1, si = a.s[i] - 1;
            const size_t slot = ring_slot(tt, a.w.Wmax), o = slot * K;
            auto combine = [&](int j) {
                const double S = __ldcg(a.w.bufP + o + j), A = __ldcg(a.w.bufA + o + j);
                const double cs = __ldcg(a.w.colsum + j), lc = __ldcg(a.w.lgnCP + j);
                double p = dlog((double)(m_s[si * K + j] - (j == zi)) + a.alpha);
                if (j != zi) {
                    p = p + S;
                    p = p - A;
                    p = p - cs;
                    p = p + lc;
                } else {
                    p = p + cs;
                    p = p - lc;
                    p = p - S;
                    p = p + A;
                }
                return p;
            };
            double mxOld;
            if (a.doSample && nd == 2 && tt < hi && gavg >= 32 && draw_recorded_zero(a.w, slot, da, db, mxOld)) {
                const double pa = combine(da), pb = combine(db);
                if (draw_still_zero(mxOld, pa, pb)) {  // carried over the commit (da, db) with an unchanged draw
                    if (a.probs && lane == 0) {
                        a.probs[i * K + da] = pa;
                        a.probs[i * K + db] = pb;
                    }
                    continue;
                }
            }
            if (lane == 0 && nd == 2 && tt < hi) ++redraws;
            for (int j = lane; j < K; j += 32) {
                const double p = combine(j);
                pl[j] = p;
                if (a.probs) a.probs[i * K + j] = p;
            }
            __syncwarp();
            int znew = zi;
            if (a.doSample) {
                double mx;
                int4 nz;
                znew = warp_sample_ll(pl, K, a.u[tt], rr2, perm, lane, mx, nz);
                if (znew < 0) {
                    if (lane == 0) atomicCAS(a.w.status, 0, znew);
                    znew = zi;
                }
                if (lane == 0) {
                    a.w.drawMx[slot] = mx;
                    reinterpret_cast<int4*>(a.w.drawNz)[slot] = nz;
                }
            }
            if (lane == 0) a.w.spec[slot] = (znew != zi) ? ((zi << 16) | znew) : -1;
            __syncwarp();
        }
        prev_end = end;
        have_spec = true;
        grid_sync(a.w.barrier, bar_target);
    }
    if (redraws) atomicAdd(reinterpret_cast<unsigned long long*>(a.w.stats + 7), (unsigned long long)redraws);
    __syncthreads();  // thread 0's last commit (a mover in the final position) precedes the write-back
    if (blockIdx.x == 0) {
        for (int q = threadIdx.x; q < K; q += blockDim.x) a.nCP[q] = nCP_s[q];
        for (int q = threadIdx.x; q < K * nS; q += blockDim.x) a.mCPByS[q] = m_s[q];
        if (threadIdx.x == 0) {
            a.w.stats[0] += rounds;
            a.w.stats[1] += units_done;
            a.w.stats[2] += movers;
        }
    }
}

inline size_t zc_sweep_smem(int K, int nS, int T, int nwarp) {
    const size_t scratchD = (size_t)2 * K + (K + 1) / 2 + 2;
    return sizeof(double) * (T + YG_TILE * 32 + nwarp * scratchD) + sizeof(long long) * K +
           sizeof(int) * (YG_TILE * 32 + (size_t)K * nS + 4);
}

// lgn = lgamma(n + beta) over a gene-major [nG][K] table (k_lgt_fill serves: layout-agnostic)

}  // namespace celda
