/*
 * cpu_arms.c — the CPU arms bench.py times next to every query: TEST / MEASUREMENT INFRASTRUCTURE,
 * never on the product path (only bench.py's cpu_baseline / --impl reference legs and tests/ call into it).
 *
 *   cpu-ref-1t   the reference's own algorithm, one thread: oracle.c (`futhark c`, the backend setup.sh:12 builds)
 *   cpu-ref-mt   the same algorithm with its data-parallel loops on all host cores (stand-in for `futhark multicore`):
 *                arm_groupby_ref_mt below = groupby.fut:8-22 (32 stable one-bit passes over whole rows) with the
 *                per-pass scan + scatter parallelised, then the segmented reduce of groupby.fut:55-58
 *   cpu-best     what a tuned columnar CPU engine would do for the same query: single-pass columnar filter with
 *                selection vectors, dense-array hash aggregate, range-trimmed parallel LSD radix sort, direct-address
 *                join probe — OpenMP on all host cores, -O3 -march=x86-64-v3 (AVX2 auto-vectorised)
 *
 * None of this restates reference source beyond what oracle.c already cites; cpu-best has no reference counterpart.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ARM_BLOCK 4096

/* ---- cpu-best, config 2: SELECT s0, s1 WHERE p0 > c0 AND p1 < c1 over f32 columns (SoA).  One pass: every thread
 * walks its row range in blocks, builds a selection vector branch-free from the two predicate columns, gathers the
 * selected columns.  The result is CHUNKED: thread t's rows start at out[r0_t] and counts[t] says how many there are
 * (a columnar engine hands such chunks on without compacting them).  Returns the total row count. */
int64_t arm_filter_best_f32(const float *p0, float c0, const float *p1, float c1, const float *s0, const float *s1,
                            int64_t n, float *out0, float *out1, int64_t *counts, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    int64_t total = 0;
#pragma omp parallel num_threads(T) reduction(+ : total)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T;
        int32_t sel[ARM_BLOCK];
        int64_t w = r0;
        for (int64_t b = r0; b < r1; b += ARM_BLOCK) {
            int len = (int)((r1 - b) < ARM_BLOCK ? (r1 - b) : ARM_BLOCK);
            int k = 0;
            for (int i = 0; i < len; i++) {
                sel[k] = i;
                k += (p0[b + i] > c0) & (p1[b + i] < c1);
            }
            for (int i = 0; i < k; i++) {
                out0[w + i] = s0[b + sel[i]];
                out1[w + i] = s1[b + sel[i]];
            }
            w += k;
        }
        counts[tid] = w - r0;
        total += w - r0;
    }
    return total;
}

/* ---- cpu-best, config 3: GROUP BY key (i32 in [kmin, kmin+range)) SUM/COUNT of val; AVG = sum / count is the
 * caller's division.  Per-thread dense tables (count, sum), merged in parallel over the key range. */
int arm_groupby_best_i32(const int32_t *key, const int32_t *val, int64_t n, int32_t kmin, int64_t range, int64_t *cnt_out,
                         int64_t *sum_out, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    int64_t *tab = (int64_t *)calloc((size_t)T * (size_t)range * 2, sizeof(int64_t));
    if (!tab) return 3;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        int64_t *c = tab + (size_t)tid * (size_t)range * 2, *s = c + range;
        int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T;
        for (int64_t i = r0; i < r1; i++) {
            int64_t k = (int64_t)key[i] - kmin;
            c[k]++;
            s[k] += val[i];
        }
#pragma omp barrier
#pragma omp for schedule(static)
        for (int64_t k = 0; k < range; k++) {
            int64_t cc = 0, ss = 0;
            for (int t = 0; t < T; t++) {
                cc += tab[(size_t)t * (size_t)range * 2 + k];
                ss += tab[(size_t)t * (size_t)range * 2 + range + k];
            }
            cnt_out[k] = cc;
            sum_out[k] = ss;
        }
    }
    free(tab);
    return 0;
}

int arm_groupby_best_f32(const int32_t *key, const float *val, int64_t n, int32_t kmin, int64_t range, int64_t *cnt_out,
                         double *sum_out, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    int64_t *ctab = (int64_t *)calloc((size_t)T * (size_t)range, sizeof(int64_t));
    double *stab = (double *)calloc((size_t)T * (size_t)range, sizeof(double));
    if (!ctab || !stab) { free(ctab); free(stab); return 3; }
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        int64_t *c = ctab + (size_t)tid * (size_t)range;
        double *s = stab + (size_t)tid * (size_t)range;
        int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T;
        for (int64_t i = r0; i < r1; i++) {
            int64_t k = (int64_t)key[i] - kmin;
            c[k]++;
            s[k] += (double)val[i];
        }
#pragma omp barrier
#pragma omp for schedule(static)
        for (int64_t k = 0; k < range; k++) {
            int64_t cc = 0;
            double ss = 0.0;
            for (int t = 0; t < T; t++) {
                cc += ctab[(size_t)t * (size_t)range + k];
                ss += stab[(size_t)t * (size_t)range + k];
            }
            cnt_out[k] = cc;
            sum_out[k] = ss;
        }
    }
    free(ctab);
    free(stab);
    return 0;
}

/* ---- cpu-ref-mt, GROUP BY: groupby.fut:8-22 with each one-bit pass parallelised the way a multicore scan + scatter
 * would run it (per-thread zero counts -> exclusive prefix -> stable scatter), then groupby.fut:55-58 sequentially.
 * rows: [n][s] u32, column 0 = key; t_cols as groupby.fut:35-41.  out [G][s]; returns G. */
static inline uint32_t arm_type_func(int32_t typ, uint32_t v1, uint32_t v2) {
    switch (typ) {
    case 1: return v1 * v2;
    case 2: return v1 + v2;
    case 3: return v1 > v2 ? v1 : v2;
    default: return v1 < v2 ? v1 : v2;
    }
}

int64_t arm_groupby_ref_mt(const uint32_t *rows, int64_t n, int64_t s, const int32_t *t_cols, uint32_t *out, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    if (n == 0) return 0;
    uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n * s));
    uint32_t *b = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n * s));
    int64_t *zeros = (int64_t *)malloc(sizeof(int64_t) * (size_t)(T + 1));
    if (!a || !b || !zeros) { free(a); free(b); free(zeros); return -1; }
    memcpy(a, rows, sizeof(uint32_t) * (size_t)(n * s));
    for (int bit = 0; bit < 32; bit++) {
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T, z = 0;
            for (int64_t i = r0; i < r1; i++) z += !((a[i * s] >> bit) & 1u);
            zeros[tid + 1] = z;
#pragma omp barrier
#pragma omp single
            {
                zeros[0] = 0;
                for (int t = 0; t < T; t++) zeros[t + 1] += zeros[t];
            }
            int64_t p0 = zeros[tid], p1 = zeros[T] + (r0 - zeros[tid]);
            for (int64_t i = r0; i < r1; i++) {
                int64_t d = ((a[i * s] >> bit) & 1u) ? p1++ : p0++;
                memcpy(b + d * s, a + i * s, sizeof(uint32_t) * (size_t)s);
            }
        }
        uint32_t *t = a; a = b; b = t;
    }
    int64_t g = -1;
    for (int64_t r = 0; r < n; r++) {
        if (r == 0 || a[(r - 1) * s] != a[r * s]) {
            g++;
            memcpy(out + g * s, a + r * s, sizeof(uint32_t) * (size_t)s);
        } else {
            for (int64_t j = 1; j < s; j++) out[g * s + j] = arm_type_func(t_cols[j - 1], out[g * s + j], a[r * s + j]);
        }
    }
    free(a); free(b); free(zeros);
    return g + 1;
}

/* ---- ORDER BY a, b (i64, ascending, signed), stable.  digit_bits = 8: cpu-best (parallel LSD radix, only the digits
 * the key ranges need, least significant first).  digit_bits = 1: the reference's rsort generalised to two signed 64-bit
 * keys (one stable split per significant bit, groupby.fut:8-22 / join.fut:9-23) — the cpu-ref arm; the reference has no
 * ORDER BY operator of its own.  Sorts (a, b) in place using (ta, tb) as the ping-pong buffers. */
int arm_orderby_i64x2(int64_t *a, int64_t *b, int64_t n, int64_t *ta, int64_t *tb, int32_t digit_bits, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    if (n < 2) return 0;
    const int D = digit_bits == 1 ? 1 : 8, R = 1 << D;
    uint64_t lo[2] = {~0ull, ~0ull}, hi[2] = {0, 0};
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = (uint64_t)a[i] ^ 0x8000000000000000ull, y = (uint64_t)b[i] ^ 0x8000000000000000ull;
        if (x < lo[0]) lo[0] = x;
        if (x > hi[0]) hi[0] = x;
        if (y < lo[1]) lo[1] = y;
        if (y > hi[1]) hi[1] = y;
    }
    int64_t *hist = (int64_t *)malloc(sizeof(int64_t) * (size_t)T * (size_t)R);
    if (!hist) return 3;
    int64_t *sa = a, *sb = b, *da = ta, *db = tb;
    for (int k = 1; k >= 0; k--) { /* least significant key first */
        uint64_t span = hi[k] - lo[k];
        int bits = 0;
        while (bits < 64 && (span >> bits)) bits++;
        for (int sh = 0; sh < bits; sh += D) {
            const uint64_t base = lo[k], mask = (uint64_t)R - 1;
#pragma omp parallel num_threads(T)
            {
#ifdef _OPENMP
                int tid = omp_get_thread_num();
#else
                int tid = 0;
#endif
                int64_t *h = hist + (size_t)tid * (size_t)R;
                memset(h, 0, sizeof(int64_t) * (size_t)R);
                int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T;
                const int64_t *src = k == 0 ? sa : sb;
                for (int64_t i = r0; i < r1; i++) h[((((uint64_t)src[i] ^ 0x8000000000000000ull) - base) >> sh) & mask]++;
#pragma omp barrier
#pragma omp single
                {
                    int64_t run = 0;
                    for (int d = 0; d < R; d++)
                        for (int t = 0; t < T; t++) {
                            int64_t c = hist[(size_t)t * (size_t)R + d];
                            hist[(size_t)t * (size_t)R + d] = run;
                            run += c;
                        }
                }
                for (int64_t i = r0; i < r1; i++) {
                    int64_t p = h[((((uint64_t)src[i] ^ 0x8000000000000000ull) - base) >> sh) & mask]++;
                    da[p] = sa[i];
                    db[p] = sb[i];
                }
            }
            int64_t *t1 = sa; sa = da; da = t1;
            int64_t *t2 = sb; sb = db; db = t2;
        }
    }
    if (sa != a) {
        memcpy(a, sa, sizeof(int64_t) * (size_t)n);
        memcpy(b, sb, sizeof(int64_t) * (size_t)n);
    }
    free(hist);
    return 0;
}

/* ---- cpu-best, config 5: fact JOIN dim ON fk = pk GROUP BY attr, SUM(val), COUNT(*).  Direct-address lookup
 * pk -> attr slot over [pk_min, pk_min + pk_span), probed by every fact row; per-thread (count, sum) tables over the
 * attr range.  lut is caller-provided scratch of pk_span int32 (filled here). */
int arm_join_groupby_best(const int32_t *fk, const int32_t *val, int64_t nf, const int32_t *pk, const int32_t *attr, int64_t nd,
                          int64_t pk_min, int64_t pk_span, int32_t g_min, int64_t g_range, int32_t *lut, int64_t *cnt_out,
                          int64_t *sum_out, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    int64_t *tab = (int64_t *)calloc((size_t)T * (size_t)g_range * 2, sizeof(int64_t));
    if (!tab) return 3;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
#pragma omp for schedule(static)
        for (int64_t i = 0; i < pk_span; i++) lut[i] = -1;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < nd; i++) lut[(int64_t)pk[i] - pk_min] = attr[i] - g_min;
        int64_t *c = tab + (size_t)tid * (size_t)g_range * 2, *s = c + g_range;
        int64_t r0 = nf * tid / T, r1 = nf * (tid + 1) / T;
        for (int64_t i = r0; i < r1; i++) {
            int64_t v = (int64_t)fk[i] - pk_min;
            if (v < 0 || v >= pk_span) continue;
            int32_t g = lut[v];
            if (g < 0) continue;
            c[g]++;
            s[g] += val[i];
        }
#pragma omp barrier
#pragma omp for schedule(static)
        for (int64_t k = 0; k < g_range; k++) {
            int64_t cc = 0, ss = 0;
            for (int t = 0; t < T; t++) {
                cc += tab[(size_t)t * (size_t)g_range * 2 + k];
                ss += tab[(size_t)t * (size_t)g_range * 2 + g_range + k];
            }
            cnt_out[k] = cc;
            sum_out[k] = ss;
        }
    }
    free(tab);
    return 0;
}

/* ---- cpu-ref-mt, config 2: the reference's data model (ONE row-major [n][m] f32 array, table.py:52-74) and what
 * `filter` + `map (sel cols)` (select.fut:18-19) would do, in a single pass per thread: a row that passes is projected
 * straight into the thread's chunk of the output (row-major [.][2]).  Same chunked result convention as
 * arm_filter_best_f32.  oracle_query_filter (oracle.c) stays the simple, generic checker. */
int64_t arm_filter_rowmajor_f32(const float *db, int64_t n, int64_t m, int32_t pc0, float c0, int32_t pc1, float c1, int32_t s0,
                                int32_t s1, float *out, int64_t *counts, int32_t threads) {
    int T = threads > 0 ? threads : 1;
    int64_t total = 0;
#pragma omp parallel num_threads(T) reduction(+ : total)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        int64_t r0 = n * tid / T, r1 = n * (tid + 1) / T, w = r0;
        for (int64_t r = r0; r < r1; r++) {
            const float *row = db + r * m;
            out[2 * w] = row[s0];
            out[2 * w + 1] = row[s1];
            w += (row[pc0] > c0) & (row[pc1] < c1);
        }
        counts[tid] = w - r0;
        total += w - r0;
    }
    return total;
}
