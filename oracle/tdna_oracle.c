/*
 * tdna_oracle.c -- CPU restatement of TaxonDNA's per-pair distance arithmetic.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (taxondna_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity pinning: the reference ships exactly three known-answer tests for this
 * path and this file is checked against all of them in tests/test_oracle_kat.py:
 *   KAT-1  Sequence.java:1913-1936  shared length 10, uncorrected p = 0.50
 *   KAT-2  Sequence.java:1942-1955  shared length 17
 *   KAT-3  Testing.java:73-87       transversion-only distance 0.2 (+ class table :40-69)
 * No expected output exists in the reference for K2P, Best Match, Pairwise Summary
 * or Cluster: for those the consumers in oracle/consumers.py are "parity unpinned"
 * by the reference's own tests.  No JVM exists in this image, so the
 * reference itself cannot be run; oracle/_ref is therefore not built.
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/src/main/java/com/ggvaidya/TaxonDNA/.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off; fp64 only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TDO_PDM_UNCORRECTED 0 /* Common/DNA/Sequence.java:53 */
#define TDO_PDM_K2P 1         /* Common/DNA/Sequence.java:54 */
#define TDO_PDM_TRANS_ONLY 2  /* Common/DNA/Sequence.java:55 */

/* ------------------------------------------------------------------ */
/* Single-character predicates                                         */
/* ------------------------------------------------------------------ */

/* Sequence.isValid, Common/DNA/Sequence.java:927-965.  Note the reference's
 * lower->upper conversion is (ch > 'a' && ch < 'z'), i.e. it skips 'a' and 'z'. */
int tdo_is_valid(int ch) {
    if (ch > 'a' && ch < 'z') ch = ch - ('a' - 'A');
    switch (ch) {
        case 'A': case 'C': case 'T': case 'G':
        case 'R': case 'Y': case 'K': case 'M': case 'S': case 'N':
        case 'B': case 'D': case 'H': case 'V': case 'W':
        case '-': case '?': case '_':
            return 1;
    }
    return 0;
}

/* Sequence.isAmbiguous, Common/DNA/Sequence.java:968-1005 */
int tdo_is_ambiguous(int ch) {
    if (ch > 'a' && ch < 'z') ch = ch - ('a' - 'A');
    switch (ch) {
        case 'R': case 'Y': case 'K': case 'M': case 'S': case 'N':
        case 'B': case 'D': case 'H': case 'V': case 'W':
            return 1;
    }
    return 0;
}

/* Sequence.getint, Common/DNA/Sequence.java:1125-1143: A=1 C=2 T=4 G=8 */
int tdo_getint(int ch) {
    int r = 0;
    if (ch >= 'a' && ch <= 'z') ch = ch - ('a' - 'A');
    if (ch == 'A' || ch == 'R' || ch == 'M' || ch == 'N' || ch == 'D' || ch == 'H' || ch == 'V' || ch == 'W') r |= 1;
    if (ch == 'C' || ch == 'Y' || ch == 'M' || ch == 'S' || ch == 'N' || ch == 'B' || ch == 'H' || ch == 'V') r |= 2;
    if (ch == 'T' || ch == 'Y' || ch == 'K' || ch == 'N' || ch == 'B' || ch == 'D' || ch == 'H' || ch == 'W') r |= 4;
    if (ch == 'G' || ch == 'R' || ch == 'K' || ch == 'S' || ch == 'N' || ch == 'B' || ch == 'D' || ch == 'V') r |= 8;
    return r;
}

/* Sequence.getcode, Common/DNA/Sequence.java:1149-1196 (inverse of getint; 0 -> '-') */
int tdo_getcode(int val) {
    static const char tbl[16] = {'-', 'A', 'C', 'M', 'T', 'W', 'Y', 'H',
                                 'G', 'R', 'S', 'V', 'K', 'D', 'B', 'N'};
    return tbl[val & 15];
}

/* Sequence.isPyrimidine / isPurine, Common/DNA/Sequence.java:1202-1216 */
int tdo_is_pyrimidine(int ch) { return ch == 'Y' || ch == 'T' || ch == 'C'; }
int tdo_is_purine(int ch) { return ch == 'R' || ch == 'A' || ch == 'G'; }
/* Sequence.isGap / isInternalGap, Common/DNA/Sequence.java:1227-1237 */
static int is_gap(int ch) { return ch == '-' || ch == '_'; }
static int is_internal_gap(int ch) { return ch == '-'; }

/* Sequence.identical(char,char), Common/DNA/Sequence.java:1250-1303.
 * ambig_allowed mirrors the static Sequence.ambiguousBasesAllowed (:62).  The
 * `ch != 'A' || ch != 'C' ...` test at :1280-1283 is always true, so with
 * ambig_allowed == 0 every base becomes 'N' and any two bases match. */
int tdo_identical(int ch1, int ch2, int ambig_allowed) {
    if (!tdo_is_valid(ch1) || !tdo_is_valid(ch2)) return 0;
    if (ch1 > 'a' && ch1 < 'z') ch1 = ch1 - ('a' - 'A');
    if (ch2 > 'a' && ch2 < 'z') ch2 = ch2 - ('a' - 'A');
    if (ch1 == '?' || ch2 == '?') return 0;
    if (is_gap(ch1) || is_gap(ch2)) {
        if (is_internal_gap(ch1) || is_internal_gap(ch2)) {
            if (ch1 == ch2) return 1;
        }
        return 0;
    }
    if (!ambig_allowed) {
        if (ch1 != 'A' || ch1 != 'C' || ch1 != 'T' || ch1 != 'G') ch1 = 'N';
        if (ch2 != 'A' || ch2 != 'C' || ch2 != 'T' || ch2 != 'G') ch2 = 'N';
    }
    if (ch1 == ch2) return 1;
    if (ambig_allowed) {
        if ((tdo_getint(ch1) & tdo_getint(ch2)) != 0) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Normalisation                                                       */
/* ------------------------------------------------------------------ */

/* Sequence.changeSequence, Common/DNA/Sequence.java:751-853 (+ doSanityChecks
 * :1727-1763, + the toUpperCase().trim() of the constructor :163-166).
 * Returns 0 and writes the normalised symbols (alphabet ACGT RYKMSWBDHVN - _ ?)
 * to out[0..*out_len); negative on the conditions that make the reference throw
 * SequenceException:  -1 nested bracket, -2 invalid char inside a bracket group,
 * -3 literal '_' in the input, -4 uracil, -5 illegal base. */
int tdo_normalise(const char *raw, int n, char *out, int *out_len) {
    char *up = (char *)malloc((size_t)n + 1);
    int lo = 0, hi = n, m = 0, i;
    /* constructor: seq.toUpperCase().trim() -- trim drops chars <= ' ' */
    while (lo < hi && (unsigned char)raw[lo] <= ' ') lo++;
    while (hi > lo && (unsigned char)raw[hi - 1] <= ' ') hi--;
    for (i = lo; i < hi; i++) {
        int c = (unsigned char)raw[i];
        if (c >= 'a' && c <= 'z') c -= 32;
        up[m++] = (char)c;
    }
    /* step 1: [..] / (..) groups -> one IUPAC letter (:761-796) */
    int k = 0;
    for (int x = 0; x < m; x++) {
        int ch = up[x];
        if (ch == '(' || ch == '[') {
            int single = 0, y;
            for (y = x + 1; y < m; y++, x++) {
                ch = up[y];
                if (ch == '(' || ch == '[') { free(up); return -1; }
                else if (ch == ')' || ch == ']') break;
                else {
                    if (!tdo_is_valid(ch)) { free(up); return -2; }
                    single |= tdo_getint(ch);
                }
            }
            x++;
            out[k++] = (char)tdo_getcode(single);
        } else {
            out[k++] = (char)ch;
        }
    }
    free(up);
    /* external gaps (:806-840) */
    int stopped = 0;
    for (int x = 0; x < k; x++) {
        if (out[x] == '_') return -3;
        else if (out[x] == '-') { stopped = x; out[x] = '_'; }
        else if (out[x] == '?') continue;
        else break;
    }
    for (int x = k - 1; x >= 0; x--) {
        if (out[x] == '_') {
            if (stopped == x) break;
            else return -3;
        } else if (out[x] == '-') out[x] = '_';
        else if (out[x] == '?') continue;
        else break;
    }
    /* doSanityChecks (:1727-1763) */
    for (int x = 0; x < k; x++) {
        if (out[x] == 'U' || out[x] == 'u') return -4;
        if (!tdo_is_valid(out[x])) return -5;
    }
    *out_len = k;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Per-pair counts                                                     */
/* ------------------------------------------------------------------ */

/* Sequence.getSharedLength, Common/DNA/Sequence.java:1414-1485; mode mirrors the
 * static pairwiseDistanceMethod the reference reads at :1446 / :1457 / :1472. */
int tdo_shared_length(const char *a, int la, const char *b, int lb, int mode) {
    int min = la < lb ? la : lb, count = 0;
    for (int x = 0; x < min; x++) {
        int ch1 = a[x], ch2 = b[x];
        if (ch1 == '?' || ch2 == '?') {
        } else if (is_gap(ch1) || is_gap(ch2)) {
            if ((is_internal_gap(ch1) && !is_internal_gap(ch2) && is_gap(ch2)) ||
                (!is_internal_gap(ch1) && is_internal_gap(ch2) && is_gap(ch1))) {
                /* one internal, one external gap: discarded (:1431-1444) */
            } else if (is_internal_gap(ch1) || is_internal_gap(ch2)) {
                if (mode == TDO_PDM_K2P) { /* ignore (:1446-1448) */ }
                else count++;
            } else {
            }
        } else {
            if (mode == TDO_PDM_K2P) {
                if (tdo_is_purine(ch1) && tdo_is_pyrimidine(ch1)) {
                } else if (tdo_is_purine(ch2) && tdo_is_pyrimidine(ch2)) {
                } else count++;
            } else if (mode == TDO_PDM_TRANS_ONLY) {
                if ((!tdo_is_purine(ch1) && !tdo_is_pyrimidine(ch1)) ||
                    (!tdo_is_purine(ch2) && !tdo_is_pyrimidine(ch2))) {
                } else count++;
            } else count++;
        }
    }
    return count;
}

/* Sequence.countIdentical, Common/DNA/Sequence.java:1322-1341 */
int tdo_count_identical(const char *a, int la, const char *b, int lb, int ambig_allowed) {
    int min = la < lb ? la : lb, count = 0;
    for (int x = 0; x < min; x++)
        if (tdo_identical(a[x], b[x], ambig_allowed)) count++;
    return count;
}

/* Sequence.countTransversions, Common/DNA/Sequence.java:1352-1371 */
int tdo_count_transversions(const char *a, int la, const char *b, int lb) {
    int min = la < lb ? la : lb, count = 0;
    for (int x = 0; x < min; x++) {
        int ch1 = a[x], ch2 = b[x];
        if ((tdo_is_purine(ch1) && tdo_is_pyrimidine(ch2)) ||
            (tdo_is_purine(ch2) && tdo_is_pyrimidine(ch1)))
            count++;
    }
    return count;
}

/* ------------------------------------------------------------------ */
/* Math.log                                                            */
/* ------------------------------------------------------------------ */

/* Natural logarithm as java.lang.StrictMath.log defines it: the fdlibm e_log.c
 * algorithm (Sun Microsystems, 1993; freely distributable), restated here from
 * its published description -- argument reduction x = 2^k (1+f), s = f/(2+f),
 * a degree-14 even polynomial in s, and a two-piece ln2.  Java's Math.log is
 * specified to be within 1 ulp of the true value (and is StrictMath.log unless
 * the JIT intrinsifies it), so this is a legitimate stand-in for
 * Sequence.java:1564's Math.log calls.  Every operation is a plain IEEE-754
 * binary64 +,-,*,/ (no FMA), so the CUDA kernel can reproduce it bit for bit. */
static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                    two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                    Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                    Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                    Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;

double tdo_log(double x) {
    union { double d; uint64_t u; } v;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k, hx, i, j;
    uint32_t lx;
    v.d = x;
    hx = (int32_t)(v.u >> 32);
    lx = (uint32_t)v.u;
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY; /* log(+-0) */
        if (hx < 0) return NAN;                              /* log(-#)  */
        k -= 54;
        x *= two54;
        v.d = x;
        hx = (int32_t)(v.u >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    v.u = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (v.u & 0xffffffffu);
    x = v.d;
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

static int g_use_libm_log = 0;
/* tests flip this to cross-check tdo_log against glibc's log */
void tdo_set_use_libm_log(int on) { g_use_libm_log = on; }
static double java_log(double x) { return g_use_libm_log ? log(x) : tdo_log(x); }

/* ------------------------------------------------------------------ */
/* Distances                                                           */
/* ------------------------------------------------------------------ */

/* Sequence.getK2PDistance, Common/DNA/Sequence.java:1498-1573.  Also returns
 * the three integer counts the loop accumulates (n, nTransitions, nTransversions). */
double tdo_k2p(const char *a, int la, const char *b, int lb, int *n_out, int *ts_out, int *tv_out) {
    int nTv = 0, nTs = 0, n = 0;
    int min = la < lb ? la : lb;
    for (int x = 0; x < min; x++) {
        int ch1 = a[x], ch2 = b[x];
        if (ch1 == '?' || ch2 == '?') {
        } else if (is_gap(ch1) || is_gap(ch2)) {
        } else {
            if (!tdo_is_purine(ch1) && !tdo_is_pyrimidine(ch1)) continue;
            else if (!tdo_is_purine(ch2) && !tdo_is_pyrimidine(ch2)) continue;
            else {
                n++;
                if (ch1 == ch2) continue;
                if ((tdo_is_purine(ch1) && tdo_is_purine(ch2)) ||
                    (tdo_is_pyrimidine(ch1) && tdo_is_pyrimidine(ch2)))
                    nTs++;
                else
                    nTv++;
            }
        }
    }
    if (n_out) *n_out = n;
    if (ts_out) *ts_out = nTs;
    if (tv_out) *tv_out = nTv;
    /* :1559-1560; Java evaluates left to right, int n is promoted to double */
    double w1 = 1.0 - (2.0 * ((double)nTs) / (double)n) - (((double)nTv) / (double)n);
    double w2 = 1.0 - (2.0 * ((double)nTv) / (double)n);
    double distance = (-0.5 * java_log(w1)) - (0.25 * java_log(w2)); /* :1564 */
    if (distance <= 0) return 0.0;                                    /* :1569-1570 */
    return distance;
}

/* Sequence.getPairwiseNoBuffer, Common/DNA/Sequence.java:1699-1717 */
double tdo_pairwise(const char *a, int la, const char *b, int lb, int mode, int min_overlap,
                    int ambig_allowed) {
    int shared = tdo_shared_length(a, la, b, lb, mode);
    if (shared >= min_overlap) {
        switch (mode) {
            case TDO_PDM_K2P:
                return tdo_k2p(a, la, b, lb, 0, 0, 0);
            case TDO_PDM_TRANS_ONLY:
                return ((double)tdo_count_transversions(a, la, b, lb)) / (double)shared;
            default:
                return 1.0 - ((double)tdo_count_identical(a, la, b, lb, ambig_allowed) / (double)shared);
        }
    }
    return -1.0;
}

/* ------------------------------------------------------------------ */
/* Settings                                                            */
/* ------------------------------------------------------------------ */

/* Java's (long) cast of a double (JLS 5.1.3): NaN -> 0, saturating. */
static int64_t java_d2l(double d) {
    if (d != d) return 0;
    if (d >= 9223372036854775807.0) return INT64_MAX;
    if (d <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)d;
}

/* Settings.makeLongFromDouble, Common/DNA/Settings.java:69-71 (accurateTo = 100000, :48) */
int64_t tdo_make_long(double d) { return java_d2l(d * 100000); }
/* Settings.makeDoubleFromLong, Common/DNA/Settings.java:79-81 */
double tdo_make_double(int64_t i) { return (double)i / (double)100000; }
/* Settings.roundOff, Common/DNA/Settings.java:90-92 */
double tdo_round_off(double d) { return tdo_make_double(tdo_make_long(d)); }
/* Settings.percentage, Common/DNA/Settings.java:84-87 */
double tdo_percentage(double x, double y) {
    if (y == 0) return 0;
    return (double)(java_d2l(tdo_round_off(x / y) * 100 * 100)) / 100;
}
/* Settings.identical, Common/DNA/Settings.java:95-97 */
int tdo_settings_identical(double d1, double d2) { return tdo_make_long(d1) == tdo_make_long(d2); }

/* ------------------------------------------------------------------ */
/* Batch drivers (tests + CPU baseline)                                */
/* ------------------------------------------------------------------ */
#include <pthread.h>

typedef struct {
    const char *rows; const int *lens; int n, stride, mode, min_overlap, ambig;
    int row_begin, row_end, upper_only;
    double *dist; int32_t *shared, *c1, *c2, *c3;
    /* work distribution: rows are claimed one at a time from a shared counter */
    volatile int *next_row;
    /* bench leg */
    int bench; double sum; int64_t cnt;
} tdo_job;

static void one_row(tdo_job *J, int i) {
    const char *a = J->rows + (size_t)i * J->stride;
    for (int j = J->upper_only ? i + 1 : 0; j < J->n; j++) {
        const char *b = J->rows + (size_t)j * J->stride;
        if (J->bench) {
            double d = tdo_pairwise(a, J->lens[i], b, J->lens[j], J->mode, J->min_overlap, J->ambig);
            if (d == d) J->sum += d;
            J->cnt++;
            continue;
        }
        size_t o = (size_t)(i - J->row_begin) * J->n + j;
        int sh = tdo_shared_length(a, J->lens[i], b, J->lens[j], J->mode);
        double d = -1.0;
        int x1 = 0, x2 = 0, x3 = 0;
        if (J->mode == TDO_PDM_K2P) {
            double dd = tdo_k2p(a, J->lens[i], b, J->lens[j], &x1, &x2, &x3);
            if (sh >= J->min_overlap) d = dd;
        } else if (J->mode == TDO_PDM_TRANS_ONLY) {
            x1 = tdo_count_transversions(a, J->lens[i], b, J->lens[j]);
            if (sh >= J->min_overlap) d = ((double)x1) / (double)sh;
        } else {
            x1 = tdo_count_identical(a, J->lens[i], b, J->lens[j], J->ambig);
            if (sh >= J->min_overlap) d = 1.0 - ((double)x1 / (double)sh);
        }
        if (J->dist) J->dist[o] = d;
        if (J->shared) J->shared[o] = sh;
        if (J->c1) J->c1[o] = x1;
        if (J->c2) J->c2[o] = x2;
        if (J->c3) J->c3[o] = x3;
    }
}

static void *worker(void *p) {
    tdo_job *J = (tdo_job *)p;
    for (;;) {
        int i = __sync_fetch_and_add(J->next_row, 1);
        if (i >= J->row_end) break;
        one_row(J, i);
    }
    return 0;
}

static void run_jobs(tdo_job *proto, int threads) {
    volatile int next = proto->row_begin;
    proto->next_row = &next;
    if (threads <= 1) { worker(proto); return; }
    if (threads > 256) threads = 256;
    pthread_t th[256];
    tdo_job jobs[256];
    for (int t = 0; t < threads; t++) { jobs[t] = *proto; pthread_create(&th[t], 0, worker, &jobs[t]); }
    for (int t = 0; t < threads; t++) { pthread_join(th[t], 0); proto->sum += jobs[t].sum; proto->cnt += jobs[t].cnt; }
}

/* All pairs (i, j) of n normalised rows (row i = rows + i*stride, length lens[i]);
 * the same two-pass-per-pair structure as getPairwiseNoBuffer, no cache.  Any
 * output pointer may be NULL; outputs are (row_end-row_begin) x n, row-major.
 * For mode K2P, c1 = n, c2 = nTs, c3 = nTv; for uncorrected c1 = ident; for
 * trans-only c1 = transversions.  Columns j > i only when upper_only != 0.
 * threads <= 1: serial, the way every reference module runs. */
void tdo_all_pairs(const char *rows, const int *lens, int n, int stride, int mode, int min_overlap,
                   int ambig_allowed, int row_begin, int row_end, int upper_only, int threads,
                   double *dist, int32_t *shared, int32_t *c1, int32_t *c2, int32_t *c3) {
    tdo_job J = {rows, lens, n, stride, mode, min_overlap, ambig_allowed, row_begin, row_end,
                 upper_only, dist, shared, c1, c2, c3, 0, 0, 0.0, 0};
    run_jobs(&J, threads);
}

/* CPU-baseline leg: getPairwiseNoBuffer over all unordered pairs i<j of rows
 * [row_begin,row_end) x all j>i, returning a checksum so the work cannot be
 * optimised away.  Counts how many pairs it evaluated in *pairs. */
double tdo_bench_pairs(const char *rows, const int *lens, int n, int stride, int mode, int min_overlap,
                       int ambig_allowed, int row_begin, int row_end, int threads, int64_t *pairs) {
    tdo_job J = {rows, lens, n, stride, mode, min_overlap, ambig_allowed, row_begin, row_end,
                 1, 0, 0, 0, 0, 0, 0, 1, 0.0, 0};
    run_jobs(&J, threads);
    if (pairs) *pairs = J.cnt;
    return J.sum;
}
