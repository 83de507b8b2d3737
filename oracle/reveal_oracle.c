/*
 * reveal_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's index-build + MUM-sweep path, used
 * ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as
 * the checker for the CUDA path.  It is never imported, linked or executed by
 * the product package (reveal_b200/), which fails loudly without its CUDA
 * library.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_*.py)
 * against the reference's own object code -- /root/reference/reveallib/{interface,
 * reveal}.c + /root/reference/divsufsort/ compiled unmodified into oracle/_ref/
 * by oracle/ref/Makefile -- and against golden vectors minted from that build
 * (tests/golden/, generator tests/golden/make_golden.py).  The reference's own
 * test-suite holds no golden vectors for this path.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/).  Index type is int32 (the reference's default saidx_t,
 * reveallib/reveal.h:11-12); the 64-bit build (reveal.h:8-9) computes the same
 * numbers in wider integers, so the Python wrapper widens on request.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t idx_t;

/* ------------------------------------------------------------------------- */
/* Suffix array.  Reference: divsufsort(T, SA, n), divsufsort/divsufsort.c:333,
 * called at reveallib/interface.c:216-218.  libdivsufsort 2.0.1 is vendored
 * third-party code; its output is THE suffix array of the byte string T
 * (plain lexicographic order over unsigned bytes, a suffix that is a proper
 * prefix of another sorts first), which is unique, so the oracle restates the
 * result with an independent linear-time construction (SA-IS, Nong/Zhang/Chan
 * 2009) rather than divsufsort's induced-copying internals.                  */

#define CHR(i) (cs == 1 ? (idx_t)((const uint8_t *)T)[i] : ((const idx_t *)T)[i])

static void bkt_count(const void *T, int cs, idx_t n, idx_t K, idx_t *cnt)
{
    idx_t i;
    memset(cnt, 0, (size_t)K * sizeof(idx_t));
    for (i = 0; i < n; i++) cnt[CHR(i)]++;
}
static void bkt_heads(const idx_t *cnt, idx_t K, idx_t *B)
{
    idx_t c, s = 0;
    for (c = 0; c < K; c++) { B[c] = s; s += cnt[c]; }
}
static void bkt_tails(const idx_t *cnt, idx_t K, idx_t *B)
{
    idx_t c, s = 0;
    for (c = 0; c < K; c++) { s += cnt[c]; B[c] = s; }
}

static void induce(const void *T, int cs, idx_t *SA, idx_t n, idx_t K, const uint8_t *isS, const idx_t *cnt, idx_t *B)
{
    idx_t i, j;
    /* L-type suffixes, left to right; the suffix n-1 precedes the virtual end-of-text */
    bkt_heads(cnt, K, B);
    j = n - 1;
    SA[B[CHR(j)]++] = j;
    for (i = 0; i < n; i++) {
        j = SA[i];
        if (j > 0 && !isS[j - 1]) { j--; SA[B[CHR(j)]++] = j; }
    }
    /* S-type suffixes, right to left */
    bkt_tails(cnt, K, B);
    for (i = n - 1; i >= 0; i--) {
        j = SA[i];
        if (j > 0 && isS[j - 1]) { j--; SA[--B[CHR(j)]] = j; }
    }
}

#define IS_LMS(i) ((i) > 0 && isS[i] && !isS[(i) - 1])

static int sais(const void *T, idx_t *SA, idx_t n, idx_t K, int cs)
{
    uint8_t *isS;
    idx_t *cnt, *B, *s1, *SA1;
    idx_t i, j, n1, name, prev;

    if (n == 0) return 0;
    if (n == 1) { SA[0] = 0; return 0; }
    isS = (uint8_t *)malloc((size_t)n);
    cnt = (idx_t *)malloc((size_t)K * sizeof(idx_t));
    B = (idx_t *)malloc((size_t)K * sizeof(idx_t));
    if (!isS || !cnt || !B) { free(isS); free(cnt); free(B); return -1; }

    isS[n - 1] = 0;
    for (i = n - 2; i >= 0; i--) {
        idx_t a = CHR(i), b = CHR(i + 1);
        isS[i] = (uint8_t)((a < b) || (a == b && isS[i + 1]));
    }
    bkt_count(T, cs, n, K, cnt);

    /* stage 1: sort the LMS substrings */
    for (i = 0; i < n; i++) SA[i] = -1;
    bkt_tails(cnt, K, B);
    for (i = 1; i < n; i++)
        if (IS_LMS(i)) SA[--B[CHR(i)]] = i;
    induce(T, cs, SA, n, K, isS, cnt, B);

    n1 = 0;
    for (i = 0; i < n; i++) {
        j = SA[i];
        if (IS_LMS(j)) SA[n1++] = j;
    }
    for (i = n1; i < n; i++) SA[i] = -1;
    name = 0;
    prev = -1;
    for (i = 0; i < n1; i++) {
        idx_t pos = SA[i], d;
        int diff = 0;
        if (prev < 0) diff = 1;
        for (d = 0; !diff; d++) {
            if (pos + d >= n || prev + d >= n) { diff = 1; break; }
            if (CHR(pos + d) != CHR(prev + d) || isS[pos + d] != isS[prev + d]) { diff = 1; break; }
            if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
        }
        if (diff) { name++; prev = pos; }
        SA[n1 + (pos >> 1)] = name - 1;
    }
    for (i = n - 1, j = n - 1; i >= n1; i--)
        if (SA[i] >= 0) SA[j--] = SA[i];

    /* stage 2: order the LMS suffixes (recursively if names collide) */
    SA1 = SA;
    s1 = SA + n - n1;
    if (name < n1) {
        if (sais(s1, SA1, n1, name, 4) != 0) { free(isS); free(cnt); free(B); return -1; }
    } else {
        for (i = 0; i < n1; i++) SA1[s1[i]] = i;
    }

    /* stage 3: induce the full order from the sorted LMS suffixes */
    for (i = 1, j = 0; i < n; i++)
        if (IS_LMS(i)) s1[j++] = i;
    for (i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
    for (i = n1; i < n; i++) SA[i] = -1;
    bkt_tails(cnt, K, B);
    for (i = n1 - 1; i >= 0; i--) {
        j = SA[i];
        SA[i] = -1;
        SA[--B[CHR(j)]] = j;
    }
    induce(T, cs, SA, n, K, isS, cnt, B);

    free(isS); free(cnt); free(B);
    return 0;
}

/* interface.c:216-218 (divsufsort call). returns 0 on success */
int orc_suffix_array(const uint8_t *T, int64_t n, int32_t *SA)
{
    if (n < 0 || n > 0x7fffffffLL) return -1;
    return sais(T, SA, (idx_t)n, 256, 1);
}

/* interface.c:235-238: SAi[SA[i]] = i */
void orc_inverse(const int32_t *SA, int64_t n, int32_t *SAi)
{
    int64_t r;
    for (r = 0; r < n; r++) SAi[SA[r]] = (int32_t)r;
}

/* interface.c:97-114 compute_lcp: Kasai et al. with the reference's barrier
 * rule: the extension loop stops when the two characters differ OR the
 * character of the text-order suffix is '$' or 'N' (interface.c:107), so a
 * match never spans a sentinel or an N.  T must be readable at T[n] (the
 * reference keeps a NUL there, interface.c:84); we pass n and test bounds.   */
void orc_compute_lcp(const uint8_t *T, const int32_t *SA, const int32_t *SAi, int32_t *LCP, int64_t n)
{
    int64_t i, h = 0;
    for (i = 0; i < n; i++) {
        int64_t r = SAi[i];
        if (r == 0) {
            LCP[0] = 0;
        } else {
            int64_t j = SA[r - 1];
            while (i + h < n && j + h < n) {
                uint8_t c = T[i + h];
                if (c != T[j + h] || c == '$' || c == 'N') break;
                h++;
            }
            LCP[r] = (int32_t)h;
        }
        if (h > 0) h--;
    }
}

/* interface.c:116-134 build_SO: sample id of every text position; nsep[k] is
 * the position of the last '$' of sample k (interface.c:36-43).             */
void orc_build_so(const int64_t *nsep, int32_t nsamples, int64_t n, uint16_t *SO)
{
    int32_t s;
    int64_t p = 0;
    for (s = 0; s < nsamples; s++) {
        int64_t last = (s == nsamples - 1) ? n - 1 : nsep[s];
        for (; p <= last && p < n; p++) SO[p] = (uint16_t)s;
    }
}

/* interface.c:136-145 comp_tab: IUPAC-aware complement, identity below '@'.
 * Restated as a rule table: pairs that swap, everything else maps to itself,
 * with the reference's quirks kept: 'U'/'u' -> 'A'/'a' and '`' (96) -> '@'.  */
static uint8_t comp_of(uint8_t c)
{
    static const char pairs[] = "ATCGBVDHKMRY"; /* A<->T C<->G B<->V D<->H K<->M R<->Y ; S,W,N,E..Z fixed */
    int k;
    if (c == 96) return 64;
    if (c == 'U') return 'A';
    if (c == 'u') return 'a';
    for (k = 0; k < 12; k++) {
        if (c == (uint8_t)pairs[k]) return (uint8_t)pairs[k ^ 1];
        if (c == (uint8_t)(pairs[k] + 32)) return (uint8_t)(pairs[k ^ 1] + 32);
    }
    return c;
}
void orc_comp_table(uint8_t *tab128)
{
    int c;
    for (c = 0; c < 128; c++) tab128[c] = comp_of((uint8_t)c);
}

/* interface.c:148-158 revcomp + :168-172: in-place reverse complement of
 * T[nsep0 .. n) (note: starts AT the sentinel that ends sample 0).           */
void orc_revcomp(uint8_t *T, int64_t n)
{
    int64_t i;
    for (i = 0; i < n / 2; i++) {
        uint8_t a = comp_of(T[i]), b = comp_of(T[n - 1 - i]);
        T[i] = b;
        T[n - 1 - i] = a;
    }
    if (n & 1) T[n / 2] = comp_of(T[n / 2]);
}

static int is_lower(uint8_t c) { return c >= 'a' && c <= 'z'; }

/* left-maximality test shared by the sweeps: reveal.c:81-85, :145-149, :246-256 */
static int left_differs(const uint8_t *T, int64_t a, int64_t b)
{
    uint8_t ca, cb;
    if (a == 0 || b == 0) return 1;
    ca = T[a - 1];
    cb = T[b - 1];
    return ca != cb || ca == 'N' || ca == '$' || is_lower(ca);
}

/* reveal.c:55-116 getmums (flavour 0) and reveal.c:119-180 getmums_rem
 * (flavour 1; differs only in the rc remap using index->n instead of nT).
 * n = entries in SA/LCP, nT = length of the original text.  Output rows
 * (l, a, b) in ascending SA rank.  Returns the number of MUMs (rows beyond
 * `cap` are counted but not stored).                                         */
int64_t orc_getmums(const uint8_t *T, const int32_t *SA, const int32_t *LCP, int64_t n, int64_t nT,
                    int64_t nsep0, int32_t rc, int32_t minl, int32_t flavour, int64_t *out, int64_t cap)
{
    int64_t i, cnt = 0;
    for (i = 1; i < n; i++) {
        int64_t l = LCP[i], a, b, before, after;
        if (l < minl) continue;
        if ((SA[i] > nsep0) == (SA[i - 1] > nsep0)) continue; /* same sample: repeat */
        a = SA[i] < SA[i - 1] ? SA[i] : SA[i - 1];
        b = SA[i] < SA[i - 1] ? SA[i - 1] : SA[i];
        if (!left_differs(T, a, b)) continue;
        before = LCP[i - 1];
        after = (i == n - 1) ? 0 : LCP[i + 1];
        if (before >= l || after >= l) continue; /* not unique */
        if (rc == 1) b = nsep0 + ((flavour ? n : nT) - b - l);
        if (cnt < cap) {
            out[3 * cnt + 0] = l;
            out[3 * cnt + 1] = a;
            out[3 * cnt + 2] = b;
        }
        cnt++;
    }
    return cnt;
}

/* sample of a text position: SO[pos] when SO is given (main nsamples > 2,
 * interface.c:265-271), else the two-sample side test of reveal.c:232-235.   */
static int32_t sample_of(const uint16_t *SO, int64_t nsep0, int64_t pos)
{
    return SO ? (int32_t)SO[pos] : (pos > nsep0 ? 1 : 0);
}

/* reveal.c:227-259 ismultimum / reveal.c:261-290 ismultimem (mem != 0).
 * flag must hold main_nsamples ints.  For mem the per-sample counts stay in
 * flag for the caller (reveal.c:334-342).                                    */
static int interval_ok(const uint8_t *T, const int32_t *SA, const uint16_t *SO, int64_t nsep0, int32_t main_nsamples,
                       int64_t l, int64_t lb, int64_t ub, int32_t *flag, int mem)
{
    int64_t j;
    if (l <= 0) return 0;
    memset(flag, 0, (size_t)main_nsamples * sizeof(int32_t));
    if (main_nsamples == 2) {
        int same = ((SA[ub] > nsep0) == (SA[lb] > nsep0));
        if (mem) flag[same]++;          /* reveal.c:267 (sic) */
        else if (same) return 0;        /* reveal.c:233-235 */
    } else {
        for (j = lb; j <= ub; j++) {
            int32_t s = (int32_t)SO[SA[j]];
            if (!mem && flag[s]) return 0; /* a sample twice: not unique */
            flag[s]++;
        }
    }
    for (j = lb; j < ub; j++)
        if (left_differs(T, SA[j], SA[j + 1])) return 1;
    return 0;
}

/* reveal.c:436-580 getmultimums (mem == 0) and reveal.c:292-434 getmultimems
 * (mem != 0): bottom-up enumeration of lcp-intervals (Abouelhoda et al. 2004)
 * with an explicit stack; an interval [lb,ub] with lcp value l is reported on
 * pop if l >= minl, size >= minn (and size <= main_nsamples for mums) and the
 * predicate above holds.  The final flush (reveal.c:538-574) closes whatever
 * is still open at ub = n-1.
 * Output (CSR): hdr rows (l, count_field, first_member) and member rows
 * (sample, pos) in SA-rank order, list order = pop order.  count_field is the
 * interval size for mums (reveal.c:497) and the number of distinct samples
 * for mems (reveal.c:353).  Returns #records via *nrec, #members via *nmem;
 * rows beyond the caps are counted but not stored.  Return value 0 / -1 (oom).
 * SO may be NULL only when main_nsamples == 2: the reference itself would
 * dereference a NULL SO there (reveal.c:489), so that sample column is
 * defined by the side test instead -- flagged in DESIGN.md.                  */
int orc_getmulti(const uint8_t *T, const int32_t *SA, const int32_t *LCP, const uint16_t *SO, int64_t n,
                 int64_t nsep0, int32_t main_nsamples, int32_t minl, int32_t minn, int32_t mem,
                 int64_t *hdr, int64_t hdr_cap, int64_t *members, int64_t mem_cap, int64_t *nrec, int64_t *nmem)
{
    int64_t cap = 1024, depth = 0, i, nr = 0, nm = 0;
    int64_t *st_l = (int64_t *)malloc(cap * sizeof(int64_t));
    int64_t *st_lb = (int64_t *)malloc(cap * sizeof(int64_t));
    int32_t *flag = (int32_t *)calloc((size_t)(main_nsamples > 2 ? main_nsamples : 2), sizeof(int32_t));
    if (!st_l || !st_lb || !flag) { free(st_l); free(st_lb); free(flag); return -1; }
    st_l[0] = 0;
    st_lb[0] = 0;
    for (i = 1; i <= n; i++) {
        /* i == n plays the role of the final flush: everything open is closed at n-1 */
        int64_t cur = (i < n) ? (int64_t)LCP[i] : -1, lb = i - 1;
        while (depth >= 0 && (i == n || cur < st_l[depth])) {
            int64_t l = st_l[depth], ilb = st_lb[depth], iub = i - 1, size = iub - ilb + 1;
            int keep_lb = 0;
            depth--;
            if (l >= minl && size >= minn && (mem || size <= main_nsamples) &&
                interval_ok(T, SA, SO, nsep0, main_nsamples, l, ilb, iub, flag, mem)) {
                int64_t field = size, x;
                int emit = 1;
                if (mem) {
                    int32_t s;
                    field = 0;
                    for (s = 0; s < main_nsamples; s++) field += flag[s] > 0;
                    /* reveal.c:340-342: this `continue` also skips the `lb = i_lb`
                     * hand-over at reveal.c:362, which the restatement keeps */
                    if (field < minn) { emit = 0; keep_lb = 1; }
                }
                if (emit) {
                    if (nr < hdr_cap) { hdr[3 * nr] = l; hdr[3 * nr + 1] = field; hdr[3 * nr + 2] = nm; }
                    for (x = 0; x < size; x++) {
                        if (nm < mem_cap) {
                            members[2 * nm] = sample_of(SO, nsep0, SA[ilb + x]);
                            members[2 * nm + 1] = SA[ilb + x];
                        }
                        nm++;
                    }
                    nr++;
                }
            }
            if (!keep_lb) lb = ilb;
        }
        if (i < n && cur > st_l[depth]) {
            depth++;
            if (depth >= cap) {
                cap *= 2;
                st_l = (int64_t *)realloc(st_l, cap * sizeof(int64_t));
                st_lb = (int64_t *)realloc(st_lb, cap * sizeof(int64_t));
                if (!st_l || !st_lb) { free(st_l); free(st_lb); free(flag); return -1; }
            }
            st_l[depth] = cur;
            st_lb[depth] = lb;
        }
    }
    free(st_l); free(st_lb); free(flag);
    *nrec = nr;
    *nmem = nm;
    return 0;
}
