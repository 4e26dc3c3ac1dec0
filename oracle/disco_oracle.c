/*
 * disco_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
 *
 * Plain-C, ASCII-string restatement of the reference BuildGraph hot path (abiswas-odu/Disco,
 * src/BuildGraph/src).  It deliberately shares nothing with the CUDA implementation: no bit packing,
 * no fingerprints -- strings, memcmp and a sorted record index, like the reference.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the reference's only golden vector (src/BuildGraph/bench_test_0_parGraph.txt),
 *   - outputs of the real reference binary (oracle/_ref/buildG, built by oracle/build_ref.sh) on the two in-tree
 *     fixtures and on seeded synthetic read sets; the vectors are committed under tests/golden/.
 *
 * Which reference code each function follows:
 *   oracle_test_read        Dataset.cpp:403-452 (testRead) + Common.h:171-181 (countSubstring)
 *   index build / lookup    HashTable.cpp:423-514 (two records per read, file order), :521-571 (getListOfReads typing)
 *   oracle_contained        OverlapGraph.cpp:333-505 (markContainedReads, sequential -t 1 semantics), :517-554
 *   oracle_edges            OverlapGraph.cpp:631-678 (insertAllEdgesOfRead, per-read capped search), :567-595 (checkOverlap),
 *                           :600-626 (twin edge), :770-784 (twinEdgeOrientation)
 *   oracle_reduce           OverlapGraph.cpp:687-723 (markTransitiveEdges), :731-761 (removeTransitiveEdges),
 *                           :808 (canonical src<dst on output)
 * Canonical (order-free) choices are the ones SURVEY.md App. A.6 defines; they coincide with the reference
 * whenever cap_fired == multi_overlap_pairs == one_sided_edges == 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAX_EDGE_PER_KMER 4 /* Common.h:62 */

/* ------------------------------------------------------------------ read filter (Dataset.cpp:403-452) */
static const char *FILTER_STRINGS[] = {
    "ACACACACACACACACACACACACACACA", "AGAGAGAGAGAGAGAGAGAGAGAGAGAGA", "ATATATATATATATATATATATATATATA",
    "CGCGCGCGCGCGCGCGCGCGCGCGCGCGC", "CTCTCTCTCTCTCTCTCTCTCTCTCTCTC", "AAGAAGAAGAAGAAGAAGAAGAAGAAGAA",
    "ATAATAATAATAATAATAATAATAATAAT", "TAATAATAATAATAATAATAATAATAATA", "AACAACAACAACAACAACAACAACAACAA",
    "ACAACAACAACAACAACAACAACAACAAC", "CAACAACAACAACAACAACAACAACAACA", "AAGAAGAAGAAGAAGAAGAAGAAGAAGAA",
    "AGAAGAAGAAGAAGAAGAAGAAGAAGAAG", "GAAGAAGAAGAAGAAGAAGAAGAAGAAGA", "TTCTTCTTCTTCTTCTTCTTCTTCTTCTT",
    "AAATAAATAAATAAATAAATAAATAAATA", "TAAATAAATAAATAAATAAATAAATAAAT", "ATAAATAAATAAATAAATAAATAAATAAA",
    "AATAAATAAATAAATAAATAAATAAATAA", "AATTAATTAATTAATTAATTAATTAATTA", "ATTAATTAATTAATTAATTAATTAATTAA",
    "TTAATTAATTAATTAATTAATTAATTAAT", "TAATTAATTAATTAATTAATTAATTAATT", "AAAGAAAGAAAGAAAGAAAGAAAGAAAGA",
    "AAAGAAAGAAAGAAAGAAAGAAAGAAAGA", "AGAAAGAAAGAAAGAAAGAAAGAAAGAAA", "GAAAGAAAGAAAGAAAGAAAGAAAGAAAG",
    "TACATACATACATACATACATACATACAT", "ACATACATACATACATACATACATACATA", "CATACATACATACATACATACATACATAC",
    "ATACATACATACATACATACATACATACA", "GTTTGTTTGTTTGTTTGTTTGTTTGTTTG", "TGTTTGTTTGTTTGTTTGTTTGTTTGTTT",
    "TTTGTTTGTTTGTTTGTTTGTTTGTTTGT", "AGGGAGGGAGGGAGGGAGGGAGGGAGGGA", "GAGGGAGGGAGGGAGGGAGGGAGGGAGGG",
    "GGAGGGAGGGAGGGAGGGAGGGAGGGAGG", "GGGAGGGAGGGAGGGAGGGAGGGAGGGAG"};
static const char *MER_STRINGS[] = {"AC", "AG", "AT", "CG", "CT", "GT", "AAT", "ATA", "TAA", "AAC",
                                    "ACA", "CAA", "AAG", "AGA", "GAA", "GGGGCC"};

static uint64_t count_substring(const char *s, uint64_t n, const char *sub)
{ /* Common.h:171-181: non-overlapping, left to right */
    uint64_t m = strlen(sub), cnt = 0, i = 0;
    if (m == 0 || n < m) return 0;
    while (i + m <= n) {
        if (memcmp(s + i, sub, m) == 0) { cnt++; i += m; }
        else i++;
    }
    return cnt;
}

/* read must already be upper-cased; returns 1 = good.  The caller applies "length > minOverlap" (Dataset.cpp:305). */
int oracle_test_read(const char *s, uint64_t n)
{
    uint64_t cnt[4] = {0, 0, 0, 0};
    if (n < 30) return 0; /* MIN_READ_SIZE, Dataset.h:15 */
    for (uint64_t i = 0; i < n; i++) {
        char c = s[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return 0;
        cnt[(c >> 1) & 3]++;
    }
    uint64_t thr = (uint64_t)((double)n * .7);
    if (cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) return 0;
    for (size_t i = 0; i < sizeof(FILTER_STRINGS) / sizeof(FILTER_STRINGS[0]); i++) {
        uint64_t len = strlen(FILTER_STRINGS[i]);
        if (n < len) return 0;
        if (memcmp(FILTER_STRINGS[i], s, len) == 0) return 0;
        if (memcmp(FILTER_STRINGS[i], s + n - len, len) == 0) return 0;
    }
    thr = (uint64_t)((double)n * .5);
    for (size_t i = 0; i < sizeof(MER_STRINGS) / sizeof(MER_STRINGS[0]); i++) {
        uint64_t rep = count_substring(s, n, MER_STRINGS[i]) * strlen(MER_STRINGS[i]);
        if (rep >= thr) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ context / index */
typedef struct {
    uint64_t n;         /* reads, ids 1..n (0-based r internally) */
    uint32_t K;         /* hashStringLength = minOverlap - 1 (HashTable.cpp:50) */
    const char *bases;  /* concatenated upper-case reads (borrowed) */
    const uint64_t *off;/* n+1 offsets (borrowed) */
    char *rc;           /* reverse complements, same offsets */
    uint64_t nrec;      /* 2n records: rec = 2*r + kind, kind 0 = prefix, 1 = suffix (file order) */
    char *canon;        /* nrec * K canonical k-mers */
    uint64_t *order;    /* records sorted by (canonical k-mer, rec) */
    uint64_t nruns;
    uint64_t *run_start;/* nruns+1 */
    uint64_t hsize;
    int64_t *hslot;     /* open addressing: run index or -1 */
    uint64_t *super_read; /* n: 0 = not contained else 1-based container id */
} octx;

static inline uint64_t rlen(const octx *c, uint64_t r) { return c->off[r + 1] - c->off[r]; }
static inline const char *rfwd(const octx *c, uint64_t r) { return c->bases + c->off[r]; }
static inline const char *rrev(const octx *c, uint64_t r) { return c->rc + c->off[r]; }

static char comp(char b) { return b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : 'A'; }
static void revcomp(const char *s, uint64_t n, char *out)
{
    for (uint64_t i = 0; i < n; i++) out[n - 1 - i] = comp(s[i]);
}

static uint64_t fnv(const char *s, uint32_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (uint32_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    return h;
}

static octx *g_sort_ctx;
static int cmp_rec(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    int c = memcmp(g_sort_ctx->canon + x * g_sort_ctx->K, g_sort_ctx->canon + y * g_sort_ctx->K, g_sort_ctx->K);
    if (c) return c;
    return x < y ? -1 : (x > y);
}

/* canonical form of a k-mer: min(kmer, rc(kmer)) -- plays the role of getHashIndex's min() (HashTable.cpp:383-391):
 * it only groups records; equality with the forward / reverse string decides the type. */
static void canonical(const char *kmer, uint32_t K, char *out, char *tmp)
{
    revcomp(kmer, K, tmp);
    memcpy(out, memcmp(kmer, tmp, K) <= 0 ? kmer : tmp, K);
}

void oracle_free(octx *c)
{
    if (!c) return;
    free(c->rc); free(c->canon); free(c->order); free(c->run_start); free(c->hslot); free(c->super_read);
    free(c);
}

octx *oracle_create(const char *bases, const uint64_t *off, uint64_t n, uint32_t min_overlap)
{
    octx *c = (octx *)calloc(1, sizeof(octx));
    c->n = n; c->K = min_overlap - 1; c->bases = bases; c->off = off;
    uint32_t K = c->K;
    c->rc = (char *)malloc(off[n] + 1);
    for (uint64_t r = 0; r < n; r++) revcomp(rfwd(c, r), rlen(c, r), c->rc + off[r]);
    c->nrec = 2 * n;
    c->canon = (char *)malloc(c->nrec * (uint64_t)K + 1);
    c->order = (uint64_t *)malloc(c->nrec * sizeof(uint64_t));
    char *tmp = (char *)malloc(K + 1);
    for (uint64_t r = 0; r < n; r++) {
        uint64_t L = rlen(c, r);
        canonical(rfwd(c, r), K, c->canon + (2 * r) * K, tmp);             /* prefix record (HashTable.cpp:430,451) */
        canonical(rfwd(c, r) + L - K, K, c->canon + (2 * r + 1) * K, tmp); /* suffix record (HashTable.cpp:431,486) */
        c->order[2 * r] = 2 * r; c->order[2 * r + 1] = 2 * r + 1;
    }
    free(tmp);
    g_sort_ctx = c;
    qsort(c->order, c->nrec, sizeof(uint64_t), cmp_rec);
    c->run_start = (uint64_t *)malloc((c->nrec + 1) * sizeof(uint64_t));
    c->nruns = 0;
    for (uint64_t i = 0; i < c->nrec; i++)
        if (i == 0 || memcmp(c->canon + c->order[i] * K, c->canon + c->order[i - 1] * K, K) != 0)
            c->run_start[c->nruns++] = i;
    c->run_start[c->nruns] = c->nrec;
    c->hsize = 16; while (c->hsize < 3 * c->nruns + 8) c->hsize <<= 1;
    c->hslot = (int64_t *)malloc(c->hsize * sizeof(int64_t));
    for (uint64_t i = 0; i < c->hsize; i++) c->hslot[i] = -1;
    for (uint64_t q = 0; q < c->nruns; q++) {
        uint64_t h = fnv(c->canon + c->order[c->run_start[q]] * K, K) & (c->hsize - 1);
        while (c->hslot[h] >= 0) h = (h + 1) & (c->hsize - 1);
        c->hslot[h] = (int64_t)q;
    }
    c->super_read = (uint64_t *)calloc(n ? n : 1, sizeof(uint64_t));
    return c;
}

/* getListOfReads (HashTable.cpp:521-571): candidates in record (= file) order, typed 0..3.
 * Returns the run [*lo,*hi) in c->order; the type of each record is computed by cand_type(). */
static int lookup(const octx *c, const char *q, char *canon_tmp, char *tmp, uint64_t *lo, uint64_t *hi)
{
    uint32_t K = c->K;
    canonical(q, K, canon_tmp, tmp);
    uint64_t h = fnv(canon_tmp, K) & (c->hsize - 1);
    while (c->hslot[h] >= 0) {
        uint64_t run = (uint64_t)c->hslot[h];
        if (memcmp(c->canon + c->order[c->run_start[run]] * K, canon_tmp, K) == 0) {
            *lo = c->run_start[run]; *hi = c->run_start[run + 1];
            return 1;
        }
        h = (h + 1) & (c->hsize - 1);
    }
    return 0;
}

/* type of record `rec` for query k-mer q: prefix record: q==x -> 0 else 3; suffix record: q==x -> 1 else 2
 * ("if ... else if", HashTable.cpp:539-565: a reverse-palindromic k-mer yields only the forward type). */
static int cand_type(const octx *c, uint64_t rec, const char *q)
{
    uint64_t r = rec >> 1;
    uint32_t K = c->K;
    if ((rec & 1) == 0) return memcmp(q, rfwd(c, r), K) == 0 ? 0 : 3;
    return memcmp(q, rfwd(c, r) + rlen(c, r) - K, K) == 0 ? 1 : 2;
}

/* checkOverlapForContainedRead (OverlapGraph.cpp:517-554) */
static int check_contained(const octx *c, uint64_t r1, uint64_t r2, int type, uint64_t j)
{
    const char *s1 = rfwd(c, r1);
    uint64_t L1 = rlen(c, r1), L2 = rlen(c, r2), K = c->K;
    const char *t = (type == 0 || type == 1) ? rfwd(c, r2) : rrev(c, r2);
    if (type == 0 || type == 2) {
        uint64_t rem1 = L1 - j - K, rem2 = L2 - K;
        if (rem1 >= rem2) return memcmp(s1 + j + K, t + K, rem2) == 0;
    } else {
        uint64_t rem1 = j, rem2 = L2 - K;
        if (rem1 >= rem2) return memcmp(s1 + j - rem2, t, rem2) == 0;
    }
    return 0;
}

/* checkOverlap (OverlapGraph.cpp:567-595) */
static int check_overlap(const octx *c, uint64_t r1, uint64_t r2, int type, uint64_t j)
{
    const char *s1 = rfwd(c, r1);
    uint64_t L1 = rlen(c, r1), L2 = rlen(c, r2), K = c->K;
    const char *t = (type == 0 || type == 1) ? rfwd(c, r2) : rrev(c, r2);
    if (type == 0 || type == 2) {
        if (L1 - j - K >= L2 - K) return 0;
        return memcmp(s1 + j + K, t + K, L1 - (j + K)) == 0;
    } else {
        if (L2 - K < j) return 0;
        return memcmp(s1, t + L2 - K - j, j) == 0;
    }
}

static void type_to_edge(int type, uint64_t L1, uint64_t K, uint64_t j, int *orient, uint64_t *ovl)
{ /* OverlapGraph.cpp:428-434 / 660-666 */
    switch (type) {
    case 0: *orient = 3; *ovl = L1 - j; break;
    case 1: *orient = 0; *ovl = K + j; break;
    case 2: *orient = 2; *ovl = L1 - j; break;
    default: *orient = 1; *ovl = K + j; break;
    }
}

/* ------------------------------------------------------------------ contained reads (sequential -t 1 semantics) */
typedef struct {
    uint64_t contained; /* 1-based read ids */
    uint64_t container;
    uint32_t orient, len2, len1, start; /* start = L1 - ovl (OverlapGraph.cpp:445) */
} oracle_crow;

/* Fills c->super_read and returns the rows in emission order (i ascending, j ascending, candidate order). */
uint64_t oracle_contained(octx *c, oracle_crow **rows_out)
{
    uint32_t K = c->K;
    uint64_t cap = 1024, nrows = 0;
    oracle_crow *rows = (oracle_crow *)malloc(cap * sizeof(oracle_crow));
    char *ct = (char *)malloc(K + 1), *tmp = (char *)malloc(K + 1);
    memset(c->super_read, 0, c->n * sizeof(uint64_t));
    for (uint64_t i = 0; i < c->n; i++) {
        if (c->super_read[i] != 0) continue;                 /* OverlapGraph.cpp:395 */
        uint64_t L1 = rlen(c, i);
        const char *s1 = rfwd(c, i);
        for (uint64_t j = 0; j < L1 - K; j++) {             /* :401 (last position excluded) */
            uint64_t lo, hi;
            if (!lookup(c, s1 + j, ct, tmp, &lo, &hi)) continue;
            for (uint64_t k = lo; k < hi; k++) {
                uint64_t rec = c->order[k], r2 = rec >> 1;
                if (c->super_read[r2] != 0) continue;        /* HashTable.cpp:533, OverlapGraph.cpp:417 */
                int type = cand_type(c, rec, s1 + j);
                if (r2 == i || !check_contained(c, i, r2, type, j)) continue; /* :421 */
                uint64_t L2 = rlen(c, r2);
                if (L1 > L2 || (L1 == L2 && i < r2)) {       /* :424, :449 */
                    int orient; uint64_t ovl;
                    type_to_edge(type, L1, K, j, &orient, &ovl);
                    c->super_read[r2] = i + 1;               /* :435-436 */
                    if (nrows == cap) { cap *= 2; rows = (oracle_crow *)realloc(rows, cap * sizeof(oracle_crow)); }
                    oracle_crow *w = &rows[nrows++];
                    w->contained = r2 + 1; w->container = i + 1; w->orient = (uint32_t)orient;
                    w->len2 = (uint32_t)L2; w->len1 = (uint32_t)L1; w->start = (uint32_t)(L1 - ovl);
                }
            }
        }
    }
    free(ct); free(tmp);
    *rows_out = rows;
    return nrows;
}

const uint64_t *oracle_super_read(const octx *c) { return c->super_read; }

/* ------------------------------------------------------------------ per-read capped dovetail search */
typedef struct {
    uint64_t src, dst;        /* 1-based */
    uint32_t orient, offset;  /* offset = L_src - ovl (OverlapGraph.cpp:667) */
} oracle_edge;

typedef struct {
    uint64_t cap_fired;           /* (read, j) positions with more than MAX_EDGE_PER_KMER insertable candidates */
    uint64_t multi_overlap_pairs; /* pairs whose two endpoints found different overlaps */
    uint64_t one_sided_edges;     /* pairs found from one endpoint only */
    uint64_t raw_directed;        /* edges found by the per-read searches (both directions counted) */
    uint64_t lookups, candidates; /* getListOfReads calls and records visited */
} oracle_stats;

static int twin_orient(int o) { return o == 0 ? 3 : o == 3 ? 0 : o; } /* OverlapGraph.cpp:770-784 */

static int cmp_edge_key(const void *a, const void *b)
{
    const oracle_edge *x = (const oracle_edge *)a, *y = (const oracle_edge *)b;
    if (x->src != y->src) return x->src < y->src ? -1 : 1;
    if (x->offset != y->offset) return x->offset < y->offset ? -1 : 1;
    if (x->dst != y->dst) return x->dst < y->dst ? -1 : 1;
    if (x->orient != y->orient) return x->orient < y->orient ? -1 : 1;
    return 0;
}
static int cmp_edge_pair(const void *a, const void *b)
{
    const oracle_edge *x = (const oracle_edge *)a, *y = (const oracle_edge *)b;
    if (x->src != y->src) return x->src < y->src ? -1 : 1;
    if (x->dst != y->dst) return x->dst < y->dst ? -1 : 1;
    return 0;
}

/* Raw finds: for every non-contained read r1 its own insertAllEdgesOfRead() without the explored-skip
 * (SURVEY App. A.6 canonical choice).  Output sorted by (src, offset, dst, orient).  Requires oracle_contained(). */
uint64_t oracle_raw_edges(octx *c, oracle_edge **out, oracle_stats *st)
{
    uint32_t K = c->K;
    uint64_t cap = 4096, ne = 0;
    oracle_edge *e = (oracle_edge *)malloc(cap * sizeof(oracle_edge));
    char *ct = (char *)malloc(K + 1), *tmp = (char *)malloc(K + 1);
    memset(st, 0, sizeof(*st));
    for (uint64_t r1 = 0; r1 < c->n; r1++) {
        if (c->super_read[r1] != 0) continue;                /* :657 */
        uint64_t L1 = rlen(c, r1), first = ne;
        const char *s1 = rfwd(c, r1);
        for (uint64_t j = 1; j < L1 - K; j++) {             /* :638 */
            uint64_t lo, hi;
            st->lookups++;
            if (!lookup(c, s1 + j, ct, tmp, &lo, &hi)) continue;
            int ctr = 0, fired = 0;
            for (uint64_t k = lo; k < hi; k++) {
                uint64_t rec = c->order[k], r2 = rec >> 1;
                st->candidates++;
                if (c->super_read[r2] != 0) continue;        /* HashTable.cpp:533 */
                if (r2 == r1) continue;                      /* :655 */
                int dup = 0;
                for (uint64_t q = first; q < ne; q++) if (e[q].dst == r2 + 1) { dup = 1; break; } /* :656 */
                if (dup) continue;
                int type = cand_type(c, rec, s1 + j);
                if (!check_overlap(c, r1, r2, type, j)) continue;
                if (ctr >= MAX_EDGE_PER_KMER) { fired = 1; continue; } /* :645 */
                int orient; uint64_t ovl;
                type_to_edge(type, L1, K, j, &orient, &ovl);
                if (ne == cap) { cap *= 2; e = (oracle_edge *)realloc(e, cap * sizeof(oracle_edge)); }
                e[ne].src = r1 + 1; e[ne].dst = r2 + 1; e[ne].orient = (uint32_t)orient; e[ne].offset = (uint32_t)(L1 - ovl);
                ne++; ctr++;
            }
            st->cap_fired += fired;
        }
    }
    free(ct); free(tmp);
    qsort(e, ne, sizeof(oracle_edge), cmp_edge_key);
    st->raw_directed = ne;
    *out = e;
    return ne;
}

/* ------------------------------------------------------------------ union over endpoints + transitive reduction */
/* in: raw directed finds (every non-contained read's own capped search, sorted by (src, offset, dst, orient)).
 * out: kept undirected edges (src<dst, from src's perspective) sorted by (src,dst).
 *
 * Canonical, order-free semantics:
 *   - the row of a read is what its OWN search found (OverlapGraph.cpp:631-678 without the explored-skip);
 *   - markTransitiveEdges (:687-723) runs for every node on those rows: neighbours in (offset, id, orientation) order,
 *     a still-INPLAY neighbour v eliminates every neighbour w of u that v's row reaches with a chaining orientation;
 *   - the undirected edge {a<b} exists when either endpoint found it, is removed when it was eliminated in any row that
 *     holds it (edge and twin are flagged together, :717-718), and is written with a's overlap when a found it, else with
 *     the twin of b's (:808, :617).
 * Whenever cap_fired == multi_overlap_pairs == one_sided_edges == 0 every pair is found from both ends with the same
 * overlap, the rows are symmetric and this is exactly the reference's graph (pinned by the goldens).  Where the cap fires
 * (or a k-mer is a reverse palindrome) the reference's own result depends on its thread schedule; there this definition is
 * the deterministic stand-in, and the GPU path is tested against it edge for edge. */
static int find_in_row(const oracle_edge *raw, const uint64_t *row, uint64_t u, uint64_t w, uint64_t *at)
{
    for (uint64_t q = row[u]; q < row[u + 1]; q++)
        if (raw[q].dst == w) { *at = q; return 1; }
    return 0;
}

uint64_t oracle_reduce(octx *c, const oracle_edge *raw, uint64_t nraw, oracle_edge **out, oracle_stats *st)
{
    /* 1. rows: raw is sorted by src, then (offset, dst, orient) = the visiting order */
    uint64_t *row = (uint64_t *)calloc(c->n + 2, sizeof(uint64_t));
    for (uint64_t i = 0; i < nraw; i++) row[raw[i].src + 1]++;
    for (uint64_t r = 1; r <= c->n + 1; r++) row[r] += row[r - 1]; /* row[id] .. row[id+1] */
    uint8_t *elim = (uint8_t *)calloc(nraw ? nraw : 1, 1);

    /* 2. markTransitiveEdges for every node */
    uint64_t maxdeg = 0;
    for (uint64_t r = 1; r <= c->n; r++) if (row[r + 1] - row[r] > maxdeg) maxdeg = row[r + 1] - row[r];
    uint8_t *state = (uint8_t *)malloc(maxdeg ? maxdeg : 1); /* 0 INPLAY, 1 ELIMINATED */
    for (uint64_t uu = 1; uu <= c->n; uu++) {
        uint64_t b = row[uu], d = row[uu + 1] - b;
        if (!d) continue;
        memset(state, 0, d);
        for (uint64_t i = 0; i < d; i++) {
            if (state[i]) continue;
            uint64_t v = raw[b + i].dst;
            uint32_t t1 = raw[b + i].orient;
            for (uint64_t q = row[v]; q < row[v + 1]; q++) {
                uint64_t w = raw[q].dst; uint32_t t2 = raw[q].orient;
                int ok = ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) || ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3));
                if (!ok) continue;
                for (uint64_t k = 0; k < d; k++)
                    if (raw[b + k].dst == w) state[k] = 1; /* a visited neighbour can still be eliminated (:712) */
            }
        }
        for (uint64_t k = 0; k < d; k++) if (state[k]) elim[b + k] = 1;
    }
    /* 3. union over both endpoints; removed when flagged in any row that holds the edge */
    oracle_edge *kept = (oracle_edge *)malloc((nraw ? nraw : 1) * sizeof(oracle_edge));
    uint64_t nk = 0;
    for (uint64_t i = 0; i < nraw; i++) {
        const oracle_edge x = raw[i];
        uint64_t at = 0;
        const int twin_found = find_in_row(raw, row, x.dst, x.src, &at);
        /* (the two counters look at entries that survive the marking of the row that holds them: that is where the
         * emission step meets them) */
        if (x.src < x.dst) {
            if (twin_found && !elim[i]) {
                uint64_t Ls = rlen(c, x.src - 1), Ld = rlen(c, x.dst - 1);
                if ((uint32_t)twin_orient((int)x.orient) != raw[at].orient || (uint32_t)(Ld + x.offset - Ls) != raw[at].offset) st->multi_overlap_pairs++;
            } else if (!twin_found && !elim[i]) st->one_sided_edges++;
            if (!elim[i] && !(twin_found && elim[at])) kept[nk++] = x;
        } else if (!twin_found) { /* only the higher id found it: written from the lower id's side as the twin */
            if (!elim[i]) st->one_sided_edges++;
            if (!elim[i]) {
                uint64_t Ls = rlen(c, x.src - 1), Ld = rlen(c, x.dst - 1);
                oracle_edge t;
                t.src = x.dst; t.dst = x.src; t.orient = (uint32_t)twin_orient((int)x.orient);
                t.offset = (uint32_t)(Ld + x.offset - Ls);   /* OverlapGraph.cpp:617 */
                kept[nk++] = t;
            }
        }
    }
    qsort(kept, nk, sizeof(oracle_edge), cmp_edge_pair);
    free(row); free(elim); free(state);
    *out = kept;
    return nk;
}

void oracle_free_buf(void *p) { free(p); }
