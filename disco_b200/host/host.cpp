// host.cpp -- host side of the BuildGraph stage: input parsing, the reference read filter, 2-bit packing, output
// formatting.  Plain C++17 + OpenMP + zlib; see include/disco_host.h for the reference lines each piece restates.
#include "../../include/disco_host.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <omp.h>
#include <string>
#include <vector>
#include <zlib.h>
#include <chrono>

namespace {
thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return -1; }

// Dataset.cpp:48-87
const char *kFilterStrings[] = {
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
// merCheckStrings (Dataset.cpp:87): AC AG AT CG CT GT AAT ATA TAA AAC ACA CAA AAG AGA GAA GGGGCC -- see short_repeat_hit()
constexpr uint64_t kMinReadSize = 30; // Dataset.h:15

uint64_t count_substring(const char *s, uint64_t n, const char *sub, uint64_t m)
{ // Common.h:171-181: std::string::find from the end of the previous match (non-overlapping)
    uint64_t cnt = 0, i = 0;
    while (i + m <= n) {
        const void *p = memmem(s + i, n - i, sub, m);
        if (!p) break;
        cnt++;
        i = (const char *)p - s + m;
    }
    return cnt;
}

// base class table: A C G T -> 0 1 2 3 (the packing code, HashTable.h:16-22), everything else 255
const struct BaseTab {
    uint8_t t[256];
    BaseTab() { for (int i = 0; i < 256; i++) t[i] = 255; t['A'] = 0; t['C'] = 1; t['G'] = 2; t['T'] = 3; }
} kBase;

// The di-/tri-mer repeat test of Dataset.cpp:431-438 for all fifteen short patterns in ONE pass over the base codes.
// countSubstring() counts greedily from the left without overlap; of the listed patterns only ATA, ACA and AGA can
// overlap themselves, so they carry a "next allowed start", the others are plain occurrence counts.
bool short_repeat_hit(const uint8_t *c, uint64_t n, uint64_t thr)
{
    uint32_t di[16] = {0}, tri[64] = {0};
    uint32_t next_xyx[16] = {0}; // indexed by (x, y) of an xyx pattern
    int c0 = c[0], c1 = n > 1 ? c[1] : 0;
    if (n > 1) di[c0 * 4 + c1]++;
    for (uint64_t i = 2; i < n; i++) {
        const int c2 = c[i];
        di[c1 * 4 + c2]++;
        const int t = (c0 * 4 + c1) * 4 + c2;
        if (c0 == c2) { // xyx patterns: greedy, non-overlapping
            const uint32_t start = (uint32_t)(i - 2);
            uint32_t &nx = next_xyx[c0 * 4 + c1];
            if (start >= nx) { tri[t]++; nx = start + 3; }
        } else tri[t]++;
        c0 = c1; c1 = c2;
    }
    // A0 C1 G2 T3 -- "AC","AG","AT","CG","CT","GT"
    const int dimers[6] = {0 * 4 + 1, 0 * 4 + 2, 0 * 4 + 3, 1 * 4 + 2, 1 * 4 + 3, 2 * 4 + 3};
    for (int d : dimers) if ((uint64_t)di[d] * 2 >= thr) return true;
    // "AAT","ATA","TAA","AAC","ACA","CAA","AAG","AGA","GAA"
    const int trimers[9] = {(0 * 4 + 0) * 4 + 3, (0 * 4 + 3) * 4 + 0, (3 * 4 + 0) * 4 + 0, (0 * 4 + 0) * 4 + 1, (0 * 4 + 1) * 4 + 0,
                            (1 * 4 + 0) * 4 + 0, (0 * 4 + 0) * 4 + 2, (0 * 4 + 2) * 4 + 0, (2 * 4 + 0) * 4 + 0};
    for (int t : trimers) if ((uint64_t)tri[t] * 3 >= thr) return true;
    return false;
}

// s = upper-cased read, codes = scratch of n bytes that receives the base codes (valid when the function returns true)
bool test_read_codes(const char *s, uint64_t n, uint8_t *codes)
{
    if (n < kMinReadSize) return false;
    uint64_t cnt[4] = {0, 0, 0, 0};
    uint8_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t c = kBase.t[(unsigned char)s[i]];
        codes[i] = c;
        bad |= c;
        cnt[c & 3]++;
    }
    if (bad & 0x80) return false; // something other than A C G T (Dataset.cpp:411)
    // (the reference counts with (ch >> 1) & 3 = A0 C1 T2 G3: the same four counters in another order)
    uint64_t thr = (uint64_t)((double)n * .7); // Dataset.cpp:415
    if (cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) return false;
    for (const char *f : kFilterStrings) {
        const uint64_t len = 29;
        if (n < len) return false;
        if (f[0] == s[0] && memcmp(f, s, len) == 0) return false;
        if (f[0] == s[n - len] && memcmp(f, s + n - len, len) == 0) return false;
    }
    thr = (uint64_t)((double)n * .5); // Dataset.cpp:431
    if (short_repeat_hit(codes, n, thr)) return false;
    if (count_substring(s, n, "GGGGCC", 6) * 6 >= thr) return false;
    return true;
}

bool test_read(const char *s, uint64_t n)
{
    std::vector<uint8_t> codes(n ? n : 1);
    return test_read_codes(s, n, codes.data());
}

void pack_codes_into(const uint8_t *c, uint64_t n, uint64_t *out)
{
    for (uint64_t w = 0; w * 32 < n; w++) {
        const uint64_t e = std::min<uint64_t>(n, w * 32 + 32);
        uint64_t v = 0;
        for (uint64_t i = w * 32; i < e; i++) v = (v << 2) | c[i];
        out[w] = v << (2 * (w * 32 + 32 - e)); // HashTable.cpp:458-470: base i at bits 62 - 2 (i mod 32)
    }
}

} // namespace

struct disco_reads {
    uint32_t min_overlap = 0;
    int threads = 1;
    uint64_t records = 0;             // file index of the last record seen
    std::vector<uint64_t> vwords;     // accepted reads, 2-bit packed back to back (variable length)
    std::vector<uint64_t> woff{0};    // count + 1 word offsets into vwords
    std::vector<uint16_t> vlen;       // accepted
    std::vector<uint64_t> file_index; // accepted
    bool finalized = false;
    uint32_t wpr = 0, min_len = 0, max_len = 0;
    std::vector<uint64_t> packed;
    std::vector<uint16_t> len;
};

namespace {
// filter a batch of raw records (pointer, length) in parallel, then pack the accepted ones, in order, straight from
// the input buffer into the variable-length 2-bit store (no per-read heap objects)
struct RawRec { const char *p; uint64_t n; bool strip_nl; };

inline uint64_t clean_into(const RawRec &rec, std::string &buf)
{
    static const struct Upper { char t[256]; Upper() { for (int i = 0; i < 256; i++) t[i] = (char)toupper(i); } } up;
    buf.resize(rec.n);
    uint64_t o = 0;
    const char *p = rec.p;
    if (rec.strip_nl) {
        for (uint64_t k = 0; k < rec.n; k++) { const char c = p[k]; if (c != '\n') buf[o++] = up.t[(unsigned char)c]; } // Dataset.cpp:276 removes only '\n'
    } else {
        for (uint64_t k = 0; k < rec.n; k++) buf[o++] = up.t[(unsigned char)p[k]];                                   // Dataset.cpp:303-304
    }
    return o;
}

void absorb(disco_reads *r, const std::vector<RawRec> &batch)
{
    const size_t m = batch.size();
    std::vector<uint16_t> clen(m, 0); // 0 = rejected
#pragma omp parallel num_threads(r->threads)
    {
        std::string buf;
        std::vector<uint8_t> codes;
#pragma omp for schedule(dynamic, 2048)
        for (size_t i = 0; i < m; i++) {
            const uint64_t o = clean_into(batch[i], buf);
            if (codes.size() < o) codes.resize(o);
            if (o > r->min_overlap && o <= 32767 && test_read_codes(buf.data(), o, codes.data())) clen[i] = (uint16_t)o; // Dataset.cpp:305
        }
    }
    // accepted records get consecutive read ids in file order (Dataset.cpp:133-134 after the file-order sort)
    const size_t n0 = r->vlen.size();
    std::vector<uint64_t> slot(m);
    uint64_t k = n0, w = r->woff.back();
    for (size_t i = 0; i < m; i++) {
        r->records++;
        slot[i] = k;
        if (clen[i]) {
            r->file_index.push_back(r->records);
            r->vlen.push_back(clen[i]);
            w += (clen[i] + 31) / 32;
            r->woff.push_back(w);
            k++;
        }
    }
    r->vwords.resize(w, 0);
#pragma omp parallel num_threads(r->threads)
    {
        std::string buf;
        std::vector<uint8_t> codes;
#pragma omp for schedule(dynamic, 2048)
        for (size_t i = 0; i < m; i++) {
            if (!clen[i]) continue;
            const uint64_t o = clean_into(batch[i], buf);
            if (codes.size() < o) codes.resize(o);
            for (uint64_t q = 0; q < o; q++) codes[q] = kBase.t[(unsigned char)buf[q]];
            pack_codes_into(codes.data(), o, r->vwords.data() + r->woff[slot[i]]);
        }
    }
}
} // namespace

extern "C" {

const char *disco_host_last_error(void) { return g_err.c_str(); }

int disco_host_test_read(const char *seq, uint64_t len) { return test_read(seq, len) ? 1 : 0; }

disco_reads *disco_reads_new(uint32_t min_overlap, int threads)
{
    disco_reads *r = new disco_reads();
    r->min_overlap = min_overlap;
    r->threads = threads > 0 ? threads : omp_get_max_threads();
    return r;
}

void disco_reads_free(disco_reads *r) { delete r; }

int disco_reads_add_records(disco_reads *r, const char *seqs, const uint64_t *off, uint64_t n)
{
    if (!r || r->finalized) return fail("reads object already finalized");
    std::vector<RawRec> batch(n);
    for (uint64_t i = 0; i < n; i++) batch[i] = RawRec{seqs + off[i], off[i + 1] - off[i], false};
    absorb(r, batch);
    return 0;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int disco_reads_add_file(disco_reads *r, const char *path)
{
    const bool trace = getenv("DISCO_HOST_TRACE") != nullptr;
    double t0 = now_s();
    if (!r || r->finalized) return fail("reads object already finalized");
    std::string data;
    {
        FILE *f = fopen(path, "rb");
        if (!f) return fail(std::string("Unable to open file: ") + path);
        unsigned char magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, f);
        const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (!gz) { // plain text: one read of the whole file
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            data.resize(sz > 0 ? (size_t)sz : 0);
            if (sz > 0 && fread(&data[0], 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return fail(std::string("read error in ") + path); }
            fclose(f);
        } else {
            fclose(f);
            gzFile fp = gzopen(path, "rb");
            if (!fp) return fail(std::string("Unable to open file: ") + path);
            gzbuffer(fp, 1 << 20);
            std::vector<char> buf(1 << 24);
            int n;
            while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) data.append(buf.data(), (size_t)n);
            if (n < 0) { gzclose(fp); return fail(std::string("read error in ") + path); }
            gzclose(fp);
        }
    }
    if (trace) { fprintf(stderr, "[host] read %.3fs\n", now_s() - t0); t0 = now_s(); }
    const uint64_t before = r->records;
    if (!data.empty()) {
        const char *b = data.data(), *e = b + data.size();
        std::vector<RawRec> batch;
        if (*b == '>') { // FASTA (Dataset.cpp:270-281): header line, then everything up to the next '>'
            const char *p = b;
            while (p < e) {
                const char *nl = (const char *)memchr(p, '\n', e - p);
                if (!nl) { batch.push_back(RawRec{e, 0, true}); break; } // header without sequence
                const char *s = nl + 1;
                const char *nx = (const char *)memchr(s, '>', e - s);
                if (!nx) nx = e;
                batch.push_back(RawRec{s, (uint64_t)(nx - s), true});
                p = nx + (nx < e ? 1 : 0);
                if (nx == e) break;
            }
        } else if (*b == '@') { // FASTQ (Dataset.cpp:282-293): four lines per record, sequence on the second
            const char *p = b;
            auto next_line = [&](const char *&ls, const char *&le) { // std::getline semantics
                if (p >= e) return false;
                ls = p;
                const char *nl = (const char *)memchr(p, '\n', e - p);
                le = nl ? nl : e;
                p = nl ? nl + 1 : e;
                return true;
            };
            const char *ls, *le;
            while (next_line(ls, le)) {
                const char *ss = e, *se = e;
                if (!next_line(ss, se)) { ss = se = e; }
                const char *d0, *d1;
                next_line(d0, d1);
                next_line(d0, d1);
                batch.push_back(RawRec{ss, (uint64_t)(se - ss), false});
            }
        } else {
            return fail("Unknown input file format."); // Dataset.cpp:267
        }
        if (trace) { fprintf(stderr, "[host] index %.3fs (%zu records)\n", now_s() - t0, batch.size()); t0 = now_s(); }
        absorb(r, batch);
        if (trace) { fprintf(stderr, "[host] absorb %.3fs\n", now_s() - t0); t0 = now_s(); }
    }
    if (r->records <= before) return fail(std::string("File empty. No reads loaded from ") + path); // Dataset.cpp:113-114
    return 0;
}

int disco_reads_finalize(disco_reads *r)
{
    if (!r) return fail("NULL");
    if (r->finalized) return 0;
    const uint64_t n = r->vlen.size();
    uint32_t mn = 0xFFFFFFFFu, mx = 0;
    for (uint16_t l : r->vlen) { mn = std::min<uint32_t>(mn, l); mx = std::max<uint32_t>(mx, l); }
    if (n == 0) { mn = mx = 0; }
    r->min_len = mn; r->max_len = mx;
    r->wpr = std::max<uint32_t>(1, (mx + 31) / 32); // exact: the GPU library re-strides on the device
    r->packed.assign(n * (uint64_t)r->wpr, 0);
    r->len = r->vlen;
#pragma omp parallel for schedule(static) num_threads(r->threads)
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t nw = r->woff[i + 1] - r->woff[i];
        memcpy(r->packed.data() + i * r->wpr, r->vwords.data() + r->woff[i], nw * sizeof(uint64_t));
    }
    std::vector<uint64_t>().swap(r->vwords);
    r->finalized = true;
    return 0;
}

uint64_t disco_reads_count(const disco_reads *r) { return r->vlen.size(); }
uint64_t disco_reads_records(const disco_reads *r) { return r->records; }
uint32_t disco_reads_words_per_read(const disco_reads *r) { return r->wpr; }
const uint64_t *disco_reads_packed(const disco_reads *r) { return r->packed.data(); }
const uint16_t *disco_reads_len(const disco_reads *r) { return r->len.data(); }
const uint64_t *disco_reads_file_index(const disco_reads *r) { return r->file_index.data(); }
uint32_t disco_reads_min_len(const disco_reads *r) { return r->min_len; }
uint32_t disco_reads_max_len(const disco_reads *r) { return r->max_len; }

int disco_host_pack_codes(const uint8_t *codes, const uint64_t *off, uint64_t n, uint32_t wpr, uint64_t *out,
                          uint16_t *len_out, int threads)
{
    if (threads <= 0) threads = omp_get_max_threads();
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(| : bad)
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t L = off[i + 1] - off[i];
        uint64_t *o = out + i * wpr;
        for (uint32_t w = 0; w < wpr; w++) o[w] = 0;
        if (L > 32767 || (L + 31) / 32 > wpr) { bad = 1; continue; }
        const uint8_t *c = codes + off[i];
        for (uint64_t k = 0; k < L; k++) o[k >> 5] |= (uint64_t)(c[k] & 3) << (62 - 2 * (k & 31));
        len_out[i] = (uint16_t)L;
    }
    return bad ? fail("read too long for words_per_read") : 0;
}

int disco_host_sort_contained(disco_crow *rows, uint64_t n, const uint16_t *len, uint32_t min_overlap)
{
    const int K = (int)min_overlap - 1;
    std::vector<std::pair<uint64_t, uint64_t>> keys(n); // (container, position | record) , index
    std::vector<uint64_t> k2(n);
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
        const disco_crow &r = rows[i];
        const int L1 = len[r.container];
        // orient 3/2 <- types 0/2: start = j ; orient 0/1 <- types 1/3: start = L1 - K - j   (OverlapGraph.cpp:428-434)
        const uint64_t j = (r.orient == 3 || r.orient == 2) ? r.start : (uint64_t)(L1 - K - (int)r.start);
        const uint64_t kind = (r.orient == 3 || r.orient == 1) ? 0 : 1;
        keys[i] = {((uint64_t)r.container << 16) | j, i};
        k2[i] = 2ULL * r.contained + kind;
    }
    std::sort(keys.begin(), keys.end(), [&](const std::pair<uint64_t, uint64_t> &a, const std::pair<uint64_t, uint64_t> &b) {
        if (a.first != b.first) return a.first < b.first;
        return k2[a.second] < k2[b.second];
    });
    std::vector<disco_crow> out(n);
    for (uint64_t i = 0; i < n; i++) out[i] = rows[keys[i].second];
    std::copy(out.begin(), out.end(), rows);
    return 0;
}

int disco_host_sort_edges(disco_edge *edges, uint64_t n)
{
    std::sort(edges, edges + n, [](const disco_edge &a, const disco_edge &b) {
        if (a.src != b.src) return a.src < b.src;
        if (a.dst != b.dst) return a.dst < b.dst;
        if (a.offset != b.offset) return a.offset < b.offset;
        return a.orient < b.orient;
    });
    return 0;
}

int disco_write_pargraph(const char *path, const disco_edge *edges, uint64_t n, const uint64_t *file_index,
                         const uint16_t *len, int flag, int append)
{
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return fail(std::string("Unable to open file: ") + path);
    std::vector<char> buf(1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (uint64_t i = 0; i < n; i++) {
        const disco_edge &e = edges[i];
        const unsigned long long sl = len[e.src], dl = len[e.dst], off = e.offset, ovl = sl - off;
        // src dst orient,ovl,0,0,srcLen,offset,srcLen-1,dstLen,0,ovl-1,NA,flag   (OverlapGraph.cpp:811-867)
        fprintf(f, "%llu\t%llu\t%u,%llu,0,0,%llu,%llu,%llu,%llu,0,%llu,NA,%d\n", (unsigned long long)file_index[e.src],
                (unsigned long long)file_index[e.dst], e.orient, ovl, sl, off, sl - 1, dl, ovl - 1, flag);
    }
    fclose(f);
    return 0;
}

int disco_write_contained(const char *path, const disco_crow *rows, uint64_t n, const uint64_t *file_index,
                          const uint16_t *len, int append)
{
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return fail(std::string("Unable to open file: ") + path);
    std::vector<char> buf(1 << 22);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (uint64_t i = 0; i < n; i++) {
        const disco_crow &r = rows[i];
        const unsigned long long l2 = len[r.contained], l1 = len[r.container], st = r.start;
        // contained container orient,L2,0,0,L2,0,L2,L1,start,start+L2   (OverlapGraph.cpp:438-447)
        fprintf(f, "%llu\t%llu\t%u,%llu,0,0,%llu,0,%llu,%llu,%llu,%llu\n", (unsigned long long)file_index[r.contained],
                (unsigned long long)file_index[r.container], r.orient, l2, l2, l2, l1, st, st + l2);
    }
    fclose(f);
    return 0;
}

} // extern "C"
