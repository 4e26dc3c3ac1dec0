// host.cpp -- host side of the BuildGraph stage: input parsing, the reference read filter, 2-bit packing, output
// formatting.  Plain C++17 + OpenMP + zlib; see include/disco_host.h for the reference lines each piece restates.
#include "../../include/disco_host.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <omp.h>
#include <parallel/algorithm>
#include <string>
#include <vector>
#include <zlib.h>
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

// ---- testRead (Dataset.cpp:403-452) on the 2-bit packed read ----------------------------------------------------------
// The reference works on std::string (count per base, 38 x 2 string compares, fifteen countSubstring() passes).  Here a
// record is packed in ONE pass over its characters (which also finds non-ACGT letters); everything else is word
// arithmetic on the packed form: base counts and pattern counts are popcounts of per-base equality masks, the two
// 29-base ends are 58-bit integers looked up among the packed filter strings.  Same decisions, ~4x fewer cycles.
constexpr uint64_t kEven = 0x5555555555555555ULL;

// A C G T (either case) -> 0 1 2 3 (the packing code, HashTable.h:16-22), everything else 255.  The reference
// upper-cases first (Dataset.cpp:303-304), so lower-case letters are accepted as their upper-case form.
const struct CodeTab {
    uint8_t t[256];
    CodeTab()
    {
        for (int i = 0; i < 256; i++) t[i] = 255;
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    }
} kCode;

// the 38 filter strings as 58-bit integers (29 bases, first base in the top bits)
const struct FilterKeys {
    uint64_t k[sizeof(kFilterStrings) / sizeof(kFilterStrings[0])];
    FilterKeys()
    {
        for (size_t f = 0; f < sizeof(k) / sizeof(k[0]); f++) {
            uint64_t v = 0;
            for (int i = 0; i < 29; i++) v = (v << 2) | kCode.t[(unsigned char)kFilterStrings[f][i]];
            k[f] = v;
        }
    }
    bool has(uint64_t key) const
    {
        bool hit = false;
        for (uint64_t x : k) hit |= (x == key);
        return hit;
    }
} kFilterKeys;

// Pack a raw record (upper- or lower-case, optionally with embedded newlines that are dropped, Dataset.cpp:276) into
// `out` (base i at bits 62 - 2 (i mod 32) of word i / 32, HashTable.cpp:458-470; the tail of the last word is zero).
// Returns the number of bases; *bad is set when a character other than A C G T was seen (its code is then garbage).
inline uint64_t pack_raw(const char *p, uint64_t n, bool strip_nl, uint64_t *out, bool *bad)
{
    uint64_t w = 0, L = 0, k = 0;
    uint8_t acc = 0;
    for (uint64_t i = 0; i < n; i++) {
        const unsigned char ch = (unsigned char)p[i];
        if (strip_nl && ch == '\n') continue;
        const uint8_t c = kCode.t[ch];
        acc |= c;
        w = (w << 2) | (c & 3);
        if ((++L & 31) == 0) { out[k++] = w; w = 0; }
    }
    if (L & 31) out[k++] = w << (2 * (32 - (L & 31)));
    *bad = (acc & 0x80) != 0;
    return L;
}

// bits 2j (j = 31 - position in the word) set where the base equals x; `valid` masks the real bases of the word
inline uint64_t eq_mask(uint64_t w, int x, uint64_t valid)
{
    const uint64_t hi = (w >> 1) & kEven, lo = w & kEven;
    const uint64_t h = (x & 2) ? hi : ~hi, l = (x & 1) ? lo : ~lo;
    return h & l & valid;
}

// 29 bases starting at base `start` as a 58-bit integer
inline uint64_t key29(const uint64_t *w, uint64_t start)
{
    const uint64_t wi = start >> 5, off = start & 31;
    uint64_t v = w[wi] << (2 * off);
    if (off > 3) v |= w[wi + 1] >> (64 - 2 * off);
    return v >> 6;
}

// testRead on a packed read of L bases (W = ceil(L / 32) words).  eq = scratch of 4 * (W + 1) words.
bool test_packed(const uint64_t *w, uint64_t L, uint64_t *eq)
{
    if (L < kMinReadSize) return false;
    const uint64_t W = (L + 31) / 32;
    // per-base equality masks, E[x][k]; one extra zero word so that shifted reads never run off the end
    uint64_t *E[4] = {eq, eq + (W + 1), eq + 2 * (W + 1), eq + 3 * (W + 1)};
    uint64_t cnt[4] = {0, 0, 0, 0};
    for (uint64_t k = 0; k < W; k++) {
        const uint64_t nb = (k + 1 < W || (L & 31) == 0) ? 32 : (L & 31);
        const uint64_t valid = nb == 32 ? kEven : (kEven << (2 * (32 - nb)));
        for (int x = 0; x < 4; x++) { E[x][k] = eq_mask(w[k], x, valid); cnt[x] += (uint64_t)__builtin_popcountll(E[x][k]); }
    }
    for (int x = 0; x < 4; x++) E[x][W] = 0;
    uint64_t thr = (uint64_t)((double)L * .7); // Dataset.cpp:415
    if (cnt[0] >= thr || cnt[1] >= thr || cnt[2] >= thr || cnt[3] >= thr) return false;
    // either end equal to one of the 38 micro-repeat strings (Dataset.cpp:420-429; L >= 30 > 29 here)
    if (kFilterKeys.has(key29(w, 0)) || kFilterKeys.has(key29(w, L - 29))) return false;
    thr = (uint64_t)((double)L * .5);          // Dataset.cpp:431
    // mask of base x at position i + s, aligned to position i (s bases further on = 2s bits lower)
    auto sh = [&](int x, uint64_t k, int s) -> uint64_t { return (E[x][k] << (2 * s)) | (E[x][k + 1] >> (64 - 2 * s)); };
    enum { A = 0, C = 1, G = 2, T = 3 };
    // countSubstring() (Common.h:171-181) counts from the left without overlap; only the x-y-x patterns (ATA ACA AGA) can
    // overlap themselves, for every other pattern that is the plain number of occurrences.  One sweep over the words
    // counts all sixteen patterns: di[] = AC AG AT CG CT GT, tri[] = AAT AAC AAG | TAA CAA GAA | ATA ACA AGA (overlapping)
    uint64_t di[6] = {0}, tri[9] = {0}, c6 = 0;
    auto pc = [](uint64_t v) { return (uint64_t)__builtin_popcountll(v); };
    for (uint64_t k = 0; k < W; k++) {
        const uint64_t e[4] = {E[A][k], E[C][k], E[G][k], E[T][k]};
        const uint64_t s1[4] = {sh(A, k, 1), sh(C, k, 1), sh(G, k, 1), sh(T, k, 1)};
        const uint64_t s2[4] = {sh(A, k, 2), sh(C, k, 2), sh(G, k, 2), sh(T, k, 2)};
        di[0] += pc(e[A] & s1[C]); di[1] += pc(e[A] & s1[G]); di[2] += pc(e[A] & s1[T]);
        di[3] += pc(e[C] & s1[G]); di[4] += pc(e[C] & s1[T]); di[5] += pc(e[G] & s1[T]);
        const uint64_t aa_ = e[A] & s1[A], _aa = s1[A] & s2[A], a_a = e[A] & s2[A];
        tri[0] += pc(aa_ & s2[T]); tri[1] += pc(aa_ & s2[C]); tri[2] += pc(aa_ & s2[G]);
        tri[3] += pc(e[T] & _aa); tri[4] += pc(e[C] & _aa); tri[5] += pc(e[G] & _aa);
        tri[6] += pc(a_a & s1[T]); tri[7] += pc(a_a & s1[C]); tri[8] += pc(a_a & s1[G]);
        c6 += pc(e[G] & s1[G] & s2[G] & sh(G, k, 3) & sh(C, k, 4) & sh(C, k, 5)); // GGGGCC
    }
    for (uint64_t d : di) if (d * 2 >= thr) return false;
    for (int t = 0; t < 6; t++) if (tri[t] * 3 >= thr) return false;
    if (c6 * 6 >= thr) return false;
    const int mid[3] = {T, C, G};
    for (int t = 0; t < 3; t++) { // the greedy count is at most the overlapping count -- walk only when that is large
        if (tri[6 + t] * 3 < thr) continue;
        uint64_t greedy = 0;
        for (uint64_t i = 0; i + 3 <= L;) {
            const uint64_t k = i >> 5, bit = 62 - 2 * (i & 31);
            const bool hit = ((E[A][k] & sh(mid[t], k, 1) & sh(A, k, 2)) >> bit) & 1;
            if (hit) { greedy++; i += 3; } else i++;
        }
        if (greedy * 3 >= thr) return false;
    }
    return true;
}

// accept / reject one raw record the way Dataset::readDataset does (Dataset.cpp:305: longer than minOverlap, then
// testRead); on acceptance `words` holds the packed read.  scratch vectors are per thread.
inline uint64_t filter_record(const char *p, uint64_t n, bool strip_nl, uint32_t min_overlap, std::vector<uint64_t> &words,
                              std::vector<uint64_t> &eq)
{
    const uint64_t wmax = n / 32 + 2;
    if (words.size() < wmax) words.resize(wmax);
    if (eq.size() < 4 * (wmax + 1)) eq.resize(4 * (wmax + 1));
    bool bad = false;
    const uint64_t L = pack_raw(p, n, strip_nl, words.data(), &bad);
    if (bad || L <= min_overlap || L > 32767) return 0;
    return test_packed(words.data(), L, eq.data()) ? L : 0;
}

bool test_read(const char *s, uint64_t n)
{
    std::vector<uint64_t> words, eq;
    bool bad = false;
    words.resize(n / 32 + 2);
    eq.resize(4 * (n / 32 + 3));
    const uint64_t L = pack_raw(s, n, false, words.data(), &bad);
    return !bad && test_packed(words.data(), L, eq.data());
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

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void absorb(disco_reads *r, const std::vector<RawRec> &batch)
{
    const bool trace = getenv("DISCO_HOST_TRACE") != nullptr;
    double t0 = trace ? now_s() : 0.0;
    const size_t m = batch.size();
    std::vector<uint16_t> clen(m, 0); // 0 = rejected
#pragma omp parallel num_threads(r->threads)
    {
        std::vector<uint64_t> words, eq;
#pragma omp for schedule(dynamic, 2048)
        for (size_t i = 0; i < m; i++)
            clen[i] = (uint16_t)filter_record(batch[i].p, batch[i].n, batch[i].strip_nl, r->min_overlap, words, eq);
    }
    if (trace) { fprintf(stderr, "[host]   filter %.3fs (%d threads)\n", now_s() - t0, r->threads); t0 = now_s(); }
    // accepted records get consecutive read ids in file order (Dataset.cpp:133-134 after the file-order sort)
    const size_t n0 = r->vlen.size();
    std::vector<uint64_t> slot(m);
    uint64_t k = n0, w = r->woff.back();
    for (size_t i = 0; i < m; i++) { slot[i] = k; k += clen[i] != 0; }
    r->file_index.resize(k); r->vlen.resize(k); r->woff.resize(k + 1);
    const uint64_t rec0 = r->records;
    for (size_t i = 0; i < m; i++) { // word offsets: a running sum over the accepted reads
        if (!clen[i]) continue;
        w += (clen[i] + 31) / 32;
        r->woff[slot[i] + 1] = w;
    }
#pragma omp parallel for schedule(static) num_threads(r->threads)
    for (size_t i = 0; i < m; i++) {
        if (!clen[i]) continue;
        r->file_index[slot[i]] = rec0 + i + 1; // fileIndex counts every record (Dataset.cpp:294)
        r->vlen[slot[i]] = clen[i];
    }
    r->records += m;
    r->vwords.resize(w, 0);
    if (trace) { fprintf(stderr, "[host]   number %.3fs\n", now_s() - t0); t0 = now_s(); }
#pragma omp parallel for schedule(dynamic, 2048) num_threads(r->threads)
    for (size_t i = 0; i < m; i++) {
        if (!clen[i]) continue;
        bool bad;
        pack_raw(batch[i].p, batch[i].n, batch[i].strip_nl, r->vwords.data() + r->woff[slot[i]], &bad); // exactly (clen + 31) / 32 words
    }
    if (trace) fprintf(stderr, "[host]   pack %.3fs\n", now_s() - t0);
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


int disco_reads_add_file(disco_reads *r, const char *path)
{
    const bool trace = getenv("DISCO_HOST_TRACE") != nullptr;
    double t0 = now_s();
    if (!r || r->finalized) return fail("reads object already finalized");
    std::string data;            // gz: the inflated text
    const char *base = nullptr;  // the text to parse: the mapping of a plain file, or data
    size_t size = 0;
    struct Mapping {
        void *p = MAP_FAILED; size_t n = 0;
        ~Mapping() { if (p != MAP_FAILED) munmap(p, n); }
    } map;
    {
        FILE *f = fopen(path, "rb");
        if (!f) return fail(std::string("Unable to open file: ") + path);
        unsigned char magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, f);
        const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        fclose(f);
        if (!gz) { // plain text: map the file, the parser works on the page cache directly
            const int fd = open(path, O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); return fail(std::string("Unable to open file: ") + path); }
            if (st.st_size > 0) {
                map.n = (size_t)st.st_size;
                map.p = mmap(nullptr, map.n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
                if (map.p != MAP_FAILED) {
                    madvise(map.p, map.n, MADV_SEQUENTIAL);
                    madvise(map.p, map.n, MADV_WILLNEED);
                    base = static_cast<const char *>(map.p); size = map.n;
                } else { // no mmap on this file system: read it
                    data.resize(map.n);
                    size_t off = 0;
                    while (off < map.n) { const ssize_t g = read(fd, &data[off], map.n - off); if (g <= 0) break; off += (size_t)g; }
                    if (off != map.n) { close(fd); return fail(std::string("read error in ") + path); }
                    base = data.data(); size = data.size();
                }
            }
            close(fd);
        } else {
            gzFile fp = gzopen(path, "rb");
            if (!fp) return fail(std::string("Unable to open file: ") + path);
            gzbuffer(fp, 1 << 20);
            std::vector<char> buf(1 << 24);
            int n;
            while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) data.append(buf.data(), (size_t)n);
            if (n < 0) { gzclose(fp); return fail(std::string("read error in ") + path); }
            gzclose(fp);
            base = data.data(); size = data.size();
        }
    }
    if (trace) { fprintf(stderr, "[host] read %.3fs\n", now_s() - t0); t0 = now_s(); }
    const uint64_t before = r->records;
    if (size) {
        const char *b = base, *e = b + size;
        std::vector<RawRec> batch;
        // every position of `ch`, found by all threads at once (the text is hundreds of MB; one memchr pass per thread)
        auto find_all = [&](char ch, std::vector<uint64_t> &pos) {
            const int T = std::max(1, r->threads);
            std::vector<std::vector<uint64_t>> part(T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int t = 0; t < T; t++) {
                const char *lo = b + size * (uint64_t)t / T, *hi = b + size * (uint64_t)(t + 1) / T;
                for (const char *q = lo; q < hi;) {
                    q = (const char *)memchr(q, ch, hi - q);
                    if (!q) break;
                    part[t].push_back((uint64_t)(q - b));
                    q++;
                }
            }
            size_t total = 0;
            std::vector<size_t> first(T + 1, 0);
            for (int t = 0; t < T; t++) { total += part[t].size(); first[t + 1] = total; }
            pos.resize(total);
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int t = 0; t < T; t++) std::copy(part[t].begin(), part[t].end(), pos.begin() + first[t]);
        };
        if (*b == '>') { // FASTA (Dataset.cpp:270-281): header line, then everything up to the next '>'
            std::vector<uint64_t> gt;
            find_all('>', gt);
            // a '>' inside a header line does not start a record: walk the candidates in order (cheap: one short memchr each)
            size_t gi = 0;
            const char *p = b;
            batch.reserve(gt.size());
            while (p < e) {
                const char *nl = (const char *)memchr(p, '\n', e - p);
                if (!nl) { batch.push_back(RawRec{e, 0, true}); break; } // header without sequence
                const char *s = nl + 1;
                const uint64_t so = (uint64_t)(s - b);
                while (gi < gt.size() && gt[gi] < so) gi++;
                const char *nx = gi < gt.size() ? b + gt[gi] : e;
                batch.push_back(RawRec{s, (uint64_t)(nx - s), true});
                if (nx == e) break;
                p = nx + 1;
            }
        } else if (*b == '@') { // FASTQ (Dataset.cpp:282-293): four lines per record, sequence on the second
            std::vector<uint64_t> nl;
            find_all('\n', nl);
            // std::getline semantics: a last line without a newline still counts
            const uint64_t lines = nl.size() + ((nl.empty() ? size > 0 : nl.back() + 1 < size) ? 1 : 0);
            const uint64_t nrec = (lines + 3) / 4;
            batch.resize(nrec);
#pragma omp parallel for schedule(static) num_threads(r->threads)
            for (uint64_t i = 0; i < nrec; i++) {
                const uint64_t j = 4 * i + 1; // the sequence line
                if (j >= lines) { batch[i] = RawRec{e, 0, false}; continue; }
                const char *ss = b + nl[j - 1] + 1;
                const char *se = j < nl.size() ? b + nl[j] : e;
                batch[i] = RawRec{ss, (uint64_t)(se - ss), false};
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
    __gnu_parallel::sort(edges, edges + n, [](const disco_edge &a, const disco_edge &b) {
        if (a.src != b.src) return a.src < b.src;
        if (a.dst != b.dst) return a.dst < b.dst;
        if (a.offset != b.offset) return a.offset < b.offset;
        return a.orient < b.orient;
    });
    return 0;
}

} // extern "C"

namespace {
// decimal digits of v appended at p; returns the new end (the writers below format ~10^8 integers per graph)
inline char *put_u64(char *p, uint64_t v)
{
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char *put_str(char *p, const char *s) { while (*s) *p++ = *s++; return p; }

// Format n lines with `line(i, p) -> new end` on all cores (each thread fills its own buffer, at most max_line bytes
// per line), then write the buffers in order.
template <typename Line>
int write_lines(const char *path, int append, uint64_t n, size_t max_line, const Line &line)
{
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return fail(std::string("Unable to open file: ") + path);
    const uint64_t block = 1 << 16; // lines per buffer
    const int T = std::max(1, omp_get_max_threads());
    std::vector<std::vector<char>> buf(T);
    std::vector<size_t> used(T, 0);
    bool ok = true;
    for (uint64_t base = 0; base < n && ok; base += block * (uint64_t)T) {
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; t++) {
            const uint64_t lo = base + block * (uint64_t)t, hi = std::min<uint64_t>(n, lo + block);
            used[t] = 0;
            if (lo >= hi) continue;
            if (buf[t].size() < (hi - lo) * max_line) buf[t].resize((hi - lo) * max_line);
            char *p = buf[t].data();
            for (uint64_t i = lo; i < hi; i++) p = line(i, p);
            used[t] = (size_t)(p - buf[t].data());
        }
        for (int t = 0; t < T && ok; t++)
            if (used[t]) ok = fwrite(buf[t].data(), 1, used[t], f) == used[t];
    }
    if (fclose(f) != 0) ok = false;
    return ok ? 0 : fail(std::string("write error in ") + path);
}
} // namespace

extern "C" {

namespace {
// src dst orient,ovl,0,0,srcLen,offset,srcLen-1,dstLen,0,ovl-1,NA,flag   (OverlapGraph.cpp:811-867)
inline char *put_edge(char *p, const disco_edge &e, const uint64_t *file_index, const uint16_t *len, int flag)
{
    const uint64_t sl = len[e.src], dl = len[e.dst], off = e.offset, ovl = sl - off;
    p = put_u64(p, file_index[e.src]); *p++ = '\t';
    p = put_u64(p, file_index[e.dst]); *p++ = '\t';
    p = put_u64(p, e.orient); *p++ = ',';
    p = put_u64(p, ovl); p = put_str(p, ",0,0,");
    p = put_u64(p, sl); *p++ = ',';
    p = put_u64(p, off); *p++ = ',';
    p = put_u64(p, sl - 1); *p++ = ',';
    p = put_u64(p, dl); p = put_str(p, ",0,");
    p = put_u64(p, ovl - 1); p = put_str(p, ",NA,");
    if (flag < 0) { *p++ = '-'; p = put_u64(p, (uint64_t)(-(long long)flag)); } else p = put_u64(p, (uint64_t)flag);
    *p++ = '\n';
    return p;
}
} // namespace

int disco_write_pargraph(const char *path, const disco_edge *edges, uint64_t n, const uint64_t *file_index,
                         const uint16_t *len, int flag, int append)
{
    return write_lines(path, append, n, 160, [&](uint64_t i, char *p) { return put_edge(p, edges[i], file_index, len, flag); });
}

// The reference's partial graphs: each BuildGraph thread finalises its own set of nodes and appends their edges to its own
// file; an edge whose two endpoints were finalised by the same thread is written once with mark flag 2, an edge between
// nodes of different threads twice -- by the source's thread with flag 0 ("only source is marked") and by the
// destination's thread with flag 1 (OverlapGraph.cpp:826-833, :852-859).  One parsimplify per file then contracts
// through the nodes marked in that file only (OverlapGraphSimple.cpp:632-641), which is what lets them run in parallel.
// Here shard t owns the reads [t * n_reads / shards, (t+1) * n_reads / shards).  `edges` sorted by (src, dst).
int disco_write_pargraph_sharded(const char *prefix, uint32_t shards, const disco_edge *edges, uint64_t n, uint64_t n_reads,
                                 const uint64_t *file_index, const uint16_t *len)
{
    if (!prefix || shards < 1) return fail("bad arguments");
    std::vector<uint64_t> bound(shards + 1);
    for (uint32_t t = 0; t <= shards; t++) bound[t] = (uint64_t)((unsigned __int128)n_reads * t / shards);
    auto shard_of = [&](uint32_t r) { return (uint32_t)(std::upper_bound(bound.begin() + 1, bound.end(), (uint64_t)r) - (bound.begin() + 1)); };
    for (uint64_t i = 1; i < n; i++)
        if (edges[i].src < edges[i - 1].src) return fail("disco_write_pargraph_sharded: edges must be sorted by source read");
    // edges that cross into a later shard, grouped by the destination's shard (sources ascending inside a group)
    std::vector<std::vector<uint64_t>> incoming(shards);
    if (shards > 1) {
        const int T = std::max(1, omp_get_max_threads());
        std::vector<std::vector<std::vector<uint64_t>>> part(T, std::vector<std::vector<uint64_t>>(shards));
#pragma omp parallel num_threads(T)
        {
            const int me = omp_get_thread_num();
            const uint64_t lo = n * (uint64_t)me / T, hi = n * (uint64_t)(me + 1) / T;
            for (uint64_t i = lo; i < hi; i++) {
                const uint32_t a = shard_of(edges[i].src), b = shard_of(edges[i].dst);
                if (a != b) part[me][b].push_back(i);
            }
        }
        for (uint32_t t = 0; t < shards; t++)
            for (int k = 0; k < T; k++) incoming[t].insert(incoming[t].end(), part[k][t].begin(), part[k][t].end());
    }
    for (uint32_t t = 0; t < shards; t++) {
        const std::string path = std::string(prefix) + "_" + std::to_string(t) + "_parGraph.txt";
        const disco_edge *first = std::lower_bound(edges, edges + n, bound[t], [](const disco_edge &e, uint64_t v) { return (uint64_t)e.src < v; });
        const disco_edge *last = std::lower_bound(edges, edges + n, bound[t + 1], [](const disco_edge &e, uint64_t v) { return (uint64_t)e.src < v; });
        const uint64_t own = (uint64_t)(last - first);
        const std::vector<uint64_t> &in = incoming[t];
        const int rc = write_lines(path.c_str(), 0, own + in.size(), 160, [&](uint64_t i, char *p) {
            if (i < own) {
                const disco_edge &e = first[i];
                return put_edge(p, e, file_index, len, e.dst < bound[t + 1] ? 2 : 0); // (src < dst: the destination cannot lie in an earlier shard)
            }
            return put_edge(p, edges[in[i - own]], file_index, len, 1);
        });
        if (rc) return rc;
    }
    return 0;
}

int disco_write_contained(const char *path, const disco_crow *rows, uint64_t n, const uint64_t *file_index,
                          const uint16_t *len, int append)
{
    // contained container orient,L2,0,0,L2,0,L2,L1,start,start+L2   (OverlapGraph.cpp:438-447)
    return write_lines(path, append, n, 160, [&](uint64_t i, char *p) {
        const disco_crow &r = rows[i];
        const uint64_t l2 = len[r.contained], l1 = len[r.container], st = r.start;
        p = put_u64(p, file_index[r.contained]); *p++ = '\t';
        p = put_u64(p, file_index[r.container]); *p++ = '\t';
        p = put_u64(p, r.orient); *p++ = ',';
        p = put_u64(p, l2); p = put_str(p, ",0,0,");
        p = put_u64(p, l2); p = put_str(p, ",0,");
        p = put_u64(p, l2); *p++ = ',';
        p = put_u64(p, l1); *p++ = ',';
        p = put_u64(p, st); *p++ = ',';
        p = put_u64(p, st + l2);
        *p++ = '\n';
        return p;
    });
}

} // extern "C"
