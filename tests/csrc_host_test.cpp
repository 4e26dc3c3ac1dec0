// Host-side unit test of disco_b200/csrc/dna.cuh (the packed-DNA primitives every kernel uses), compiled with g++.
// Compares each primitive against plain string code on random reads.  Exit code 0 = pass.
#include "../disco_b200/csrc/dna.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
using namespace disco;

static std::string rc(const std::string &s)
{
    std::string r(s.size(), 'A');
    for (size_t i = 0; i < s.size(); i++) {
        char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    }
    return r;
}
static std::vector<uint64_t> pack(const std::string &s)
{
    std::vector<uint64_t> w((s.size() + 31) / 32, 0);
    for (size_t i = 0; i < s.size(); i++) {
        uint64_t c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
        w[i / 32] |= c << (62 - 2 * (i % 32));
    }
    return w;
}
static std::vector<uint32_t> padded(const std::vector<uint64_t> &w)
{
    std::vector<uint32_t> p(padded_u32((int)w.size()), 0);
    for (size_t i = 0; i < w.size(); i++) pstore(p.data(), (int)i + 1, w[i]);
    return p;
}
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

// string versions of the reference checks (OverlapGraph.cpp:517-595)
static bool ref_dovetail(const std::string &s1, const std::string &s2, int type, int j, int K)
{
    std::string t = (type == 0 || type == 1) ? s2 : rc(s2);
    int L1 = s1.size(), L2 = t.size();
    if (type == 0 || type == 2) {
        if (L1 - j - K >= L2 - K) return false;
        return s1.substr(j + K) == t.substr(K, L1 - (j + K)) && s1.substr(j, K) == t.substr(0, K);
    }
    if (L2 - K < j) return false;
    return s1.substr(0, j) == t.substr(L2 - K - j, j) && s1.substr(j, K) == t.substr(L2 - K, K);
}
static bool ref_contained(const std::string &s1, const std::string &s2, int type, int j, int K)
{
    std::string t = (type == 0 || type == 1) ? s2 : rc(s2);
    int L1 = s1.size(), L2 = t.size();
    if (type == 0 || type == 2) {
        if (L1 - j - K < L2 - K) return false;
        return s1.substr(j, L2) == t;
    }
    if (j < L2 - K) return false;
    return s1.substr(j - (L2 - K), L2) == t;
}

int main()
{
    std::mt19937_64 rng(12345);
    auto rnd_read = [&](int L) { std::string s(L, 'A'); for (auto &c : s) c = "ACGT"[rng() & 3]; return s; };
    // 1. rc_word / fetch64
    for (int L = 30; L <= 300; L += 7) {
        std::string s = rnd_read(L), r = rc(s);
        auto A = padded(pack(s));
        int W = (L + 31) / 32;
        std::vector<uint32_t> R(padded_u32(W), 0);
        for (int w = 0; w < W; w++) pstore(R.data(), w + 1, rc_word(A.data(), L, W, w));
        auto Rref = padded(pack(r));
        CHECK(R == Rref);
        for (int p = -32; p < L; p++) {
            uint64_t x = fetch64(A.data(), p);
            for (int b = 0; b < 32; b++) {
                int pos = p + b;
                uint64_t got = (x >> (62 - 2 * b)) & 3;
                uint64_t exp = (pos >= 0 && pos < L) ? (s[pos] == 'A' ? 0 : s[pos] == 'C' ? 1 : s[pos] == 'G' ? 2 : 3) : 0;
                if (pos < 32 * W) CHECK(got == exp);
            }
        }
    }
    // 2. canonical hash: equal for k-mer and its reverse complement found anywhere in any read; orientation flags
    for (int K : {29, 31, 32, 33, 34, 49, 63, 64, 65, 74, 120}) {
        for (int it = 0; it < 200; it++) {
            int L1 = K + 2 + rng() % 150, L2 = K + 2 + rng() % 150;
            std::string s1 = rnd_read(L1), s2 = rnd_read(L2);
            int j1 = rng() % (L1 - K + 1), j2 = rng() % (L2 - K + 1);
            bool flip = rng() & 1;
            std::string km = s1.substr(j1, K);
            s2.replace(j2, K, flip ? rc(km) : km);
            auto A1 = padded(pack(s1)), R1 = padded(pack(rc(s1))), A2 = padded(pack(s2)), R2 = padded(pack(rc(s2)));
            int f1, f2;
            uint64_t h1 = canon_kmer_hash(A1.data(), R1.data(), L1, j1, K, &f1);
            uint64_t h2 = canon_kmer_hash(A2.data(), R2.data(), L2, j2, K, &f2);
            CHECK(h1 == h2);
            std::string k2 = s2.substr(j2, K);
            bool pal = km == rc(km);
            if (!pal) CHECK((f1 == f2) == (km == k2));
            else CHECK(f1 == 1 && f2 == 1);
            // different k-mer -> different hash (overwhelmingly)
            int j3 = rng() % (L1 - K + 1);
            if (s1.substr(j3, K) != km && s1.substr(j3, K) != rc(km)) {
                int f3; CHECK(canon_kmer_hash(A1.data(), R1.data(), L1, j3, K, &f3) != h1);
            }
        }
    }
    // reverse palindrome (even K): typed forward on both sides
    {
        std::string half = rnd_read(17), pal = half + rc(half);
        std::string s = rnd_read(20) + pal + rnd_read(25);
        auto A = padded(pack(s)), R = padded(pack(rc(s)));
        int f; canon_kmer_hash(A.data(), R.data(), s.size(), 20, 34, &f);
        CHECK(f == 1);
    }
    // 3. dovetail / containment checks against the string versions, planted and random
    long n_dove = 0, n_cont = 0;
    for (int K : {29, 34, 49, 74}) {
        for (int it = 0; it < 3000; it++) {
            int L1 = K + 2 + rng() % 200;
            std::string s1 = rnd_read(L1);
            std::string s2;
            int mode = rng() % 4;
            if (mode == 0) { // s2 overlaps the right end of s1
                int ov = K + 1 + rng() % (L1 - K - 1);
                s2 = s1.substr(L1 - ov) + rnd_read(1 + rng() % 100);
            } else if (mode == 1) { // s2 overlaps the left end
                int ov = K + 1 + rng() % (L1 - K - 1);
                s2 = rnd_read(1 + rng() % 100) + s1.substr(0, ov);
            } else if (mode == 2) { // s2 contained
                int L2 = K + 2 + rng() % (L1 - K - 1);
                if (L2 > L1) L2 = L1;
                s2 = s1.substr(rng() % (L1 - L2 + 1), L2);
            } else s2 = rnd_read(K + 2 + rng() % 200);
            if ((int)s2.size() < K + 2) continue;
            if (rng() & 1) s2 = rc(s2);
            if (rng() % 8 == 0) s2[rng() % s2.size()] = "ACGT"[rng() & 3]; // occasional mismatch
            int L2 = s2.size();
            auto A = padded(pack(s1)), R = padded(pack(rc(s1)));
            auto w2 = pack(s2);
            auto ldf = [&](int w) { return w2[w]; };
            LoaderMatcher<decltype(ldf)> ld{ldf};
            for (int j = 0; j <= L1 - K; j++)
                for (int type = 0; type < 4; type++) {
                    // only call with a genuine k-mer anchor, as the kernels do
                    std::string t = (type == 0 || type == 1) ? s2 : rc(s2);
                    std::string anchor = (type == 0 || type == 2) ? t.substr(0, K) : t.substr(L2 - K, K);
                    bool anch = s1.substr(j, K) == anchor;
                    bool d = check_dovetail(A.data(), R.data(), L1, j, K, type, L2, ld);
                    bool c = check_contained(A.data(), R.data(), L1, j, K, type, L2, ld);
                    bool dref = anch && ref_dovetail(s1, s2, type, j, K);
                    bool cref = anch && ref_contained(s1, s2, type, j, K);
                    CHECK(d == dref);
                    CHECK(c == cref);
                    n_dove += d; n_cont += c;
                }
        }
    }
    CHECK(n_dove > 1000 && n_cont > 1000);
    // 4. encodings
    for (int it = 0; it < 1000; it++) {
        uint64_t nb = rng() % (1ULL << 40); int off = rng() % 32768, o = rng() & 3;
        uint64_t e = make_entry(off, nb, o);
        CHECK(entry_offset(e) == off && entry_nbr(e) == nb && entry_orient(e) == o && !(e & kElimBit));
        CHECK(entry_offset(e | kElimBit) == off);
        uint64_t ri = make_rowinfo(nb, off); CHECK(rowinfo_start(ri) == nb && rowinfo_deg(ri) == (uint32_t)off);
    }
    for (int t1 = 0; t1 < 4; t1++) for (int t2 = 0; t2 < 4; t2++) {
        bool ref = ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) || ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3));
        CHECK(chain_ok(t1, t2) == ref);
    }
    printf("OK dovetail=%ld contained=%ld\n", n_dove, n_cont);
    return 0;
}
