// dna.cuh -- 2-bit packed DNA primitives shared by every kernel (and unit-tested on the host by tests/csrc_host_test.cpp).
//
// Packed layout = the reference's record payload (src/BuildGraph/src/HashTable.cpp:456-477):
//   base i of a read lives in bits [62-2*(i%32), 63-2*(i%32)] of 64-bit word i/32 (MSB first), A=0 C=1 G=2 T=3
//   (HashTable.h:16-22); unused tail bits are 0.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DHD __host__ __device__ __forceinline__
#else
#define DHD inline
#endif

namespace disco {

constexpr uint64_t kEmptySlot = ~0ULL;

DHD uint64_t brev64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
#endif
}

// reverse complement of the 32 bases held in one word (complement = bitwise NOT with A0 C1 G2 T3)
DHD uint64_t revcomp64(uint64_t x)
{
    uint64_t y = brev64(~x);
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
}

// ---- padded read arrays --------------------------------------------------------------------------------------------
// A padded array P holds a read for random base-offset access as 32-bit half-words in BASE ORDER (16 bases each):
//   P[2k] = high half, P[2k+1] = low half of padded word k;  padded word 0 = 0, words 1..W = the read, word W+1 = 0.
// 32-bit granularity lets every unaligned fetch be two single-instruction funnel shifts (SHF) instead of 64-bit
// shift/or sequences.
DHD uint32_t fsl32(uint32_t lo, uint32_t hi, int s) // upper 32 bits of (hi:lo) << s, 0 <= s < 32
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
DHD uint64_t pword(const uint32_t *P, int k) { return ((uint64_t)P[2 * k] << 32) | P[2 * k + 1]; }
DHD void pstore(uint32_t *P, int k, uint64_t w) { P[2 * k] = (uint32_t)(w >> 32); P[2 * k + 1] = (uint32_t)w; }
DHD int padded_u32(int words) { return 2 * (words + 2); }

// 32 bases starting at base position p (-32 <= p < 32*W) of a padded array; positions outside the read give 0 bits
DHD uint64_t fetch64(const uint32_t *P, int p)
{
    const int q = p + 32;
    const int i = q >> 4, s = (q & 15) * 2;
    const uint32_t w0 = P[i], w1 = P[i + 1], w2 = P[i + 2];
    return ((uint64_t)fsl32(w1, w0, s) << 32) | fsl32(w2, w1, s);
}

// mask selecting bases [lo, hi) of a word, 0 <= lo < hi <= 32
DHD uint64_t base_mask(int lo, int hi)
{
    uint64_t m = (lo == 0) ? ~0ULL : (~0ULL >> (2 * lo));
    if (hi < 32) m &= ~(~0ULL >> (2 * hi));
    return m;
}

// word w (0-based) of the reverse complement of a read of L bases / W words held in padded array A.
// Z[i] = revcomp64(word W-1-i) spells [32W-L pad T's][rc(read)]; shift the pad out.
DHD uint64_t rc_word(const uint32_t *A, int L, int W, int w)
{
    int sh = (32 * W - L) * 2;
    uint64_t z0 = revcomp64(pword(A, W - w));
    if (sh == 0) return z0;
    uint64_t z1 = (w + 1 < W) ? revcomp64(pword(A, W - w - 1)) : 0ULL;
    return (z0 << sh) | (z1 >> (64 - sh));
}

DHD uint64_t mix64(uint64_t h)
{ // murmur3 fmix64
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

// Canonical fingerprint of the K-mer starting at base j of the read (A = forward padded, R = reverse-complement padded,
// L = read length).  Canonical form = the smaller of the k-mer and its reverse complement in packed (lexicographic)
// order -- the role getHashIndex()'s min() plays in the reference (HashTable.cpp:383-391).  *fwd_is_canon tells
// which one won (ties -- reverse palindromes -- count as forward, matching the reference's "if / else if" typing).
// final fold of the per-word multiply/xorshift chain
DHD uint64_t finish_hash(uint64_t h)
{   // (every word already went through a multiply and a fold of the high half into the low one; a second multiply here
    // was 7% of the probe kernel's instructions and bought nothing measurable in chain lengths or tag collisions)
    h ^= h >> 29;
    return h;
}

// K-mer of KW words (KW static): both strands in registers, decide, hash the winner
template <int KW>
DHD uint64_t canon_kmer_hash_kw(const uint32_t *A, const uint32_t *R, int j, int jr, uint64_t tmask, uint64_t seed, int *fwd_is_canon)
{
    uint64_t x[KW], y[KW];
#pragma unroll
    for (int i = 0; i < KW; i++) { x[i] = fetch64(A, j + 32 * i); y[i] = fetch64(R, jr + 32 * i); }
    x[KW - 1] &= tmask; y[KW - 1] &= tmask;
    bool fwd = true, decided = false;
#pragma unroll
    for (int i = 0; i < KW; i++) {
        if (!decided && x[i] != y[i]) { fwd = x[i] < y[i]; decided = true; }
    }
    uint64_t h = seed;
#pragma unroll
    for (int i = 0; i < KW; i++) { // one multiply per word
        h = (h ^ (fwd ? x[i] : y[i])) * 0xD6E8FEB86659FD93ULL;
        h ^= h >> 32;
    }
    *fwd_is_canon = fwd;
    return finish_hash(h);
}

DHD uint64_t canon_kmer_hash(const uint32_t *A, const uint32_t *R, int L, int j, int K, int *fwd_is_canon)
{
    const int KW = (K + 31) >> 5;
    const int tail = K - 32 * (KW - 1);
    const uint64_t tmask = base_mask(0, tail);
    const int jr = L - j - K; // the k-mer's reverse complement starts here in rc(read)
    const uint64_t seed = 0x9E3779B97F4A7C15ULL ^ (uint64_t)K;
    switch (KW) { // uniform branch; the common sizes run fully unrolled in registers
    case 1: return canon_kmer_hash_kw<1>(A, R, j, jr, tmask, seed, fwd_is_canon);
    case 2: return canon_kmer_hash_kw<2>(A, R, j, jr, tmask, seed, fwd_is_canon);
    case 3: return canon_kmer_hash_kw<3>(A, R, j, jr, tmask, seed, fwd_is_canon);
    case 4: return canon_kmer_hash_kw<4>(A, R, j, jr, tmask, seed, fwd_is_canon);
    default: break;
    }
    int fwd = 1;
    for (int i = 0; i < KW; i++) {
        uint64_t x = fetch64(A, j + 32 * i), y = fetch64(R, jr + 32 * i);
        if (i == KW - 1) { x &= tmask; y &= tmask; }
        if (x != y) { fwd = x < y; break; }
    }
    const uint32_t *S = fwd ? A : R;
    const int s = fwd ? j : jr;
    uint64_t h = seed;
    for (int i = 0; i < KW; i++) {
        uint64_t x = fetch64(S, s + 32 * i);
        if (i == KW - 1) x &= tmask;
        h = (h ^ x) * 0xD6E8FEB86659FD93ULL;
        h ^= h >> 32;
    }
    *fwd_is_canon = fwd;
    return finish_hash(h);
}

// Compare n bases: padded array P (query side) from base a, against plain word array s2 (candidate, forward strand)
// from base b.  Returns true when all n bases agree.
template <typename WordLoader>
DHD bool match_window(const uint32_t *P, int a, WordLoader s2, int b, int n)
{
    if (n <= 0) return true;
    const int wlo = b >> 5, whi = (b + n - 1) >> 5;
    uint64_t diff = 0;
    for (int w = wlo; w <= whi; w++) {
        int lo = b - 32 * w; if (lo < 0) lo = 0;
        int hi = b + n - 32 * w; if (hi > 32) hi = 32;
        uint64_t x = fetch64(P, a + 32 * w - b) ^ s2(w);
        diff |= x & base_mask(lo, hi);
    }
    return diff == 0;
}

// --- overlap geometry (types as in HashTable::getListOfReads, HashTable.cpp:535-566) -------------------------------
// type 0: query k-mer == prefix(s2)      type 3: == rc(prefix(s2))  (t = rc(s2) ends with it)
// type 1: query k-mer == suffix(s2)      type 2: == rc(suffix(s2))  (t = rc(s2) starts with it)
DHD int cand_type(int kind /*0 prefix rec, 1 suffix rec*/, bool same_strand)
{
    return kind == 0 ? (same_strand ? 0 : 3) : (same_strand ? 1 : 2);
}

// edge orientation and overlap length from (type, j): OverlapGraph.cpp:660-666
DHD void type_to_edge(int type, int L1, int K, int j, int *orient, int *ovl)
{
    switch (type) {
    case 0: *orient = 3; *ovl = L1 - j; break;
    case 1: *orient = 0; *ovl = K + j; break;
    case 2: *orient = 2; *ovl = L1 - j; break;
    default: *orient = 1; *ovl = K + j; break;
    }
}

// A "matcher" m(P, a, b, n) answers: do bases [a, a+n) of padded array P equal bases [b, b+n) of the candidate read?
// (LoaderMatcher walks the candidate's words through a loader; the kernels use a register-resident variant.)
template <typename WordLoader>
struct LoaderMatcher {
    WordLoader s2;
    DHD bool operator()(const uint32_t *P, int a, int b, int n) const { return match_window(P, a, s2, b, n); }
};

// Window to compare for a candidate: bases [a, a+n) of the query (reverse-complement array when *use_rc) against bases
// [b, b+n) of the candidate's forward strand.  Returns false when the geometry already rules the candidate out.
// One parameter set + ONE matcher call keeps a warp converged whatever mix of types its lanes hold.
//
// checkOverlap (OverlapGraph.cpp:567-595): dovetail test of the query read (length L1) at k-mer position j against a
// candidate of length L2.  The whole overlap -- k-mer included -- is compared, so a fingerprint collision can never
// produce an edge.
DHD bool dovetail_window(int type, int L1, int j, int K, int L2, int *use_rc, int *a, int *b, int *n)
{
    if (type == 0 || type == 2) {
        if (L1 - j >= L2) return false; // the overlap must run to the end of read1 and stop inside read2
        const int ov = L1 - j;
        *n = ov;
        if (type == 0) { *use_rc = 0; *a = j; *b = 0; }      // s1[j..L1) == s2[0..ov)
        else { *use_rc = 1; *a = 0; *b = L2 - ov; }          // rc(s1)[0..ov) == s2[L2-ov..L2)
        return true;
    }
    if (L2 - K < j) return false;
    const int ov = K + j;
    *n = ov;
    if (type == 1) { *use_rc = 0; *a = 0; *b = L2 - ov; }    // s1[0..ov) == s2[L2-ov..L2)
    else { *use_rc = 1; *a = L1 - ov; *b = 0; }              // rc(s1)[L1-ov..L1) == s2[0..ov)
    return true;
}

// checkOverlapForContainedRead (OverlapGraph.cpp:517-554): is the whole of s2 (or its reverse complement) inside s1,
// anchored by the k-mer hit at position j?
DHD bool contained_window(int type, int L1, int j, int K, int L2, int *use_rc, int *a, int *b, int *n)
{
    *b = 0; *n = L2;
    if (type == 0 || type == 2) {
        if (j + L2 > L1) return false;
        if (type == 0) { *use_rc = 0; *a = j; } else { *use_rc = 1; *a = L1 - j - L2; }
        return true;
    }
    if (j < L2 - K) return false;
    if (type == 1) { *use_rc = 0; *a = j - (L2 - K); } else { *use_rc = 1; *a = L1 - j - K; }
    return true;
}

template <typename Matcher>
DHD bool check_dovetail(const uint32_t *A, const uint32_t *R, int L1, int j, int K, int type, int L2, const Matcher &m)
{
    int use_rc, a, b, n;
    if (!dovetail_window(type, L1, j, K, L2, &use_rc, &a, &b, &n)) return false;
    return m(use_rc ? R : A, a, b, n);
}

template <typename Matcher>
DHD bool check_contained(const uint32_t *A, const uint32_t *R, int L1, int j, int K, int type, int L2, const Matcher &m)
{
    int use_rc, a, b, n;
    if (!contained_window(type, L1, j, K, L2, &use_rc, &a, &b, &n)) return false;
    return m(use_rc ? R : A, a, b, n);
}

// --- hash table slot / CSR entry / key encodings ---------------------------------------------------------------------
// slot  : [63..33 tag][32 kmer-is-canonical-forward][31..0 rec = 2*read + kind]      (EMPTY = all ones)
DHD uint64_t make_slot(uint64_t hash, int fwd_is_canon, uint32_t rec)
{
    return ((hash & 0x7FFFFFFFULL) << 33) | ((uint64_t)(fwd_is_canon & 1) << 32) | rec;
}
DHD uint32_t slot_tag(uint64_t hash) { return (uint32_t)(hash & 0x7FFFFFFFULL); }
DHD uint64_t bucket_of(uint64_t hash, uint64_t nbuckets)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(hash, nbuckets);
#else
    return (uint64_t)(((unsigned __int128)hash * nbuckets) >> 64);
#endif
}

// CSR entry: [63 eliminated][62..48 offset][47..2 neighbour (0-based)][1..0 orientation]; sorts by (offset, id, orient)
constexpr uint64_t kElimBit = 1ULL << 63;
DHD uint64_t make_entry(int offset, uint64_t nbr, int orient)
{
    return ((uint64_t)offset << 48) | (nbr << 2) | (uint64_t)orient;
}
DHD int entry_offset(uint64_t e) { return (int)((e >> 48) & 0x7FFF); }
DHD uint64_t entry_nbr(uint64_t e) { return (e >> 2) & 0x3FFFFFFFFFFFULL; }
DHD int entry_orient(uint64_t e) { return (int)(e & 3); }

// twin edge (OverlapGraph.cpp:617, :770-784)
DHD int twin_orient(int o) { return o == 0 ? 3 : (o == 3 ? 0 : o); }

// Myers chain test (OverlapGraph.cpp:705-708)
DHD bool chain_ok(int t1, int t2)
{
    return (((t1 == 0) | (t1 == 2)) & ((t2 == 0) | (t2 == 1))) | (((t1 == 1) | (t1 == 3)) & ((t2 == 2) | (t2 == 3)));
}

// row info: [63 exact-path pending][62 crowded position][61..20 first entry][19..0 degree]
// (the two flag bits only live between the kernels of the edge pass; finished rows have them clear)
constexpr uint64_t kInfoExact = 1ULL << 63;   // redo this read with the exact sequential search
constexpr uint64_t kInfoCrowded = 1ULL << 62; // some position has more than `cap` tag matches: count per position
DHD uint64_t make_rowinfo(uint64_t start, uint32_t deg) { return (start << 20) | deg; }
DHD uint64_t rowinfo_start(uint64_t ri) { return (ri >> 20) & ((1ULL << 42) - 1); }
DHD uint32_t rowinfo_deg(uint64_t ri) { return (uint32_t)(ri & 0xFFFFF); }

// containment key: smallest (container, j, record kind) wins == the reference's sequential -t 1 attribution
// [63..20 container (0-based)][19..4 j][3..2 kind][1..0 type]
DHD uint64_t make_ckey(uint64_t r1, int j, int kind, int type)
{
    return (r1 << 20) | ((uint64_t)j << 4) | ((uint64_t)kind << 2) | (uint64_t)type;
}

} // namespace disco
