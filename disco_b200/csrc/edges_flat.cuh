// edges_flat.cuh -- the edge pass in its flat, batch-of-32-reads formulation (included by kernels.cu; short reads, the
// default).  Same semantics as k_edges_probe / k_edges_verify (insertAllEdgesOfRead, OverlapGraph.cpp:631-678, and
// checkOverlap, :567-595), different shape:
//
//   k_probe_flat   one warp = 32 consecutive query reads, ONE LANE PER READ while hashing: every lane slides the k-mer
//                  window of its own read one base per step (forward k-mer shifted left, reverse complement shifted
//                  right -- a few funnel shifts instead of re-extracting both strands at each of ~100 positions), so all
//                  32 lanes stay busy whatever the read length.  Positions that pass the presence filter are
//                  ballot-compacted into a warp queue of (fingerprint, position, read); whenever 32 are queued the warp
//                  switches to ONE LANE PER PROBE: 32 independent bucket loads, branch-free tag compare, the tag matches
//                  prefix-summed and appended -- coalesced -- to the batch's candidate list in global memory.
//   k_verify_flat  one warp = the same batch: the 32 query reads staged once (forward + reverse complement), then ONE
//                  LANE PER CANDIDATE straight down the batch's list across read boundaries (two candidates per lane in
//                  flight): 64-byte row fetch, windowed 2-bit compare, survivors scattered into their read's row.
//
// Everything that needs the reference's sequential semantics -- a position with more than `cap` partners, a neighbour
// reached through two positions (OverlapGraph.cpp:656), a chain of more than kFlatWalkMax buckets, more than kFlatParkMax
// candidates -- only FLAGS the read here; k_edges_exact then redoes it exactly as the reference inserts.
#pragma once

namespace disco {

constexpr int kFlatSlice = 8192;    // candidate entries a warp reserves per global atomic
constexpr int kFlatQueue = 64;      // warp queue of filter-passing positions (flushed 32 at a time; a step adds <= 32)
constexpr int kFlatParkMax = 1024;  // candidates one read may have before it is left to the exact path
constexpr int kFlatSet = 2048;      // slots of the verify kernel's (read, neighbour) set
constexpr int kFlatGroupMax = 1400; // candidates verified against one filling of that set
constexpr int kFlatSegs = 4;        // segments a batch's candidate list may consist of
constexpr int kFlatWalkMax = 127;   // buckets of one chain a queued probe may walk (7 bits of its queue entry)

// candidate: [63..59 read within the batch][58..44 position j][33..2 record][1..0 type]
__device__ __forceinline__ uint64_t make_cand(uint32_t local, int j, uint32_t rec, int type)
{
    return ((uint64_t)local << 59) | ((uint64_t)j << 44) | ((uint64_t)rec << 2) | (uint64_t)type;
}
__device__ __forceinline__ uint32_t cand_local(uint64_t c) { return (uint32_t)(c >> 59); }
__device__ __forceinline__ int cand_j(uint64_t c) { return (int)((c >> 44) & 0x7FFF); }
__device__ __forceinline__ uint32_t cand_read(uint64_t c) { return (uint32_t)(c >> 3) & 0x7FFFFFFFu; }
__device__ __forceinline__ int cand_type(uint64_t c) { return (int)(c & 3); }
// segment of a batch's candidate list: [63..20 first entry][19..0 entries]
__device__ __forceinline__ uint64_t make_seg(uint64_t start, uint32_t n) { return (start << 20) | n; }

// ---- sliding k-mer window (KW words, left aligned, bits beyond K bases zero) ----------------------------------------
template <int KW>
__device__ __forceinline__ void slide_fwd(uint64_t (&x)[KW], uint64_t in, int last_sh)
{   // drop the first base, append `in` as base K-1
#pragma unroll
    for (int i = 0; i < KW - 1; i++) x[i] = (x[i] << 2) | (x[i + 1] >> 62);
    x[KW - 1] = (x[KW - 1] << 2) | (in << last_sh);
}
template <int KW>
__device__ __forceinline__ void slide_rc(uint64_t (&y)[KW], uint64_t cin, uint64_t tmask)
{   // reverse complement of the same window: drop the last base, prepend the complement of the new one
#pragma unroll
    for (int i = KW - 1; i > 0; i--) y[i] = (y[i] >> 2) | (y[i - 1] << 62);
    y[0] = (y[0] >> 2) | (cin << 62);
    y[KW - 1] &= tmask;
}
// == canon_kmer_hash_kw (dna.cuh) on a window held in registers: the same fingerprint the table was built with
template <int KW>
__device__ __forceinline__ uint64_t window_hash(const uint64_t (&x)[KW], const uint64_t (&y)[KW], uint64_t seed, int *fwd_is_canon)
{
    bool fwd = true, decided = false;
#pragma unroll
    for (int i = 0; i < KW; i++)
        if (!decided && x[i] != y[i]) { fwd = x[i] < y[i]; decided = true; }
    uint64_t h = seed;
#pragma unroll
    for (int i = 0; i < KW; i++) {
        h = (h ^ (fwd ? x[i] : y[i])) * 0xD6E8FEB86659FD93ULL;
        h ^= h >> 32;
    }
    *fwd_is_canon = fwd;
    return finish_hash(h);
}

__host__ __device__ inline int flat_row_u64(int stride) { return stride | 1; }           // lane-private read words, odd pitch
__host__ __device__ inline size_t probe_flat_words_per_warp(int stride)
{   // read words, queue (fingerprint + meta), per-read counters, segments, control
    return (size_t)32 * flat_row_u64(stride) + kFlatQueue + kFlatQueue / 2 + 16 + kFlatSegs + 2;
}
__host__ __device__ inline int flat_query_u32(int max_len) { return 4 * (((max_len + 31) >> 5) + 2) + 1; } // A + R of one read, odd pitch
__host__ __device__ inline size_t contain_flat_words_per_warp(int max_len)
{   // staged queries + their lengths
    return ((size_t)32 * flat_query_u32(max_len) + 1) / 2 + 16;
}
__host__ __device__ inline size_t verify_flat_words_per_warp(int max_len)
{   // staged queries, neighbour set, row starts, lengths + counters, control
    return ((size_t)32 * flat_query_u32(max_len) + 1) / 2 + kFlatSet / 2 + 32 + 32 + 4 + 96;
}

// dovetail_window + type_to_edge (dna.cuh) without branches: a warp's lanes hold candidates of all four types
__device__ __forceinline__ bool flat_window(int type, int L1, int j, int K, int L2, int *use_rc, int *a, int *b, int *n)
{
    const bool to_end = (type & 1) == 0;            // types 0, 2: the overlap runs to the end of read 1
    const int ov = to_end ? L1 - j : K + j;
    const bool ok = to_end ? (L1 - j < L2) : (j <= L2 - K);
    *use_rc = type >> 1;
    *a = type == 0 ? j : (type == 3 ? L1 - ov : 0);
    *b = (type == 1 || type == 2) ? L2 - ov : 0;    // types 1, 2: the candidate's suffix overlaps
    *n = ov;
    return ok;
}
__device__ __forceinline__ int flat_orient(int type) { return (0x63 >> (2 * type)) & 3; } // 0->3 1->0 2->2 3->1 (OverlapGraph.cpp:660-666)

// bases [a, a+n) of padded array P against bases [b, b+n) of NW candidate words held in registers (RegMatcher's compare)
template <int NW>
__device__ __forceinline__ bool match_regs(const uint64_t (&v)[NW], const uint32_t *P, int a, int b, int n, int p_u32)
{
    const int q0 = a - b + 32;
    const int i0 = q0 >> 4, sh = (q0 & 15) * 2;
    const int wlo = b >> 5, whi = (b + n - 1) >> 5;
    const uint64_t head = ~0ULL >> (2 * (b & 31));
    const int tb = (b + n) & 31;
    const uint64_t tail = tb ? ~(~0ULL >> (2 * tb)) : ~0ULL;
    uint64_t diff = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        int i = i0 + 2 * w;
        i = max(0, min(i, p_u32 - 3));
        const uint32_t w0 = P[i], w1 = P[i + 1], w2 = P[i + 2];
        const uint64_t x = (((uint64_t)fsl32(w1, w0, sh) << 32) | fsl32(w2, w1, sh)) ^ v[w];
        uint64_t m = (w >= wlo && w <= whi) ? ~0ULL : 0ULL;
        if (w == wlo) m &= head;
        if (w == whi) m &= tail;
        diff |= x & m;
    }
    return diff == 0;
}
__device__ __forceinline__ void load_sector(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d, uint64_t pol)
{
    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
}

// ---------------------------------------------------------------------------------------------------------------------
// CONTAIN: the containment pass (markContainedReads, OverlapGraph.cpp:333-505) through the same machinery: every read is a
// query, positions [0, L-K) pruned to those where a read of the shortest length could fit, no exact path (a batch whose
// candidate list outgrows its segments is left to the warp-per-read kernel, flagged through its row infos).
template <int KW, bool SHARDED, bool CONTAIN>
__global__ void __launch_bounds__(kThreads, 4) k_probe_flat(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = p.K, cap = p.cap;
    const int RS = flat_row_u64(p.reads.stride);
    uint64_t *w0 = smem + wib * probe_flat_words_per_warp(p.reads.stride);
    uint64_t *rw = w0 + lane * RS;                                   // this lane's read
    uint64_t *qh = w0 + 32 * RS;                                     // queue: fingerprints
    uint32_t *qm = reinterpret_cast<uint32_t *>(qh + kFlatQueue);    // queue: [31..17 j][16 canonical-is-forward][15..11 lane][10..4 buckets walked][3..0 matches so far]
    uint32_t *cnt = qm + kFlatQueue;                                 // tag matches per read of the batch
    uint64_t *segs = reinterpret_cast<uint64_t *>(cnt + 32);
    uint32_t *ctrl = reinterpret_cast<uint32_t *>(segs + kFlatSegs); // [0] reads flagged for the exact path (bit per lane)
    const unsigned lt_mask = (1u << lane) - 1;
    const uint64_t nbuckets = p.table.nbuckets;
    const int tail = K - 32 * (KW - 1);
    const uint64_t tmask = base_mask(0, tail);
    const int last_sh = 62 - 2 * ((K - 1) & 31);
    const uint64_t seed = 0x9E3779B97F4A7C15ULL ^ (uint64_t)K;
    const uint64_t nbatches = (p.q_hi - p.q_lo + 31) >> 5;
    unsigned n_queries = 0, n_probes = 0, n_buckets = 0;
    unsigned long long blk_cur = 0, blk_end = 0; // this warp's reserved slice of the candidate buffer
    for (;;) {
        unsigned long long bi = 0;
        if (lane == 0) bi = atomicAdd(p.work_counter, 1ULL);
        bi = __shfl_sync(FULL, bi, 0);
        if (bi >= nbatches) break;
        const uint64_t r0 = p.q_lo + (bi << 5), r1 = r0 + lane;
        const bool inrange = r1 < p.q_hi;
        const bool valid = inrange && (CONTAIN || !((__ldg(p.contained_bits + (r1 >> 5)) >> (r1 & 31)) & 1)); // OverlapGraph.cpp:657
        const int L1 = valid ? read_len(p.reads, r1) : 0;
        if (inrange) {
            const uint64_t *src = p.reads.words + r1 * (uint64_t)p.reads.stride;
            for (int w = 0; w < p.reads.stride; w += 2) { // rows are 16-byte aligned
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(src + w));
                rw[w] = v.x; rw[w + 1] = v.y;
            }
        } else {
            for (int w = 0; w < p.reads.stride; w++) rw[w] = 0;
        }
        cnt[lane] = CONTAIN ? (uint32_t)L1 : 0u; // (containment: the batch's read lengths, for the geometry test in flush())
        if (lane == 0) ctrl[0] = 0;
        n_queries += valid;
        __syncwarp();
        // the window at j = 0 and its reverse complement
        uint64_t x[KW], y[KW];
#pragma unroll
        for (int i = 0; i < KW; i++) x[i] = rw[i];
        x[KW - 1] &= tmask;
        {
            uint64_t z[KW + 1];
#pragma unroll
            for (int i = 0; i < KW; i++) z[i] = revcomp64(x[KW - 1 - i]);
            z[KW] = 0;
            const int sh = (32 * KW - K) * 2; // pad bases of the last word, reversed to the front: shift them out
#pragma unroll
            for (int i = 0; i < KW; i++) y[i] = sh ? ((z[i] << sh) | (z[i + 1] >> (64 - sh))) : z[i];
            y[KW - 1] &= tmask;
        }
        const int jmax = L1 - K; // positions [1, L1-K) (OverlapGraph.cpp:638)
        int jtop = jmax;
        for (int o = 16; o; o >>= 1) jtop = max(jtop, __shfl_xor_sync(FULL, jtop, o));
        uint64_t wi = rw[K >> 5] << (2 * (K & 31)); // incoming bases, the next one in the top two bits
        int qn = 0;                                 // queued positions (warp-uniform)
        int nseg = 0;
        bool dead = false;                          // more candidates than kFlatSegs segments hold: whole batch to the exact path
        unsigned long long seg_start = blk_cur;
        unsigned seg_cnt = 0;

        // one lane per queued probe: bucket walk, tag matches appended to the candidate list
        // (a probe whose bucket has no hole goes back into the queue for the next bucket of its chain instead of
        // keeping the whole warp in a second round for one or two lanes)
        auto flush = [&]() {
            const int n = qn < 32 ? qn : 32;
            const bool has = lane < n;
            const int qi = qn - n + lane;
            const uint64_t h = has ? qh[qi] : 0;
            const uint32_t meta = has ? qm[qi] : 0;
            const int j = (int)(meta >> 17), fq = (int)((meta >> 16) & 1);
            const uint32_t src = (meta >> 11) & 31;
            const int walked = (int)((meta >> 4) & 127);
            int pushed = (int)(meta & 15);
            const uint32_t self = (uint32_t)(r0 + src);
            const uint32_t tag = slot_tag(h);
            const uint64_t *slots = p.table.slots;
            uint64_t b;
            if (SHARDED) { const Home home = home_of(p.table, h); slots = home.slots; b = home.b; }
            else b = bucket_of(h, nbuckets);
            b += (uint64_t)walked;
            if (b >= nbuckets) b -= nbuckets;
            uint64_t v[4];
            unsigned mbits = 0;
            bool hole = true;
            if (has) {
                load_bucket(slots, b, v, policy_evict_first());
                n_buckets++;
                hole = false;
#pragma unroll
                for (int q = 0; q < 4; q++) { // branch-free classification of the four slots
                    const bool empty = v[q] == kEmptySlot;
                    hole |= empty;
                    const bool match = !empty && (uint32_t)(v[q] >> 33) == tag && ((uint32_t)v[q] >> 1) != self; // :655
                    mbits |= (unsigned)match << q;
                }
                if (!CONTAIN && mbits && p.skip_contained) { // "ignore contained reads" (HashTable.cpp:533); the bitmap sits in L2
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (((mbits >> q) & 1) && is_contained(p.contained_bits, (uint32_t)v[q] >> 1)) mbits &= ~(1u << q);
                }
                if (!CONTAIN && mbits && cnt[src] > (uint32_t)kFlatParkMax) mbits = 0; // this read goes to the exact path anyway
                if (CONTAIN && mbits) {
                    // only candidates that can be contained at this position are worth listing: read 1 longer, or equal
                    // and earlier in the file (OverlapGraph.cpp:424, :449), and the window inside read 1 (:517-554)
                    const int Lq = (int)cnt[src];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (!((mbits >> q) & 1)) continue;
                        const uint32_t rec = (uint32_t)v[q], r2 = rec >> 1;
                        const int L2 = read_len(p.reads, r2);
                        int use_rc, a, bb, nn;
                        const bool ok = (Lq > L2 || (Lq == L2 && self < r2)) &&
                                        contained_window(cand_type(rec & 1, (int)((v[q] >> 32) & 1) == fq), Lq, j, K, L2, &use_rc, &a, &bb, &nn);
                        if (!ok) mbits &= ~(1u << q);
                    }
                }
            }
            const int c = __popc(mbits);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
            const int tot = __shfl_sync(FULL, incl, 31);
            if (tot && !dead) {
                if (blk_cur + tot > blk_end) { // next slice of the candidate buffer: the list continues in a new segment
                    if (seg_cnt) {
                        if (nseg < kFlatSegs) { if (lane == 0) segs[nseg] = make_seg(seg_start, seg_cnt); nseg++; }
                        else dead = true;
                    }
                    const unsigned long long want = tot > kFlatSlice ? (unsigned long long)tot : (unsigned long long)kFlatSlice;
                    if (lane == 0) blk_cur = atomicAdd(p.cands_cursor, want);
                    blk_cur = __shfl_sync(FULL, blk_cur, 0);
                    blk_end = blk_cur + want;
                    seg_start = blk_cur; seg_cnt = 0;
                }
                if (!dead) {
                    if (blk_cur + tot <= p.cands_cap) {
                        unsigned long long at = blk_cur + (unsigned)(incl - c);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if ((mbits >> q) & 1) {
                                const uint32_t rec = (uint32_t)v[q];
                                p.cands[at++] = make_cand(src, j, rec, cand_type(rec & 1, (int)((v[q] >> 32) & 1) == fq));
                            }
                        }
                    } else if (lane == 0) {
                        atomicOr(p.stats + ST_OVERFLOW, 2ULL);
                    }
                    if (!CONTAIN && c) atomicAdd(&cnt[src], (uint32_t)c);
                    blk_cur += tot; seg_cnt += tot;
                }
            }
            pushed = min(pushed + c, 15);
            bool cont = has && !hole;
            // A chain longer than the queue entry can count (very many copies of one k-mer): containment redoes the batch
            // with the warp-per-read kernel, the edge pass leaves the read to the exact path.  (Long chains as such are no
            // reason for the exact path here -- a re-queued probe costs one more lane slot; with the contained reads still
            // in the table, duplicate-rich data has many chains of a dozen buckets.)
            if (cont && walked + 1 == kFlatWalkMax) { atomicOr(&ctrl[0], 1u << src); cont = false; }
            if (!CONTAIN && has && !cont && pushed > cap) atomicOr(&ctrl[0], 1u << src); // MAX_EDGE_PER_KMER may fire here: exact path
            const unsigned cm = __ballot_sync(FULL, cont); // (also orders this round's queue reads before the writes below)
            if (cont) {
                const int pos = qn - n + __popc(cm & lt_mask);
                qh[pos] = h;
                qm[pos] = (meta & 0xFFFFF800u) | ((uint32_t)(walked + 1) << 4) | (uint32_t)pushed;
            }
            qn = qn - n + __popc(cm);
            __syncwarp();
        };

        // (one loop for the position steps and the final drain of the queue, so that flush() is inlined once)
        // containment: positions [0, L1-K) (OverlapGraph.cpp:401) where some read can fit -- types 0/2 need j + L2 <= L1,
        // types 1/3 need j >= L2 - K, and L2 >= min_len; edges: positions [1, L1-K) (:638)
        const int fit_lo = L1 - p.reads.min_len, fit_hi = p.reads.min_len - K;
        for (int j = CONTAIN ? 0 : 1; j < jtop || qn > 0; j++) {
            if (j < jtop) {
                if (j > 0) {
                    const int t = j + K - 1; // incoming base
                    if ((t & 31) == 0) wi = rw[t >> 5];
                    const uint64_t in = wi >> 62;
                    wi <<= 2;
                    slide_fwd<KW>(x, in, last_sh);
                    slide_rc<KW>(y, 3 - in, tmask);
                }
                bool pass = false;
                uint64_t h = 0;
                int fq = 0;
                if (j < jmax && (!CONTAIN || j <= fit_lo || j >= fit_hi)) {
                    h = window_hash<KW>(x, y, seed, &fq);
                    n_probes++;
                    pass = filter_test(p.table, h);
                }
                const unsigned m = __ballot_sync(FULL, pass);
                if (pass) {
                    const int pos = qn + __popc(m & lt_mask);
                    qh[pos] = h;
                    qm[pos] = ((uint32_t)j << 17) | ((uint32_t)fq << 16) | ((uint32_t)lane << 11);
                }
                qn += __popc(m);
                __syncwarp();
            }
            while (qn >= 32 || (j >= jtop && qn > 0)) flush();
        }
        // ---- close the batch: candidate list segments, one row reserved per read (as long as its candidate count; the
        // verify kernel writes the survivors there and shortens the row), reads for the exact path flagged
        if (seg_cnt) {
            if (nseg < kFlatSegs) { if (lane == 0) segs[nseg] = make_seg(seg_start, seg_cnt); nseg++; }
            else dead = true;
        }
        __syncwarp();
        const uint32_t flagged = ctrl[0];
        uint32_t c = cnt[lane];
        if (CONTAIN) {
            // nothing to reserve: the verify kernel only takes minima.  A batch that could not be listed completely is
            // flagged read by read for the warp-per-read kernel (the row infos are free until the edge pass).
            const bool redo = dead || flagged != 0;
            if (inrange && redo) p.rowinfo[r1] = kInfoExact;
            if (lane < kFlatSegs) p.batchinfo[bi * kFlatSegs + lane] = (!redo && lane < nseg) ? segs[lane] : 0ULL;
            __syncwarp();
            continue;
        }
        const bool exact = valid && (dead || ((flagged >> lane) & 1) || c > (uint32_t)kFlatParkMax);
        if (exact || !valid) c = 0;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        unsigned long long base = 0;
        if (lane == 0 && tot) base = atomicAdd(p.rows_cursor, (unsigned long long)tot);
        base = __shfl_sync(FULL, base, 0);
        if (base + tot > p.rows_cap) {
            if (lane == 0) atomicOr(p.stats + ST_OVERFLOW, 1ULL);
            c = 0; // nothing may be written there
        }
        if (inrange) p.rowinfo[r1] = exact ? kInfoExact : (c ? make_rowinfo(base + (incl - c), c) : 0ULL);
        if (lane < kFlatSegs) p.batchinfo[bi * kFlatSegs + lane] = (!dead && lane < nseg) ? segs[lane] : 0ULL;
        __syncwarp();
    }
    warp_stat_add(p.stats, ST_QUERIES, n_queries);
    warp_stat_add(p.stats, ST_PROBES, n_probes);
    warp_stat_add(p.stats, ST_BUCKETS, n_buckets);
}

// ---------------------------------------------------------------------------------------------------------------------
// SECT: every read has the same length (<= 256 bases) and the tail-sector copy exists -> sector-wise candidate fetches
template <int NW, bool SECT>
__global__ void __launch_bounds__(kThreads, 2) k_verify_flat(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = p.K;
    const int WP = ((p.reads.max_len + 31) >> 5) + 2, PU = 2 * WP, QS = flat_query_u32(p.reads.max_len);
    uint64_t *w0 = smem + wib * verify_flat_words_per_warp(p.reads.max_len);
    uint32_t *qbase = reinterpret_cast<uint32_t *>(w0);                           // read l: A at qbase + l*QS, R at + PU
    uint32_t *set = reinterpret_cast<uint32_t *>(w0 + ((size_t)32 * QS + 1) / 2); // (read, neighbour) keys of verified overlaps
    uint64_t *rstart = reinterpret_cast<uint64_t *>(set + kFlatSet);              // where each read's row starts
    uint32_t *rlen = reinterpret_cast<uint32_t *>(rstart + 32);                   // query lengths
    uint32_t *cnt = rlen + 32;                                                    // survivors per read
    uint32_t *ctrl = cnt + 32;                                                    // [0] reads with a neighbour seen twice
    uint64_t *cq = reinterpret_cast<uint64_t *>(ctrl + 8);                        // compaction queue: 96 candidates
    const uint64_t nbatches = (p.q_hi - p.q_lo + 31) >> 5;
    const uint64_t pol_stream = policy_evict_first();
    const int UL = p.reads.uniform_len;
    // a buffer overflowed in the probe kernel: its candidate lists are incomplete (and may point past the buffer); the host
    // grows the buffers and repeats the pass
    if (*reinterpret_cast<volatile unsigned long long *>(p.stats + ST_OVERFLOW)) return;
    unsigned n_verified = 0, n_hits = 0, maxdeg = 0;
    unsigned long long n_entries = 0;
    for (;;) {
        unsigned long long bi = 0;
        if (lane == 0) bi = atomicAdd(p.work_counter + 1, 1ULL);
        bi = __shfl_sync(FULL, bi, 0);
        if (bi >= nbatches) break;
        const uint64_t r1 = p.q_lo + (bi << 5) + lane;
        const bool inrange = r1 < p.q_hi;
        const uint64_t ri = inrange ? p.rowinfo[r1] : 0ULL;
        const uint32_t c = (ri & kInfoExact) ? 0u : rowinfo_deg(ri);
        const unsigned live = __ballot_sync(FULL, c != 0); // reads whose candidates are verified here
        if (!live) continue;
        uint64_t mysegs = lane < kFlatSegs ? p.batchinfo[bi * kFlatSegs + lane] : 0ULL;
        // stage the batch: lane l packs read l (forward and reverse complement, padded: see dna.cuh)
        uint32_t *A = qbase + lane * QS, *R = A + PU;
        const int L1 = c ? read_len(p.reads, r1) : 0;
        if (c) {
            const int W = (L1 + 31) >> 5;
            const uint64_t *src = p.reads.words + r1 * (uint64_t)p.reads.stride;
            pstore(A, 0, 0ULL); pstore(R, 0, 0ULL);
            for (int w = 1; w <= W; w++) pstore(A, w, __ldg(src + (w - 1)));
            for (int w = W + 1; w < WP; w++) { pstore(A, w, 0ULL); pstore(R, w, 0ULL); }
            for (int w = 1; w <= W; w++) pstore(R, w, rc_word(A, L1, W, w - 1));
        }
        rstart[lane] = rowinfo_start(ri);
        rlen[lane] = (uint32_t)L1;
        cnt[lane] = 0;
        if (lane == 0) ctrl[0] = 0;
        // groups of consecutive reads whose candidates fit one filling of the set
        uint32_t pre = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
        __syncwarp();
        int g0 = 0;
        while (g0 < 32) {
            const uint32_t before = g0 ? __shfl_sync(FULL, pre, g0 - 1) : 0u;
            const unsigned fits = __ballot_sync(FULL, lane >= g0 && pre - before <= (uint32_t)kFlatGroupMax);
            // lanes >= g0 that fit form a prefix of [g0, 32) (pre is non-decreasing); a single read always fits
            int g1 = g0 + __popc(fits);
            if (g1 == g0) g1 = g0 + 1;
            const unsigned gmask = (g1 >= 32 ? 0xFFFFFFFFu : ((1u << g1) - 1u)) & ~((1u << g0) - 1u) & live;
            const uint32_t gtot = __shfl_sync(FULL, pre, g1 - 1) - before;
            g0 = g1;
            if (gtot == 0) continue;
            for (int k = lane; k < kFlatSet; k += 32) set[k] = 0xFFFFFFFFu;
            __syncwarp();
            // The group's candidates are compacted through a small queue (a batch whose reads have many candidates is
            // verified in several groups, each of which picks its own reads' candidates out of the batch's list), so that
            // the expensive part below always runs with every lane busy: 64 candidates at a time, two per lane.
            int qn = 0;
            auto process = [&](const int take) {
                // two candidates per lane: both rows requested before either is compared.  Equal-length reads with
                // the tail copy: an overlap of up to 128 bases is one 32-byte sector -- the head of the candidate's
                // row (its prefix overlaps) or its tail sector (its suffix overlaps); otherwise the whole row.
                    uint64_t cd[2], v[2][NW];
                    bool act[2];
                    int ua[2], ub[2], un[2], urc[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        act[u] = 32 * u + lane < take;
                        cd[u] = act[u] ? cq[qn - take + 32 * u + lane] : 0ULL;
#pragma unroll
                        for (int w = 0; w < NW; w++) v[u][w] = 0;
                        if (act[u]) {
                            const uint32_t local = cand_local(cd[u]), r2 = cand_read(cd[u]);
                            const uint64_t *row = p.reads.words + (uint64_t)r2 * (uint64_t)p.reads.stride;
                            if constexpr (SECT && NW >= 6) {
                                const int type = cand_type(cd[u]);
                                act[u] = flat_window(type, UL, cand_j(cd[u]), K, UL, &urc[u], &ua[u], &ub[u], &un[u]);
                                const bool sfx = ub[u] != 0; // (types 1, 2; b = L2 - ov > 0 since ov < L2)
                                if (act[u] && un[u] <= 128) {
                                    if (sfx) { row = p.reads.tails + (uint64_t)r2 * 4; ub[u] = 128 - un[u]; }
                                    load_sector(row, v[u][0], v[u][1], v[u][2], v[u][3], pol_stream);
                                } else if (act[u]) {
                                    load_sector(row, v[u][0], v[u][1], v[u][2], v[u][3], pol_stream);
                                    if constexpr (NW == 6) asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v[u][4]), "=l"(v[u][5]) : "l"(row + 4), "l"(pol_stream));
                                    if constexpr (NW == 8) load_sector(row + 4, v[u][4], v[u][5], v[u][6], v[u][7], pol_stream);
                                }
                                (void)local;
                            } else {
#pragma unroll
                                for (int w = 0; w < NW / 2; w++)
                                    asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v[u][2 * w]), "=l"(v[u][2 * w + 1]) : "l"(row + 2 * w), "l"(pol_stream));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        if (!act[u]) continue;
                        const uint32_t local = cand_local(cd[u]), r2 = cand_read(cd[u]);
                        const int j = cand_j(cd[u]), type = cand_type(cd[u]);
                        const int Lq = SECT ? UL : (int)rlen[local];
                        n_verified++;
                        if (!SECT) {
                            const int L2 = read_len(p.reads, r2);
                            if (!flat_window(type, Lq, j, K, L2, &urc[u], &ua[u], &ub[u], &un[u])) continue;
                        }
                        const uint32_t *P = qbase + local * QS + (urc[u] ? PU : 0);
                        if (!match_regs<NW>(v[u], P, ua[u], ub[u], un[u], PU)) continue;
                        // first hit per neighbour (OverlapGraph.cpp:656): a second verified overlap with the same read
                        // sends the query to the exact path
                        const uint32_t key = (r2 * 0x9E3779B1u) ^ (local * 0x85EBCA6Bu);
                        uint32_t hh = (key >> 11) & (kFlatSet - 1);
                        for (int tries = 0;; tries++) {
                            const uint32_t old = atomicCAS(&set[hh], 0xFFFFFFFFu, key);
                            if (old == 0xFFFFFFFFu) break;
                            if (old == key || tries == kFlatSet) { atomicOr(&ctrl[0], 1u << local); break; }
                            hh = (hh + 1) & (kFlatSet - 1);
                        }
                        const uint32_t idx = atomicAdd(&cnt[local], 1u);
                        const int ovl = (type & 1) ? K + j : Lq - j;
                        p.rows[rstart[local] + idx] = make_entry(Lq - ovl, r2, flat_orient(type));
                    }
                __syncwarp();
                qn -= take;
            };
            int sg = 0;
            uint32_t i0 = 0;
            bool more = true;
            while (more || qn > 0) {
                if (more) {
                    const uint64_t seg = __shfl_sync(FULL, mysegs, sg);
                    const uint32_t sn = (uint32_t)(seg & 0xFFFFF);
                    if (!sn) { more = false; }
                    else {
                        const uint32_t i = i0 + lane;
                        const uint64_t cnd = i < sn ? __ldg(p.cands + (seg >> 20) + i) : 0ULL;
                        const bool mine = i < sn && ((gmask >> cand_local(cnd)) & 1);
                        const unsigned m = __ballot_sync(FULL, mine);
                        if (mine) cq[qn + __popc(m & ((1u << lane) - 1))] = cnd;
                        qn += __popc(m);
                        __syncwarp();
                        i0 += 32;
                        if (i0 >= sn) { i0 = 0; if (++sg == kFlatSegs) more = false; }
                    }
                }
                if (qn >= 64 || (!more && qn > 0)) process(qn < 64 ? qn : 64);
            }
            __syncwarp();
        }
        // ---- rows are final: shorten them to the survivors; reads with a doubly reached neighbour go to the exact path
        if (c) {
            const bool dup = (ctrl[0] >> lane) & 1;
            const uint32_t deg = cnt[lane];
            p.rowinfo[r1] = dup ? kInfoExact : (deg ? make_rowinfo(rowinfo_start(ri), deg) : 0ULL);
            if (!dup) {
                n_hits += deg; n_entries += deg;
                if (deg > maxdeg) maxdeg = deg;
            }
        }
        __syncwarp();
    }
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_HITS, n_hits);
    warp_stat_add(p.stats, ST_ENTRIES, n_entries);
    for (int o = 16; o; o >>= 1) { unsigned t = __shfl_xor_sync(FULL, maxdeg, o); if (t > maxdeg) maxdeg = t; }
    if (lane == 0 && maxdeg) atomicMax(p.stats + ST_MAXDEG, (unsigned long long)maxdeg);
}

// checkOverlapForContainedRead (OverlapGraph.cpp:517-554) for the candidates k_probe_flat<CONTAIN> listed: one lane per
// candidate, two in flight; a verified containment elects the container with one atomicMin (dna.cuh: make_ckey).
template <int NW>
__global__ void __launch_bounds__(kThreads, 3) k_contain_verify_flat(SearchParams p)
{
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = p.K;
    const int WP = ((p.reads.max_len + 31) >> 5) + 2, PU = 2 * WP, QS = flat_query_u32(p.reads.max_len);
    uint64_t *w0 = smem + wib * contain_flat_words_per_warp(p.reads.max_len);
    uint32_t *qbase = reinterpret_cast<uint32_t *>(w0);
    uint32_t *rlen = reinterpret_cast<uint32_t *>(w0 + ((size_t)32 * QS + 1) / 2);
    const uint64_t nbatches = (p.q_hi - p.q_lo + 31) >> 5;
    const uint64_t pol_stream = policy_evict_first();
    unsigned n_verified = 0, n_hits = 0;
    if (*reinterpret_cast<volatile unsigned long long *>(p.stats + ST_OVERFLOW)) return; // candidate lists incomplete: the pass is repeated
    for (;;) {
        unsigned long long bi = 0;
        if (lane == 0) bi = atomicAdd(p.work_counter + 1, 1ULL);
        bi = __shfl_sync(FULL, bi, 0);
        if (bi >= nbatches) break;
        const uint64_t mysegs = lane < kFlatSegs ? p.batchinfo[bi * kFlatSegs + lane] : 0ULL;
        if (!__any_sync(FULL, mysegs != 0)) continue;
        const uint64_t r0 = p.q_lo + (bi << 5), r1 = r0 + lane;
        const bool inrange = r1 < p.q_hi;
        uint32_t *A = qbase + lane * QS, *R = A + PU;
        const int L1 = inrange ? read_len(p.reads, r1) : 0;
        if (inrange) {
            const int W = (L1 + 31) >> 5;
            const uint64_t *src = p.reads.words + r1 * (uint64_t)p.reads.stride;
            pstore(A, 0, 0ULL); pstore(R, 0, 0ULL);
            for (int w = 1; w <= W; w++) pstore(A, w, __ldg(src + (w - 1)));
            for (int w = W + 1; w < WP; w++) { pstore(A, w, 0ULL); pstore(R, w, 0ULL); }
            for (int w = 1; w <= W; w++) pstore(R, w, rc_word(A, L1, W, w - 1));
        }
        rlen[lane] = (uint32_t)L1;
        __syncwarp();
        for (int sg = 0; sg < kFlatSegs; sg++) {
            const uint64_t seg = __shfl_sync(FULL, mysegs, sg);
            const uint32_t sn = (uint32_t)(seg & 0xFFFFF);
            if (!sn) break;
            const uint64_t *cl = p.cands + (seg >> 20);
            for (uint32_t i0 = 0; i0 < sn; i0 += 64) {
                uint64_t cd[2], v[2][NW];
                int l2[2];
                bool act[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint32_t i = i0 + 32 * u + lane;
                    act[u] = i < sn;
                    cd[u] = act[u] ? __ldg(cl + i) : 0ULL;
                    l2[u] = 0;
#pragma unroll
                    for (int w = 0; w < NW; w++) v[u][w] = 0;
                    if (act[u]) {
                        const uint32_t r2 = cand_read(cd[u]);
                        l2[u] = read_len(p.reads, r2);
                        const uint64_t *row = p.reads.words + (uint64_t)r2 * (uint64_t)p.reads.stride;
#pragma unroll
                        for (int w = 0; w < NW / 2; w++)
                            asm volatile("ld.global.nc.L2::cache_hint.L2::64B.v2.u64 {%0,%1}, [%2], %3;" : "=l"(v[u][2 * w]), "=l"(v[u][2 * w + 1]) : "l"(row + 2 * w), "l"(pol_stream));
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (!act[u]) continue;
                    const uint32_t local = cand_local(cd[u]), r2 = cand_read(cd[u]);
                    const uint64_t rq = r0 + local;
                    const int j = cand_j(cd[u]), type = cand_type(cd[u]);
                    const int Lq = (int)rlen[local], L2 = l2[u];
                    // read1 must be longer, or equal and earlier in the file (OverlapGraph.cpp:424, :449)
                    if (!(Lq > L2 || (Lq == L2 && rq < r2))) continue;
                    n_verified++;
                    int use_rc, a, b, n;
                    if (!contained_window(type, Lq, j, K, L2, &use_rc, &a, &b, &n)) continue;
                    const uint32_t *P = qbase + local * QS + (use_rc ? PU : 0);
                    if (!match_regs<NW>(v[u], P, a, b, n, PU)) continue;
                    n_hits++;
                    const uint32_t rec = (uint32_t)(cd[u] >> 2);
                    atomicMin(p.best + r2, (unsigned long long)make_ckey(rq, j, (int)(rec & 1), type));
                }
            }
        }
        __syncwarp();
    }
    warp_stat_add(p.stats, ST_VERIFIED, n_verified);
    warp_stat_add(p.stats, ST_HITS, n_hits);
}

} // namespace disco
