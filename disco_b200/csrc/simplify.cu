// simplify.cu -- the first consumer step of the hot path's output, on the reduced edge list the context still holds in HBM:
// parsimplify's composite-edge contraction and dead-end removal (src/SimplifyGraph/src/OverlapGraphSimple.cpp:
// contractParCompositeEdges :313-505, contractParCompositeEdges_Serial :69-112, removeParDeadEndNodes :135-218, the
// constructor's loop :236-244, printEdge :658-690; EdgeSimple.cpp: Add / merge_forward_edges / mergeList :160-245,
// is_mergeable :247-259, make_nonComposite_reverseEdge :102-111).
//
// The reference merges edge objects pairwise and rebuilds read lists on every merge.  Here nothing is merged: the graph
// after any number of contractions is determined by the set of ALIVE original edges ("atoms", two directed half-edges
// each) -- a node the reference would have contracted away is a node with exactly two alive half-edges whose
// orientations pass through it (is_mergeable), and a composite edge is a maximal chain of atoms through such nodes.  So a
// round is: degrees + contractible flags (one thread per node), list ranking of the chains by pointer jumping (one
// thread per half-edge: atoms to the end of the chain, offset sum, last atom), dead-end test per node on the chain
// aggregates, removal of the chains that touch a dead-end node; rounds repeat until nothing is removed (the reference's
// do { contract; remove dead ends } while (changed)).  The output pass places every inner read by its rank in its chain.
//
// Not reproduced: chains that close on themselves without a branching node (isolated cycles).  The reference breaks them
// wherever its node order happens to start; here their atoms are emitted unmerged and counted (cycle_atoms).
#include "../../include/disco_gpu.h"
#include "dna.cuh"
#include "kernels.cuh"
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace disco {

namespace {

// one half-edge's view of its chain: successor, last atom, atoms and offset sum up to the end -- one 32-byte sector, so
// that a pointer-jumping step costs one random access per half-edge
struct __align__(32) SLink {
    int32_t nxt, last;
    uint32_t cnt, pad;
    uint64_t sum, pad2;
};

__device__ __forceinline__ int twin_o(int o) { return ((o >> 1) ^ 1) | (((o & 1) ^ 1) << 1); } // EdgeSimple.cpp:261-267
__device__ __forceinline__ int rlen_of(const uint16_t *len, int uniform, uint32_t r) { return uniform ? uniform : (int)len[r]; }

// half-edge 2e = edge e as stored (src -> dst), 2e+1 = its reverse (make_nonComposite_reverseEdge)
__global__ void k_s_atoms(const disco_edge *e, uint64_t ne, const uint16_t *len, int uniform, uint32_t min_ovl,
                          uint32_t *a_src, uint32_t *a_dst, uint32_t *a_off, uint8_t *a_or, uint8_t *alive, unsigned long long *deg_all)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const disco_edge x = e[i];
    const int ls = rlen_of(len, uniform, x.src), ld = rlen_of(len, uniform, x.dst);
    const uint8_t ok = (uint32_t)(ls - (int)x.offset) >= min_ovl; // loadParEdgesFromEdgeFile: overlapLength >= m_minOvl (:572)
    a_src[2 * i] = x.src; a_dst[2 * i] = x.dst; a_off[2 * i] = x.offset; a_or[2 * i] = (uint8_t)x.orient; alive[2 * i] = ok;
    a_src[2 * i + 1] = x.dst; a_dst[2 * i + 1] = x.src; a_off[2 * i + 1] = (uint32_t)(ld + (int)x.offset - ls);
    a_or[2 * i + 1] = (uint8_t)twin_o((int)x.orient); alive[2 * i + 1] = ok;
    atomicAdd(deg_all + x.src, 1ULL);
    atomicAdd(deg_all + x.dst, 1ULL);
}

// CSR fill: half-edges grouped by source node (order inside a node is irrelevant)
__global__ void k_s_fill(const uint32_t *a_src, uint64_t nh, const uint64_t *row, uint32_t *fillc, uint32_t *adj)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    const uint32_t s = a_src[h];
    adj[row[s] + atomicAdd(fillc + s, 1u)] = (uint32_t)h;
}

// per node: alive degree, its first two alive half-edges, and whether the reference would contract it away
// (exactly two edges, orientation passes through: is_mergeable(into v, out of v), EdgeSimple.cpp:247-259)
__global__ void k_s_nodes(uint64_t n, const uint64_t *row, const uint32_t *adj, const uint8_t *alive, const uint8_t *a_or,
                          uint32_t *deg, int32_t *slot, uint8_t *contractible)
{
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint32_t d = 0;
    int32_t s0 = -1, s1 = -1;
    for (uint64_t k = row[v]; k < row[v + 1]; k++) {
        const uint32_t h = adj[k];
        if (!alive[h]) continue;
        if (d == 0) s0 = (int32_t)h; else if (d == 1) s1 = (int32_t)h;
        d++;
    }
    deg[v] = d; slot[2 * v] = s0; slot[2 * v + 1] = s1;
    uint8_t c = 0;
    if (d == 2) {
        const int o_in = a_or[s0 ^ 1], o_out = a_or[s1]; // (a -> v) = reverse of (v -> a)
        c = (o_in & 1) == ((o_out >> 1) & 1);
    }
    contractible[v] = c;
}

// chain links: the half-edge that continues h through a contractible destination
__global__ void k_s_links(uint64_t nh, const uint32_t *a_dst, const uint32_t *a_off, const uint8_t *alive, const uint8_t *contractible,
                          const int32_t *slot, SLink *L)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    int32_t nx = -1;
    if (alive[h]) {
        const uint32_t y = a_dst[h];
        if (contractible[y]) nx = slot[2 * y] == (int32_t)(h ^ 1) ? slot[2 * y + 1] : slot[2 * y];
    }
    SLink l;
    l.nxt = nx; l.last = (int32_t)h; l.cnt = 1; l.pad = 0; l.sum = a_off[h]; l.pad2 = 0;
    L[h] = l;
}

// one pointer-jumping step (ping-pong buffers); *active counts the warps that still hold a half-edge with a successor
__global__ void k_s_jump(uint64_t nh, const SLink *L, SLink *L2, unsigned long long *active)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool more = false;
    if (h < nh) {
        SLink a = L[h];
        if (a.nxt >= 0) {
            const SLink b = L[a.nxt];
            a.nxt = b.nxt; a.last = b.last; a.cnt += b.cnt; a.sum += b.sum;
            more = b.nxt >= 0;
        }
        L2[h] = a;
    }
    const unsigned m = __ballot_sync(0xffffffffu, more);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(active, 1ULL);
}

// removeParDeadEndNodes (:135-218) on the chain aggregates: a node all of whose (composite) edges are short, are no loops
// and point the same way
__global__ void k_s_deadends(uint64_t n, const uint64_t *row, const uint32_t *adj, const uint8_t *alive, const uint8_t *contractible,
                             const uint32_t *a_dst, const uint8_t *a_or, const SLink *L, const uint16_t *len, int uniform,
                             uint32_t min_reads, uint32_t min_len, uint8_t *dead)
{
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint8_t isdead = 0;
    if (!contractible[v]) {
        uint32_t in = 0, out = 0;
        bool good = false;
        for (uint64_t k = row[v]; k < row[v + 1] && !good; k++) {
            const uint32_t h = adj[k];
            if (!alive[h]) continue;
            const SLink l = L[h];
            const uint32_t far = a_dst[l.last];
            if (l.nxt >= 0) { good = true; break; }                           // (runs into a closed chain: leave it alone)
            if (l.cnt - 1 >= min_reads) good = true;                          // composite edge with enough reads (:176)
            else if (l.sum + (uint64_t)rlen_of(len, uniform, far) >= min_len) good = true; // long enough (:181)
            else if (far == (uint32_t)v) good = true;                         // loop (:186)
            else if ((a_or[h] >> 1) & 1) out++; else in++;                    // (:193-196)
        }
        isdead = !good && in * out == 0 && in + out > 0;
    }
    dead[v] = isdead;
}

// the edges of a dead-end node go, i.e. every atom of every chain that starts or ends there
__global__ void k_s_remove(uint64_t nh, const uint32_t *a_dst, const SLink *L, const uint8_t *dead, uint8_t *alive, unsigned long long *removed)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh || !alive[h]) return;
    const uint32_t w = a_dst[L[h].last], u = a_dst[L[h ^ 1].last]; // far ends of the chain in both directions
    if (dead[u] || dead[w]) {
        alive[h] = 0;
        if (!(h & 1)) atomicAdd(removed, 1ULL);
    }
}

// output, pass 1: chain heads (source not contractible) on the side printEdge writes (:663): reserve the record and the
// inner-read slots.  Atoms of closed chains are written one by one.
__global__ void k_s_heads(uint64_t nh, const uint32_t *a_src, const uint32_t *a_dst, const uint32_t *a_off, const uint8_t *a_or, const uint8_t *alive,
                          const uint8_t *contractible, const SLink *L,
                          disco_cedge *out, uint64_t cap, uint64_t *base_of, unsigned long long *cursors /* [0] edges [1] inner [2] cycle atoms */)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    base_of[h] = ~0ULL;
    if (!alive[h]) return;
    const uint32_t u = a_src[h];
    uint32_t w, natoms;
    uint64_t total;
    int o;
    const SLink l = L[h];
    if (l.nxt >= 0) { // atom of a closed chain: written unmerged, once, from its smaller end
        w = a_dst[h]; natoms = 1; total = a_off[h]; o = a_or[h];
        if (!(u < w)) return;
        atomicAdd(cursors + 2, 1ULL);
    } else {
        if (contractible[u]) return;
        const int32_t t = l.last;
        w = a_dst[t]; natoms = l.cnt; total = l.sum;
        o = (a_or[h] & 2) | (a_or[t] & 1); // mergedEdgeOrientation (EdgeSimple.cpp:256-259) along the chain
        if (!(u < w || (u == w && h < (uint64_t)(t ^ 1)))) return;
    }
    const unsigned long long at = atomicAdd(cursors, 1ULL);
    const unsigned long long ib = atomicAdd(cursors + 1, (unsigned long long)(natoms - 1));
    base_of[h] = ib;
    if (at < cap) {
        disco_cedge r;
        r.src = u; r.dst = w; r.orient = (uint32_t)o; r.n_inner = natoms - 1;
        r.offset_total = total; r.inner_start = ib;
        out[at] = r;
    }
}

// output, pass 2: every atom of an emitted chain writes the read it leads to at its rank: (read | offset << 32 | strand << 63)
// = mergeList's entry (EdgeSimple.cpp:226-230)
__global__ void k_s_inner(uint64_t nh, const uint32_t *a_dst, const uint32_t *a_off, const uint8_t *a_or, const uint8_t *alive,
                          const SLink *L, const uint64_t *base_of, uint64_t *inner, uint64_t cap)
{
    const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh || !alive[h] || L[h].nxt >= 0) return;
    const int32_t head = L[h ^ 1].last ^ 1;        // the reverse chain ends at the reverse of this chain's first atom
    const uint64_t b = base_of[head];
    if (b == ~0ULL) return;                        // not the side that is written
    const uint32_t nhead = L[head].cnt;
    const uint32_t rank = nhead - L[h].cnt;        // atoms in front of this one
    if (rank + 1 >= nhead) return;                 // the last atom leads to the end node, not to an inner read
    const uint64_t at = b + rank;
    if (at < cap) inner[at] = (uint64_t)a_dst[h] | ((uint64_t)a_off[h] << 32) | ((uint64_t)(a_or[h] & 1) << 63);
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct SimplifyBuffers {
    uint32_t *a_src = nullptr, *a_dst = nullptr, *a_off = nullptr, *deg = nullptr, *fillc = nullptr, *adj = nullptr;
    unsigned long long *deg_all = nullptr;
    void *scan_tmp = nullptr;
    uint8_t *a_or = nullptr, *alive = nullptr, *contractible = nullptr, *dead = nullptr;
    uint64_t *row = nullptr, *base_of = nullptr;
    int32_t *slot = nullptr;
    SLink *L[2] = {nullptr, nullptr};
    unsigned long long *counters = nullptr; // [0] active / removed, [1..3] output cursors
};

// ---- the reduced edge list in the order the files want it (src, dst ascending: the writers and parsimplify's loader walk it
// node by node) -- one radix sort of the 64-bit key (src << 32 | dst) with the 16-byte records as values.  A pair of reads
// has at most one reduced edge (the lower id's overlap wins), so the key is a total order.
__global__ void k_edge_keys(const disco_edge *e, uint64_t n, uint64_t *keys)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)e[i].src << 32) | e[i].dst;
}

struct Edge16 { uint64_t a, b; }; // disco_edge as an opaque 16-byte value

cudaError_t sort_edges_device(disco_edge *d_edges, uint64_t n, uint64_t n_reads, cudaStream_t s, unsigned long long *launches)
{
    if (n < 2) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    uint64_t *k0 = nullptr, *k1 = nullptr;
    Edge16 *v1 = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int hi_bit = 33; // bits of the key that can be set: 32 for dst + what the largest read id needs
    while (hi_bit < 64 && (n_reads >> (hi_bit - 32))) hi_bit++;
    Edge16 *v0 = reinterpret_cast<Edge16 *>(d_edges);
    if ((err = cudaMalloc(&k0, n * 8)) != cudaSuccess) goto out;
    if ((err = cudaMalloc(&k1, n * 8)) != cudaSuccess) goto out;
    if ((err = cudaMalloc(&v1, n * 16)) != cudaSuccess) goto out;
    k_edge_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_edges, n, k0);
    if ((err = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, (int64_t)n, 0, hi_bit, s)) != cudaSuccess) goto out;
    if ((err = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) goto out;
    if ((err = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int64_t)n, 0, hi_bit, s)) != cudaSuccess) goto out;
    if ((err = cudaMemcpyAsync(d_edges, v1, n * 16, cudaMemcpyDeviceToDevice, s)) != cudaSuccess) goto out;
    err = cudaStreamSynchronize(s);
    if (launches) *launches += 2 + (unsigned long long)((hi_bit + 7) / 8) * 2; // key kernel + copy + the sort's passes (histogram / onesweep)
out:
    cudaFree(k0); cudaFree(k1); cudaFree(v1); cudaFree(tmp);
    return err != cudaSuccess ? err : cudaGetLastError();
}

#define SCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = e__; goto done; } } while (0)

cudaError_t run_simplify(const disco_edge *d_edges, uint64_t ne, uint64_t n_reads, const uint16_t *d_len, int uniform_len,
                         uint32_t min_ovl, uint32_t min_reads, uint32_t min_len, cudaStream_t s,
                         disco_cedge **d_out, uint64_t *n_out, uint64_t **d_inner, uint64_t *n_inner, uint64_t *rounds, uint64_t *removed_edges,
                         uint64_t *cycle_atoms, unsigned long long *launches)
{
    cudaError_t err = cudaSuccess;
    SimplifyBuffers b;
    const uint64_t nh = 2 * ne, n = n_reads;
    const unsigned T = 256;
    auto blocks = [&](uint64_t k) { return (unsigned)((k + T - 1) / T); };
    int cur = 0;
    uint64_t tot_removed = 0, nrounds = 0;
    const bool trace = getenv("DISCO_SIMPLIFY_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "simplify: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    unsigned long long hc[4] = {0, 0, 0, 0};
    *d_out = nullptr; *d_inner = nullptr; *n_out = *n_inner = 0;
    if (ne == 0) { *rounds = 0; *removed_edges = 0; *cycle_atoms = 0; return cudaSuccess; }
    SCK(cudaMalloc(&b.a_src, nh * 4)); SCK(cudaMalloc(&b.a_dst, nh * 4)); SCK(cudaMalloc(&b.a_off, nh * 4));
    SCK(cudaMalloc(&b.a_or, nh)); SCK(cudaMalloc(&b.alive, nh)); SCK(cudaMalloc(&b.adj, nh * 4)); SCK(cudaMalloc(&b.base_of, nh * 8));
    SCK(cudaMalloc(&b.deg_all, (n + 1) * 8)); SCK(cudaMalloc(&b.deg, n * 4)); SCK(cudaMalloc(&b.fillc, n * 4)); SCK(cudaMalloc(&b.row, (n + 1) * 8));
    SCK(cudaMalloc(&b.slot, n * 8)); SCK(cudaMalloc(&b.contractible, n)); SCK(cudaMalloc(&b.dead, n));
    for (int k = 0; k < 2; k++) SCK(cudaMalloc(&b.L[k], nh * sizeof(SLink)));
    SCK(cudaMalloc(&b.counters, 4 * sizeof(unsigned long long)));
    lap("allocations");
    SCK(cudaMemsetAsync(b.deg_all, 0, (n + 1) * 8, s)); SCK(cudaMemsetAsync(b.fillc, 0, n * 4, s));
    k_s_atoms<<<blocks(ne), T, 0, s>>>(d_edges, ne, d_len, uniform_len, min_ovl, b.a_src, b.a_dst, b.a_off, b.a_or, b.alive, b.deg_all);
    {   // CSR offsets: exclusive prefix sum of the n + 1 counters (the last one is zero) -- library scan, runs once per call
        size_t tmp_bytes = 0;
        SCK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b.deg_all, reinterpret_cast<unsigned long long *>(b.row), n + 1, s));
        SCK(cudaMalloc(&b.scan_tmp, tmp_bytes ? tmp_bytes : 8));
        SCK(cub::DeviceScan::ExclusiveSum(b.scan_tmp, tmp_bytes, b.deg_all, reinterpret_cast<unsigned long long *>(b.row), n + 1, s));
    }
    k_s_fill<<<blocks(nh), T, 0, s>>>(b.a_src, nh, b.row, b.fillc, b.adj);
    *launches += 3;
    lap("atoms + CSR");
    for (;;) {
        nrounds++;
        k_s_nodes<<<blocks(n), T, 0, s>>>(n, b.row, b.adj, b.alive, b.a_or, b.deg, b.slot, b.contractible);
        cur = 0;
        k_s_links<<<blocks(nh), T, 0, s>>>(nh, b.a_dst, b.a_off, b.alive, b.contractible, b.slot, b.L[0]);
        *launches += 2;
        for (int it = 0; it < 40; it++) { // chains double per step; closed chains never finish and are left after 2^40
            SCK(cudaMemsetAsync(b.counters, 0, sizeof(unsigned long long), s));
            k_s_jump<<<blocks(nh), T, 0, s>>>(nh, b.L[cur], b.L[cur ^ 1], b.counters);
            *launches += 1;
            cur ^= 1;
            SCK(cudaMemcpyAsync(hc, b.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            SCK(cudaStreamSynchronize(s));
            if (!hc[0]) break;
            if (it >= 34) break; // only closed chains are left
        }
        lap("nodes + links + jumping");
        k_s_deadends<<<blocks(n), T, 0, s>>>(n, b.row, b.adj, b.alive, b.contractible, b.a_dst, b.a_or, b.L[cur], d_len, uniform_len, min_reads, min_len, b.dead);
        SCK(cudaMemsetAsync(b.counters, 0, sizeof(unsigned long long), s));
        k_s_remove<<<blocks(nh), T, 0, s>>>(nh, b.a_dst, b.L[cur], b.dead, b.alive, b.counters);
        *launches += 2;
        SCK(cudaMemcpyAsync(hc, b.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        SCK(cudaStreamSynchronize(s));
        tot_removed += hc[0];
        lap("dead ends + removal");
        if (!hc[0]) break; // nothing removed: the chains of this round are the result
        if (nrounds > 10000) break;
    }
    // ---- output: count, allocate, fill
    for (int pass = 0; pass < 2; pass++) {
        SCK(cudaMemsetAsync(b.counters, 0, 4 * sizeof(unsigned long long), s));
        k_s_heads<<<blocks(nh), T, 0, s>>>(nh, b.a_src, b.a_dst, b.a_off, b.a_or, b.alive, b.contractible, b.L[cur],
                                         *d_out, pass ? *n_out : 0, b.base_of, b.counters + 1);
        *launches += 1;
        SCK(cudaMemcpyAsync(hc, b.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        SCK(cudaStreamSynchronize(s));
        if (pass == 0) {
            *n_out = hc[1]; *n_inner = hc[2]; *cycle_atoms = hc[3];
            SCK(cudaMalloc(d_out, (hc[1] ? hc[1] : 1) * sizeof(disco_cedge)));
            SCK(cudaMalloc(d_inner, (hc[2] ? hc[2] : 1) * sizeof(uint64_t)));
        }
    }
    k_s_inner<<<blocks(nh), T, 0, s>>>(nh, b.a_dst, b.a_off, b.a_or, b.alive, b.L[cur], b.base_of, *d_inner, *n_inner);
    *launches += 1;
    SCK(cudaStreamSynchronize(s));
    lap("output");
    *rounds = nrounds; *removed_edges = tot_removed;
done:
    cudaFree(b.a_src); cudaFree(b.a_dst); cudaFree(b.a_off); cudaFree(b.a_or); cudaFree(b.alive); cudaFree(b.adj); cudaFree(b.base_of);
    cudaFree(b.deg_all); cudaFree(b.deg); cudaFree(b.fillc); cudaFree(b.row); cudaFree(b.slot); cudaFree(b.contractible); cudaFree(b.dead);
    for (int k = 0; k < 2; k++) cudaFree(b.L[k]);
    cudaFree(b.counters); cudaFree(b.scan_tmp);
    if (err != cudaSuccess) { cudaFree(*d_out); cudaFree(*d_inner); *d_out = nullptr; *d_inner = nullptr; }
    return err;
}

} // namespace disco
