/*
 * disco_host.h -- C ABI of the host side of the BuildGraph stage (libdisco_host.so, plain C++/OpenMP, no CUDA).
 *
 * It restates what the reference does around the hot path so that buildG stays a drop-in:
 *   disco_host_test_read      Dataset::testRead                         (src/BuildGraph/src/Dataset.cpp:403-452)
 *   disco_reads_add_file      Dataset::readDataset + file-order numbering (Dataset.cpp:161-380, :133-134) and the
 *                             2-bit packing of HashTable::insertIntoTable (HashTable.cpp:456-477), in ONE pass over
 *                             the input instead of the reference's three (Dataset.cpp:161, HashTable.cpp:119, :236)
 *   disco_write_pargraph      OverlapGraph::saveParGraphToFile line format (OverlapGraph.cpp:808-868)
 *   disco_write_contained     markContainedReads row format             (OverlapGraph.cpp:438-447)
 * All functions return 0 on success, negative on error (disco_host_last_error()).
 */
#ifndef DISCO_HOST_H_
#define DISCO_HOST_H_
#include <stdint.h>
#include "disco_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct disco_reads disco_reads;

const char *disco_host_last_error(void);

/* 1 = the (already upper-cased) read passes the reference quality filter.  The "length > minOverlap" test of
 * Dataset.cpp:305 is separate (disco_reads_* apply both). */
int disco_host_test_read(const char *seq, uint64_t len);

disco_reads *disco_reads_new(uint32_t min_overlap, int threads);
void disco_reads_free(disco_reads *r);
/* Parse one FASTA/FASTQ(.gz) file; every record advances the file index, accepted ones get the next read id. */
int disco_reads_add_file(disco_reads *r, const char *path);
/* Same for in-memory records (tests, synthetic data): seqs = concatenated raw sequences, off = n+1 offsets. */
int disco_reads_add_records(disco_reads *r, const char *seqs, const uint64_t *off, uint64_t n);
/* Build the packed arrays (stride = words for the longest accepted read). */
int disco_reads_finalize(disco_reads *r);
uint64_t disco_reads_count(const disco_reads *r);          /* accepted reads */
uint64_t disco_reads_records(const disco_reads *r);        /* all records seen = last file index */
uint32_t disco_reads_words_per_read(const disco_reads *r);
const uint64_t *disco_reads_packed(const disco_reads *r);  /* count * words_per_read */
const uint16_t *disco_reads_len(const disco_reads *r);     /* count */
const uint64_t *disco_reads_file_index(const disco_reads *r); /* count, 1-based index among all records */
uint32_t disco_reads_min_len(const disco_reads *r);
uint32_t disco_reads_max_len(const disco_reads *r);

/* Pack 2-bit base codes (0..3, concatenated, off = n+1 offsets) into out[n * words_per_read]; len_out[n]. */
int disco_host_pack_codes(const uint8_t *codes, const uint64_t *off, uint64_t n, uint32_t words_per_read,
                          uint64_t *out, uint16_t *len_out, int threads);

/* Sort contained rows into the reference's emission order at -t 1: container ascending, then k-mer position, then
 * record (prefix before suffix) -- rows of one container end up consecutive, which SimplifyGraph expects
 * (SimplifyGraph/src/DataSet.cpp:316-335).  min_overlap as given to the GPU run. */
int disco_host_sort_contained(disco_crow *rows, uint64_t n, const uint16_t *len, uint32_t min_overlap);
/* Sort edges by (src, dst, offset, orient): the canonical order used when files must be reproducible. */
int disco_host_sort_edges(disco_edge *edges, uint64_t n);

/* flag = trailing mark field of every line (2 = both endpoints finalised in this file, OverlapGraph.cpp:826-833) */
int disco_write_pargraph(const char *path, const disco_edge *edges, uint64_t n, const uint64_t *file_index,
                         const uint16_t *len, int flag, int append);
int disco_write_contained(const char *path, const disco_crow *rows, uint64_t n, const uint64_t *file_index,
                          const uint16_t *len, int append);
/* <prefix>_<t>_parGraph.txt for t = 0..shards-1, the reference's partial graphs (one file per BuildGraph thread): shard t
 * owns the reads [t*n_reads/shards, (t+1)*n_reads/shards); an edge inside one shard is written once with mark flag 2, an
 * edge between shards by the source's shard with flag 0 and by the destination's with flag 1 (OverlapGraph.cpp:826-859),
 * so that one parsimplify per file can run in parallel (OverlapGraphSimple.cpp:632-641).  edges sorted by (src, dst). */
int disco_write_pargraph_sharded(const char *prefix, uint32_t shards, const disco_edge *edges, uint64_t n, uint64_t n_reads,
                                 const uint64_t *file_index, const uint16_t *len);

#ifdef __cplusplus
}
#endif
#endif
