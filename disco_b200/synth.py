"""Seeded synthetic read sets for the BuildGraph hot path.

All generators return a ReadSet: 2-bit base codes (A=0 C=1 G=2 T=3, the reference's packing code,
HashTable.h:16-22) concatenated in one uint8 array plus offsets.  Uniform-random genomes never trip the
reference read filter (Dataset.cpp:403-452) in practice; `ReadSet.strings()` lets tests confirm that.
"""
from dataclasses import dataclass
import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class ReadSet:
    codes: np.ndarray      # uint8, concatenated base codes 0..3
    off: np.ndarray        # int64, n+1 offsets into codes
    name: str = "synthetic"

    @property
    def n(self) -> int:
        return len(self.off) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.off)

    def strings(self):
        a = _ASCII[self.codes].tobytes().decode()
        o = self.off
        return [a[o[i]:o[i + 1]] for i in range(self.n)]

    def ascii(self) -> np.ndarray:
        return _ASCII[self.codes]

    def write_fasta(self, path, start=0, stop=None):
        stop = self.n if stop is None else stop
        a = _ASCII[self.codes[self.off[start]:self.off[stop]]].tobytes()
        base = int(self.off[start])
        with open(path, "wb") as f:
            chunk = []
            for i in range(start, stop):
                chunk.append(b">%d\n" % (i + 1))
                chunk.append(a[int(self.off[i]) - base:int(self.off[i + 1]) - base])
                chunk.append(b"\n")
                if len(chunk) > 30000:
                    f.write(b"".join(chunk)); chunk = []
            f.write(b"".join(chunk))

    def subset(self, idx) -> "ReadSet":
        idx = np.asarray(idx)
        lens = self.lengths[idx]
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        codes = np.empty(int(off[-1]), dtype=np.uint8)
        for k, i in enumerate(idx):
            codes[off[k]:off[k + 1]] = self.codes[self.off[i]:self.off[i + 1]]
        return ReadSet(codes, off, self.name + "_subset")


def revcomp_codes(x: np.ndarray) -> np.ndarray:
    return (3 - x[..., ::-1]).astype(np.uint8)


def random_genome(rng: np.random.Generator, n: int) -> np.ndarray:
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def _uniform_reads(genome, starts, read_len, flip):
    """(n, read_len) matrix of reads taken at `starts`, reverse-complemented where flip is set."""
    n = len(starts)
    out = np.empty((n, read_len), dtype=np.uint8)
    step = 1 << 18
    ar = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        m = genome[starts[lo:hi, None] + ar[None, :]]
        f = flip[lo:hi]
        m[f] = 3 - m[f][:, ::-1]
        out[lo:hi] = m
    return out


def from_matrix(mat: np.ndarray, name="synthetic") -> ReadSet:
    n, L = mat.shape
    return ReadSet(np.ascontiguousarray(mat).reshape(-1), np.arange(n + 1, dtype=np.int64) * L, name)


def from_strings(reads, name="strings") -> ReadSet:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    lens = np.array([len(r) for r in reads], dtype=np.int64)
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    codes = lut[np.frombuffer("".join(reads).encode(), dtype=np.uint8)]
    assert (codes < 4).all(), "from_strings expects filtered upper-case ACGT reads"
    return ReadSet(codes, off, name)


def single_genome(n_reads: int, read_len: int = 150, coverage: float = 30.0, seed: int = 2,
                  genome_len: int = None) -> ReadSet:
    """Config 2 shape: uniform-random genome, uniform starts, strand ~ Bernoulli(1/2), error-free."""
    rng = np.random.default_rng(seed)
    if genome_len is None:
        genome_len = max(read_len + 1, int(round(n_reads * read_len / coverage)))
    g = random_genome(rng, genome_len)
    starts = rng.integers(0, genome_len - read_len + 1, size=n_reads, dtype=np.int64)
    flip = rng.random(n_reads) < 0.5
    return from_matrix(_uniform_reads(g, starts, read_len, flip), f"single_genome_{n_reads}x{read_len}_seed{seed}")


def metagenome(n_reads: int, n_genomes: int = 200, genome_len: int = 2_500_000, read_len: int = 150,
               sigma: float = 1.0, seed: int = 3) -> ReadSet:
    """Config 3/5 shape: many random genomes with log-normal abundances (mean coverage = n*L / total bases)."""
    rng = np.random.default_rng(seed)
    ab = rng.lognormal(0.0, sigma, size=n_genomes)
    ab /= ab.sum()
    g = random_genome(rng, n_genomes * genome_len)
    which = rng.choice(n_genomes, size=n_reads, p=ab).astype(np.int64)
    starts = which * genome_len + rng.integers(0, genome_len - read_len + 1, size=n_reads, dtype=np.int64)
    flip = rng.random(n_reads) < 0.5
    return from_matrix(_uniform_reads(g, starts, read_len, flip), f"metagenome_{n_reads}x{read_len}_g{n_genomes}_seed{seed}")


def dup_contained(n_reads: int, read_len: int = 150, coverage: float = 60.0, dup_frac: float = 0.3,
                  trunc_frac: float = 0.2, min_len: int = 100, seed: int = 4) -> ReadSet:
    """Config 4 shape: dup_frac of the reads re-emitted as exact / reverse-complement duplicates and trunc_frac
    truncated to [min_len, read_len) so that they are contained in an overlapping full-length read."""
    rng = np.random.default_rng(seed)
    n_base = max(1, int(round(n_reads / (1.0 + dup_frac))))
    genome_len = max(read_len + 1, int(round(n_base * read_len / coverage)))
    g = random_genome(rng, genome_len)
    starts = rng.integers(0, genome_len - read_len + 1, size=n_base, dtype=np.int64)
    flip = rng.random(n_base) < 0.5
    mat = _uniform_reads(g, starts, read_len, flip)
    n_dup = n_reads - n_base
    src = rng.integers(0, n_base, size=n_dup)
    dups = mat[src].copy()
    rcd = rng.random(n_dup) < 0.5
    dups[rcd] = 3 - dups[rcd][:, ::-1]
    allm = np.concatenate([mat, dups], axis=0)
    perm = rng.permutation(n_reads)
    allm = allm[perm]
    lens = np.full(n_reads, read_len, dtype=np.int64)
    tr = rng.random(n_reads) < trunc_frac
    lens[tr] = rng.integers(min_len, read_len, size=int(tr.sum()))
    head = rng.random(n_reads) < 0.5           # drop bases from the head or the tail
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    codes = np.empty(int(off[-1]), dtype=np.uint8)
    for i in range(n_reads):
        L = lens[i]
        codes[off[i]:off[i + 1]] = allm[i, read_len - L:] if (tr[i] and head[i]) else allm[i, :L]
    return ReadSet(codes, off, f"dup_contained_{n_reads}x{read_len}_seed{seed}")


def paired_genome(n_pairs: int, read_len: int = 250, insert: int = 500, insert_sd: int = 50,
                  genome_len: int = None, coverage: float = 30.0, seed: int = 1) -> ReadSet:
    """Config 1 stand-in (E. coli-like): interleaved pairs, 2 x read_len, insert ~ N(insert, sd), both strands."""
    rng = np.random.default_rng(seed)
    if genome_len is None:
        genome_len = int(round(2 * n_pairs * read_len / coverage))
    g = random_genome(rng, genome_len)
    ins = np.clip(np.rint(rng.normal(insert, insert_sd, size=n_pairs)).astype(np.int64), read_len, genome_len)
    s = rng.integers(0, genome_len - ins + 1, dtype=np.int64)
    flip = rng.random(n_pairs) < 0.5
    left = _uniform_reads(g, s, read_len, np.zeros(n_pairs, dtype=bool))
    right = _uniform_reads(g, s + ins - read_len, read_len, np.ones(n_pairs, dtype=bool))
    mat = np.empty((2 * n_pairs, read_len), dtype=np.uint8)
    mat[0::2] = np.where(flip[:, None], right, left)
    mat[1::2] = np.where(flip[:, None], left, right)
    return from_matrix(mat, f"paired_{n_pairs}x2x{read_len}_seed{seed}")


def repeats(n_reads: int, read_len: int = 150, n_copies: int = 12, unique_len: int = 2000, rep_len: int = 1000,
            divergence: float = 0.01, seed: int = 11) -> ReadSet:
    """Probe P3 shape: near-identical repeats at high coverage; MAX_EDGE_PER_KMER fires."""
    rng = np.random.default_rng(seed)
    rep = random_genome(rng, rep_len)
    parts = []
    for _ in range(n_copies):
        parts.append(random_genome(rng, unique_len))
        r = rep.copy()
        mut = rng.random(rep_len) < divergence
        r[mut] = (r[mut] + rng.integers(1, 4, size=int(mut.sum()))) % 4
        parts.append(r.astype(np.uint8))
    g = np.concatenate(parts)
    starts = rng.integers(0, len(g) - read_len + 1, size=n_reads, dtype=np.int64)
    flip = rng.random(n_reads) < 0.5
    return from_matrix(_uniform_reads(g, starts, read_len, flip), f"repeats_{n_reads}x{read_len}_seed{seed}")
