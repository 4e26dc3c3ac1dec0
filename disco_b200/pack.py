"""2-bit packing of base codes into the reference record-payload layout (HashTable.cpp:456-477):
base i -> bits [62-2*(i%32), 63-2*(i%32)] of word i//32, A=0 C=1 G=2 T=3, zero padded."""
import numpy as np

_SHIFTS = (62 - 2 * np.arange(32, dtype=np.uint64)).astype(np.uint64)


def words_for(length: int) -> int:
    return (int(length) + 31) // 32


def pack_codes(codes: np.ndarray, off: np.ndarray, words_per_read: int = None):
    """codes: uint8 0..3 concatenated; off: n+1 offsets.  Returns (packed uint64[n, wpr], len uint16[n])."""
    off = np.asarray(off, dtype=np.int64)
    n = len(off) - 1
    lens = np.diff(off)
    assert n == 0 or (lens.max() <= 32767), "read longer than the 15-bit length field"
    wpr = words_per_read or max(1, words_for(lens.max() if n else 1))
    out = np.zeros((n, wpr), dtype=np.uint64)
    if n == 0:
        return out, lens.astype(np.uint16)
    uniform = bool((lens == lens[0]).all())
    step = 1 << 16
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        m = hi - lo
        buf = np.zeros((m, wpr * 32), dtype=np.uint64)
        if uniform:
            L = int(lens[0])
            buf[:, :L] = codes[off[lo]:off[hi]].reshape(m, L)
        else:
            idx = np.arange(off[lo], off[hi], dtype=np.int64)
            row = np.repeat(np.arange(m, dtype=np.int64), lens[lo:hi])
            col = idx - np.repeat(off[lo:hi], lens[lo:hi])
            buf[row, col] = codes[off[lo]:off[hi]]
        out[lo:hi] = (buf.reshape(m, wpr, 32) << _SHIFTS[None, None, :]).sum(axis=2, dtype=np.uint64)
    return out, lens.astype(np.uint16)
