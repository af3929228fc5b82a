"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the dropout mask function of
interactron_b200/csrc/itn_philox.cuh (Philox4x32, 7 rounds): keep(seed, site, row, col)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_7(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(7):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return c0, c1, c2, c3


def keep_mask(seed, site, rows, cols, p, row0=0):
    """bool [rows, cols]: True where the element is kept (probability 1 - p)."""
    thr = np.uint64(int(p * 4294967296.0))
    r = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
    g = np.arange((cols + 3) // 4, dtype=np.uint64)[None, :]
    r, g = np.broadcast_arrays(r, g)
    w = philox4x32_7(r & np.uint64(MASK), r >> np.uint64(32), g, np.full_like(g, site),
                     seed & MASK, (seed >> 32) & MASK)
    words = np.stack(w, axis=-1).reshape(rows, -1)[:, :cols]
    return words >= thr
