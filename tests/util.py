import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def csc_from(g, prefix, n):
    return sp.csc_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=(n, n))


def canon(m):
    m = sp.csc_matrix(m).copy()
    m.eliminate_zeros()
    m.sort_indices()
    return m


def planted_blocks(n_blocks, block, seed, noise=0.0, strength=30.0):
    """Sparse symmetric count matrix with planted clusters; O(n * block) entries; self loops = 1."""
    rng = np.random.default_rng(seed)
    n = n_blocks * block
    perm = rng.permutation(n)
    rows, cols, vals = [], [], []
    for b in range(n_blocks):
        base = b * block
        i, j = np.triu_indices(block, 1)
        lam = strength / (1.0 + np.abs(i - j))
        c = rng.poisson(lam)
        nz = c > 0
        rows.append(perm[base + i[nz]])
        cols.append(perm[base + j[nz]])
        vals.append(c[nz])
    if noise > 0:
        m = int(noise * n)
        a = rng.integers(0, n, m)
        b2 = rng.integers(0, n, m)
        ok = a != b2
        rows.append(a[ok])
        cols.append(b2[ok])
        vals.append(np.ones(int(ok.sum()), dtype=np.int64))
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    v = np.concatenate(vals).astype(np.float32)
    m = sp.coo_matrix((np.concatenate([v, v]), (np.concatenate([r, c]), np.concatenate([c, r]))), shape=(n, n)).tocsc()
    m.setdiag(0)
    m = m + sp.identity(n, dtype=np.float32, format="csc")
    m = sp.csc_matrix(m, dtype=np.float32)
    m.sort_indices()
    truth = np.empty(n, dtype=np.int64)
    truth[perm] = np.repeat(np.arange(n_blocks), block)
    return m, truth


def bam_batches_py(path, name_to_id, inter_only=True, batch_bytes=64 << 20):
    """Pure-Python BAM decoder (gzip module + numpy), the independent check of the native reader hh_bam_*."""
    import gzip
    import struct
    from haphic_b200.hicio import read_bam_header
    with gzip.open(path, "rb") as f:
        hdr = read_bam_header(f)
        ref_to_id = np.array([name_to_id[n] for n in hdr.ref_names] + [-1], dtype=np.int32)   # refID -1 -> last slot
        carry = b""
        while True:
            chunk = f.read(batch_bytes)
            buf = carry + chunk
            if not buf:
                break
            offs = []
            p, n = 0, len(buf)
            while p + 4 <= n:
                (bs,) = struct.unpack_from("<i", buf, p)
                if p + 4 + bs > n:
                    break
                offs.append(p + 4)
                p += 4 + bs
            carry = buf[p:]
            if not chunk and carry:
                raise EOFError("truncated BAM record")
            if offs:
                a = np.frombuffer(buf, dtype=np.uint8)
                o = np.asarray(offs, dtype=np.int64)

                def i32(at):
                    return a[(o + at)[:, None] + np.arange(4)].copy().view("<i4").ravel()

                def u16(at):
                    return a[(o + at)[:, None] + np.arange(2)].copy().view("<u2").ravel()

                refid, pos, flag, mrefid, mpos = i32(0), i32(4), u16(14), i32(20), i32(24)
                sel = (flag & 0x40) != 0
                if inter_only:
                    sel &= refid != mrefid
                if sel.any():
                    yield np.ascontiguousarray(np.stack([ref_to_id[refid[sel]], pos[sel], ref_to_id[mrefid[sel]], mpos[sel]], axis=1),
                                               dtype=np.int32)
            if not chunk:
                break
