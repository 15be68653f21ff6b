"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's text -> CSR conversion (src/converters/converter.cc:27-60,314-420,
GraphT::write_to_file src/common/graph.cc:467-508): 1-based ids, self-loops dropped, one sorted SET of neighbours per vertex, the reverse
edge added for symmetric inputs. Pinned to the reference converter's outputs by tests/test_converter.py (tests/golden/converter.json)."""
import numpy as np


def read_mtx_pairs(path):
    with open(path) as f:
        head = f.readline().split()
        assert head[0] == "%%MatrixMarket" and head[1] == "matrix" and head[2] == "coordinate" and head[3] == "pattern"
        symmetric = head[4] == "symmetric"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, _ = (int(x) for x in line.split())
        assert m == n
        uv = np.array([[int(t) for t in ln.split()[:2]] for ln in f if ln.strip()], np.int64).reshape(-1, 2) - 1
    return m, uv[:, 0], uv[:, 1], symmetric


def coo_to_csr(nv, src, dst, symmetrize):
    """(rowptr int64[nv+1], colidx uint32[ne]) of the edge SET: self-loops and out-of-range ids dropped, mirror added if `symmetrize`."""
    src, dst = np.asarray(src, np.int64), np.asarray(dst, np.int64)
    keep = (src != dst) & (src < nv) & (dst < nv)
    s, d = src[keep], dst[keep]
    if symmetrize:
        s, d = np.concatenate([s, d]), np.concatenate([d, s])
    keys = np.unique(s * (1 << 32) + d)
    rows, cols = keys >> 32, keys & 0xFFFFFFFF
    rowptr = np.zeros(nv + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=nv), out=rowptr[1:])
    return rowptr, cols.astype(np.uint32)
