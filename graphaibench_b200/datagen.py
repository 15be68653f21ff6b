"""Synthetic inputs in the reference's on-disk conventions: R-MAT power-law graphs, symmetrised, sorted,
deduplicated, no self-loops; N(0,1) features; uniform labels; range splits. Host-side numpy (input generation only)."""
from __future__ import annotations

import os

import numpy as np


def rmat_csr(n_vertices: int, target_nnz: int, seed: int = 1, abcd=(0.57, 0.19, 0.19, 0.05), permute: bool = False):
    """Undirected R-MAT: draw directed pairs, drop loops, symmetrise, sort, dedupe. Returns (rowptr int64[n+1], colidx u32[nnz]).
    nnz lands within a few percent of the target (duplicates are removed, so we oversample and trim whole pairs)."""
    rng = np.random.default_rng(seed)
    scale = int(np.ceil(np.log2(max(n_vertices, 2))))
    a, b, c, _ = abcd
    want_pairs = target_nnz // 2
    keys = np.empty(0, np.int64)
    draw = int(want_pairs * 1.25) + 16
    for _ in range(8):
        src = np.zeros(draw, np.int64)
        dst = np.zeros(draw, np.int64)
        for _lvl in range(scale):
            r = rng.random(draw)
            sb = (r >= a + b).astype(np.int64)                       # lower half (c or d quadrant)
            db = (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64)  # right half (b or d quadrant)
            src = (src << 1) | sb
            dst = (dst << 1) | db
        ok = (src < n_vertices) & (dst < n_vertices) & (src != dst)
        lo = np.minimum(src[ok], dst[ok])
        hi = np.maximum(src[ok], dst[ok])
        keys = np.unique(np.concatenate([keys, lo * n_vertices + hi]))
        if len(keys) >= want_pairs:
            break
        draw = int((want_pairs - len(keys)) * 1.6) + 16
    if len(keys) > want_pairs:
        keys = keys[np.sort(rng.choice(len(keys), want_pairs, replace=False))]
    lo, hi = keys // n_vertices, keys % n_vertices
    if permute:
        perm = rng.permutation(n_vertices)
        lo, hi = perm[lo], perm[hi]
    s = np.concatenate([lo, hi])
    d = np.concatenate([hi, lo])
    order = np.lexsort((d, s))
    s, d = s[order], d[order]
    rowptr = np.zeros(n_vertices + 1, np.int64)
    np.cumsum(np.bincount(s, minlength=n_vertices), out=rowptr[1:])
    return rowptr, d.astype(np.uint32)


def features(n, f, seed=2):
    return np.random.default_rng(seed).standard_normal((n, f), dtype=np.float32)


def labels(n, ncls, seed=3):
    return np.random.default_rng(seed).integers(0, ncls, n, dtype=np.uint8)


def split_ranges(n):
    """first 50% train / next 25% val / rest test, as graph.meta.txt ranges (begin, end, count) x 3."""
    t, v = n // 2, n // 2 + n // 4
    return np.array([0, t, t, t, v, v - t, v, n, n - v], np.int64)


def write_dataset(path, rowptr64, colidx, feats, labs, ncls, split9):
    """Write the reference .bin set (reader.cpp:414-457): graph.meta.txt, graph.vertex.bin (int64), graph.edge.bin (u32),
    graph.feats.bin (f32), graph.vlabel.bin (u8)."""
    os.makedirs(path, exist_ok=True)
    nv, ne = len(rowptr64) - 1, len(colidx)
    deg = np.diff(rowptr64)
    meta = [nv, ne, 4, 8, 1, 2, int(deg.max()) if nv else 0, feats.shape[1], ncls, 0] + [int(x) for x in split9]
    with open(os.path.join(path, "graph.meta.txt"), "w") as f:
        f.write("\n".join(str(m) for m in meta[:10]) + "\n")
        f.write(" ".join(str(m) for m in meta[10:13]) + "\n" + " ".join(str(m) for m in meta[13:16]) + "\n" + " ".join(str(m) for m in meta[16:19]) + "\n")
    np.asarray(rowptr64, np.int64).tofile(os.path.join(path, "graph.vertex.bin"))
    np.asarray(colidx, np.uint32).tofile(os.path.join(path, "graph.edge.bin"))
    np.asarray(feats, np.float32).tofile(os.path.join(path, "graph.feats.bin"))
    np.asarray(labs, np.uint8).tofile(os.path.join(path, "graph.vlabel.bin"))


def _rmat_pairs_torch(n_vertices, target_nnz, seed, abcd, device):
    """Unique undirected R-MAT pairs (lo < hi) as int64 keys lo * n + hi; ~target_nnz / 2 of them."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    scale = int(np.ceil(np.log2(max(n_vertices, 2))))
    a, b, c, _ = abcd
    want_pairs = target_nnz // 2
    keys = torch.empty(0, dtype=torch.int64, device=device)
    keep = (n_vertices / float(1 << scale)) ** 2   # share of draws whose two ends are both < n_vertices
    draw = int(want_pairs * 1.25 / max(keep, 0.05)) + 16
    for _ in range(8):
        src = torch.zeros(draw, dtype=torch.int64, device=device)
        dst = torch.zeros(draw, dtype=torch.int64, device=device)
        for _lvl in range(scale):
            r = torch.rand(draw, generator=g, device=device)
            sb = (r >= a + b).to(torch.int64)
            db = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
            src = (src << 1) | sb
            dst = (dst << 1) | db
        ok = (src < n_vertices) & (dst < n_vertices) & (src != dst)
        lo = torch.minimum(src[ok], dst[ok])
        hi = torch.maximum(src[ok], dst[ok])
        keys = torch.unique(torch.cat([keys, lo * n_vertices + hi]))
        del src, dst, r, sb, db, ok, lo, hi
        if keys.numel() >= want_pairs:
            break
        draw = int((want_pairs - keys.numel()) * 1.6 / max(keep, 0.05)) + 16
    if keys.numel() > want_pairs:
        sel = torch.randperm(keys.numel(), generator=g, device=device)[:want_pairs]
        keys = keys[torch.sort(sel).values]
    return keys, g


def rmat_csr_torch(n_vertices: int, target_nnz: int, seed: int = 1, abcd=(0.57, 0.19, 0.19, 0.05), device="cuda", permute=False, rows=None):
    """Same construction as rmat_csr, vectorised with torch on `device` (input generation only; the multi-million-edge
    bench graphs take seconds instead of minutes). Returns (rowptr int64, colidx int64) tensors on `device`.
    permute: relabel the vertices with a seeded random permutation (balances a contiguous 1D partition: natural R-MAT ids
    put most edges into the low id ranges). rows=(first, last): return only those rows (rowptr rebased to 0, global column ids)
    — what one rank of a 1D partition loads."""
    import torch
    keys, g = _rmat_pairs_torch(n_vertices, target_nnz, seed, abcd, device)
    lo, hi = keys // n_vertices, keys % n_vertices
    del keys
    if permute:
        perm = torch.randperm(n_vertices, generator=g, device=device)
        lo, hi = perm[lo], perm[hi]
    first, last = (0, n_vertices) if rows is None else rows
    if lo.is_cuda and n_vertices < (1 << 31) and os.environ.get("GAI_DATAGEN_TORCH", "0") != "1":
        # CSR construction on the device (csrc/convert.cu: keys, radix sort, unique, offsets — the reference converter's edge SET), the
        # same routine gpu_converter uses; the CSR of an edge set is unique, so this returns what the torch construction below does
        from . import ops
        rowptr, colidx = ops.coo_to_csr(n_vertices, lo.to(torch.int32), hi.to(torch.int32), symmetrize=True)
        del lo, hi
        if rows is not None:
            b, e = int(rowptr[first]), int(rowptr[last])
            rowptr, colidx = rowptr[first:last + 1] - b, colidx[b:e]
        return rowptr, colidx.to(torch.int64)
    s = torch.cat([lo, hi])
    d = torch.cat([hi, lo])
    del lo, hi
    if rows is not None:
        sel = (s >= first) & (s < last)
        s, d = s[sel], d[sel]
    order = torch.argsort(s * n_vertices + d)
    s, d = s[order], d[order]
    rowptr = torch.zeros(last - first + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(s - first, minlength=last - first), 0)
    return rowptr, d
