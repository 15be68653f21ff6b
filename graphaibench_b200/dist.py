"""1D vertex partition with halo exchange for full-graph GCN / GraphSAGE training on P GPUs.

The reference GNN path is single-GPU; what it does have is the 1D master+halo partitioner
(PartitionedGraph::edgecut_induced_partition1D, src/partitioner/graph_partition.cc:128-178), whose ownership rule is kept
bit for bit: S = ceil(N / P), rank p owns the contiguous global range [p*S, min((p+1)*S, N)).

One process per GPU. A rank holds only the CSR rows of its masters (global column ids) and builds, on its device:

  local id space   [ interior masters | boundary masters | halo block of peer 0 | halo block of peer 1 | ... ]
                   interior = every neighbour is a master of this rank; the halo block of a peer lists the distinct
                   neighbours it owns, ascending global id. Edge order inside each row stays the global one (ascending
                   global id), so every aggregated row is bit-identical to the single-GPU result.
  send lists       the masters each peer needs (one all-to-all of global ids at plan time).
  norms            from GLOBAL degrees (halo degrees are fetched once through the same exchange).

Per aggregation: the interior rows run on a side stream while the main stream packs the rows each peer needs
(gai_gather_rows), exchanges them with ONE all-to-all-v whose receive buffer is the halo block of the gathered matrix
itself, and then aggregates the boundary rows. Weights and Adam state are replicated; dW and the loss statistics are
all-reduced. Layer schedules, optimiser quirks and the loss follow the single-GPU host classes
(graphaibench_b200/host/gai_layers.cpp, which cite the reference line by line).

torch / torch.distributed are plumbing (device memory, streams, NCCL); all arithmetic runs in libgai_b200.so.
"""
from __future__ import annotations

import threading

import numpy as np
import torch


def owner_range(nv: int, nparts: int, part: int):
    """graph_partition.cc:131-140: S = ceil(nv / P); part p owns [p*S, min((p+1)*S, nv))."""
    S = (nv + nparts - 1) // nparts
    first = min(S * part, nv)
    return S, first, min(first + S, nv)


# ---- communicators --------------------------------------------------------------------------------------------------

class TorchComm:
    """torch.distributed process group (NCCL over NVLink on GPUs; gloo for the CPU tests of the host-side logic)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_to_all_rows(self, send, send_counts, recv, recv_counts):
        self.dist.all_to_all_single(recv, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=self.group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def barrier(self):
        self.dist.barrier(group=self.group)


class _ThreadGroup:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world
        self.lock = threading.Lock()  # one rank launches at a time (the library's per-device scratch is shared by the threads)


class ThreadComm:
    """P ranks as P threads of one process sharing one device: exercises partition, plan, packing and the row-range
    kernels on a single GPU (tests). Collectives are plain copies between the ranks' tensors around thread barriers;
    a rank holds the group's launch lock except while it waits inside a collective."""

    def __init__(self, group: _ThreadGroup, rank: int):
        self.g, self.rank, self.world = group, rank, group.world

    @staticmethod
    def run(world, fn):
        """Run fn(comm) on `world` threads; returns the list of results by rank (re-raises the first failure)."""
        g = _ThreadGroup(world)
        res, err = [None] * world, [None] * world

        def body(r):
            with g.lock:
                try:
                    res[r] = fn(ThreadComm(g, r))
                except BaseException as e:  # noqa: BLE001 - surfaced to the caller below
                    err[r] = e
                    g.bar.abort()

        ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return res

    def _wait(self, t=None):
        if t is not None and t.is_cuda:
            torch.cuda.synchronize()
        self.g.lock.release()
        try:
            self.g.bar.wait()
        finally:
            self.g.lock.acquire()

    def all_to_all_rows(self, send, send_counts, recv, recv_counts):
        self.g.slots[self.rank] = (send, list(send_counts))
        self._wait(send)
        off = 0
        for q in range(self.world):
            s, sc = self.g.slots[q]
            so = sum(sc[: self.rank])
            n = sc[self.rank]
            assert n == recv_counts[q]
            if n:
                recv[off:off + n].copy_(s[so:so + n])
            off += n
        self._wait(recv)

    def all_reduce_sum(self, t):
        self.g.slots[self.rank] = t
        self._wait(t)
        tot = self.g.slots[0].clone()
        for q in range(1, self.world):  # fixed rank order on every rank: identical results everywhere
            tot += self.g.slots[q]
        self._wait(tot)
        t.copy_(tot)
        self._wait(t)

    def barrier(self):
        self._wait()


class SelfComm:
    """world == 1."""
    rank, world = 0, 1

    def all_to_all_rows(self, send, send_counts, recv, recv_counts):
        if recv.numel():
            recv.copy_(send)

    def all_reduce_sum(self, t):
        pass

    def barrier(self):
        pass


# ---- plan -----------------------------------------------------------------------------------------------------------

class HaloPlan:
    """Partition-local graph + exchange lists of one rank. All index arrays are torch tensors on `device`.

    rows_rowptr  int64[n_loc+1]   offsets of the masters' rows (row r = global vertex first+r)
    rows_colidx  int/uint32[nnz]  GLOBAL column ids, rows complete (every neighbour of a master, sorted as in the file)
    """

    def __init__(self, comm, nv, rows_rowptr, rows_colidx, device=None):
        dev = torch.device(device) if device is not None else rows_colidx.device
        P, rank = comm.world, comm.rank
        self.comm, self.nv, self.device = comm, int(nv), dev
        S, first, last = owner_range(self.nv, P, rank)
        self.S, self.first, self.last = S, first, last
        n_loc = last - first
        rp = torch.as_tensor(rows_rowptr).to(dev, torch.int64)
        col = torch.as_tensor(rows_colidx).to(dev, torch.int64)
        assert rp.numel() == n_loc + 1 and int(rp[-1]) == col.numel()
        nnz = col.numel()
        deg = rp[1:] - rp[:-1]
        own = (col >= first) & (col < last)
        row_of_edge = torch.repeat_interleave(torch.arange(n_loc, device=dev), deg)
        remote_per_row = torch.bincount(row_of_edge[~own], minlength=n_loc) if nnz else torch.zeros(n_loc, dtype=torch.int64, device=dev)
        boundary = remote_per_row > 0
        # local order of the masters: interior first, then boundary, ascending global id inside each class
        order = torch.argsort(boundary.to(torch.int8), stable=True)      # order[new] = old (row offset into the range)
        new_of_old = torch.empty(n_loc, dtype=torch.int64, device=dev)
        new_of_old[order] = torch.arange(n_loc, device=dev)
        self.n_loc = n_loc
        self.n_int = int((~boundary).sum()) if n_loc else 0
        self.master_gids = order + first                                  # global id of local row r, r < n_loc
        # halo: distinct remote neighbours, ascending global id == grouped by owner rank
        halo = torch.unique(col[~own]) if nnz else torch.empty(0, dtype=torch.int64, device=dev)
        self.halo_gids = halo
        self.n_halo = int(halo.numel())
        self.m = n_loc + self.n_halo
        self.recv_counts = torch.bincount(halo // S, minlength=P).tolist() if self.n_halo else [0] * P
        # local CSR: rows in the new order, columns in local ids, edge order untouched
        lcol = torch.where(own, new_of_old[(col - first).clamp_(0, max(n_loc - 1, 0))] if n_loc else col,
                           n_loc + torch.searchsorted(halo, col) if self.n_halo else col)
        deg_new = deg[order]
        rp_new = torch.zeros(n_loc + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg_new, 0, out=rp_new[1:])
        edge_src = torch.repeat_interleave(rp[:-1][order] - rp_new[:-1], deg_new) + torch.arange(nnz, device=dev)
        self.colidx = lcol[edge_src].to(torch.int32)
        self.rowptr = torch.cat([rp_new, rp_new[-1:].expand(self.n_halo)]).to(torch.int32)  # halo rows are empty
        self.nnz = nnz
        assert nnz < 2 ** 31 and self.m < 2 ** 31
        # send lists: tell every owner which of its masters this rank needs
        rc = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev).reshape(P, 1)
        sc = torch.empty_like(rc)
        comm.all_to_all_rows(rc, [1] * P, sc, [1] * P)
        self.send_counts = sc.reshape(-1).tolist()
        want = torch.empty(sum(self.send_counts), dtype=torch.int64, device=dev)
        comm.all_to_all_rows(halo.contiguous(), self.recv_counts, want, self.send_counts)
        assert bool(((want >= first) & (want < last)).all())
        self.send_ids = new_of_old[want - first].to(torch.int32)         # local rows to pack, grouped by destination rank
        self.n_send = int(self.send_ids.numel())
        # global degrees of masters + halo (the normalisers use full-graph degrees, lgraph.cpp:22-34)
        dloc = torch.empty(self.m, 1, dtype=torch.int64, device=dev)
        dloc[:n_loc, 0] = deg_new
        self.exchange_setup(dloc)
        self.degree = dloc.reshape(-1)

    def exchange_setup(self, buf):
        """Plan-time halo exchange of a [m, W] tensor of any dtype (torch indexing; not on the hot path)."""
        send = buf[self.send_ids.long()].contiguous()
        self.comm.all_to_all_rows(send, self.send_counts, buf[self.n_loc:], self.recv_counts)
        return buf

    def norms(self):
        """(norm_gcn, norm_mean) over the local id space — the reference's mixed float/double expressions
        (lgraph.cpp:29-32, sage_aggregator.cpp:17,41), evaluated on the global degrees."""
        fdeg = self.degree.to(torch.float32)
        # sqrtf: a double square root rounded to float is the correctly rounded float result (53 >= 2*24+2 bits), which
        # torch's vectorised float sqrt on the CPU is not for every input
        t = torch.sqrt(fdeg.to(torch.float64)).to(torch.float32)
        ngcn = torch.where(t == 0, torch.zeros_like(t), (1.0 / t.to(torch.float64)).to(torch.float32))
        nmean = (1.0 / fdeg.to(torch.float64)).to(torch.float32)
        return ngcn, nmean

    def halo_bytes(self, width):
        return 4 * width * self.n_halo


def rows_of_rank(rowptr64, colidx, nparts, part):
    """Slice a host CSR (numpy) into the rows one rank loads: (rowptr int64[n_loc+1] rebased to 0, colidx of those rows)."""
    nv = len(rowptr64) - 1
    _, first, last = owner_range(nv, nparts, part)
    rp = np.asarray(rowptr64[first:last + 1], np.int64)
    return rp - rp[0], np.asarray(colidx[rp[0]:rp[-1]])


def add_selfloop_rows(rows_rowptr, rows_colidx, first):
    """LearningGraph::add_selfloop (lgraph.h:185-218) on a block of rows: global id first+r enters row r at its sorted place."""
    rp = torch.as_tensor(rows_rowptr).to(torch.int64)
    col = torch.as_tensor(rows_colidx).to(rp.device, torch.int64)
    n = rp.numel() - 1
    deg = rp[1:] - rp[:-1]
    row = torch.repeat_interleave(torch.arange(n, device=rp.device), deg)
    smaller = torch.bincount(row[col < row + first], minlength=n) if col.numel() else torch.zeros(n, dtype=torch.int64, device=rp.device)
    rp2 = rp + torch.arange(n + 1, device=rp.device)
    out = torch.empty(col.numel() + n, dtype=torch.int64, device=rp.device)
    pos_loop = rp2[:-1] + smaller
    out[pos_loop] = torch.arange(n, device=rp.device) + first
    shift = (col > row + first).to(torch.int64)                          # edges after the loop move one further
    out[torch.arange(col.numel(), device=rp.device) + row + shift] = col
    return rp2, out


# ---- trainer --------------------------------------------------------------------------------------------------------

def _ceil4(x):
    return (x + 3) // 4 * 4


def _pitch(x):
    """Row pitch of every tall buffer (host/gai_layers.h row_pitch): rows padded to 4 floats."""
    return _ceil4(x)


class _Adam:
    """adam (optimizer.cpp:22-35) as the host classes drive it: beta powers advance once per update() call on the object,
    moments keyed by the weight tensor."""

    def __init__(self, lr):
        self.lr = np.float32(lr)
        self.b1, self.b2 = np.float32(0.9), np.float32(0.999)
        self.b1_t, self.b2_t = np.float32(0.9), np.float32(0.999)
        self.moments = {}

    def update(self, dW, W):
        from . import ops
        mv = self.moments.get(id(W))
        if mv is None:
            mv = self.moments[id(W)] = (torch.zeros_like(W), torch.zeros_like(W))
        ops.adam_update(dW, W, mv[0], mv[1], float(self.lr), float(self.b1_t), float(self.b2_t))
        self.b1_t = np.float32(self.b1_t * self.b1)
        self.b2_t = np.float32(self.b2_t * self.b2)


class OpTimer:
    """Per-op device times with CUDA events on the launching stream, in the format of the host library's profile buckets
    (bucket, shape, calls, ms, algorithmic bytes, flops)."""

    def __init__(self):
        self.rec = []

    def scope(self, bucket, shape, nbytes, flops):
        return _OpScope(self, bucket, shape, nbytes, flops)

    def collect(self):
        torch.cuda.synchronize()
        agg = {}
        for bucket, shape, nbytes, flops, e0, e1 in self.rec:
            r = agg.setdefault((bucket, shape), dict(bucket=bucket, shape=shape, calls=0, ms=0.0, bytes=0.0, flops=0.0))
            r["calls"] += 1
            r["ms"] += e0.elapsed_time(e1)
            r["bytes"] += nbytes
            r["flops"] += flops
        self.rec = []
        return list(agg.values())


class _OpScope:
    def __init__(self, timer, *meta):
        self.t, self.meta = timer, meta

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record(torch.cuda.current_stream())

    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream())
        self.t.rec.append((*self.meta, self.e0, e1))


class _NoScope:
    def __enter__(self):
        pass

    def __exit__(self, *a):
        pass


class DistGnn:
    """Model<GCN_layer> / Model<SAGE_layer> (net.cpp:361-615) on one rank of a 1D partition.

    feats / labels / train_mask are the rows of this rank's masters in PLAN order (index with plan.master_gids).
    dims = [dim_init, hidden, ..., num_cls].  n_train_global = the reference's (train_end - train_begin).
    For GCN pass a plan built on the self-looped rows (add_selfloop_rows), as Model::load_data does.
    """

    def __init__(self, arch, plan: HaloPlan, feats, labels, train_mask, n_train_global, dims, lr=0.01, static_input_halo=True, overlap=True):
        from . import model as gmodel
        from . import ops
        assert arch in ("gcn", "sage")
        self.ops, self.arch, self.plan, self.comm = ops, arch, plan, plan.comm
        self.dims, self.L = list(dims), len(dims) - 1
        self.n_train_global = int(n_train_global)
        self.static_input_halo, self.overlap = static_input_halo, overlap
        self.timer = None   # OpTimer() to collect per-op device times (bench.py); timing serialises the interior/halo overlap
        self.exchanges = self.exchange_bytes = 0
        dev = feats.device
        p = plan
        self.graph = ops.DeviceGraph(p.rowptr, p.colidx, device_arrays=True)
        ngcn, nmean = p.norms()
        self.graph.set_norms(ngcn, nmean)
        segs = [(0, p.n_loc)]
        if self.comm.world > 1:
            segs += [(0, p.n_int), (p.n_int, p.n_loc)]
        self.graph.set_row_segments(segs)
        self.labels = labels.to(dev, torch.uint8).contiguous()
        self.mask = train_mask.to(dev, torch.uint8).contiguous()
        self.side = torch.cuda.Stream()
        n, m = p.n_loc, p.m

        def buf(rows, width):
            return torch.zeros(max(rows, 1), _pitch(width), dtype=torch.float32, device=dev)

        self.W, self.Ws, self.dW, self.dWs = [], [], [], []
        self.feat_in, self.grad_in, self.T, self.A, self.Tm, self.D, self.bits = [], [], [], [], [], [], []
        # every weight gradient is a view of one flat buffer: ONE all-reduce per step instead of one per matrix
        n_w = sum(dims[l] * dims[l + 1] for l in range(self.L)) * (2 if arch == "sage" else 1)
        self.dW_flat = torch.zeros(n_w, dtype=torch.float32, device=dev)
        off = 0

        def wview(din, dout):
            nonlocal off
            v = self.dW_flat[off: off + din * dout].view(din, dout)
            off += din * dout
            return v
        for l in range(self.L):
            din, dout = dims[l], dims[l + 1]
            tf = din > dout
            self.W.append(torch.from_numpy(gmodel.glorot(din, dout, 1)).to(dev))        # seeds: graph_conv_layer.cpp:13,18
            self.dW.append(wview(din, dout))
            if arch == "sage":
                self.Ws.append(torch.from_numpy(gmodel.glorot(din, dout, 2)).to(dev))
                self.dWs.append(wview(din, dout))
            # the matrices an aggregation gathers from carry the halo rows (m rows); row stride is a multiple of 4 floats so
            # every gather is a 128-bit load without a staging copy
            self.feat_in.append(buf(m if not tf else n, din))
            self.grad_in.append(buf(m if tf else n, dout))
            self.T.append(buf(m, dout) if tf else None)            # X·W before aggregation (transform first)
            self.A.append(buf(n, din) if not tf else None)         # Â·X (aggregate first), kept for dW
            self.Tm.append(buf(m, din) if (not tf and l > 0) else None)  # G·Wᵀ before the transposed aggregation
            self.D.append(buf(n, dout) if tf else None)            # Âᵀ·G
            # sign bits of the activation (aggregate-first layers apply ReLU in a transform epilogue that also emits them): the layer
            # above masks its input gradient with 1 bit per element instead of re-reading the activation
            self.bits.append(torch.zeros(max(n, 1), (dout + 31) // 32, dtype=torch.int32, device=dev) if (not tf and l < self.L - 1) else None)
        ncls = dims[-1]
        self.logits = buf(n, ncls)[:, :ncls]     # rows padded to 4 floats like every tall buffer; the loss kernels take the pitch
        self.probs = buf(n, ncls)[:, :ncls]
        self.losses = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.sendbuf = torch.empty(max(p.n_send, 1), _pitch(max(dims)), dtype=torch.float32, device=dev)
        self.feat_in[0][:n, :dims[0]].copy_(feats)
        if self.comm.world > 1 and static_input_halo and dims[0] <= dims[1]:
            self._exchange(self.feat_in[0])   # layer-0 input is constant: its halo rows are fetched once (replicated features)
        self.opt = _Adam(lr)                                   # Model's shared optimiser (net.cpp:362), used by GCN layers
        self.layer_opt = [_Adam(lr) for _ in range(self.L)]    # graph_conv_layer::optm, used by SAGE layers
        self.exchanges = self.exchange_bytes = 0   # counted from here on (the one-time input halo fetch above is setup)

    def _scope(self, bucket, shape, nbytes=0.0, flops=0.0):
        return self.timer.scope(bucket, shape, nbytes, flops) if self.timer is not None else _NoScope()

    # -- pieces --
    def _pack(self, B):
        """Owners copy the rows their peers need into the contiguous send buffer (gai_gather_rows)."""
        p = self.plan
        ld = B.shape[1]
        send = self.sendbuf.reshape(-1)[: max(p.n_send, 1) * ld].view(-1, ld)
        if p.n_send:
            self.ops.gather_rows(p.send_ids, B, out=send[: p.n_send])
        return send

    def _a2a(self, B, send):
        """One all-to-all-v straight into B's halo block."""
        p = self.plan
        self.comm.all_to_all_rows(send[: p.n_send], p.send_counts, B[p.n_loc:], p.recv_counts)
        self.exchanges += 1
        self.exchange_bytes += 4 * B.shape[1] * p.n_halo

    def _exchange(self, B):
        """Halo rows of B[m, ld] <- the owners' rows. Pack + one all-to-all-v straight into B's halo block."""
        p = self.plan
        if self.comm.world == 1:
            return
        with self._scope("HALO", f"pack+all-to-all W={B.shape[1]}", 4.0 * B.shape[1] * p.n_halo):
            self._a2a(B, self._pack(B))

    def _spmm(self, B, F, out, transposed, flags, addend, rows):
        if rows[0] == rows[1]:
            return
        if self.arch == "gcn":
            self.ops.spmm_gcn(self.graph, B[:, :F], out=out, flags=flags, addend=addend, rows=rows)
        else:
            self.ops.spmm_mean(self.graph, B[:, :F], out=out, transposed=transposed, flags=flags, addend=addend, rows=rows)

    def _aggregate(self, B, F, out, transposed=False, flags=0, addend=None, exchange=True):
        p = self.plan
        out = out[:, :F]
        if addend is not None:
            addend = addend[:, :F]
        kind = "gcn" if self.arch == "gcn" else ("meanT" if transposed else "mean")
        # gather model of SURVEY.md §8d over this rank's rows
        nbytes = 4.0 * (p.nnz * F + p.n_loc * F * (2 if addend is not None else 1) + p.nnz + 2 * p.n_loc + 1)
        if self.comm.world == 1 or not exchange:
            with self._scope("AGGR", f"{kind} F={F}", nbytes, 2.0 * p.nnz * F):
                self._spmm(B, F, out, transposed, flags, addend, (0, p.n_loc))
            return
        if self.timer is not None:
            self._exchange(B)
            with self._scope("AGGR", f"{kind} F={F}", nbytes, 2.0 * p.nnz * F):
                self._spmm(B, F, out, transposed, flags, addend, (0, p.n_loc))
            return
        if self.overlap and p.n_int:
            # Order matters: the persistent aggregation kernel fills every SM, so a collective launched after it would only start
            # when it drains. The all-to-all is therefore enqueued FIRST (its few CTAs become resident as soon as the pack is done)
            # and the interior rows — which need nothing from the peers — start behind the pack on a side stream and take the
            # remaining SMs while the halo rows travel over NVLink.
            main = torch.cuda.current_stream()
            send = self._pack(B)
            ev = torch.cuda.Event()
            ev.record(main)
            self._a2a(B, send)
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                self._spmm(B, F, out, transposed, flags, addend, (0, p.n_int))
                done = torch.cuda.Event()
                done.record(self.side)
            self._spmm(B, F, out, transposed, flags, addend, (p.n_int, p.n_loc))
            main.wait_event(done)
        else:
            self._exchange(B)
            self._spmm(B, F, out, transposed, flags, addend, (0, p.n_loc))

    def _mm(self, A, B, out, transA=False, transB=False, accum=False, flags=0, mask=None, mask_bits=None, relu_bits=None):
        x, z = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
        y = out.shape[1]
        tag = " bitmask" if mask_bits is not None else (" mask" if mask is not None else (" relu+bits" if relu_bits is not None else ""))
        with self._scope("LINEAR", f"{x}x{y}x{z}" + (" TA" if transA else "") + (" TB" if transB else "") + tag,
                         4.0 * (x * z + z * y + x * y * (2 if accum or mask is not None else 1)), 2.0 * x * y * z):
            if mask_bits is not None:
                self.ops.matmul_mask(A, B, mask_bits, out=out, transB=transB, flags=self._pad(out) | self.ops.EPI_BITMASK)
            elif mask is not None:
                self.ops.matmul_mask(A, B, mask, out=out, transB=transB, flags=self._pad(out))
            elif relu_bits is not None:
                self.ops.matmul_relu_bits(A, B, relu_bits, out=out, flags=self._pad(out))
            else:
                self.ops.matmul(A, B, out=out, transA=transA, transB=transB, accum=accum, flags=flags | (0 if transA else self._pad(out)))

    def _pad(self, out):
        # every tall buffer of this class has rows padded to 4 floats (buf()): the transforms may use 128-bit stores throughout
        return self.ops.EPI_PADDED if out.stride(0) % 4 == 0 and out.stride(0) >= _ceil4(out.shape[1]) else 0

    def _mm_kcat(self, A1, B1, A2, B2, out, transB=False, flags=0, mask=None, mask_bits=None, relu_bits=None):
        x, z1, z2, y = A1.shape[0], A1.shape[1], A2.shape[1], out.shape[1]
        tag = " bitmask" if mask_bits is not None else (" mask" if mask is not None else "")
        with self._scope("LINEAR", f"{x}x{y}x{z1}+{z2}" + (" TB" if transB else "") + " kcat" + tag,
                         4.0 * (x * (z1 + z2) + (z1 + z2) * y + x * y * (2 if mask is not None else 1)), 2.0 * x * y * (z1 + z2)):
            f = flags | self._pad(out)
            if mask_bits is not None:
                f |= self.ops.EPI_MASK | self.ops.EPI_BITMASK
            elif mask is not None:
                f |= self.ops.EPI_MASK
            self.ops.matmul_kcat(A1, B1, A2, B2, out=out, transB=transB, flags=f, mask=mask_bits if mask_bits is not None else mask,
                                 relu_bits=relu_bits)

    def _out_buffer(self, l):
        return self.logits if l == self.L - 1 else self.feat_in[l + 1]

    def _transform_first(self, l):
        return self.dims[l] > self.dims[l + 1]

    def _premasked(self, l):
        """grad_in[l] arrives already multiplied by d_relu: layer l+1 ends its backward with a dense transform whose epilogue
        applies the mask (host/gai_layers.cpp: can_mask_grad_out)."""
        return l < self.L - 1 and self._transform_first(l + 1)

    # -- layers (schedules of GCN_layer / SAGE_layer in host/gai_layers.cpp; reference gcn_layer.cpp:5-60, sage_layer.cpp:5-53) --
    def _forward_layer(self, l):
        ops, n = self.ops, self.plan.n_loc
        din, dout = self.dims[l], self.dims[l + 1]
        relu = ops.EPI_RELU if l < self.L - 1 else 0
        X = self.feat_in[l]
        out = self._out_buffer(l)[:n, :dout]
        if din > dout:
            T = self.T[l]
            if self.arch == "sage":
                with self._scope("LINEAR", f"{n}x{dout}x{din} ncat2", 4.0 * (n * din + 2 * din * dout + 2 * n * dout), 4.0 * n * dout * din):
                    ops.matmul_ncat(X[:n, :din], self.W[l], self.Ws[l], out1=T[:n, :dout], out2=out,
                                    flags=self._pad(T[:n, :dout]) & self._pad(out))
                self._aggregate(T, dout, out, flags=ops.EPI_ADD | relu, addend=out)
            else:
                self._mm(X[:n, :din], self.W[l], T[:n, :dout])
                self._aggregate(T, dout, out, flags=relu)
        else:
            A = self.A[l]
            static = l == 0 and self.static_input_halo
            self._aggregate(X, din, A[:n], exchange=not static)
            bits = self.bits[l][:n] if self.bits[l] is not None else None
            if self.arch == "sage":
                self._mm_kcat(A[:n, :din], self.W[l], X[:n, :din], self.Ws[l], out, flags=relu, relu_bits=bits)
            elif bits is not None:
                self._mm(A[:n, :din], self.W[l], out, relu_bits=bits)
            else:
                self._mm(A[:n, :din], self.W[l], out, flags=relu)

    def _backward_layer(self, l):
        ops, n = self.ops, self.plan.n_loc
        din, dout = self.dims[l], self.dims[l + 1]
        X = self.feat_in[l]
        G = self.grad_in[l]
        if l < self.L - 1 and not self._premasked(l):
            Y = self.feat_in[l + 1]
            assert Y.shape[1] == G.shape[1]
            with self._scope("RELU", f"d_relu n={n * G.shape[1]}", 12.0 * n * G.shape[1]):
                ops.d_relu(G[:n], Y[:n], out=G[:n])
        gout = self.grad_in[l - 1][:n, :din] if l > 0 else None
        sage = self.arch == "sage"
        if din > dout:
            D = self.D[l]
            self._aggregate(G, dout, D[:n], transposed=True)
            if sage:
                with self._scope("LINEAR", f"{din}x{dout}x{n} TA two_b", 4.0 * (n * (din + 2 * dout) + 2 * din * dout), 4.0 * n * dout * din):
                    ops.wgrad_two_b(X[:n, :din], G[:n, :dout], D[:n, :dout], out1=self.dWs[l], out2=self.dW[l])
            else:
                self._mm(X[:n, :din], D[:n, :dout], self.dW[l], transA=True)
            if l > 0:
                mask = X[:n, :din] if self._premasked(l - 1) else None   # X = output of layer l-1 (post-ReLU)
                mbits = self.bits[l - 1][:n] if (mask is not None and self.bits[l - 1] is not None) else None
                if mbits is not None:
                    mask = None
                if sage:
                    self._mm_kcat(D[:n, :dout], self.W[l], G[:n, :dout], self.Ws[l], gout, transB=True, mask=mask, mask_bits=mbits)
                else:
                    self._mm(D[:n, :dout], self.W[l], gout, transB=True, mask=mask, mask_bits=mbits)
        else:
            A = self.A[l]
            if sage:
                with self._scope("LINEAR", f"{din}x{dout}x{n} TA two_a", 4.0 * (n * (2 * din + dout) + 2 * din * dout), 4.0 * n * dout * din):
                    ops.wgrad_two_a(A[:n, :din], X[:n, :din], G[:n, :dout], out1=self.dW[l], out2=self.dWs[l])
            else:
                self._mm(A[:n, :din], G[:n, :dout], self.dW[l], transA=True)
            if l > 0:
                Tm = self.Tm[l]
                self._mm(G[:n, :dout], self.W[l], Tm[:n, :din], transB=True)
                if sage:   # self term first, the neighbour term is added on top by the aggregation epilogue
                    self._mm(G[:n, :dout], self.Ws[l], gout, transB=True)
                    self._aggregate(Tm, din, gout, transposed=True, flags=ops.EPI_ADD, addend=gout)
                else:
                    self._aggregate(Tm, din, gout, transposed=True)

    def forward(self):
        """forward_prop (net.cpp:458-476): returns (mean train loss, train accuracy) over the GLOBAL train range."""
        ops, n = self.ops, self.plan.n_loc
        for l in range(self.L):
            self._forward_layer(l)
        if n:
            with self._scope("LOSS", "fwd+reduce", 4.0 * n * (3 * self.dims[-1] + 2)):
                ops.softmax_ce_forward_stats(self.logits[:n], self.labels, self.mask, 0, n, self.probs, self.losses, self.stats)
        st = self.stats.to(torch.float64)
        cnt = st[2] if n else torch.zeros((), dtype=torch.float64, device=st.device)
        tot = torch.stack([st[0] * cnt, st[1] * cnt, cnt]) if n else torch.zeros(3, dtype=torch.float64, device=st.device)
        tot = torch.nan_to_num(tot)
        self.comm.all_reduce_sum(tot)
        self._last = tot
        return tot

    def backward(self):
        ops, n = self.ops, self.plan.n_loc
        G = self.grad_in[self.L - 1]
        if n:
            # rows outside the train range keep their initial zero gradient (softmax_loss_layer.cpp:24-36)
            with self._scope("LOSS", "bwd", 8.0 * n * self.dims[-1]):
                ops.softmax_ce_backward_scaled(self.probs[:n], self.labels, self.mask, 0, n, G, self.n_train_global)
        for l in range(self.L - 1, -1, -1):
            self._backward_layer(l)
        if self.comm.world > 1:
            with self._scope("ALLREDUCE", "dW", 4.0 * self.dW_flat.numel()):
                self.comm.all_reduce_sum(self.dW_flat)

    def update(self):
        for l in range(self.L):
            if self.arch == "gcn":
                self.opt.update(self.dW[l], self.W[l])            # shared optimiser (gcn_layer.cpp:62-66)
            else:
                self.layer_opt[l].update(self.dW[l], self.W[l])   # the layer's own, neighbour then self (sage_layer.cpp:55-59)
                self.layer_opt[l].update(self.dWs[l], self.Ws[l])

    def train_epoch_async(self):
        """One epoch (net.cpp:373-383) enqueued on the current stream; returns the device tensor {loss·count, correct, count}."""
        tot = self.forward()
        self.backward()
        self.update()
        return tot

    def train_epoch(self):
        tot = self.train_epoch_async().cpu()
        cnt = float(tot[2])
        return (float(tot[0]) / cnt, float(tot[1]) / cnt) if cnt else (0.0, 0.0)
