"""Query-sharded FindNeighbors over one process per GPU (torch.distributed).

The path shards naturally with ONE exchange step:

  rank g computes the kNN rows of its share       -- kNN rows are independent
  one collective on the int32 kNN index           -- the only exchange (4*n*k bytes)
  every rank transposes the full index locally    -- K3, cheap, HBM-bound
  rank g emits SNN rows [g*ceil(n/G), ...)        -- K4 rows are independent

Which kNN rows a rank computes is the library's choice when the backend offers
``knn_owned`` (b200nn_opts.shard_rank / shard_world): whole groups of the tile-pruned
search, so that every rank keeps full work items; the shares partition the rows and are
combined by a sum all-reduce over zero-filled index buffers.  Otherwise (query != reference
calls, the CPU test backend) rows are split contiguously and the index is all-gathered.

The reference embedding is replicated on every GPU (<= 1 GB at 4M x 30).
Distances never travel.  `torch.distributed` (NCCL over NVLink on GPUs, gloo
in the CPU tests) is plumbing only; all compute goes through the C ABI of
libb200nn via the backend object.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row ranges, ceil(n / world) rows each (the tail may be short or empty)."""
    per = math.ceil(n / world) if n > 0 else 0
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


class CudaBackend:
    """Compute stages on this rank's GPU through libb200nn's device-buffer entry points."""

    def __init__(self, device: int, knn_algo: int = _lib.KNN_AUTO, n_cand: int = 0, scan_mode: int = 0):
        self.device = torch.device("cuda", device)
        self.ctx = _lib.Context(device)
        self.knn_algo, self.n_cand, self.scan_mode = knn_algo, n_cand, scan_mode
        self.last_stats: dict = {}

    def knn_shard(self, ref: torch.Tensor, q_begin: int, q_end: int, k: int, qry: torch.Tensor | None = None):
        """ref: float64 CUDA [nr, d] row-major; queries are rows [q_begin, q_end) of qry (or ref)."""
        src = ref if qry is None else qry
        nr, d = ref.shape
        rows = q_end - q_begin
        idx0 = torch.empty((rows, k), dtype=torch.int32, device=self.device)
        dst = torch.empty((rows, k), dtype=torch.float64, device=self.device)
        if rows > 0:
            q = src[q_begin:q_end]
            torch.cuda.current_stream(self.device).synchronize()
            self.last_stats = self.ctx.dev_knn(ref, nr, q, rows, d, k, idx0, dst, self.knn_algo, self.n_cand,
                                               self.scan_mode)
        return idx0, dst

    def knn_owned(self, ref: torch.Tensor, rank: int, world: int, k: int):
        """Self search, share `rank` of `world`.  Returns (idx0 int32 [n, k] zero except the rows of this
        share, dist float64 [n, k] NaN except those rows); the buffers are reused between calls."""
        n, d = ref.shape
        key = (n, k)
        if getattr(self, "_owned_key", None) != key:
            self._owned_idx = torch.empty((n, k), dtype=torch.int32, device=self.device)
            self._owned_dist = torch.empty((n, k), dtype=torch.float64, device=self.device)
            self._owned_key = key
        self._owned_idx.zero_()
        self._owned_dist.fill_(float("nan"))
        torch.cuda.current_stream(self.device).synchronize()
        self.last_stats = self.ctx.dev_knn(ref, n, None, n, d, k, self._owned_idx, self._owned_dist, self.knn_algo,
                                           self.n_cand, self.scan_mode, rank, world)
        return self._owned_idx, self._owned_dist

    def graph_rows(self, idx0_full: torch.Tensor, k: int, prune: float, compute_snn: bool,
                   row_begin: int, row_end: int) -> dict:
        n = idx0_full.shape[0]
        torch.cuda.current_stream(self.device).synchronize()
        st = self.ctx.dev_graphs(idx0_full, n, k, prune, compute_snn, row_begin, row_end)
        out = {"stats": st}
        for name, which in (("nn", 0), ("snn", 1)):
            if which == 1 and not compute_snn:
                continue
            r = self.ctx.dev_result(which)
            p = torch.empty(r.n_rows + 1, dtype=torch.int64, device=self.device)
            i = torch.empty(r.nnz, dtype=torch.int32, device=self.device)
            x = torch.empty(r.nnz, dtype=torch.float64, device=self.device)
            self.ctx.dev_copy_result(which, p, i, x)
            out[name] = (p, i, x)
        return out


def all_gather_index(idx_local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather of the row-sharded kNN index.  Shards are padded to the common
    length ceil(n / world) so that a single equal-count collective suffices."""
    world = dist.get_world_size(group)
    if world == 1:
        return idx_local
    k = idx_local.shape[1]
    per = math.ceil(n / world)
    send = idx_local
    if idx_local.shape[0] != per:
        send = torch.zeros((per, k), dtype=idx_local.dtype, device=idx_local.device)
        send[: idx_local.shape[0]] = idx_local
    send = send.contiguous()
    recv = torch.empty((world * per, k), dtype=idx_local.dtype, device=idx_local.device)
    dist.all_gather(list(recv.view(world, per, k).unbind(0)), send, group=group)
    return recv[:n]


def all_reduce_index(idx_owned: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the zero-filled per-rank index buffers = the full index (the shares partition the rows)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(idx_owned, op=dist.ReduceOp.SUM, group=group)
    return idx_owned


def find_neighbors_sharded(ref: torch.Tensor, k: int, prune: float = 1 / 15, compute_snn: bool = True,
                           backend=None, group=None, qry: torch.Tensor | None = None) -> dict:
    """One rank's part of FindNeighbors on a replicated embedding.

    Returns this rank's rows of the kNN result (0-based index, distances), its
    SNN rows as a local CSR block (global column ids; offsets start at 0) and
    the full nn Graph (identical on every rank).  With ``qry`` given only the
    kNN stage runs (the ``FindNeighbors(query=)`` path, config E).
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = ref.shape[0]
    nq = n if qry is None else qry.shape[0]
    b, e = shard_bounds(nq, world)[rank]
    if qry is None and world > 1 and hasattr(backend, "knn_owned"):
        # group-aligned shares: this rank's kNN rows are scattered over [0, n); "dist" is the full-size
        # buffer (NaN rows belong to other ranks), "owned" the mask of the rows computed here
        idx_owned, dist_owned = backend.knn_owned(ref, rank, world, k)
        owned = ~torch.isnan(dist_owned[:, 0])
        idx_full = all_reduce_index(idx_owned, group)
        out = {"rows": (b, e), "idx0": idx_full[b:e], "dist": dist_owned, "owned": owned}
        g = backend.graph_rows(idx_full, k, prune, compute_snn, b, e)
        out.update(g)
        out["idx0_full"] = idx_full
        return out
    idx_local, dist_local = backend.knn_shard(ref, b, e, k, qry)
    out = {"rows": (b, e), "idx0": idx_local, "dist": dist_local}
    if qry is not None:
        return out
    idx_full = all_gather_index(idx_local, n, group) if world > 1 else idx_local
    g = backend.graph_rows(idx_full, k, prune, compute_snn, b, e)
    out.update(g)
    out["idx0_full"] = idx_full
    return out


def gather_csr_rows(p: torch.Tensor, i: torch.Tensor, x: torch.Tensor, group=None, dst: int = 0):
    """Concatenate row-sharded CSR blocks on rank `dst` (prefix-offsetting p).
    Returns (p, i, x) CPU tensors on `dst`, None elsewhere."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return p.cpu(), i.cpu(), x.cpu()
    meta = torch.tensor([p.numel() - 1, i.numel()], dtype=torch.int64, device=p.device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    if rank != dst:
        dist.send(p.contiguous(), dst, group=group)
        if i.numel():
            dist.send(i.contiguous(), dst, group=group)
            dist.send(x.contiguous(), dst, group=group)
        return None
    ps, is_, xs = [], [], []
    off = 0
    for r in range(world):
        rows, nnz = int(metas[r][0]), int(metas[r][1])
        if r == dst:
            pr, ir, xr = p, i, x
        else:
            pr = torch.empty(rows + 1, dtype=p.dtype, device=p.device)
            ir = torch.empty(nnz, dtype=i.dtype, device=p.device)
            xr = torch.empty(nnz, dtype=x.dtype, device=p.device)
            dist.recv(pr, r, group=group)
            if nnz:
                dist.recv(ir, r, group=group)
                dist.recv(xr, r, group=group)
        ps.append(pr[:-1].cpu() + off if r < world - 1 else pr.cpu() + off)
        is_.append(ir.cpu())
        xs.append(xr.cpu())
        off += nnz
    return torch.cat(ps), torch.cat(is_), torch.cat(xs)
