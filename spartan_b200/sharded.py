"""Intra-proof sharding of the two data-parallel primitives across GPUs, one process per GPU:

* dense sumcheck: every table is partitioned CYCLICALLY (global index i lives on rank i mod W at local index i div W).  Binding the top
  variable pairs i with i + len/2, and len/2 is a multiple of W while the local length is >= 2, so both partners of every pair live on the
  same rank for all those rounds: the fold is local, no bulk data ever moves.  Per round each rank contributes 2-3 partial field elements
  (96 bytes); one all-gather + a local mod-q add gives the round evaluations ("scalar-add allreduce": NCCL has no field reduction).
  When one element per rank is left, the W survivors are all-gathered and the last log2(W) rounds run replicated.
* row commitments (DensePolynomial::commit_inner): rows are independent MSMs over the same generators -> rank r commits rows
  [r*L/W, (r+1)*L/W) and the 32-byte compressed commitments are all-gathered.

* variable-base MSM (`sharded_msm_var`): the points (and their scalars) are partitioned by index range, every rank runs the bucket MSM on its
  slice, and the W partial results meet in a "point-add allreduce": an all-gather of the 32-byte encodings followed by a W-term sum.

The Fiat-Shamir transcript is deterministic, so every rank replays it on identical bytes and derives identical challenges: no broadcast.

`backend` supplies the local device operations (the GPU API in production — `GpuBackend` below; the CPU tests inject an oracle-backed
stand-in to check the partition / collective logic under gloo)."""
import numpy as np


class GpuBackend:
    """local operations on this rank's GPU through libspartan_b200.so"""

    def __init__(self, ctx=None):
        from . import api
        self.api = api
        self.ctx = ctx or api.default_context()

    def poly(self, table):
        return self.api.DensePolynomial(table, ctx=self.ctx)

    def length(self, p):
        return p.len()

    def to_numpy(self, p):
        return p.to_numpy()

    def sc_eval(self, kind, polys):
        return self.api.sumcheck_eval(kind, polys)

    def sc_fold_eval(self, kind, polys, r):
        return self.api.sumcheck_fold_eval(kind, polys, r)

    def fold(self, polys, r):
        self.api.fold_top(polys, r)

    def commit_rows(self, table, gens, L, R, blinds):
        return self.api.DensePolynomial(table, ctx=self.ctx).commit(gens, L, R, blinds)

    def points_derive(self, label, lo, hi):
        """MultiCommitGens::new(hi, label).G[lo:hi] on this rank's GPU (the SHAKE stream is sequential: the prefix is squeezed and dropped)"""
        full = self.api.Points.derive(hi, label, ctx=self.ctx)
        return full, lo

    def msm_var(self, points, scalars):
        pts, off = points
        return pts.msm(scalars, offset=off)

    def sum_points(self, encodings):
        """sum of a few points given by their encodings: an MSM with all scalars one"""
        P = self.api.Points(list(encodings), ctx=self.ctx)
        ones = np.tile(self.api.scalar_from_bytes(b"\x01" + bytes(31)), (len(encodings), 1))
        return P.msm(ones)

    def add(self, a, b):
        import ctypes as C
        out = np.zeros(4, dtype=np.uint64)
        self.api.lib.sp_scalar_add(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out


class Collective:
    """all-gather of small byte buffers over torch.distributed (NCCL on the GPUs, gloo in the CPU tests)"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = "cuda" if (dist.is_initialized() and dist.get_backend() == "nccl") else "cpu"

    def all_gather_bytes(self, payload):
        """payload: bytes of equal length on every rank -> list of `world` byte strings, rank order"""
        if self.world == 1:
            return [bytes(payload)]
        t = self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8).to(self.device)
        out = self.torch.empty(self.world * len(payload), dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        b = out.cpu().numpy().tobytes()
        n = len(payload)
        return [b[i * n:(i + 1) * n] for i in range(self.world)]


def cyclic_shard(table, rank, world):
    """global index i -> rank i mod W, local index i div W"""
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
    assert len(t) % world == 0
    return np.ascontiguousarray(t[rank::world])


def allreduce_scalars(backend, coll, vals):
    """vals: (k,4) uint64 Montgomery limbs -> element-wise field sum over ranks (all-gather + local adds)"""
    vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    parts = coll.all_gather_bytes(vals.tobytes())
    acc = np.frombuffer(parts[0], dtype=np.uint64).reshape(-1, 4).copy()
    for p in parts[1:]:
        other = np.frombuffer(p, dtype=np.uint64).reshape(-1, 4)
        for i in range(len(acc)):
            acc[i] = backend.add(acc[i], other[i])
    return acc


def sharded_sumcheck_rounds(backend, coll, kind, full_tables, challenges):
    """Runs the round-polynomial evaluations of a dense sumcheck (kind 0: A*B, 1: A*B*C, 2: A*(B*C-D)) over cyclically sharded tables.
    full_tables: list of (2^n, 4) arrays (every rank passes the same full tables; it keeps only its shard);  challenges: n field elements.
    Returns (list of per-round [e0, e2, e3] as (3,4) arrays, final table values as (ntables,4))."""
    W, rank = coll.world, coll.rank
    n_total = len(full_tables[0])
    polys = [backend.poly(cyclic_shard(t, rank, W)) for t in full_tables]
    num_rounds = int(np.log2(n_total))
    evals = []
    local_len = n_total // W
    j = 0
    e = backend.sc_eval(kind, polys) if local_len >= 2 else None
    while local_len >= 2:
        evals.append(allreduce_scalars(backend, coll, e))
        r = challenges[j]
        j += 1
        if local_len >= 4:
            e = backend.sc_fold_eval(kind, polys, r)
        else:
            backend.fold(polys, r)
        local_len //= 2
    # one element per rank left: gather the W survivors (global index == rank) and finish replicated
    finals = []
    survivors = []
    for p in polys:
        mine = backend.to_numpy(p)[:1]
        parts = coll.all_gather_bytes(mine.tobytes())
        survivors.append(np.concatenate([np.frombuffer(x, dtype=np.uint64).reshape(1, 4) for x in parts]))
    if W > 1:
        rep = [backend.poly(np.ascontiguousarray(s)) for s in survivors]
        ln = W
        e = backend.sc_eval(kind, rep)
        while ln >= 2:
            evals.append(np.ascontiguousarray(e))
            r = challenges[j]
            j += 1
            if ln >= 4:
                e = backend.sc_fold_eval(kind, rep, r)
            else:
                backend.fold(rep, r)
            ln //= 2
        finals = [backend.to_numpy(p)[0] for p in rep]
    else:
        finals = [s[0] for s in survivors]
    assert j == num_rounds
    return evals, np.stack(finals)


def sharded_commit_rows(backend, coll, table, gens, L, R, blinds=None):
    """DensePolynomial::commit_inner with the rows split across ranks; returns the L compressed commitments on every rank"""
    W, rank = coll.world, coll.rank
    assert L % W == 0
    t = np.ascontiguousarray(table, dtype=np.uint64).reshape(L, R, 4)
    lo, hi = rank * (L // W), (rank + 1) * (L // W)
    bl = None if blinds is None else np.ascontiguousarray(blinds, dtype=np.uint64).reshape(L, 4)[lo:hi]
    mine = backend.commit_rows(np.ascontiguousarray(t[lo:hi]).reshape(-1, 4), gens, L // W, R, bl)
    parts = coll.all_gather_bytes(b"".join(mine))
    out = []
    for p in parts:
        out += [p[32 * i:32 * i + 32] for i in range(L // W)]
    return out


def sharded_msm_var(backend, coll, points, scalars_slice):
    """GroupElement::vartime_multiscalar_mul over a point set split by index range: `points` / `scalars_slice` are this rank's slice
    (backend-specific handle, (m,4) Montgomery limbs).  Returns the 32-byte encoding of the full sum on every rank."""
    mine = backend.msm_var(points, scalars_slice)
    parts = coll.all_gather_bytes(mine)
    if coll.world == 1:
        return parts[0]
    return backend.sum_points(parts)


def index_range(n, rank, world):
    """contiguous slice [lo, hi) of n items for `rank` (sizes differ by at most one)"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
