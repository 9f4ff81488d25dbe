"""CPU model of the intra-proof sharding (spartan_b200/csrc/comm.cu, prover.cpp, snark.cpp) with the oracle as the arithmetic: the index logic the
CUDA path relies on — cyclic shards keep bound_poly_var_top local, a rank's slice of an eq table is eq over the leading variables times the
factor its low index bits fix (shard_eq_scale), product-circuit layers can be built from local entries, partial round sums add up to the
unsharded round polynomial before AND after the shards are gathered into replicated tables.  No GPU, no process group: W ranks are W slices."""
import numpy as np
import pytest

from oracle.spartan_ref import core as oc

Q = oc.Q


def shard(t, W):
    return [np.ascontiguousarray(t[r::W]) for r in range(W)]


def shard_eq_scale(r, W, rank):
    """engine.hpp: the factor of eq(r, j*W + rank) contributed by the last log2(W) variables"""
    logW = W.bit_length() - 1
    c = 1
    for k in range(logW):
        rj = r[len(r) - logW + k]
        c = c * (rj if (rank >> (logW - 1 - k)) & 1 else (1 - rj)) % Q
    return c


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("ell", [4, 7, 10])
def test_eq_table_slice(W, ell):
    r = oc.to_ints(oc.prg_scalars("eqr", ell, ell + W))
    full = oc.to_ints(oc.eq_evals(r))
    logW = W.bit_length() - 1
    lead = oc.to_ints(oc.eq_evals(r[:ell - logW]))
    for rank in range(W):
        c = shard_eq_scale(r, W, rank)
        assert [v * c % Q for v in lead] == full[rank::W]


@pytest.mark.parametrize("W", [2, 8])
@pytest.mark.parametrize("kind,nt", [(0, 2), (1, 3), (2, 4)])
def test_sharded_rounds_then_gather(W, kind, nt):
    """rounds on cyclic shards (partial sums added over the ranks), then — once the shards are `small` — one local bind, a cyclic all-gather
    into replicated tables and the remaining rounds replicated: every round polynomial and the final evaluations equal the unsharded ones"""
    logn, small = 9, 8
    n = 1 << logn
    tabs = [oc.prg_scalars("t%d" % k, n, 3 * kind + k) for k in range(nt)]
    ch = oc.to_ints(oc.prg_scalars("r", logn, kind))

    def ev(ts):
        if kind == 0:
            e0, e2 = oc.sc_eval_quad(ts[0], ts[1])
            return [e0, e2, 0]
        return oc.sc_eval_cubic(ts[0], ts[1], ts[2], ts[3] if kind == 2 else None)
    want, cur = [], [t.copy() for t in tabs]
    for j in range(logn):
        want.append(ev(cur))
        cur = [oc.bound_top(t, ch[j]).copy() for t in cur]
    finals = [oc.to_ints(t)[0] for t in cur]
    loc = [[s.copy() for s in shard(t, W)] for t in tabs]     # loc[table][rank]
    rep, got = None, []
    for j in range(logn):
        if rep is None:
            parts = [ev([loc[t][rk] for t in range(nt)]) for rk in range(W)]
            got.append([sum(p[k] for p in parts) % Q for k in range(3)])
            loc = [[oc.bound_top(loc[t][rk], ch[j]).copy() for rk in range(W)] for t in range(nt)]
            if len(loc[0][0]) <= small and j + 1 < logn:      # leave the sharded stage: gather (element j of rank rk -> global j*W + rk)
                rep = []
                for t in range(nt):
                    g = oc.zeros(len(loc[t][0]) * W)
                    for rk in range(W):
                        g[rk::W] = loc[t][rk]
                    rep.append(g)
        else:
            got.append(ev(rep))
            rep = [oc.bound_top(t, ch[j]).copy() for t in rep]
    assert got == want
    assert [oc.to_ints(t)[0] for t in rep] == finals


@pytest.mark.parametrize("W", [2, 4])
def test_product_circuit_layers_from_local_entries(W):
    """product_tree.rs:18-56 on cyclic shards: layer k+1 = left half * right half pairs i with i + len/2, both on rank i mod W, so the local
    Hadamard of the local halves IS the shard of the next layer; the root equals the product of all inputs"""
    n = 256
    x = oc.prg_scalars("pc", n, W)
    layers = [x]
    while len(layers[-1]) > 2:
        h = len(layers[-1]) // 2
        layers.append(oc.hadamard(layers[-1][:h].copy(), layers[-1][h:].copy()))
    loc = shard(x, W)
    k = 0
    while len(loc[0]) >= 2 and len(layers[k]) // 2 >= W:
        for rk in range(W):
            assert np.array_equal(loc[rk], layers[k][rk::W])
        nxt = []
        for rk in range(W):
            h = len(loc[rk]) // 2
            nxt.append(oc.hadamard(loc[rk][:h].copy(), loc[rk][h:].copy()))
        loc, k = nxt, k + 1
    assert k >= 3


def test_row_partition_of_commitments():
    """dense_mlpoly.rs:165-177 split by rows: the concatenation of the ranks' row commitments is the commitment of the whole table"""
    L, R, W = 8, 8, 4
    gens = oc.MultiCommitGens.new(R, b"rows-shard")
    Z = oc.prg_scalars("Zrows", L * R, 1)
    blinds = oc.to_ints(oc.prg_scalars("bl", L, 2))
    whole = oc.commit_rows(Z, L, R, blinds, gens)
    parts = []
    for rk in range(W):
        lo, hi = rk * L // W, (rk + 1) * L // W
        parts += oc.commit_rows(np.ascontiguousarray(Z[lo * R:hi * R]), hi - lo, R, blinds[lo:hi], gens)
    assert parts == whole


@pytest.mark.parametrize("W", [1, 2, 4, 8])
@pytest.mark.parametrize("logN", [12, 16, 20])
def test_background_derefs_commitment_row_partition(W, logN):
    """snark.cpp (background-stream commitment of the dereferenced values): the 8N-entry derefs polynomial is an L x R matrix; rows [0, n_part) depend on rx,
    [n_part, 2 n_part) on ry, the rest are zero.  Rank r commits rows [p n_part + r cnt, p n_part + (r+1) cnt) of part p; the all-gather delivers, per rank, its
    part-0 rows followed by its part-1 rows; the host puts them back in matrix order.  Every non-zero row must be committed exactly once and land at its own index."""
    N = 1 << logN
    ell = (8 * N).bit_length() - 1
    L, R = 1 << (ell // 2), 1 << (ell - ell // 2)
    assert L * R == 8 * N and N % R == 0
    n_part = 3 * N // R
    shard_rows = W > 1 and n_part % W == 0
    cnt = n_part // W if shard_rows else n_part
    ranks = W if shard_rows else 1
    gathered = []                                             # rank-major: [rank][part][cnt] row ids, as allgather_block lays the 32-byte encodings out
    for r in range(ranks):
        for p in range(2):
            first = p * n_part + (r * cnt if shard_rows else 0)
            gathered.extend(range(first, first + cnt))
    C = [None] * L
    for r in range(ranks):
        for p in range(2):
            for i in range(cnt):
                C[p * n_part + r * cnt + i] = gathered[(r * 2 + p) * cnt + i]
    assert all(C[i] == i for i in range(2 * n_part)) and all(c is None for c in C[2 * n_part:])
    # the rows of part 0 hold exactly the row-derefs (first 3N entries), part 1 the column-derefs (next 3N)
    assert n_part * R == 3 * N
