"""GPU parity for the variable-base MSM (bucket method, kernels_pip.cu) through the C ABI: `sp_points_*`, `sp_msm_var*`.
Oracle: oracle/ Straus/Pippenger restatement of dalek's vartime_multiscalar_mul (group.rs:98-117).  Run on the B200 box: pytest -m gpu."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.spartan_ref import core as oc  # noqa: E402


@pytest.fixture(scope="module")
def sb():
    import spartan_b200 as m
    m.default_context()
    return m


@pytest.fixture(scope="module")
def pool(sb):
    """one derived point set shared by the tests (device copy + oracle copy)"""
    n = 1 << 16
    from spartan_b200 import api
    return api.Points.derive(n, b"msm-bench"), oc.MultiCommitGens.new(n, b"msm-bench")


def edge_scalars(n, tag):
    sc = oc.prg_scalars(tag, n, n)
    edges = [0, 1, oc.Q - 1, 2, (1 << 252) + 129, 1 << 12, (1 << 12) - 1, 1 << 15, (1 << 16) - 1, (1 << 64) - 1, oc.Q - (1 << 200)]
    for k, e in enumerate(edges):
        if k < n:
            sc[(k * 5) % n] = oc.to_arr([e])[0]
    return sc


def test_points_derive_and_export(sb, pool):
    """MultiCommitGens::new (commitments.rs:15-33) squeezed in slabs == the oracle's one-shot SHAKE stream"""
    P, ref = pool
    assert len(P) == 1 << 16
    got = P.export(0, 300)
    assert got == [ref.g(i).compress() for i in range(300)]
    assert P.export(65000, 17) == [ref.g(65000 + i).compress() for i in range(17)]


def test_points_upload(sb, pool):
    from spartan_b200 import api
    P, ref = pool
    enc = P.export(10, 40) + [bytes(32)]             # the identity encoding is a valid point
    Q = api.Points(enc)
    assert Q.export() == enc
    sc = edge_scalars(41, "up")
    G = np.concatenate([ref.G[10:50], oc.Point.identity().buf[None, :]])
    assert Q.msm(sc) == oc.msm(sc, G).compress()
    bad = enc[:5] + [hashlib.sha256(b"bad0").digest()] + enc[5:]
    assert oc.Point.decompress(bad[5]) is None
    with pytest.raises(api.SpartanB200Error, match="error 8"):
        api.Points(bad)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 190, 1000, 4096, 20000, 65536])
def test_msm_var_vs_oracle(sb, pool, n):
    P, ref = pool
    sc = edge_scalars(n, "msmv")
    assert P.msm(sc) == oc.msm(sc, ref.G).compress()
    small = oc.from_u64(np.arange(n, dtype=np.uint64) % 3)          # 0/1/2-valued vectors
    assert P.msm(small) == oc.msm(small, ref.G).compress()
    rng = np.random.default_rng(n)
    u64 = oc.from_u64(rng.integers(0, 1 << 63, size=n, dtype=np.uint64))   # the 64-bit scalars that dominate in-protocol
    assert P.msm(u64) == oc.msm(u64, ref.G).compress()
    zeros = oc.from_u64(np.zeros(n, dtype=np.uint64))
    assert P.msm(zeros) == bytes(32)


def test_msm_var_empty_and_offset(sb, pool):
    P, ref = pool
    assert P.msm(oc.to_arr([])) == bytes(32)
    sc = edge_scalars(777, "off")
    assert P.msm(sc, offset=12345) == oc.msm(sc, ref.G[12345:12345 + 777]).compress()
    with pytest.raises(Exception):
        P.msm(sc, offset=(1 << 16) - 100)


@pytest.mark.parametrize("c", [2, 5, 6, 8, 11, 13, 15, 16])
def test_msm_var_every_window_width(sb, pool, c):
    """all window widths give the same point (11 divides 253: the case that needs the extra top window)"""
    P, ref = pool
    n = 5000
    sc = edge_scalars(n, "win")
    os.environ["SP_PIP_WINDOW"] = str(c)
    try:
        got = P.msm(sc)
    finally:
        del os.environ["SP_PIP_WINDOW"]
    assert got == oc.msm(sc, ref.G).compress()


def test_msm_var_repeated_and_negated_points(sb, pool):
    """collisions: the same point many times and P with -P (bucket sums that cancel to the identity)"""
    from spartan_b200 import api
    P, ref = pool
    g0, g1 = ref.g(0), ref.g(1)
    neg0 = g0 * (oc.Q - 1)
    enc = [g0.compress()] * 500 + [neg0.compress()] * 500 + [g1.compress()] * 24
    Q = api.Points(enc)
    G = np.stack([oc.Point.decompress(e).buf for e in enc])
    ones = oc.from_u64(np.ones(len(enc), dtype=np.uint64))
    assert Q.msm(ones) == (g1 * 24).compress()
    sc = edge_scalars(len(enc), "rep")
    assert Q.msm(sc) == oc.msm(sc, G).compress()


def test_msm_var_large_properties(sb):
    """2^20 points: split property MSM[0,n) = MSM[0,n/2) + MSM[n/2,n) (two different window plans) and linearity in the scalars;
    one direct oracle comparison at 2^18"""
    from spartan_b200 import api
    n = 1 << 20
    P = api.Points.derive(n, b"msm-large")
    s = oc.prg_scalars("msm", n, 1)
    t = oc.prg_scalars("msm", n, 2)
    a = 0x1234567890abcdef1234567
    S, T = sb.DensePolynomial(s), sb.DensePolynomial(t)
    full = oc.Point.decompress(P.msm(S))
    lo = oc.Point.decompress(P.msm(s[: n // 2]))
    hi = oc.Point.decompress(P.msm(s[n // 2:], offset=n // 2))
    assert (lo + hi).compress() == full.compress()
    comb = oc.to_arr([(a * x + y) % oc.Q for x, y in zip(oc.to_ints(s[:4096]), oc.to_ints(t[:4096]))])
    lhs = P.msm(comb)
    rhs = oc.Point.decompress(P.msm(s[:4096])) * a + oc.Point.decompress(P.msm(t[:4096]))
    assert lhs == rhs.compress()
    m = 1 << 18
    ref = oc.MultiCommitGens.new(m, b"msm-large")
    assert P.msm(t[:m]) == oc.msm(t[:m], ref.G).compress()
    del T


def test_msm_var_2p24_matches_oracle_golden(sb):
    """BASELINE.json configs[2] exactly as bench.py runs it — 2^24 points of the b"msm-bench" generator stream, the scalars of
    numpy.random.default_rng(0) — against the oracle's result for the same inputs (tests/golden/msm_2p24.json, made by
    tests/golden/make_msm_golden.py on the CPU), both for the whole vector and for the eight index-range slices the 8-GPU run computes;
    the slices also have to add up to the whole (point-add all-reduce)."""
    import json
    from spartan_b200 import api
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msm_2p24.json")))
    n = fx["points"]
    P = api.Points.derive(n, b"msm-bench")
    assert hashlib.sha256(b"".join(P.export(i, 1)[0] for i in (0, 1, n // 2, n - 1))).hexdigest() == fx["generator_digest"]
    rng = np.random.default_rng(0)
    t = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64)
    t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    assert P.msm(sb.DensePolynomial(t)).hex() == fx["encoding"]
    per = n // 8
    acc = oc.Point.identity()
    for k in range(8):
        enc = P.msm(sb.DensePolynomial(t[k * per:(k + 1) * per]), offset=k * per)
        assert enc.hex() == fx["slices8"][k], k
        acc = acc + oc.Point.decompress(enc)
    assert acc.compress().hex() == fx["encoding"]
