"""The arithmetic the kernels run (spartan_b200/csrc/field.cuh + curve.cuh, 32-bit-limb device formulation compiled for the
host with -DSP_FORCE_PORTABLE) and the host fast paths, checked against the oracle WITHOUT a GPU.  The oracle is the checker."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from oracle.spartan_ref import core as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**255 - 19
Q = oc.Q


def _lib(kind):
    path = os.path.join(ROOT, "spartan_b200", "libsp_hosttest_%s.so" % kind)
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build_hosttest()
    return C.CDLL(path)


@pytest.fixture(params=["portable", "fast"])
def lib(request):
    l = _lib(request.param)
    assert l.spt_portable() == (1 if request.param == "portable" else 0)
    return l


def b32(x):
    return int(x).to_bytes(32, "little")


def call2(lib, name, a, b):
    out = C.create_string_buffer(32)
    getattr(lib, name)(C.c_char_p(a), C.c_char_p(b), out)
    return out.raw


def call1(lib, name, a):
    out = C.create_string_buffer(32)
    getattr(lib, name)(C.c_char_p(a), out)
    return out.raw


EDGE_Q = [0, 1, 2, Q - 1, Q - 2, oc.R_MONT, (oc.R_MONT * oc.R_MONT) % Q, 2**252, 2**252 - 1, 2**128, 2**64 - 1, 2**32 - 1, 2**32]


def test_fq_against_oracle(lib):
    rng = np.random.default_rng(3)
    vals = EDGE_Q + [int.from_bytes(rng.bytes(32), "little") % Q for _ in range(300)]
    for i in range(len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        A, B = oc.mont_bytes(a), oc.mont_bytes(b)
        assert oc.from_mont_bytes(call2(lib, "spt_fq_mul", A, B)) == a * b % Q
        assert oc.from_mont_bytes(call2(lib, "spt_fq_add", A, B)) == (a + b) % Q
        assert oc.from_mont_bytes(call2(lib, "spt_fq_sub", A, B)) == (a - b) % Q
        assert int.from_bytes(call1(lib, "spt_fq_from_mont", A), "little") == a
    for a in vals[:40]:
        got = oc.from_mont_bytes(call1(lib, "spt_fq_inv", oc.mont_bytes(a)))
        assert got == (pow(a, -1, Q) if a else 0)
    for _ in range(50):
        w = rng.bytes(64)
        assert oc.from_mont_bytes(call1(lib, "spt_fq_from_wide", w)) == int.from_bytes(w, "little") % Q
    assert oc.from_mont_bytes(call1(lib, "spt_fq_from_wide", b"\xff" * 64)) == (2**512 - 1) % Q   # ristretto255.rs:995-1005
    out = C.create_string_buffer(32)
    lib.spt_fq_from_u64(C.c_uint64(2**64 - 1), out)
    assert oc.from_mont_bytes(out.raw) == 2**64 - 1


def test_fq_fold_const_equals_bound_poly_var_top(lib):
    """the constant-multiplier fold a0 + r*(a1 - a0) (table of r*2^(32j) mod q, one fold through 2^252 = -c) returns the very limbs of
    dense_mlpoly.rs:218 computed the long way, including at the edges of the field"""
    rng = np.random.default_rng(17)
    vals = EDGE_Q + [int.from_bytes(rng.bytes(32), "little") % Q for _ in range(200)]
    out = C.create_string_buffer(32)
    for i in range(len(vals)):
        a0, a1, r = vals[i], vals[(3 * i + 1) % len(vals)], vals[(7 * i + 5) % len(vals)]
        lib.spt_fq_fold_const(C.c_char_p(oc.mont_bytes(a0)), C.c_char_p(oc.mont_bytes(a1)), C.c_char_p(oc.mont_bytes(r)), out)
        assert out.raw == oc.mont_bytes((a0 + r * (a1 - a0)) % Q), (a0, a1, r)


def test_fq_mul_bit_exact_with_reference_restatement(lib):
    """same Montgomery limbs as oracle/csrc/fq.c (the restatement of ristretto255.rs:690-714) on raw limb inputs"""
    rng = np.random.default_rng(5)
    for _ in range(300):
        a = int.from_bytes(rng.bytes(32), "little") % Q
        b = int.from_bytes(rng.bytes(32), "little") % Q
        r = np.zeros(4, dtype=np.uint64)
        A = np.frombuffer(b32(a), dtype=np.uint64).copy()
        B = np.frombuffer(b32(b), dtype=np.uint64).copy()
        oc.lib.fq_mul(oc._ptr(r), oc._ptr(A), oc._ptr(B))
        assert call2(lib, "spt_fq_mul", b32(a), b32(b)) == r.tobytes()


def test_fp_against_python_ints(lib):
    rng = np.random.default_rng(4)
    edge = [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2**255 - 1, 2 * P, 2 * P + 1]
    vals = edge + [int.from_bytes(rng.bytes(32), "little") for _ in range(300)]
    for i in range(len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        assert int.from_bytes(call2(lib, "spt_fp_mul", b32(a), b32(b)), "little") == a * b % P
        assert int.from_bytes(call2(lib, "spt_fp_add", b32(a), b32(b)), "little") == (a + b) % P
        assert int.from_bytes(call2(lib, "spt_fp_sub", b32(a), b32(b)), "little") == (a - b) % P
        assert int.from_bytes(call1(lib, "spt_fp_canon", b32(a)), "little") == a % P
    for a in vals[:30]:
        if a % P:
            assert int.from_bytes(call1(lib, "spt_fp_inv", b32(a)), "little") == pow(a, -1, P)


def test_fe51_against_python_ints(lib):
    """host_fe51.hpp: radix-2^51 arithmetic mod 2^255-19 on arbitrary 256-bit inputs (incl. non-canonical ones), results frozen to [0, p)"""
    rng = np.random.default_rng(51)
    out = C.create_string_buffer(160)
    edge = [0, 1, 2, 19, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**51 - 1, 2**51, 2**102 - 1]
    vals = edge + [int.from_bytes(rng.bytes(32), "little") for _ in range(200)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        lib.spt_fe51_ops(C.c_char_p(b32(a)), C.c_char_p(b32(b)), out)
        got = [int.from_bytes(out.raw[32 * k:32 * k + 32], "little") for k in range(5)]
        assert got == [a * b % P, a * a % P, (a + b) % P, (a - b) % P, (2 * a + b) * (2 * a + 2 * b) % P], (a, b)


def test_group_against_oracle(lib):
    out = C.create_string_buffer(32)
    pts = []
    for i in range(24):
        h = hashlib.sha512(b"hm%d" % i).digest()
        lib.spt_from_uniform(C.c_char_p(h), out)
        ref = oc.Point.from_uniform_bytes(h)
        assert out.raw == ref.compress()
        assert lib.spt_decode_encode(C.c_char_p(out.raw), out) == 1 and out.raw == ref.compress()
        pts.append(ref)
    ident = bytes(32)
    assert lib.spt_decode_encode(C.c_char_p(ident), out) == 1 and out.raw == ident
    for i in range(0, 24, 2):
        a, b = pts[i], pts[i + 1]
        assert lib.spt_add(C.c_char_p(a.compress()), C.c_char_p(b.compress()), out) == 1
        assert out.raw == (a + b).compress()
        assert lib.spt_dbl(C.c_char_p(a.compress()), out) == 1 and out.raw == (a + a).compress()
        assert lib.spt_add(C.c_char_p(a.compress()), C.c_char_p(ident), out) == 1 and out.raw == a.compress()
    o64 = C.create_string_buffer(64)
    for i in range(0, 24, 2):     # the two-at-a-time encoder used on the prover's host side gives the same bytes
        a, b = pts[i], pts[i + 1]
        assert lib.spt_compress2(C.c_char_p(a.compress()), C.c_char_p(b.compress()), o64) == 1
        assert o64.raw == (a + b).compress() + (a + a).compress()
    assert lib.spt_compress2(C.c_char_p(ident), C.c_char_p(ident), o64) == 1 and o64.raw == ident + ident
    rng = np.random.default_rng(9)
    o96 = C.create_string_buffer(96)
    for i, k in enumerate([0, 1, 8, 9, 15, 16, 2**252 + 9, Q - 1, 2**255 - 1] + [int.from_bytes(rng.bytes(32), "little") % Q for _ in range(6)]):
        a, b = pts[i % 24], pts[(i + 5) % 24]      # host radix-2^51 point arithmetic: windowed scalar mult, add, double
        assert lib.spt_hge_ops(C.c_char_p(b32(k)), C.c_char_p(a.compress()), C.c_char_p(b.compress()), o96) == 1
        assert o96.raw == (a * (k % Q)).compress() + (a + b).compress() + (a + a).compress(), k
    for i in range(6):
        k = int.from_bytes(rng.bytes(32), "little") % Q
        assert lib.spt_scalarmul(C.c_char_p(b32(k)), C.c_char_p(pts[i].compress()), out) == 1
        assert out.raw == (pts[i] * k).compress()
    for k in [0, 1, 127, 128, 129, 255, 256, 2**8 * 129, Q - 1, 2**252, (1 << 252) - 1, int.from_bytes(rng.bytes(32), "little") % Q]:
        assert lib.spt_fixed_base_mul(C.c_char_p(oc.mont_bytes(k)), C.c_char_p(pts[0].compress()), out) == 1
        assert out.raw == (pts[0] * k).compress(), k
    # bad encodings are rejected exactly like the oracle (RFC 9496 A.2 + random strings)
    for i in range(200):
        b = hashlib.sha256(b"bad%d" % i).digest()
        assert bool(lib.spt_decode_encode(C.c_char_p(b), out)) == (oc.Point.decompress(b) is not None)


def test_host_transcript_and_shake(lib):
    out = C.create_string_buffer(32)
    lib.spt_transcript_kat(out)
    assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    rng = np.random.default_rng(2)
    for n in [1, 5, 40]:
        lens = [int(x) for x in rng.integers(0, 400, size=n)]
        msgs = [rng.bytes(l) for l in lens]
        o64 = C.create_string_buffer(64)
        arr = (C.c_size_t * n)(*lens)
        lib.spt_transcript_run(C.c_char_p(b"lbl"), C.c_size_t(3), C.c_char_p(b"".join(msgs)), arr, C.c_size_t(n), o64)
        t = oc.Transcript(b"lbl")
        for m in msgs:
            t.append_message(b"m", m)
        assert o64.raw == t.challenge_bytes(b"c", 64)
    for n_in, n_out in [(0, 64), (135, 136), (136, 300), (500, 1000)]:
        msg = rng.bytes(n_in)
        o = C.create_string_buffer(n_out)
        lib.spt_shake256(o, C.c_size_t(n_out), C.c_char_p(msg), C.c_size_t(n_in))
        assert o.raw == hashlib.shake_256(msg).digest(n_out)
