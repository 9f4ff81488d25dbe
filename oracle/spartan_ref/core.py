"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; it is the checker, never the thing measured as the product or shipped.

CPU restatement of microsoft/Spartan (libspartan 0.9.0, /root/reference) for the prover hot path.
Protocol logic is Python; every O(n) loop and all field / group arithmetic is plain C in
oracle/csrc (liboracle.so).  Scalars at protocol level are Python ints holding the canonical value
in [0, q); vectors are numpy uint64 arrays of shape (n, 4) holding Montgomery-form limbs, i.e. the
exact memory layout of a Rust `Vec<Scalar>` (src/scalar/ristretto255.rs:195-199).

Parity status: F_q is pinned by the reference's own known-answer tests
(src/scalar/ristretto255.rs:777-1201); the group / transcript layers live in third-party crates
(curve25519-dalek ^4.1.1, merlin ^3.0.0, sha3 ^0.8.2 — not under /root/reference) and are pinned
against RFC 9496 vectors, libsodium and the Merlin conformance vector.  The reference holds NO golden
proof bytes, and cannot be compiled here (no Rust toolchain): proof-byte parity is
"parity unpinned" beyond those anchors plus the restated verifier accepting every proof.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

Q = 2**252 + 27742317777372353535851937790883648493
R_MONT = (1 << 256) % Q
R_INV = pow(R_MONT, -1, Q)

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.dirname(_HERE)
_LIB_PATH = os.path.join(_ORACLE_DIR, "liboracle.so")


def build(force=False):
    """Compile oracle/csrc -> oracle/liboracle.so (plain C, gcc)."""
    srcs = [os.path.join(_ORACLE_DIR, "csrc", f) for f in os.listdir(os.path.join(_ORACLE_DIR, "csrc"))]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _ORACLE_DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    if not os.path.exists(_LIB_PATH):
        build()
    return C.CDLL(_LIB_PATH)


lib = _load()
_vp = C.c_void_p
_sz = C.c_size_t


def _ptr(a):
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_vp)
    return a


# ----------------------------------------------------------------------------- scalars
def mont_bytes(v):
    """canonical int -> 32 bytes of Montgomery limbs (what bincode writes for a Scalar)"""
    return ((v % Q) * R_MONT % Q).to_bytes(32, "little")


def from_mont_bytes(b):
    return int.from_bytes(bytes(b), "little") * R_INV % Q


def to_arr(vals):
    """list of canonical ints -> (n,4) uint64 Montgomery array"""
    buf = b"".join(mont_bytes(v) for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def to_ints(arr):
    b = np.ascontiguousarray(arr).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") * R_INV % Q for i in range(0, len(b), 32)]


def arr_get(arr, i):
    return from_mont_bytes(arr[i].tobytes())


def zeros(n):
    return np.zeros((n, 4), dtype=np.uint64)


def scalar_to_bytes(v):
    """Scalar::to_bytes (ristretto255.rs:419): canonical little-endian"""
    return (v % Q).to_bytes(32, "little")


def scalar_from_bytes_wide(b):
    """Scalar::from_bytes_wide (ristretto255.rs:435) through the C restatement"""
    out = np.zeros(4, dtype=np.uint64)
    lib.fq_from_bytes_wide(_ptr(out), C.c_char_p(bytes(b)))
    return from_mont_bytes(out.tobytes())


def inv(v):
    return pow(v % Q, -1, Q)


def _fqp(v):
    """canonical int -> pointer-able Montgomery buffer"""
    return np.frombuffer(mont_bytes(v), dtype=np.uint64).copy()


# ----------------------------------------------------------------------------- deterministic inputs
def prg_scalars(tag, n, seed=0):
    """SHAKE256("spartan-b200/v1/" || tag || LE64(seed)), 64 bytes per scalar -> from_bytes_wide.
    Same map as Scalar::random (ristretto255.rs:374-380).  Returns an (n,4) Montgomery array."""
    raw = hashlib.shake_256(b"spartan-b200/v1/" + tag.encode() + int(seed).to_bytes(8, "little")).digest(64 * n)
    out = zeros(n)
    lib.fq_from_bytes_wide_batch(_ptr(out), C.c_char_p(raw), _sz(n))
    return out


# ----------------------------------------------------------------------------- vector ops (C loops)
def bound_top(Z, r):
    """DensePolynomial::bound_poly_var_top (dense_mlpoly.rs:215-223); returns the halved array"""
    Z = np.ascontiguousarray(Z)
    lib.poly_bound_top(_ptr(Z), _sz(len(Z)), _ptr(_fqp(r)))
    return Z[: len(Z) // 2]


def bound_bot_ints(vals, r):
    """DensePolynomial::bound_poly_var_bot (dense_mlpoly.rs:225-233) on a short list"""
    n = len(vals) // 2
    return [(vals[2 * i] + r * (vals[2 * i + 1] - vals[2 * i])) % Q for i in range(n)]


def eq_evals(r):
    """EqPolynomial::evals (dense_mlpoly.rs:68-84)"""
    out = zeros(1 << len(r))
    rr = to_arr(r) if len(r) else zeros(0)
    lib.poly_eq_evals(_ptr(out), _ptr(rr), _sz(len(r)))
    return out


def dot(a, b):
    out = np.zeros(4, dtype=np.uint64)
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert len(a) == len(b)
    lib.poly_dot(_ptr(out), _ptr(a), _ptr(b), _sz(len(a)))
    return from_mont_bytes(out.tobytes())


def dot3(a, b, c):
    out = np.zeros(4, dtype=np.uint64)
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b); c = np.ascontiguousarray(c)
    lib.poly_dot3(_ptr(out), _ptr(a), _ptr(b), _ptr(c), _sz(len(a)))
    return from_mont_bytes(out.tobytes())


def bound_rows(Z, L, L_size, R_size):
    """DensePolynomial::bound (dense_mlpoly.rs:206-213)"""
    out = zeros(R_size)
    Z = np.ascontiguousarray(Z); L = np.ascontiguousarray(L)
    lib.poly_bound_rows(_ptr(out), _ptr(Z), _ptr(L), _sz(L_size), _sz(R_size))
    return out


def sc_eval_quad(A, B):
    out = zeros(2)
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    lib.sc_eval_quad(_ptr(out), _ptr(A), _ptr(B), _sz(len(A)))
    return to_ints(out)


def sc_eval_cubic(A, B, Cc, D=None):
    out = zeros(3)
    A = np.ascontiguousarray(A); B = np.ascontiguousarray(B); Cc = np.ascontiguousarray(Cc)
    if D is not None:
        D = np.ascontiguousarray(D)
    lib.sc_eval_cubic(_ptr(out), _ptr(A), _ptr(B), _ptr(Cc), _ptr(D) if D is not None else None, _sz(len(A)))
    return to_ints(out)


def hadamard(a, b):
    out = zeros(len(a))
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    lib.poly_hadamard(_ptr(out), _ptr(a), _ptr(b), _sz(len(a)))
    return out


def lincomb3(A, B, Cc, ra, rb, rc):
    out = zeros(len(A))
    lib.poly_lincomb3(_ptr(out), _ptr(np.ascontiguousarray(A)), _ptr(np.ascontiguousarray(B)), _ptr(np.ascontiguousarray(Cc)),
                      _ptr(_fqp(ra)), _ptr(_fqp(rb)), _ptr(_fqp(rc)), _sz(len(A)))
    return out


def from_u64(vals):
    v = np.ascontiguousarray(np.asarray(vals, dtype=np.uint64))
    out = zeros(len(v))
    lib.poly_from_u64(_ptr(out), _ptr(v), _sz(len(v)))
    return out


def evaluate(Z, r):
    """DensePolynomial::evaluate (dense_mlpoly.rs:236-242)"""
    assert len(Z) == 1 << len(r)
    return dot(Z, eq_evals(r))


# ----------------------------------------------------------------------------- group
class Point:
    """GroupElement = RistrettoPoint (group.rs:6); 160-byte extended-coordinates blob owned by C."""
    __slots__ = ("buf",)

    def __init__(self, buf=None):
        self.buf = buf if buf is not None else np.zeros(20, dtype=np.uint64)

    @staticmethod
    def identity():
        p = Point(); lib.ge_identity(_ptr(p.buf)); return p

    @staticmethod
    def decompress(b):
        p = Point()
        ok = lib.ristretto_decode(_ptr(p.buf), C.c_char_p(bytes(b)))
        return p if ok else None

    @staticmethod
    def from_uniform_bytes(b):
        p = Point(); lib.ristretto_from_uniform_bytes(_ptr(p.buf), C.c_char_p(bytes(b))); return p

    def compress(self):
        out = C.create_string_buffer(32)
        lib.ristretto_encode(out, _ptr(self.buf))
        return out.raw

    def __add__(self, o):
        p = Point(); lib.ge_add(_ptr(p.buf), _ptr(self.buf), _ptr(o.buf)); return p

    def __sub__(self, o):
        p = Point(); lib.ge_sub(_ptr(p.buf), _ptr(self.buf), _ptr(o.buf)); return p

    def __mul__(self, k):
        """Scalar * Point via canonical bytes (group.rs:33-46, scalar/mod.rs:32-36)"""
        p = Point(); lib.ge_scalarmul_bytes(_ptr(p.buf), C.c_char_p(scalar_to_bytes(k)), _ptr(self.buf)); return p

    __rmul__ = __mul__

    def __eq__(self, o):
        return bool(lib.ge_eq(_ptr(self.buf), _ptr(o.buf)))


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")


def points_array(points):
    return np.concatenate([p.buf for p in points]).reshape(-1, 20) if points else np.zeros((0, 20), dtype=np.uint64)


def msm(scalars, G_arr):
    """GroupElement::vartime_multiscalar_mul (group.rs:98-117).  scalars: Montgomery array or list of ints"""
    if not isinstance(scalars, np.ndarray):
        scalars = to_arr(scalars)
    scalars = np.ascontiguousarray(scalars)
    n = len(scalars)
    assert len(G_arr) >= n
    p = Point()
    lib.ge_msm(_ptr(p.buf), _ptr(scalars), _ptr(G_arr), _sz(n))
    return p


class MultiCommitGens:
    """commitments.rs:8-67.  G is a (n,20) uint64 array of extended points, h a Point."""

    def __init__(self, n, G, h):
        self.n, self.G, self.h = n, G, h

    @staticmethod
    def new(n, label):
        # commitments.rs:15-33: SHAKE256(label || basepoint) read 64 bytes at a time, n+1 times
        uniform = hashlib.shake_256(label + BASEPOINT_COMPRESSED).digest(64 * (n + 1))
        pts = np.zeros((n + 1, 20), dtype=np.uint64)
        lib.gens_from_uniform(_ptr(pts), C.c_char_p(uniform), _sz(n + 1))
        return MultiCommitGens(n, pts[:n].copy(), Point(pts[n].copy()))

    def split_at(self, mid):
        return (MultiCommitGens(mid, self.G[:mid].copy(), self.h), MultiCommitGens(self.n - mid, self.G[mid:].copy(), self.h))

    def scale(self, s):
        G = np.stack([(Point(self.G[i].copy()) * s).buf for i in range(self.n)]) if self.n else self.G
        return MultiCommitGens(self.n, G, self.h)

    def g(self, i):
        return Point(self.G[i].copy())


def commit_scalar(x, blind, gens):
    """impl Commitments for Scalar (commitments.rs:73-78)"""
    assert gens.n == 1
    return msm([x, blind], np.concatenate([gens.G[:1], gens.h.buf.reshape(1, 20)]))


def commit_vec(x, blind, gens):
    """impl Commitments for [Scalar] (commitments.rs:80-92)"""
    n = len(x)
    assert gens.n == n
    return msm(x, gens.G) + gens.h * blind


def commit_rows(Z, L_size, R_size, blinds, gens):
    """DensePolynomial::commit_inner (dense_mlpoly.rs:148-177) -> list of 32-byte compressed points"""
    out = C.create_string_buffer(32 * L_size)
    Z = np.ascontiguousarray(Z)
    bl = to_arr(blinds)
    lib.poly_commit_rows(out, _ptr(Z), _sz(L_size), _sz(R_size), _ptr(bl), _ptr(gens.G), _ptr(gens.h.buf))
    return [out.raw[32 * i:32 * i + 32] for i in range(L_size)]


# ----------------------------------------------------------------------------- transcript
class Transcript:
    """merlin::Transcript + the ProofTranscript extension trait (transcript.rs:5-37)."""

    def __init__(self, label):
        self.st = C.create_string_buffer(lib.merlin_sizeof())
        lib.merlin_init(self.st, C.c_char_p(label), _sz(len(label)))

    def append_message(self, label, msg):
        lib.merlin_append_message(self.st, C.c_char_p(label), _sz(len(label)), C.c_char_p(bytes(msg)), _sz(len(msg)))

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        out = C.create_string_buffer(n)
        lib.merlin_challenge_bytes(self.st, C.c_char_p(label), _sz(len(label)), out, _sz(n))
        return out.raw

    # ProofTranscript
    def append_protocol_name(self, name):
        self.append_message(b"protocol-name", name)

    def append_scalar(self, label, s):
        self.append_message(label, scalar_to_bytes(s))

    def append_point(self, label, p):
        assert len(p) == 32
        self.append_message(label, p)

    def challenge_scalar(self, label):
        return scalar_from_bytes_wide(self.challenge_bytes(label, 64))

    def challenge_vector(self, label, n):
        return [self.challenge_scalar(label) for _ in range(n)]

    # AppendToTranscript for [Scalar] (transcript.rs:49-57)
    def append_scalars(self, label, vals):
        self.append_message(label, b"begin_append_vector")
        if isinstance(vals, np.ndarray):
            buf = C.create_string_buffer(32 * len(vals))
            v = np.ascontiguousarray(vals)
            lib.fq_to_bytes_batch(buf, _ptr(v), _sz(len(v)))
            lib.merlin_append_many32(self.st, C.c_char_p(label), _sz(len(label)), buf, _sz(len(v)))
        else:
            for s in vals:
                self.append_scalar(label, s)
        self.append_message(label, b"end_append_vector")


class RandomTape:
    """random.rs:6-28, with the OsRng seed scalar made an explicit input."""

    def __init__(self, name, seed_scalar):
        self.tape = Transcript(name)
        self.tape.append_scalar(b"init_randomness", seed_scalar)

    def random_scalar(self, label):
        return self.tape.challenge_scalar(label)

    def random_vector(self, label, n):
        return self.tape.challenge_vector(label, n)


def log_2(x):
    """Math::log_2 (math.rs:21-29): floor for powers of two, ceil otherwise"""
    assert x != 0
    return x.bit_length() - 1 if x & (x - 1) == 0 else x.bit_length()


def next_pow2(x):
    return 1 if x <= 1 else 1 << (x - 1).bit_length()
