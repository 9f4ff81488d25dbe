"""ctypes mirror of the reference's public API over libspartan_b200.so (see include/spartan_b200.h)."""
import ctypes as C
import hashlib
import os
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspartan_b200%s.so" % os.environ.get("SP_LIB_TAG", ""))  # SP_LIB_TAG: tuning builds only

SP_OK, SP_ERR_NO_DEVICE, SP_ERR_CUDA, SP_ERR_INVALID_ARG, SP_ERR_INVALID_INDEX, SP_ERR_INVALID_SCALAR, SP_ERR_INVALID_INPUTS, SP_ERR_INTERNAL = range(8)
SP_ERR_INVALID_POINT, SP_ERR_VERIFY, SP_ERR_DECOMPRESS = 8, 9, 10


class SpartanB200Error(RuntimeError):
    pass


class ProofVerifyError(SpartanB200Error):
    """errors.rs:5-12: InternalError (a check failed) or DecompressionError (a point of the proof does not decompress)"""


class R1CSError(SpartanB200Error):
    """errors.rs:28-41: InvalidIndex / InvalidScalar / InvalidNumberOfInputs"""


def _load():
    if not os.path.exists(_LIB_PATH):
        raise SpartanB200Error("libspartan_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    return C.CDLL(_LIB_PATH)


lib = _load()
lib.sp_last_error.restype = C.c_char_p
lib.sp_kernel_launches.restype = C.c_ulonglong
lib.sp_poly_len.restype = C.c_size_t
lib.sp_points_len.restype = C.c_size_t
_vp, _sz = C.c_void_p, C.c_size_t


def kernel_launches():
    return int(lib.sp_kernel_launches())


def _p(a):
    return a.ctypes.data_as(_vp) if isinstance(a, np.ndarray) else a


def _limbs(a):
    """(n,4) uint64 contiguous view of Montgomery limbs"""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, 4)


class Context:
    """One per GPU / host thread.  Fails loudly when no CUDA device is present."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib.sp_ctx_create(C.c_int(device), C.byref(h))
        if rc != SP_OK:
            raise SpartanB200Error("sp_ctx_create failed (%d): %s" % (rc, lib.sp_last_error(None).decode()))
        self.h = h
        self.device = device

    def check(self, rc):
        if rc == SP_OK:
            return
        msg = lib.sp_last_error(self.h).decode()
        if rc == SP_ERR_INVALID_INDEX:
            raise R1CSError("InvalidIndex: " + msg)
        if rc == SP_ERR_INVALID_SCALAR:
            raise R1CSError("InvalidScalar: " + msg)
        if rc == SP_ERR_INVALID_INPUTS:
            raise R1CSError("InvalidNumberOfInputs: " + msg)
        if rc == SP_ERR_VERIFY:
            raise ProofVerifyError("InternalError: " + msg)
        if rc == SP_ERR_DECOMPRESS:
            raise ProofVerifyError("DecompressionError: " + msg)
        raise SpartanB200Error("spartan_b200 error %d: %s" % (rc, msg))

    def connect_peers(self, rank, world, allgather):
        """Join `world` single-GPU processes into one sharded prover (sp_comm_export / sp_comm_connect).  `allgather(b: bytes) -> list[bytes]`
        is the host-side exchange (every rank contributes its handle, all get the list in rank order): spartan_b200.dist.allgather_bytes under
        torch.distributed.  Afterwards NIZK.prove / SNARK.prove must be called by every rank with the same arguments."""
        lib.sp_comm_handle_bytes.restype = C.c_size_t
        hb = int(lib.sp_comm_handle_bytes())
        mine = C.create_string_buffer(hb)
        self.check(lib.sp_comm_export(self.h, mine))
        handles = allgather(mine.raw)
        assert len(handles) == world and all(len(h) == hb for h in handles)
        self.check(lib.sp_comm_connect(self.h, C.c_int(rank), C.c_int(world), C.c_char_p(b"".join(handles))))
        self.rank, self.world = rank, world

    def set_sharding(self, enabled):
        """switch sharded proving off / on for a connected context (every rank must switch together)"""
        lib.sp_comm_set_enabled(self.h, C.c_int(1 if enabled else 0))

    def set_overlap(self, enabled):
        """background-stream commitment of the dereferenced values in SNARK.prove on / off (off: every kernel on one stream, for per-kernel profiling)"""
        lib.sp_ctx_set_overlap(self.h, C.c_int(1 if enabled else 0))

    def timings(self):
        buf = C.create_string_buffer(8192)
        lib.sp_timings(self.h, buf, _sz(8192))
        out = {}
        for item in buf.value.decode().split(";"):
            if "=" in item:
                k, v = item.split("=")
                out[k] = float(v)
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if lib.sp_device_count() > 1 else 0)
    return _default_ctx


# ----------------------------------------------------------------------------- scalars
def scalar_from_bytes(b):
    """Scalar::from_bytes (ristretto255.rs:391): canonical 32 bytes -> Montgomery limbs; raises on non-canonical input"""
    out = np.zeros(4, dtype=np.uint64)
    if lib.sp_scalar_from_bytes(C.c_char_p(bytes(b)), _p(out)) != SP_OK:
        raise R1CSError("InvalidScalar")
    return out


def scalar_to_bytes(limbs):
    out = C.create_string_buffer(32)
    lib.sp_scalar_to_bytes(_p(np.ascontiguousarray(limbs, dtype=np.uint64)), out)
    return out.raw


def prg_scalars(tag, n, seed=0):
    """DESIGN.md deterministic inputs: SHAKE256("spartan-b200/v1/" || tag || LE64(seed)) -> from_bytes_wide, as Montgomery limbs"""
    raw = hashlib.shake_256(b"spartan-b200/v1/" + tag.encode() + int(seed).to_bytes(8, "little")).digest(64 * n)
    out = np.zeros((n, 4), dtype=np.uint64)
    for i in range(n):
        lib.sp_scalar_from_bytes_wide(C.c_char_p(raw[64 * i:64 * i + 64]), _p(out[i]))
    return out


def tape_seed(seed=0):
    """a DETERMINISTIC RandomTape seed (tests / benchmarks / byte-parity runs only): every blind of the proof becomes a public function
    of `seed`, so a proof made with it is not zero-knowledge.  Real proofs use random_tape_seed() (the default of NIZK.prove / SNARK.prove)."""
    return prg_scalars("tape", 1, seed)[0]


def random_tape_seed():
    """the scalar RandomTape::new draws from OsRng (random.rs:13-15): 64 bytes of OS randomness -> Scalar::from_bytes_wide"""
    out = np.zeros(4, dtype=np.uint64)
    lib.sp_scalar_from_bytes_wide(C.c_char_p(os.urandom(64)), _p(out))
    return out


# ----------------------------------------------------------------------------- operator level
class DensePolynomial:
    """dense_mlpoly.rs:18-22 with Z resident on the device"""

    def __init__(self, Z, ctx=None):
        self.ctx = ctx or default_context()
        Z = _limbs(Z)
        h = _vp()
        self.ctx.check(lib.sp_poly_upload(self.ctx.h, _p(Z), _sz(len(Z)), C.byref(h)))
        self.h = h

    @classmethod
    def _wrap(cls, ctx, h):
        o = cls.__new__(cls)
        o.ctx, o.h = ctx, h
        return o

    def len(self):
        return int(lib.sp_poly_len(self.h))

    def to_numpy(self):
        out = np.zeros((self.len(), 4), dtype=np.uint64)
        self.ctx.check(lib.sp_poly_download(self.ctx.h, self.h, _p(out), _sz(len(out))))
        return out

    def bound_poly_var_top(self, r):
        arr = (_vp * 1)(self.h)
        self.ctx.check(lib.sp_fold_top(self.ctx.h, arr, C.c_int(1), _p(np.ascontiguousarray(r, dtype=np.uint64))))

    def evaluate(self, r):
        r = _limbs(r)
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib.sp_poly_evaluate(self.ctx.h, self.h, _p(r), _sz(len(r)), _p(out)))
        return out

    def bound(self, L):
        L = _limbs(L)
        h = _vp()
        self.ctx.check(lib.sp_poly_bound_rows(self.ctx.h, self.h, _p(L), _sz(len(L)), C.byref(h)))
        return DensePolynomial._wrap(self.ctx, h)

    def dot(self, other):
        out = np.zeros(4, dtype=np.uint64)
        self.ctx.check(lib.sp_dot(self.ctx.h, self.h, other.h, _p(out)))
        return out

    def commit(self, gens, L, R, blinds=None):
        """DensePolynomial::commit_inner (dense_mlpoly.rs:148-177) -> list of L compressed points"""
        out = C.create_string_buffer(32 * L)
        bl = _limbs(blinds) if blinds is not None else None
        self.ctx.check(lib.sp_commit_rows(self.ctx.h, gens.h, self.h, _sz(L), _sz(R), _p(bl) if bl is not None else None, out))
        return [out.raw[32 * i:32 * i + 32] for i in range(L)]

    @staticmethod
    def eq_evals(r, ctx=None):
        ctx = ctx or default_context()
        r = _limbs(r)
        h = _vp()
        ctx.check(lib.sp_eq_evals(ctx.h, _p(r), _sz(len(r)), C.byref(h)))
        return DensePolynomial._wrap(ctx, h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_poly_free(self.h)
                self.h = None
        except Exception:
            pass


def sumcheck_eval(kind, polys):
    ctx = polys[0].ctx
    arr = (_vp * len(polys))(*[p.h for p in polys])
    out = np.zeros((3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_eval(ctx.h, C.c_int(kind), arr, _p(out)))
    return out


def sumcheck_fold_eval(kind, polys, r):
    ctx = polys[0].ctx
    arr = (_vp * len(polys))(*[p.h for p in polys])
    out = np.zeros((3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_fold_eval(ctx.h, C.c_int(kind), arr, _p(np.ascontiguousarray(r, dtype=np.uint64)), _p(out)))
    return out


def sumcheck_eval_sharded(kind, polys):
    """sumcheck_eval on cyclic shards held by the ranks of a connected context: returns the evaluations of the whole tables on every rank"""
    ctx = polys[0].ctx
    out = np.zeros((3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_eval_sharded(ctx.h, C.c_int(kind), (_vp * len(polys))(*[p.h for p in polys]), _p(out)))
    return out


def sumcheck_fold_eval_sharded(kind, polys, r):
    ctx = polys[0].ctx
    out = np.zeros((3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_fold_eval_sharded(ctx.h, C.c_int(kind), (_vp * len(polys))(*[p.h for p in polys]), _p(np.ascontiguousarray(r, dtype=np.uint64)), _p(out)))
    return out


def _handles(polys):
    return (_vp * len(polys))(*[p.h for p in polys])


def sumcheck_batched_eval(As, Bs, Cs):
    """prove_cubic_batched evaluation loops (sumcheck.rs:290-357) for len(As) instances of A*B*C; returns (ninst, 3, 4) limbs [e0, e2, e3]"""
    ctx = As[0].ctx
    out = np.zeros((len(As), 3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_batched_eval(ctx.h, C.c_int(len(As)), _handles(As), _handles(Bs), _handles(Cs), _p(out)))
    return out


def sumcheck_batched_fold_eval(As, Bs, Cs, r):
    """bind the top variable of every table with r (a C shared by several instances once) and evaluate the next round"""
    ctx = As[0].ctx
    out = np.zeros((len(As), 3, 4), dtype=np.uint64)
    ctx.check(lib.sp_sumcheck_batched_fold_eval(ctx.h, C.c_int(len(As)), _handles(As), _handles(Bs), _handles(Cs), _p(np.ascontiguousarray(r, dtype=np.uint64)), _p(out)))
    return out


def fold_top(polys, r):
    ctx = polys[0].ctx
    arr = (_vp * len(polys))(*[p.h for p in polys])
    ctx.check(lib.sp_fold_top(ctx.h, arr, C.c_int(len(polys)), _p(np.ascontiguousarray(r, dtype=np.uint64))))


class MultiCommitGens:
    """commitments.rs:8-33"""

    def __init__(self, n, label, ctx=None):
        self.ctx = ctx or default_context()
        self.n = n
        h = _vp()
        self.ctx.check(lib.sp_gens_create(self.ctx.h, C.c_char_p(label), _sz(len(label)), _sz(n), C.byref(h)))
        self.h = h

    @classmethod
    def from_points(cls, compressed, ctx=None):
        """the caller's own generators: n+1 ristretto255 encodings, G[0..n) then h (e.g. a `scale`d MultiCommitGens, commitments.rs:43-49)"""
        o = cls.__new__(cls)
        o.ctx = ctx or default_context()
        o.n = len(compressed) - 1
        h = _vp()
        o.ctx.check(lib.sp_gens_upload(o.ctx.h, C.c_char_p(b"".join(compressed)), _sz(o.n), C.byref(h)))
        o.h = h
        return o

    def export(self):
        out = C.create_string_buffer(32 * (self.n + 1))
        self.ctx.check(lib.sp_gens_export(self.ctx.h, self.h, out))
        return [out.raw[32 * i:32 * i + 32] for i in range(self.n + 1)]

    def msm(self, scalars):
        """GroupElement::vartime_multiscalar_mul(scalars, G).compress()"""
        s = _limbs(scalars)
        out = C.create_string_buffer(32)
        self.ctx.check(lib.sp_msm(self.ctx.h, self.h, _p(s), _sz(len(s)), out))
        return out.raw

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_gens_free(self.h)
                self.h = None
        except Exception:
            pass


class BulletReduction:
    """operator-level BulletReductionProof::prove (nizk/bullet.rs:32-132): the device keeps a, b and the (implicitly folded) generators, the
    caller keeps the transcript.  round_LR() -> (L, R) compressed; fold(u, u_inv); finish() -> (a_hat, b_hat, G_hat compressed)."""

    def __init__(self, gens, a_vec, b_vec):
        self.ctx = gens.ctx
        h = _vp()
        self.ctx.check(lib.sp_ipa_begin(self.ctx.h, gens.h, a_vec.h, b_vec.h, C.byref(h)))
        self.h = h

    def round_LR(self, Q, H, blind_L, blind_R):
        L, R = C.create_string_buffer(32), C.create_string_buffer(32)
        self.ctx.check(lib.sp_ipa_round_LR(self.ctx.h, self.h, C.c_char_p(Q), C.c_char_p(H), _p(np.ascontiguousarray(blind_L, dtype=np.uint64)),
                                           _p(np.ascontiguousarray(blind_R, dtype=np.uint64)), L, R))
        return L.raw, R.raw

    def fold(self, u, u_inv):
        self.ctx.check(lib.sp_ipa_fold(self.ctx.h, self.h, _p(np.ascontiguousarray(u, dtype=np.uint64)), _p(np.ascontiguousarray(u_inv, dtype=np.uint64))))

    def finish(self):
        a, b, g = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64), C.create_string_buffer(32)
        self.ctx.check(lib.sp_ipa_finish(self.ctx.h, self.h, _p(a), _p(b), g))
        return a, b, g.raw

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_ipa_free(self.h)
                self.h = None
        except Exception:
            pass


class Points:
    """A caller-supplied `&[GroupElement]` (group.rs:8-9) resident on the device, for the variable-base MSM (bucket method, no tables)."""

    def __init__(self, compressed, ctx=None):
        """compressed: list of 32-byte ristretto255 encodings; raises if any does not decompress"""
        self.ctx = ctx or default_context()
        h = _vp()
        self.ctx.check(lib.sp_points_upload(self.ctx.h, C.c_char_p(b"".join(compressed)), _sz(len(compressed)), C.byref(h)))
        self.h = h

    @classmethod
    def derive(cls, n, label, ctx=None):
        """MultiCommitGens::new(n, label).G without window tables (commitments.rs:15-33)"""
        o = cls.__new__(cls)
        o.ctx = ctx or default_context()
        h = _vp()
        o.ctx.check(lib.sp_points_derive(o.ctx.h, C.c_char_p(label), _sz(len(label)), _sz(n), C.byref(h)))
        o.h = h
        return o

    def __len__(self):
        return int(lib.sp_points_len(self.h))

    def export(self, offset=0, n=None):
        n = len(self) - offset if n is None else n
        out = C.create_string_buffer(32 * n + 1)
        self.ctx.check(lib.sp_points_export(self.ctx.h, self.h, _sz(offset), _sz(n), out))
        return [out.raw[32 * i:32 * i + 32] for i in range(n)]

    def msm(self, scalars, offset=0):
        """GroupElement::vartime_multiscalar_mul(scalars, points[offset:offset+len(scalars)]).compress()  (group.rs:98-117)"""
        out = C.create_string_buffer(32)
        if isinstance(scalars, DensePolynomial):
            self.ctx.check(lib.sp_msm_var_resident(self.ctx.h, self.h, _sz(offset), scalars.h, out))
        else:
            s = _limbs(scalars)
            self.ctx.check(lib.sp_msm_var(self.ctx.h, self.h, _sz(offset), _p(s), _sz(len(s)), out))
        return out.raw

    def msm_sharded(self, scalars, offset=0):
        """this rank's slice of a vector split by index range over the ranks of a connected context -> encoding of the whole sum (on every rank)"""
        out = C.create_string_buffer(32)
        self.ctx.check(lib.sp_msm_var_sharded(self.ctx.h, self.h, _sz(offset), scalars.h, out))
        return out.raw

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_points_free(self.h)
                self.h = None
        except Exception:
            pass


def point_roundtrip(points, ctx=None):
    ctx = ctx or default_context()
    n = len(points)
    out = C.create_string_buffer(32 * n)
    ctx.check(lib.sp_point_roundtrip(ctx.h, C.c_char_p(b"".join(points)), _sz(n), out))
    return [out.raw[32 * i:32 * i + 32] for i in range(n)]


def point_decompress_check(points, ctx=None):
    ctx = ctx or default_context()
    n = len(points)
    ok = (C.c_int * n)()
    ctx.check(lib.sp_point_decompress_check(ctx.h, C.c_char_p(b"".join(points)), _sz(n), ok))
    return [bool(x) for x in ok]


class Transcript:
    """merlin::Transcript as the caller-owned object the reference's prove / verify mutate (`transcript: &mut Transcript`, lib.rs:339-347):
    the 203-byte STROBE-128 state, advanced in place by NIZK.prove / SNARK.prove / verify when passed instead of a label."""

    def __init__(self, label):
        self.state = C.create_string_buffer(203)
        lib.sp_transcript_new(C.c_char_p(label), _sz(len(label)), self.state)

    def append_message(self, label, msg):
        msg = bytes(msg)
        lib.sp_transcript_append_message(self.state, C.c_char_p(label), _sz(len(label)), C.c_char_p(msg), _sz(len(msg)))

    def challenge_bytes(self, label, n):
        out = C.create_string_buffer(n)
        lib.sp_transcript_challenge_bytes(self.state, C.c_char_p(label), _sz(len(label)), out, _sz(n))
        return out.raw


# ----------------------------------------------------------------------------- public API of lib.rs
class Assignment:
    """lib.rs:57-104: a vector of scalars given as canonical 32-byte strings"""

    def __init__(self, assignment):
        if isinstance(assignment, np.ndarray):
            self.limbs = _limbs(assignment).copy()
        else:
            self.limbs = np.zeros((len(assignment), 4), dtype=np.uint64)
            for i, b in enumerate(assignment):
                self.limbs[i] = scalar_from_bytes(b)  # raises R1CSError::InvalidScalar

    @staticmethod
    def new(assignment):
        return Assignment(assignment)

    def __len__(self):
        return len(self.limbs)


VarsAssignment = Assignment
InputsAssignment = Assignment


class Instance:
    """lib.rs:111-274"""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h
        nc, nv, ni = _sz(), _sz(), _sz()
        lib.sp_instance_dims(h, C.byref(nc), C.byref(nv), C.byref(ni))
        self.num_cons, self.num_vars, self.num_inputs = nc.value, nv.value, ni.value
        self._digest = None

    @staticmethod
    def new(num_cons, num_vars, num_inputs, A, B, Cm, ctx=None):
        """Instance::new (lib.rs:121-227); A/B/C: lists of (row, col, 32 canonical bytes)"""
        ctx = ctx or default_context()

        def split(t):
            rows = np.array([r for r, _, _ in t], dtype=np.uint64)
            cols = np.array([c for _, c, _ in t], dtype=np.uint64)
            vals = b"".join(bytes(v) for _, _, v in t)
            return rows, cols, vals, len(t)
        a, b, c = split(A), split(B), split(Cm)
        h = _vp()
        ctx.check(lib.sp_instance_create(ctx.h, _sz(num_cons), _sz(num_vars), _sz(num_inputs), _p(a[0]), _p(a[1]), C.c_char_p(a[2]), _sz(a[3]),
                                         _p(b[0]), _p(b[1]), C.c_char_p(b[2]), _sz(b[3]), _p(c[0]), _p(c[1]), C.c_char_p(c[2]), _sz(c[3]), C.byref(h)))
        return Instance(ctx, h)

    @staticmethod
    def produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=0, ctx=None):
        """Instance::produce_synthetic_r1cs (lib.rs:262-274) with a seeded generator instead of OsRng"""
        ctx = ctx or default_context()
        h = _vp()
        vars_out = np.zeros((num_vars, 4), dtype=np.uint64)
        inputs_out = np.zeros((num_inputs, 4), dtype=np.uint64)
        ctx.check(lib.sp_instance_synthetic(ctx.h, _sz(num_cons), _sz(num_vars), _sz(num_inputs), C.c_uint64(seed), C.byref(h), _p(vars_out), _p(inputs_out)))
        return Instance(ctx, h), Assignment(vars_out), Assignment(inputs_out)

    def bincode(self):
        out, n = C.POINTER(C.c_ubyte)(), _sz()
        lib.sp_instance_bincode(self.h, C.byref(out), C.byref(n))
        b = C.string_at(out, n.value)
        lib.sp_free(out)
        return b

    @property
    def digest(self):
        """R1CSShape::get_digest (r1cs.rs:154-158): the caller-supplied bytes, else the library's miniz-level-6 zlib stream of bincode(shape)"""
        if self._digest is None:
            out, n = C.POINTER(C.c_ubyte)(), _sz()
            lib.sp_instance_digest(self.h, C.byref(out), C.byref(n))
            self._digest = _take_bytes(out, n)
        return self._digest

    def set_digest(self, digest):
        """absorb these bytes instead (e.g. the digest another compressor produced); b"" returns to the library's own"""
        self._digest = bytes(digest) if digest else None
        lib.sp_instance_set_digest(self.h, C.c_char_p(bytes(digest)), _sz(len(digest)))

    def export(self, matrix):
        n = _sz()
        lib.sp_instance_nnz(self.h, C.c_int(matrix), C.byref(n))
        row = np.zeros(n.value, dtype=np.uint64)
        col = np.zeros(n.value, dtype=np.uint64)
        val = np.zeros((n.value, 4), dtype=np.uint64)
        lib.sp_instance_export(self.h, C.c_int(matrix), _p(row), _p(col), _p(val))
        return row, col, val

    def is_sat(self, vars, inputs):
        sat = C.c_int()
        self.ctx.check(lib.sp_instance_is_sat(self.ctx.h, self.h, _p(vars.limbs), _sz(len(vars)), _p(inputs.limbs), _sz(len(inputs)), C.byref(sat)))
        return bool(sat.value)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_instance_free(self.h)
                self.h = None
        except Exception:
            pass


def _take_bytes(out, n):
    b = C.string_at(out, n.value)
    lib.sp_free(out)
    return b


class NIZKGens:
    """lib.rs:468-486"""

    def __init__(self, num_cons, num_vars, num_inputs, ctx=None):
        self.ctx = ctx or default_context()
        h = _vp()
        self.ctx.check(lib.sp_nizk_gens_create(self.ctx.h, _sz(num_cons), _sz(num_vars), _sz(num_inputs), C.byref(h)))
        self.h = h

    @staticmethod
    def new(num_cons, num_vars, num_inputs, ctx=None):
        return NIZKGens(num_cons, num_vars, num_inputs, ctx)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_nizk_gens_free(self.h)
                self.h = None
        except Exception:
            pass


class NIZK:
    """lib.rs:488-591.  `bytes` = bincode::serialize(&NIZK)"""

    def __init__(self, data):
        self.bytes = data

    @staticmethod
    def prove(inst, vars, inputs, gens, transcript_label, seed=None):
        """NIZK::prove(&inst, vars, &inputs, &gens, &mut Transcript::new(transcript_label)).  `seed`: RandomTape seed scalar limbs; None (the
        default) draws it from the OS like the reference (random.rs:13-15).  Pass tape_seed(k) only for reproducible test / bench runs."""
        ctx = inst.ctx
        seed = random_tape_seed() if seed is None else np.ascontiguousarray(seed, dtype=np.uint64)
        out, n = C.POINTER(C.c_ubyte)(), _sz()
        if isinstance(transcript_label, Transcript):   # the caller's `&mut Transcript`: advanced in place
            v = vars.to_numpy() if isinstance(vars, DensePolynomial) else vars.limbs
            ctx.check(lib.sp_nizk_prove_t(ctx.h, inst.h, _p(v), _sz(len(v)), _p(inputs.limbs), _sz(len(inputs)), gens.h, transcript_label.state, _p(seed),
                                          C.byref(out), C.byref(n)))
        elif isinstance(vars, DensePolynomial):
            ctx.check(lib.sp_nizk_prove_resident(ctx.h, inst.h, vars.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, C.c_char_p(transcript_label),
                                                 _sz(len(transcript_label)), _p(seed), C.byref(out), C.byref(n)))
        else:
            ctx.check(lib.sp_nizk_prove(ctx.h, inst.h, _p(vars.limbs), _sz(len(vars)), _p(inputs.limbs), _sz(len(inputs)), gens.h,
                                        C.c_char_p(transcript_label), _sz(len(transcript_label)), _p(seed), C.byref(out), C.byref(n)))
        return NIZK(_take_bytes(out, n))

    def verify(self, inst, inputs, transcript_label, gens):
        """NIZK::verify(&self, &inst, &inputs, &mut Transcript::new(transcript_label), &gens): returns None, raises ProofVerifyError"""
        ctx = inst.ctx
        if isinstance(transcript_label, Transcript):
            ctx.check(lib.sp_nizk_verify_t(ctx.h, inst.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, transcript_label.state, C.c_char_p(bytes(self.bytes)), _sz(len(self.bytes))))
            return
        ctx.check(lib.sp_nizk_verify(ctx.h, inst.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, C.c_char_p(transcript_label), _sz(len(transcript_label)),
                                     C.c_char_p(bytes(self.bytes)), _sz(len(self.bytes))))


class SNARKGens:
    """lib.rs:277-309"""

    def __init__(self, num_cons, num_vars, num_inputs, num_nz_entries, ctx=None):
        self.ctx = ctx or default_context()
        h = _vp()
        self.ctx.check(lib.sp_snark_gens_create(self.ctx.h, _sz(num_cons), _sz(num_vars), _sz(num_inputs), _sz(num_nz_entries), C.byref(h)))
        self.h = h

    @staticmethod
    def new(num_cons, num_vars, num_inputs, num_nz_entries, ctx=None):
        return SNARKGens(num_cons, num_vars, num_inputs, num_nz_entries, ctx)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_snark_gens_free(self.h)
                self.h = None
        except Exception:
            pass


class ComputationCommitment:
    """(ComputationCommitment, ComputationDecommitment) pair of SNARK::encode (lib.rs:44-55): the decommitment stays on the device"""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    @classmethod
    def from_bytes(cls, data, ctx=None):
        """a verifier's view: bincode(ComputationCommitment) only (cannot be used to prove)"""
        ctx = ctx or default_context()
        h = _vp()
        ctx.check(lib.sp_snark_commitment_load(ctx.h, C.c_char_p(bytes(data)), _sz(len(data)), C.byref(h)))
        return cls(ctx, h)

    def commitment_bytes(self):
        out, n = C.POINTER(C.c_ubyte)(), _sz()
        lib.sp_snark_commitment_bytes(self.h, C.byref(out), C.byref(n))
        return _take_bytes(out, n)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.sp_snark_encoding_free(self.h)
                self.h = None
        except Exception:
            pass


class SNARK:
    """lib.rs:311-465.  `bytes` = bincode::serialize(&SNARK)"""

    def __init__(self, data):
        self.bytes = data

    @staticmethod
    def encode(inst, gens):
        """SNARK::encode (lib.rs:325-336) -> ComputationCommitment (holding the decommitment too)"""
        h = _vp()
        inst.ctx.check(lib.sp_snark_encode(inst.ctx.h, inst.h, gens.h, C.byref(h)))
        return ComputationCommitment(inst.ctx, h)

    @staticmethod
    def prove(inst, comm, vars, inputs, gens, transcript_label, seed=None):
        """SNARK::prove(&inst, &comm, &decomm, vars, &inputs, &gens, &mut Transcript::new(transcript_label)).  `seed` as in NIZK.prove:
        None = fresh OS randomness (zero-knowledge); tape_seed(k) = deterministic, for tests and benchmarks only."""
        ctx = inst.ctx
        seed = random_tape_seed() if seed is None else np.ascontiguousarray(seed, dtype=np.uint64)
        out, n = C.POINTER(C.c_ubyte)(), _sz()
        if isinstance(transcript_label, Transcript):   # the caller's `&mut Transcript`: advanced in place
            v = vars.to_numpy() if isinstance(vars, DensePolynomial) else vars.limbs
            ctx.check(lib.sp_snark_prove_t(ctx.h, inst.h, comm.h, _p(v), _sz(len(v)), _p(inputs.limbs), _sz(len(inputs)), gens.h, transcript_label.state, _p(seed),
                                           C.byref(out), C.byref(n)))
        elif isinstance(vars, DensePolynomial):
            ctx.check(lib.sp_snark_prove_resident(ctx.h, inst.h, comm.h, vars.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, C.c_char_p(transcript_label),
                                                  _sz(len(transcript_label)), _p(seed), C.byref(out), C.byref(n)))
        else:
            ctx.check(lib.sp_snark_prove(ctx.h, inst.h, comm.h, _p(vars.limbs), _sz(len(vars)), _p(inputs.limbs), _sz(len(inputs)), gens.h,
                                         C.c_char_p(transcript_label), _sz(len(transcript_label)), _p(seed), C.byref(out), C.byref(n)))
        return SNARK(_take_bytes(out, n))

    def verify(self, comm, inputs, transcript_label, gens):
        """SNARK::verify(&self, &comm, &inputs, &mut Transcript::new(transcript_label), &gens): returns None, raises ProofVerifyError"""
        ctx = comm.ctx
        if isinstance(transcript_label, Transcript):
            ctx.check(lib.sp_snark_verify_t(ctx.h, comm.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, transcript_label.state, C.c_char_p(bytes(self.bytes)), _sz(len(self.bytes))))
            return
        ctx.check(lib.sp_snark_verify(ctx.h, comm.h, _p(inputs.limbs), _sz(len(inputs)), gens.h, C.c_char_p(transcript_label), _sz(len(transcript_label)),
                                      C.c_char_p(bytes(self.bytes)), _sz(len(self.bytes))))


# ----------------------------------------------------------------------------- measurement helpers (bench.py)
def io_bytes():
    a, b = C.c_ulonglong(), C.c_ulonglong()
    lib.sp_io_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def prof_enable(on=True):
    lib.sp_prof_enable(C.c_int(1 if on else 0))


def prof_report():
    """{kernel family: {"launches", "ms", "bytes"}} measured with CUDA events on the launching stream"""
    buf = C.create_string_buffer(1 << 16)
    lib.sp_prof_report(buf, _sz(1 << 16))
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, n, ms, by, bb, bms = item.split(":")
            out[name] = {"launches": int(n), "ms": float(ms), "bytes": float(by), "largest_bytes": float(bb), "largest_ms": float(bms)}
    return out


def timer_start(ctx=None):
    ctx = ctx or default_context()
    ctx.check(lib.sp_timer_start(ctx.h))


def timer_stop_ms(ctx=None):
    ctx = ctx or default_context()
    ms = C.c_float()
    ctx.check(lib.sp_timer_stop_ms(ctx.h, C.byref(ms)))
    return float(ms.value)
