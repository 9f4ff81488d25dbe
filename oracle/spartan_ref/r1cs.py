"""
ORACLE — TEST INFRASTRUCTURE ONLY (see core.py header).

Restatement of src/r1cs.rs (R1CSShape) and the public API of src/lib.rs (Instance, Assignment,
NIZKGens / NIZK, SNARKGens / SNARK).  The three OsRng sites of the reference (r1cs.rs:169,183-185;
random.rs:13-15) are replaced by the SHAKE256 generator of SURVEY.md §8d so that instances, witnesses
and the prover's RandomTape seed are explicit, reproducible inputs.
"""
import ctypes as C
import zlib
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import core as oc
from . import protocol as pr
from .core import Q


class R1CSError(Exception):
    pass


class SparseMat:
    """SparseMatPolynomial (sparse_mlpoly.rs:31-37) as COO arrays: row, col (uint64) and val (Montgomery)"""

    def __init__(self, num_vars_x, num_vars_y, row, col, val):
        self.num_vars_x, self.num_vars_y = num_vars_x, num_vars_y
        self.row = np.ascontiguousarray(row, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint64)
        self.val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 4)

    def nnz(self):
        return len(self.row)

    def get_num_nz_entries(self):
        return oc.next_pow2(self.nnz())

    def multiply_vec(self, num_rows, num_cols, z):
        """sparse_mlpoly.rs:454-464"""
        assert len(z) == num_cols
        out = oc.zeros(num_rows)
        oc.lib.sparse_multiply_vec(oc._ptr(out), C.c_size_t(num_rows), oc._ptr(self.row), oc._ptr(self.col), oc._ptr(self.val),
                                   C.c_size_t(self.nnz()), oc._ptr(np.ascontiguousarray(z)))
        return out

    def compute_eval_table_sparse(self, rx, num_rows, num_cols):
        """sparse_mlpoly.rs:466-481"""
        assert len(rx) == num_rows
        out = oc.zeros(num_cols)
        oc.lib.sparse_eval_table(oc._ptr(out), C.c_size_t(num_cols), oc._ptr(self.row), oc._ptr(self.col), oc._ptr(self.val),
                                 C.c_size_t(self.nnz()), oc._ptr(np.ascontiguousarray(rx)))
        return out

    def evaluate_with_tables(self, trx, try_):
        """sparse_mlpoly.rs:426-438"""
        out = np.zeros(4, dtype=np.uint64)
        oc.lib.sparse_evaluate(oc._ptr(out), oc._ptr(self.row), oc._ptr(self.col), oc._ptr(self.val), C.c_size_t(self.nnz()),
                               oc._ptr(np.ascontiguousarray(trx)), oc._ptr(np.ascontiguousarray(try_)))
        return oc.from_mont_bytes(out.tobytes())

    def bincode(self):
        """SparseMatPolynomial{num_vars_x,num_vars_y,M:Vec<{row,col,val}>}"""
        n = self.nnz()
        rec = np.zeros((n, 6), dtype=np.uint64)
        rec[:, 0] = self.row
        rec[:, 1] = self.col
        rec[:, 2:] = self.val
        return (self.num_vars_x.to_bytes(8, "little") + self.num_vars_y.to_bytes(8, "little") + n.to_bytes(8, "little") + rec.tobytes())


class R1CSShape:
    """r1cs.rs:19-26"""

    def __init__(self, num_cons, num_vars, num_inputs, A, B, Cm):
        assert oc.next_pow2(num_cons) == num_cons and oc.next_pow2(num_vars) == num_vars and num_inputs < num_vars
        self.num_cons, self.num_vars, self.num_inputs = num_cons, num_vars, num_inputs
        self.A, self.B, self.C = A, B, Cm

    @staticmethod
    def from_triples(num_cons, num_vars, num_inputs, A, B, Cm):
        """R1CSShape::new (r1cs.rs:88-140); triples are (row, col, canonical int)"""
        nx, ny = oc.log_2(num_cons), oc.log_2(2 * num_vars)

        def mk(t):
            return SparseMat(nx, ny, [r for r, _, _ in t], [c for _, c, _ in t], oc.to_arr([v for _, _, v in t]) if t else oc.zeros(0))
        return R1CSShape(num_cons, num_vars, num_inputs, mk(A), mk(B), mk(Cm))

    def get_digest(self):
        """r1cs.rs:154-158: zlib(bincode(shape)).  flate2's default backend is miniz_oxide, whose byte stream
        is not guaranteed equal to system zlib's — the digest is treated as an opaque input."""
        raw = (self.num_cons.to_bytes(8, "little") + self.num_vars.to_bytes(8, "little") + self.num_inputs.to_bytes(8, "little")
               + self.A.bincode() + self.B.bincode() + self.C.bincode())
        return zlib.compress(raw, 6)

    def multiply_vec(self, num_rows, num_cols, z):
        """r1cs.rs:268-282"""
        assert num_rows == self.num_cons and len(z) == num_cols and num_cols > self.num_vars
        return (self.A.multiply_vec(num_rows, num_cols, z), self.B.multiply_vec(num_rows, num_cols, z), self.C.multiply_vec(num_rows, num_cols, z))

    def compute_eval_table_sparse(self, num_rows, num_cols, evals):
        """r1cs.rs:284-298"""
        assert num_rows == self.num_cons and num_cols > self.num_vars
        return tuple(M.compute_eval_table_sparse(evals, num_rows, num_cols) for M in (self.A, self.B, self.C))

    def evaluate(self, rx, ry):
        """r1cs.rs:300-303 -> SparseMatPolynomial::multi_evaluate (sparse_mlpoly.rs:440-452)"""
        trx, try_ = oc.eq_evals(rx), oc.eq_evals(ry)
        return tuple(M.evaluate_with_tables(trx, try_) for M in (self.A, self.B, self.C))

    def is_sat(self, vars_arr, input_ints):
        """r1cs.rs:240-266"""
        z = np.concatenate([vars_arr, oc.to_arr([1] + list(input_ints))])
        nc = self.num_vars + self.num_inputs + 1
        Az = oc.to_ints(self.A.multiply_vec(self.num_cons, nc, z))
        Bz = oc.to_ints(self.B.multiply_vec(self.num_cons, nc, z))
        Cz = oc.to_ints(self.C.multiply_vec(self.num_cons, nc, z))
        return all(a * b % Q == c for a, b, c in zip(Az, Bz, Cz))


def produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=0):
    """R1CSShape::produce_synthetic_r1cs (r1cs.rs:160-238) with Z drawn from prg_scalars("Z").
    Returns (shape, vars array, inputs list of ints)."""
    assert 1 << oc.log_2(num_cons) == num_cons and 1 << oc.log_2(num_vars) == num_vars and num_inputs < num_vars
    size_z = num_vars + num_inputs + 1
    Zarr = oc.prg_scalars("Z", size_z, seed)
    Zarr[num_vars] = oc.to_arr([1])[0]
    i = np.arange(num_cons, dtype=np.uint64)
    A_idx = i % size_z
    B_idx = (i + 2) % size_z
    C_idx = (i + 3) % size_z
    one = oc.to_arr([1])[0]
    ones = np.tile(one, (num_cons, 1))
    AB = oc.hadamard(Zarr[A_idx.astype(np.int64)], Zarr[B_idx.astype(np.int64)])
    Cv = Zarr[C_idx.astype(np.int64)].copy()
    zero_mask = ~Cv.any(axis=1)
    Cinv = Cv.copy()
    Cinv[zero_mask] = one
    oc.lib.fq_batch_invert(oc._ptr(Cinv), C.c_size_t(num_cons), None)
    C_val = oc.hadamard(AB, Cinv)
    C_col = C_idx.copy()
    C_val[zero_mask] = AB[zero_mask]          # r1cs.rs:208-209
    C_col[zero_mask] = num_vars
    nx, ny = oc.log_2(num_cons), oc.log_2(2 * num_vars)
    inst = R1CSShape(num_cons, num_vars, num_inputs, SparseMat(nx, ny, i, A_idx, ones), SparseMat(nx, ny, i, B_idx, ones.copy()),
                     SparseMat(nx, ny, i, C_col, C_val))
    return inst, Zarr[:num_vars].copy(), oc.to_ints(Zarr[num_vars + 1:])


class Instance:
    """lib.rs:111-274"""

    def __init__(self, inst, digest=None):
        self.inst = inst
        self.digest = digest if digest is not None else inst.get_digest()

    @staticmethod
    def new(num_cons, num_vars, num_inputs, A, B, Cm):
        """lib.rs:121-227; A/B/C are lists of (row, col, 32 canonical bytes)"""
        num_vars_padded = oc.next_pow2(max(num_vars, num_inputs + 1))
        num_cons_padded = num_cons
        if num_cons_padded in (0, 1):
            num_cons_padded = 2
        if oc.next_pow2(num_cons) != num_cons:
            num_cons_padded = oc.next_pow2(num_cons)

        def conv(tups):
            mat = []
            for row, col, vb in tups:
                if row >= num_cons:
                    raise R1CSError("InvalidIndex")
                if col >= num_vars + 1 + num_inputs:
                    raise R1CSError("InvalidIndex")
                v = int.from_bytes(vb, "little")
                if v >= Q:
                    raise R1CSError("InvalidScalar")
                mat.append((row, col + num_vars_padded - num_vars if col >= num_vars else col, v))
            if num_cons in (0, 1):
                for i in range(len(tups), num_cons_padded):
                    mat.append((i, num_vars, 0))
            return mat
        return Instance(R1CSShape.from_triples(num_cons_padded, num_vars_padded, num_inputs, conv(A), conv(B), conv(Cm)))

    @staticmethod
    def produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed=0):
        inst, vars_arr, inputs = produce_synthetic_r1cs(num_cons, num_vars, num_inputs, seed)
        return Instance(inst), vars_arr, inputs

    def is_sat(self, vars_arr, inputs):
        if len(vars_arr) > self.inst.num_vars or len(inputs) != self.inst.num_inputs:
            raise R1CSError("InvalidNumberOfInputs")
        return self.inst.is_sat(pad_vars(vars_arr, self.inst.num_vars), inputs)


def assignment_from_bytes(items):
    """Assignment::new (lib.rs:64-88): canonical 32-byte scalars -> Montgomery array"""
    out = oc.zeros(len(items))
    ok = oc.lib.fq_from_bytes_batch(oc._ptr(out), C.c_char_p(b"".join(items)), C.c_size_t(len(items)))
    if not ok:
        raise R1CSError("InvalidScalar")
    return out


def pad_vars(vars_arr, n):
    """Assignment::pad (lib.rs:91-104)"""
    if len(vars_arr) < n:
        return np.concatenate([vars_arr, oc.zeros(n - len(vars_arr))])
    return vars_arr


def _padded_num_vars(num_vars, num_inputs):
    return oc.next_pow2(max(num_vars, num_inputs + 1))


class NIZKGens:
    """lib.rs:468-486"""

    def __init__(self, num_cons, num_vars, num_inputs):
        self.gens_r1cs_sat = pr.R1CSGens(b"gens_r1cs_sat", num_cons, _padded_num_vars(num_vars, num_inputs))


def tape_seed(seed=0):
    """the OsRng scalar of RandomTape::new (random.rs:13-15), made explicit: prg tag "tape" """
    return oc.arr_get(oc.prg_scalars("tape", 1, seed), 0)


@dataclass
class NIZK:
    r1cs_sat_proof: pr.R1CSProof
    r: Tuple[List[int], List[int]]

    @staticmethod
    def prove(inst, vars_arr, inputs, gens, T, seed_scalar):
        """lib.rs:501-546"""
        tape = oc.RandomTape(b"proof", seed_scalar)
        T.append_protocol_name(b"Spartan NIZK proof")
        T.append_message(b"R1CSShapeDigest", inst.digest)
        padded = pad_vars(vars_arr, inst.inst.num_vars)
        proof, rx, ry = pr.R1CSProof.prove(inst.inst, padded, inputs, gens.gens_r1cs_sat, T, tape)
        return NIZK(proof, (rx, ry))

    def verify(self, inst, inputs, T, gens):
        """lib.rs:549-591"""
        T.append_protocol_name(b"Spartan NIZK proof")
        T.append_message(b"R1CSShapeDigest", inst.digest)
        claimed_rx, claimed_ry = self.r
        inst_evals = inst.inst.evaluate(claimed_rx, claimed_ry)
        assert len(inputs) == inst.inst.num_inputs
        rx, ry = self.r1cs_sat_proof.verify(inst.inst.num_vars, inst.inst.num_cons, inputs, inst_evals, T, gens.gens_r1cs_sat)
        assert rx == claimed_rx and ry == claimed_ry

    def ser(self):
        return pr.ser(self)
