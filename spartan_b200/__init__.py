"""spartan_b200 — B200-native prover for the Spartan zkSNARK (drop-in for the prover hot path of microsoft/Spartan).

The product is the C-ABI library `libspartan_b200.so` (include/spartan_b200.h): hand-written sm_100a kernels plus a C++ host
prover.  This package is a thin ctypes mirror of the reference's public API (src/lib.rs): Instance, Assignment, NIZKGens, NIZK,
SNARKGens, SNARK — same names, argument meaning and error behaviour.  There is no CPU fallback: without the CUDA library or
without a GPU every entry point raises.
"""
from .api import (  # noqa: F401
    Assignment, Context, DensePolynomial, InputsAssignment, Instance, MultiCommitGens, NIZK, NIZKGens, ProofVerifyError, R1CSError, SNARK, SNARKGens,
    SpartanB200Error, Transcript, VarsAssignment, default_context, kernel_launches, lib, random_tape_seed, scalar_from_bytes, scalar_to_bytes, tape_seed, prg_scalars,
)
