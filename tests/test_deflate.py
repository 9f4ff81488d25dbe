"""spartan_b200/csrc/deflate.cpp (the zlib stream NIZK::prove absorbs as R1CSShapeDigest, src/r1cs.rs:154-158 + src/lib.rs:514) against the C
miniz that PyTorch bundles (libtorch_cpu.so exports mz_compress2): bit-identical level-6 streams on inputs that exercise every block type —
empty, tiny (static block), text, runs, incompressible bytes (raw-block fallback), multi-block inputs, and real bincode(R1CSShape) bytes.
miniz_oxide (what flate2 uses in the reference) is the Rust port of this compressor; its own bytes cannot be produced here (no Rust toolchain)."""
import ctypes as C
import glob
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _miniz():
    import torch
    for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cpu.so")):
        lib = C.CDLL(p)
        if hasattr(lib, "mz_compress2"):
            lib.mz_compress2.argtypes = [C.c_char_p, C.POINTER(C.c_ulong), C.c_char_p, C.c_ulong, C.c_int]
            return lib
    return None


def _mine():
    path = os.path.join(ROOT, "spartan_b200", "libsp_hosttest_fast.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build_hosttest()
    lib = C.CDLL(path)
    lib.spt_zlib6.restype = C.c_size_t
    lib.spt_zlib6.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    return lib


def mine_compress(lib, data):
    out = C.create_string_buffer(len(data) + len(data) // 8 + 256)
    n = lib.spt_zlib6(data, len(data), out)
    return out.raw[:n]


def miniz_compress(lib, data):
    cap = C.c_ulong(len(data) + len(data) // 8 + 256)
    out = C.create_string_buffer(cap.value)
    assert lib.mz_compress2(out, C.byref(cap), data, len(data), 6) == 0
    return out.raw[:cap.value]


def cases():
    rng = np.random.default_rng(42)
    yield b""
    yield b"a"
    yield b"abc" * 5
    yield b"hello world, hello world, hello world!" * 3
    yield bytes(1000)
    yield bytes(70000)
    yield b"\xff" * 300 + b"\x00" * 300
    yield rng.bytes(10)
    yield rng.bytes(47)
    yield rng.bytes(48)
    yield rng.bytes(5000)          # incompressible: stored block
    yield rng.bytes(100000)        # several stored blocks
    words = [b"spartan", b"sumcheck", b"ristretto", b"commit", b"transcript", b"r1cs", b" ", b"\n", b"0123456789"]
    yield b"".join(words[i] for i in rng.integers(0, len(words), 60000))   # text-like, several dynamic blocks
    yield bytes(rng.integers(0, 4, 200000, dtype=np.uint8))                 # low-entropy bytes
    yield bytes((np.arange(300000) % 251).astype(np.uint8))                 # long-distance periodic matches
    mix = bytearray()
    for _ in range(200):
        mix += rng.bytes(int(rng.integers(1, 400))) + bytes(int(rng.integers(1, 600))) + b"abcdefgh" * int(rng.integers(1, 40))
    yield bytes(mix)


def test_bit_identical_to_miniz_level6():
    mz, me = _miniz(), _mine()
    if mz is None:
        pytest.skip("no miniz in this torch build")
    for k, data in enumerate(cases()):
        a, b = mine_compress(me, data), miniz_compress(mz, data)
        assert zlib.decompress(a) == data, "case %d does not round-trip" % k
        assert a == b, "case %d (%d bytes): %d vs %d compressed bytes, first difference at %s" % (
            k, len(data), len(a), len(b), next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), "length"))


def test_r1cs_shape_digest_bytes():
    """the real input: bincode(R1CSShape) of synthetic instances (indices + 32-byte Montgomery scalars), as Instance::new / get_digest build it"""
    mz, me = _miniz(), _mine()
    if mz is None:
        pytest.skip("no miniz in this torch build")
    from oracle.spartan_ref import r1cs
    for (nc, nv, ni, seed) in [(16, 16, 3, 0), (1024, 1024, 10, 1), (4096, 2048, 10, 2), (1 << 15, 1 << 15, 10, 3)]:
        inst, _, _ = r1cs.Instance.produce_synthetic_r1cs(nc, nv, ni, seed)
        sh = inst.inst
        data = sh.num_cons.to_bytes(8, "little") + sh.num_vars.to_bytes(8, "little") + sh.num_inputs.to_bytes(8, "little") + sh.A.bincode() + sh.B.bincode() + sh.C.bincode()
        a, b = mine_compress(me, data), miniz_compress(mz, data)
        assert zlib.decompress(a) == data and a == b, (nc, len(data), len(a), len(b))
