"""Merlin v1.0 / STROBE-128 / Keccak restatement (oracle/csrc/merlin.c) — SURVEY.md §8c item 4."""
import ctypes as C
import hashlib

from oracle.spartan_ref import core as oc


def test_merlin_conformance_vector():
    # the vector of merlin's own transcript test ("test protocol" / "some label" / "some data" / "challenge")
    t = oc.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_keccak_via_shake256_matches_hashlib():
    for n_in, n_out in [(0, 32), (1, 64), (135, 136), (136, 137), (137, 500), (1000, 3000)]:
        msg = bytes((i * 7 + 3) & 0xFF for i in range(n_in))
        out = C.create_string_buffer(n_out)
        oc.lib.shake256(out, C.c_size_t(n_out), C.c_char_p(msg), C.c_size_t(n_in))
        assert out.raw == hashlib.shake_256(msg).digest(n_out)


def test_transcript_is_order_and_length_sensitive():
    def run(msgs):
        t = oc.Transcript(b"x")
        for l, m in msgs:
            t.append_message(l, m)
        return t.challenge_bytes(b"c", 64)
    a = run([(b"a", b"12"), (b"b", b"3")])
    assert a != run([(b"a", b"1"), (b"b", b"23")])
    assert a != run([(b"b", b"3"), (b"a", b"12")])
    assert a == run([(b"a", b"12"), (b"b", b"3")])
    # long messages cross the 166-byte STROBE rate several times
    assert run([(b"l", bytes(1000))]) != run([(b"l", bytes(999))])


def test_challenge_scalar_is_wide_reduction():
    t1, t2 = oc.Transcript(b"y"), oc.Transcript(b"y")
    raw = t1.challenge_bytes(b"ch", 64)
    assert t2.challenge_scalar(b"ch") == int.from_bytes(raw, "little") % oc.Q


def test_vector_framing_matches_manual():
    # transcript.rs:49-57
    vals = [5, 7, oc.Q - 1]
    t1, t2 = oc.Transcript(b"z"), oc.Transcript(b"z")
    t1.append_scalars(b"v", vals)
    t2.append_message(b"v", b"begin_append_vector")
    for v in vals:
        t2.append_message(b"v", v.to_bytes(32, "little"))
    t2.append_message(b"v", b"end_append_vector")
    assert t1.challenge_bytes(b"c", 32) == t2.challenge_bytes(b"c", 32)
    t3 = oc.Transcript(b"z")
    t3.append_scalars(b"v", oc.to_arr(vals))
    t4 = oc.Transcript(b"z")
    t4.append_scalars(b"v", vals)
    assert t3.challenge_bytes(b"c", 32) == t4.challenge_bytes(b"c", 32)
