"""F_q restatement (oracle/csrc/fq.c) against every known-answer test the reference holds:
/root/reference/src/scalar/ristretto255.rs:777-1201."""
import ctypes as C

import numpy as np

from oracle.spartan_ref import core as oc

lib = oc.lib
Q = oc.Q
MODULUS = [0x5812631A5CF5D3ED, 0x14DEF9DEA2F79CD6, 0x0, 0x1000000000000000]
R = [0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF]
R2 = [0xA40611E3449C0F01, 0xD00E1BA768859347, 0xCEEC73D217F5BE65, 0x0399411B7C309A3D]
R3 = [0x2A9E49687B83A2DB, 0x278324E6AEF7F3EC, 0x8065DC6C04EC5B65, 0x0E530B773599CEC7]
LARGEST = [0x5812631A5CF5D3EC, 0x14DEF9DEA2F79CD6, 0x0, 0x1000000000000000]
INV = 0xD2B51DA312547E1B


def L(x):
    return np.array(x, dtype=np.uint64)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def op2(name, a, b):
    r = np.zeros(4, dtype=np.uint64)
    getattr(lib, name)(P(r), P(L(a)), P(L(b)))
    return [int(x) for x in r]


def op1(name, a):
    r = np.zeros(4, dtype=np.uint64)
    getattr(lib, name)(P(r), P(L(a)))
    return [int(x) for x in r]


def to_bytes(a):
    out = C.create_string_buffer(32)
    lib.fq_to_bytes(out, P(L(a)))
    return out.raw


def from_bytes(b):
    r = np.zeros(4, dtype=np.uint64)
    ok = lib.fq_from_bytes(P(r), C.c_char_p(bytes(b)))
    return ok, [int(x) for x in r]


def from_wide(b):
    r = np.zeros(4, dtype=np.uint64)
    lib.fq_from_bytes_wide(P(r), C.c_char_p(bytes(b)))
    return [int(x) for x in r]


def from_raw(v):
    return op2("fq_mul", v, R2)


def limbs_int(a):
    return sum(int(x) << (64 * i) for i, x in enumerate(a))


def test_inv_constant():  # ristretto255.rs:777-789
    inv = 1
    for _ in range(63):
        inv = (inv * inv) % 2**64
        inv = (inv * MODULUS[0]) % 2**64
    assert (-inv) % 2**64 == INV


def test_constants_consistent():
    assert limbs_int(MODULUS) == Q
    assert limbs_int(R) == 2**256 % Q and limbs_int(R2) == 2**512 % Q and limbs_int(R3) == 2**768 % Q


def test_to_bytes():  # ristretto255.rs:819-849 (also the Debug strings :793-806)
    assert to_bytes([0, 0, 0, 0]) == bytes(32)
    assert to_bytes(R) == bytes([1] + [0] * 31)
    r2b = bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15])
    assert to_bytes(R2) == r2b
    assert to_bytes(R2)[::-1].hex() == "0ffffffffffffffffffffffffffffffec6ef5bf4737dcf70d6ec31748d98951d"
    m1 = bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16])
    assert to_bytes(op1("fq_neg", R)) == m1


def test_from_bytes():  # ristretto255.rs:852-932
    assert from_bytes(bytes(32)) == (1, [0, 0, 0, 0])
    assert from_bytes(bytes([1] + [0] * 31)) == (1, R)
    r2b = bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15])
    assert from_bytes(r2b) == (1, R2)
    m1 = bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16])
    assert from_bytes(m1)[0] == 1
    bad = [
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [2, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 58, 51, 72, 125, 157, 41, 83, 167, 237, 115],
        [1, 0, 0, 0, 255, 255, 255, 255, 254, 91, 254, 255, 2, 164, 189, 83, 5, 216, 161, 9, 8, 216, 57, 51, 72, 125, 157, 41, 83, 167, 237, 116],
    ]
    for b in bad:
        assert from_bytes(bytes(b))[0] == 0
    # the modulus itself and modulus+1 are not canonical
    assert from_bytes(Q.to_bytes(32, "little"))[0] == 0
    assert from_bytes((Q + 1).to_bytes(32, "little"))[0] == 0
    assert from_bytes((Q - 1).to_bytes(32, "little"))[0] == 1


def test_from_u512():  # ristretto255.rs:935-1005
    w = lambda limbs: b"".join(int(x).to_bytes(8, "little") for x in limbs)
    assert from_wide(w(MODULUS + [0, 0, 0, 0])) == [0, 0, 0, 0]
    assert from_wide(w([1, 0, 0, 0, 0, 0, 0, 0])) == R
    assert from_wide(w([0, 0, 0, 0, 1, 0, 0, 0])) == R2
    mx = 0xFFFFFFFFFFFFFFFF
    assert from_wide(w([mx] * 8)) == op2("fq_sub", R3, R)
    r2b = bytes([29, 149, 152, 141, 116, 49, 236, 214, 112, 207, 125, 115, 244, 91, 239, 198, 254] + [255] * 14 + [15])
    assert from_wide(r2b + bytes(32)) == R2
    m1 = bytes([236, 211, 245, 92, 26, 99, 18, 88, 214, 156, 247, 162, 222, 249, 222, 20] + [0] * 15 + [16])
    assert from_wide(m1 + bytes(32)) == op1("fq_neg", R)
    assert from_wide(bytes([0xFF] * 64)) == from_raw([0xA40611E3449C0F00, 0xD00E1BA768859347, 0xCEEC73D217F5BE65, 0x0399411B7C309A3D])


def test_zero_add_neg_sub():  # ristretto255.rs:1008-1069
    Z = [0, 0, 0, 0]
    assert op1("fq_neg", Z) == Z and op2("fq_add", Z, Z) == Z and op2("fq_sub", Z, Z) == Z and op2("fq_mul", Z, Z) == Z
    assert op2("fq_add", LARGEST, LARGEST) == [0x5812631A5CF5D3EB, 0x14DEF9DEA2F79CD6, 0, 0x1000000000000000]
    assert op2("fq_add", LARGEST, [1, 0, 0, 0]) == Z
    assert op1("fq_neg", LARGEST) == [1, 0, 0, 0]
    assert op1("fq_neg", [1, 0, 0, 0]) == LARGEST
    assert op2("fq_sub", LARGEST, LARGEST) == Z
    assert op2("fq_sub", Z, LARGEST) == op2("fq_sub", MODULUS, LARGEST)


def _double_and_add(cur):
    tmp2 = [0, 0, 0, 0]
    for byte in to_bytes(cur)[::-1]:
        for i in range(7, -1, -1):
            tmp2 = op2("fq_add", tmp2, tmp2)
            if (byte >> i) & 1:
                tmp2 = op2("fq_add", tmp2, cur)
    return tmp2


def test_multiplication_and_squaring():  # ristretto255.rs:1072-1127
    cur = LARGEST
    for _ in range(100):
        assert op2("fq_mul", cur, cur) == _double_and_add(cur)
        assert op1("fq_square", cur) == _double_and_add(cur)
        cur = op2("fq_add", cur, LARGEST)


def test_inversion():  # ristretto255.rs:1130-1172
    r = np.zeros(4, dtype=np.uint64)
    assert lib.fq_invert(P(r), P(L([0, 0, 0, 0]))) == 0
    assert op1("fq_invert", R) == R
    m1 = op1("fq_neg", R)
    assert op1("fq_invert", m1) == m1
    tmp = R2
    for _ in range(100):
        assert op2("fq_mul", op1("fq_invert", tmp), tmp) == R
        tmp = op2("fq_add", tmp, R2)
    q_minus_2 = L([0x5812631A5CF5D3EB, 0x14DEF9DEA2F79CD6, 0, 0x1000000000000000])
    r1 = R
    for _ in range(100):
        a = op1("fq_invert", r1)
        b = np.zeros(4, dtype=np.uint64)
        lib.fq_pow_vartime(P(b), P(L(r1)), P(q_minus_2))
        assert a == [int(x) for x in b]
        r1 = op2("fq_add", a, R)


def test_from_raw_and_double():  # ristretto255.rs:1175-1201
    assert from_raw([0xD6EC31748D98951C, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF]) == from_raw([0xFFFFFFFFFFFFFFFF] * 4)
    assert from_raw(MODULUS) == [0, 0, 0, 0]
    assert from_raw([1, 0, 0, 0]) == R
    a = from_raw([0x1FFF3231233FFFFD, 0x4884B7FA00034802, 0x998C4FEFECBC4FF3, 0x1824B159ACC50562])
    assert op2("fq_add", a, a) == op2("fq_mul", a, from_raw([2, 0, 0, 0]))


def test_c_matches_python_ints():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = int.from_bytes(rng.bytes(32), "little") % Q
        b = int.from_bytes(rng.bytes(32), "little") % Q
        A, B = oc.to_arr([a])[0], oc.to_arr([b])[0]
        assert oc.from_mont_bytes(L(op2("fq_mul", A, B)).tobytes()) == a * b % Q
        assert oc.from_mont_bytes(L(op2("fq_add", A, B)).tobytes()) == (a + b) % Q
        assert oc.from_mont_bytes(L(op2("fq_sub", A, B)).tobytes()) == (a - b) % Q
    w = rng.bytes(64)
    assert oc.scalar_from_bytes_wide(w) == int.from_bytes(w, "little") % Q


def test_batch_invert():
    vals = [3, 5, Q - 1, 12345678901234567890]
    arr = oc.to_arr(vals)
    allinv = np.zeros(4, dtype=np.uint64)
    lib.fq_batch_invert(P(arr), C.c_size_t(len(vals)), P(allinv))
    assert oc.to_ints(arr) == [pow(v, -1, Q) for v in vals]
