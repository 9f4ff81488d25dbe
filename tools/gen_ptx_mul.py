"""Generator + bit-exact emulator for the device multiplications of spartan_b200/csrc/field.cuh.

Emits spartan_b200/csrc/mul_ptx.cuh: fq_mul_ptx (Montgomery product mod q, R = 2^256) and fp_mul_ptx (product mod 2^255-19, loose result)
as single inline-PTX blocks built from mad.lo.cc / madc.hi.cc carry chains:
  * 8x8-limb product: products a[j]*b[i] with i+j even accumulate into an "even" array E (position i+j), those with i+j odd into an "odd"
    array O (position i+j-1), so every row is one uninterrupted carry chain per array (no carry ripples); E and O are merged by one add chain.
  * Fq: word-serial Montgomery reduction exploiting q = 2^252 + c (limbs 4..6 of q are zero, limb 7 is 2^28); carries that would ripple
    through untouched limbs are parked in small counters C[k] and folded in when limb k is next needed.
  * Fp: hi*38 folded into lo with two chains.
Before writing the header the instruction lists are run through a PTX-subset emulator on random and edge inputs and compared with Python
integers, so the arithmetic is checked without a GPU (the GPU parity tests then check the build)."""
import os
import random
import sys

Q = 2**252 + 27742317777372353535851937790883648493
P = 2**255 - 19
QL = [(Q >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
QINV32 = (-pow(Q, -1, 2**32)) % 2**32
M32 = 0xFFFFFFFF


class Prog:
    def __init__(self):
        self.ins = []      # (op, dst, a, b, c)
        self.regs = set()

    def r(self, name):
        self.regs.add(name)
        return name

    def emit(self, op, dst, *src):
        self.regs.add(dst)
        self.ins.append((op, dst) + tuple(src))


def val(regs, x):
    return x if isinstance(x, int) else regs[x]


def run(prog, regs):
    cc = 0
    for ins in prog.ins:
        op, dst, src = ins[0], ins[1], ins[2:]
        v = [val(regs, s) for s in src]
        if op == "mul.lo":
            res = (v[0] * v[1]) & M32
        elif op == "mul.hi":
            res = (v[0] * v[1]) >> 32
        elif op in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc", "madc.hi", "madc.lo", "mad.hi.cc", "mad.lo"):
            prod = v[0] * v[1]
            part = (prod & M32) if ".lo" in op else (prod >> 32)
            cin = cc if op.startswith("madc") else 0
            t = part + v[2] + cin
            res = t & M32
            if op.endswith(".cc"):
                cc = t >> 32
        elif op in ("add.cc", "addc.cc", "addc", "add"):
            cin = cc if op.startswith("addc") else 0
            t = v[0] + v[1] + cin
            res = t & M32
            if op.endswith(".cc"):
                cc = t >> 32
        elif op in ("sub.cc", "subc.cc", "subc"):
            bin_ = cc if op.startswith("subc") else 0   # PTX: CC holds the borrow for sub
            t = v[0] - v[1] - bin_
            res = t & M32
            if op.endswith(".cc"):
                cc = 1 if t < 0 else 0
        elif op == "shl":
            res = (v[0] << v[1]) & M32
        elif op == "shr":
            res = v[0] >> v[1]
        elif op == "mov":
            res = v[0]
        elif op == "and":
            res = v[0] & v[1]
        elif op == "or":
            res = v[0] | v[1]
        else:
            raise ValueError(op)
        regs[dst] = res
    return regs


def wide_mul(p, a, b):
    """returns list T[0..15] of register names holding a*b"""
    E, O = {}, {}

    def chain(arr, name, start, i, js):
        first = True
        for k in range(8):
            j = js[k // 2]
            idx = start + k
            hi = k & 1
            addend = arr.get(idx, 0)
            dst = p.r("%s%d" % (name, idx))
            fresh_chain = first and all((start + kk) not in arr for kk in range(8))
            if fresh_chain or (first and addend == 0 and False):
                pass
            if first:
                op = "mad.lo.cc"
            else:
                op = "madc.hi.cc" if hi else "madc.lo.cc"
            p.emit(op, dst, a[j], b[i], addend)
            arr[idx] = dst
            first = False
        # carry out
        idx = start + 8
        if idx <= 15 or (name == "o" and idx <= 14):
            dst = p.r("%s%d" % (name, idx))
            p.emit("addc", dst, arr.get(idx, 0), 0)
            arr[idx] = dst

    for i in range(8):
        js_e = [j for j in range(8) if (i + j) % 2 == 0]
        js_o = [j for j in range(8) if (i + j) % 2 == 1]
        if i == 0:
            # fresh rows: plain mul.lo / mul.hi, no carries
            for k, j in enumerate(js_e):
                lo, hi = p.r("e%d" % (2 * k)), p.r("e%d" % (2 * k + 1))
                p.emit("mul.lo", lo, a[j], b[0]); p.emit("mul.hi", hi, a[j], b[0])
                E[2 * k], E[2 * k + 1] = lo, hi
            for k, j in enumerate(js_o):
                lo, hi = p.r("o%d" % (2 * k)), p.r("o%d" % (2 * k + 1))
                p.emit("mul.lo", lo, a[j], b[0]); p.emit("mul.hi", hi, a[j], b[0])
                O[2 * k], O[2 * k + 1] = lo, hi
            continue
        chain(E, "e", i + js_e[0], i, js_e)
        chain(O, "o", i + js_o[0] - 1, i, js_o)
    # merge: T[k] = E[k] + O[k-1]
    T = [E[0]]
    for k in range(1, 16):
        dst = p.r("t%d" % k)
        op = "add.cc" if k == 1 else ("addc.cc" if k < 15 else "addc")
        p.emit(op, dst, E.get(k, 0), O.get(k - 1, 0))
        T.append(dst)
    return T


def gen_fp_mul():
    p = Prog()
    a = ["a%d" % i for i in range(8)]
    b = ["b%d" % i for i in range(8)]
    T = wide_mul(p, a, b)
    r = [p.r("r%d" % i) for i in range(8)]
    # r = lo + 38*hi : lo parts of the products, then hi parts one limb up
    for k in range(8):
        p.emit("mad.lo.cc" if k == 0 else "madc.lo.cc", r[k], T[8 + k], 38, T[k])
    p.emit("addc", "c1", 0, 0)
    for k in range(1, 8):
        p.emit("mad.hi.cc" if k == 1 else "madc.hi.cc", r[k], T[8 + k - 1], 38, r[k])
    p.emit("madc.hi", "c1", T[15], 38, "c1")
    # fold c1*38 (c1 < 2^7) and a possible last carry
    p.emit("mad.lo.cc", r[0], "c1", 38, r[0])
    for k in range(1, 8):
        p.emit("addc.cc", r[k], r[k], 0)
    p.emit("addc", "c2", 0, 0)
    p.emit("mad.lo", r[0], "c2", 38, r[0])
    return p, a, b, r


def gen_fq_mul():
    p = Prog()
    a = ["a%d" % i for i in range(8)]
    b = ["b%d" % i for i in range(8)]
    T = wide_mul(p, a, b)
    T = list(T) + [0]
    C = {}
    for k in range(1, 17):
        C[k] = 0

    def bump(k):
        dst = p.r("c%d" % k)
        p.emit("addc", dst, C[k], 0)
        C[k] = dst
    # Reduction rows are chained through a value that is always zero but that ptxas cannot prove zero (the word each row has just cleared,
    # and-ed with the last word it wrote; for row 0 the top bit of T[15], clear because a, b < 2^255): m_i = T[i]*QINV + dep_i.  Without it
    # ptxas overlaps all eight rows (and the 8x8 product), keeps ~20 carry predicates live and spills them bit by bit into a register
    # (120 LOP3 per multiplication on sm_100a); with it at most one row's chains are in flight and no predicate is spilled.
    dep = p.r("dep0")
    p.emit("shr", dep, T[15], 31)
    for i in range(8):
        if i > 0 and C[i] != 0:
            d = p.r("u%d" % i)
            p.emit("add.cc", d, T[i], C[i])
            T[i] = d
            bump(i + 1)
        m = p.r("m%d" % i)
        p.emit("mad.lo", m, T[i], QINV32, dep)
        zeroed = None
        # chain A: q0 at (i, i+1), q2 at (i+2, i+3)
        for k, (ql, part) in enumerate([(0, "lo"), (0, "hi"), (2, "lo"), (2, "hi")]):
            d = p.r("x%d_%d" % (i, k))
            op = "mad.lo.cc" if k == 0 else ("madc.%s.cc" % part)
            p.emit(op, d, m, QL[ql], T[i + k])
            T[i + k] = d
            if k == 0:
                zeroed = d      # T[i] + lo(m*q0) == 0 mod 2^32 by construction
        bump(i + 4)
        # chain B: q1 at (i+1, i+2), q3 at (i+3, i+4)
        for k, (ql, part) in enumerate([(1, "lo"), (1, "hi"), (3, "lo"), (3, "hi")]):
            d = p.r("y%d_%d" % (i, k))
            op = "mad.lo.cc" if k == 0 else ("madc.%s.cc" % part)
            p.emit(op, d, m, QL[ql], T[i + 1 + k])
            T[i + 1 + k] = d
        bump(i + 5)
        # q7 = 2^28 at (i+7, i+8)
        lo, hi = p.r("s%d" % i), p.r("h%d" % i)
        p.emit("shl", lo, m, 28)
        p.emit("shr", hi, m, 4)
        d = p.r("z%d_0" % i); p.emit("add.cc", d, T[i + 7], lo); T[i + 7] = d
        d = p.r("z%d_1" % i); p.emit("addc.cc", d, T[i + 8], hi); T[i + 8] = d
        bump(i + 9)
        if i < 7:
            dep = p.r("dep%d" % (i + 1))
            p.emit("and", dep, d, zeroed)
    r = [p.r("r%d" % i) for i in range(8)]
    for k in range(8):
        p.emit("add.cc" if k == 0 else ("addc.cc" if k < 7 else "addc"), r[k], T[8 + k], C[8 + k])
    return p, a, b, r


def gen_fq_fold_const():
    """a0 + r*(a1 - a0) for a per-launch constant r, as a0 + sum_k d_k * RK[k] with d = a1 - a0 + q (< 2q) and the host-made table
    RK[k] = r * 2^(32k) mod q (plain integers, r out of Montgomery form): 8 rows of 8 wide products land on the SAME nine limbs (the reduction
    mod q of the shifted partial products is already inside the table), then one fold of the top 36 bits through 2^252 = -c (mod q).
    72 wide products instead of the 112 of a Montgomery multiplication, and the addition of a0 rides in the first row's carry chain.
    Inputs: a (= a0, canonical), b (= d, < 2q), k[0..63] (RK, row-major).  Output < 2q, congruent to a0 + r*(a1-a0)."""
    p = Prog()
    a = ["a%d" % i for i in range(8)]
    d = ["b%d" % i for i in range(8)]
    K = ["k%d" % i for i in range(64)]
    E, O = {}, {}
    for k in range(8):
        # even columns j = 0,2,4,6 -> limbs (j, j+1); odd columns j = 1,3,5,7 -> limbs (j, j+1) = O index (j-1, j)
        for n, j in enumerate([0, 2, 4, 6]):
            for part, idx in (("lo", j), ("hi", j + 1)):
                dst = p.r("e%d" % idx)
                addend = a[idx] if k == 0 else E[idx]
                first = n == 0 and part == "lo"
                op = "mad.lo.cc" if first else "madc.%s.cc" % part
                p.emit(op, dst, d[k], K[8 * k + j], addend)
                E[idx] = dst
        dst = p.r("e8")
        p.emit("addc", dst, E.get(8, 0), 0)
        E[8] = dst
        for n, j in enumerate([1, 3, 5, 7]):
            for part, idx in (("lo", j - 1), ("hi", j)):
                dst = p.r("o%d" % idx)
                if k == 0:
                    p.emit("mul.%s" % part, dst, d[k], K[8 * k + j])
                else:
                    first = n == 0 and part == "lo"
                    last = n == 3 and part == "hi"
                    op = "mad.lo.cc" if first else ("madc.hi" if last else "madc.%s.cc" % part)   # O < 2^256: no carry out of its top limb
                    p.emit(op, dst, d[k], K[8 * k + j], O[idx])
                O[idx] = dst
    T = [E[0]]
    for i in range(1, 9):
        dst = p.r("t%d" % i)
        p.emit("add.cc" if i == 1 else ("addc.cc" if i < 8 else "addc"), dst, E[i], O[i - 1])
        T.append(dst)
    # T = T_hi * 2^252 + T_lo,  T_hi < 2^36:  T = T_lo - T_hi * c (mod q), c = q - 2^252 (limbs QL[0..3])
    p.emit("shr", "h0a", T[7], 28); p.emit("shl", "h0b", T[8], 4); p.emit("or", "hi0", "h0a", "h0b")
    p.emit("shr", "hi1", T[8], 28)
    p.emit("and", "t7m", T[7], 0x0FFFFFFF)
    p.regs.update(["h0a", "h0b", "hi0", "hi1", "t7m"])
    p.emit("mul.lo", "x0", "hi0", QL[0]); p.emit("mul.hi", "x1", "hi0", QL[0])
    p.emit("mul.lo", "x2", "hi0", QL[2]); p.emit("mul.hi", "x3", "hi0", QL[2])
    p.emit("mul.lo", "y1", "hi0", QL[1]); p.emit("mul.hi", "y2", "hi0", QL[1])
    p.emit("mul.lo", "y3", "hi0", QL[3]); p.emit("mul.hi", "y4", "hi0", QL[3])
    # hi1 * c * 2^32: c0 -> limbs (1,2), c2 -> (3,4) join Y; c1 -> (2,3), c3 -> (4,5) join X
    p.emit("mad.lo.cc", "y1", "hi1", QL[0], "y1"); p.emit("madc.hi.cc", "y2", "hi1", QL[0], "y2")
    p.emit("madc.lo.cc", "y3", "hi1", QL[2], "y3"); p.emit("madc.hi.cc", "y4", "hi1", QL[2], "y4")
    p.emit("addc", "y5", 0, 0)
    p.emit("mad.lo.cc", "x2", "hi1", QL[1], "x2"); p.emit("madc.hi.cc", "x3", "hi1", QL[1], "x3")
    p.emit("madc.lo.cc", "x4", "hi1", QL[3], 0); p.emit("madc.hi", "x5", "hi1", QL[3], 0)
    p.emit("add.cc", "x1", "x1", "y1"); p.emit("addc.cc", "x2", "x2", "y2"); p.emit("addc.cc", "x3", "x3", "y3")
    p.emit("addc.cc", "x4", "x4", "y4"); p.emit("addc", "x5", "x5", "y5")
    # U = T_lo + q - X   in (0, 2q)
    r = [p.r("r%d" % i) for i in range(8)]
    lo = T[:7] + ["t7m"]
    for i in range(8):
        p.emit("add.cc" if i == 0 else ("addc.cc" if i < 7 else "addc"), r[i], lo[i], QL[i])
    X = ["x0", "x1", "x2", "x3", "x4", "x5", 0, 0]
    for i in range(8):
        p.emit("sub.cc" if i == 0 else ("subc.cc" if i < 7 else "subc"), r[i], r[i], X[i])
    return p, a, d, K, r


def emit_fold_c(name, prog, a, d, K, r, comment):
    temps = sorted(x for x in prog.regs if x not in a and x not in d and x not in K)
    lines = ["// " + comment, "__device__ __forceinline__ u256 %s(const u256& a, const u256& b, const FqConst& rc) {" % name, "  u256 r;", "  asm(\"{\\n\\t\""]
    for i in range(0, len(temps), 24):
        lines.append("      \".reg .u32 %s;\\n\\t\"" % ", ".join(temps[i:i + 24]))

    def opnd(x):
        if isinstance(x, int):
            return str(x)
        if x in a:
            return "%%%d" % (8 + a.index(x))
        if x in d:
            return "%%%d" % (16 + d.index(x))
        if x in K:
            return "%%%d" % (24 + K.index(x))
        return x
    ptxop = {"mul.lo": "mul.lo.u32", "mul.hi": "mul.hi.u32", "shl": "shl.b32", "shr": "shr.u32", "mov": "mov.u32", "and": "and.b32", "or": "or.b32"}
    for ins in prog.ins:
        op, dst, src = ins[0], ins[1], ins[2:]
        lines.append("      \"%s %s, %s;\\n\\t\"" % (ptxop.get(op, op + ".u32"), dst, ", ".join(opnd(s_) for s_ in src)))
    for i in range(8):
        lines.append("      \"mov.u32 %%%d, %s;\\n\\t\"" % (i, r[i]))
    lines.append("      \"}\"")
    lines.append("      : " + ", ".join("\"=r\"(r.v[%d])" % i for i in range(8)))
    lines.append("      : " + ", ".join("\"r\"(a.v[%d])" % i for i in range(8)) + ", " + ", ".join("\"r\"(b.v[%d])" % i for i in range(8)) + ",")
    for i in range(0, 64, 16):
        lines.append("        " + ", ".join("\"r\"(rc.k[%d])" % j for j in range(i, i + 16)) + ("," if i < 48 else ");"))
    lines.append("  return r;")
    lines.append("}")
    return "\n".join(lines)


def check_fold(prog, a, d, K, r, nrand=4000):
    rnd = random.Random(7)
    Rinv = pow(2**256, -1, Q)
    edge = [0, 1, 2, Q - 1, Q - 2, 2**252, 2**252 - 1, 2**128, 2**32 - 1]
    cases = [(x, y, z) for x in edge for y in edge for z in edge[:6]] + [(rnd.getrandbits(256) % Q, rnd.getrandbits(256) % Q, rnd.getrandbits(256) % Q) for _ in range(nrand)]
    worst = 0
    for a0, a1, rm in cases:
        rplain = rm * Rinv % Q                       # r out of Montgomery form
        rk = [(rplain << (32 * k)) % Q for k in range(8)]
        dd = a1 - a0 + Q
        regs = {}
        for i, v in enumerate(limbs(a0)):
            regs[a[i]] = v
        for i, v in enumerate(limbs(dd)):
            regs[d[i]] = v
        for k in range(8):
            for j, v in enumerate(limbs(rk[k])):
                regs[K[8 * k + j]] = v
        run(prog, regs)
        got = sum(regs[r[i]] << (32 * i) for i in range(8))
        want = (a0 + rm * (a1 - a0) * Rinv) % Q      # Montgomery-domain a0 + r*(a1-a0), exactly what fq_add(a0, fq_mul(r, fq_sub(a1,a0))) returns
        assert 0 < got < 2 * Q and got % Q == want, (hex(a0), hex(a1), hex(rm), hex(got), hex(want))
        worst = max(worst, got)
    return worst


def limbs(x):
    return [(x >> (32 * i)) & M32 for i in range(8)]


def check(prog, a, b, r, ref, nrand=20000):
    rnd = random.Random(1)
    edge = [0, 1, 2, Q - 1, Q - 2, P - 1, 2**255 - 1, 2**256 - 1, 2**256 - 38, 2**252, 2**128, 2**32 - 1, 2**64 - 1, 38, 19]
    cases = [(x, y) for x in edge for y in edge] + [(rnd.getrandbits(256), rnd.getrandbits(256)) for _ in range(nrand)]
    for x, y in cases:
        x, y = ref["dom"](x), ref["dom"](y)
        regs = {}
        for i, v in enumerate(limbs(x)):
            regs[a[i]] = v
        for i, v in enumerate(limbs(y)):
            regs[b[i]] = v
        run(prog, regs)
        got = sum(regs[r[i]] << (32 * i) for i in range(8))
        ref["check"](x, y, got)


def emit_c(name, prog, a, b, r, comment):
    temps = sorted(x for x in prog.regs if x not in a and x not in b)
    lines = ["// " + comment, "__device__ __forceinline__ u256 %s(const u256& a, const u256& b) {" % name, "  u256 r;", "  asm(\"{\\n\\t\""]
    decl = ", ".join(temps)
    # split declaration into chunks to keep lines short
    for i in range(0, len(temps), 24):
        lines.append("      \".reg .u32 %s;\\n\\t\"" % ", ".join(temps[i:i + 24]))

    def opnd(x):
        if isinstance(x, int):
            return str(x)
        if x in a:
            return "%%%d" % (8 + a.index(x))
        if x in b:
            return "%%%d" % (16 + b.index(x))
        return x
    ptxop = {"mul.lo": "mul.lo.u32", "mul.hi": "mul.hi.u32", "shl": "shl.b32", "shr": "shr.u32", "mov": "mov.u32", "and": "and.b32"}
    for ins in prog.ins:
        op, dst, src = ins[0], ins[1], ins[2:]
        o = ptxop.get(op, op + ".u32")
        lines.append("      \"%s %s, %s;\\n\\t\"" % (o, dst, ", ".join(opnd(s) for s in src)))
    for i in range(8):
        lines.append("      \"mov.u32 %%%d, %s;\\n\\t\"" % (i, r[i]))
    lines.append("      \"}\"")
    lines.append("      : " + ", ".join("\"=r\"(r.v[%d])" % i for i in range(8)))
    lines.append("      : " + ", ".join("\"r\"(a.v[%d])" % i for i in range(8)) + ", " + ", ".join("\"r\"(b.v[%d])" % i for i in range(8)) + ");")
    lines.append("  return r;")
    lines.append("}")
    return "\n".join(lines)


def main():
    pfq, a, b, r = gen_fq_mul()
    Rinv = pow(2**256, -1, Q)

    def chk_fq(x, y, got):
        want = x * y * Rinv % Q
        assert got < 2 * Q and got % Q == want, (hex(x), hex(y), hex(got), hex(want))
    check(pfq, a, b, r, {"dom": lambda v: v % Q, "check": chk_fq})
    # the same routine also accepts lazily reduced operands (sums / differences of two residues): a, b < 4q still give a result < 2q, because
    # a*b/2^256 < 16 q^2 / 2^256 < q.  Not relied upon yet (every caller passes canonical residues); checked here so a lazy-reduction variant of
    # the sumcheck kernels can build on it.
    check(pfq, a, b, r, {"dom": lambda v: v % (4 * Q), "check": chk_fq}, nrand=5000)
    pfp, a2, b2, r2 = gen_fp_mul()

    def chk_fp(x, y, got):
        assert got < 2**256 and got % P == x * y % P, (hex(x), hex(y), hex(got))
    check(pfp, a2, b2, r2, {"dom": lambda v: v % 2**256, "check": chk_fp})
    # lazily reduced operands of the evaluation products: up to 5q (x3 = 3*hi - 2*lo + 2q) still gives < 2^256 and the right residue
    def chk_fq_loose(x, y, got):
        assert got < 2**256 and got % Q == x * y * Rinv % Q
    check(pfq, a, b, r, {"dom": lambda v: v % (5 * Q), "check": chk_fq_loose}, nrand=3000)
    pfo, fa, fd, fK, fr = gen_fq_fold_const()
    check_fold(pfo, fa, fd, fK, fr)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spartan_b200", "csrc", "mul_ptx.cuh")
    with open(out, "w") as f:
        f.write("// GENERATED by tools/gen_ptx_mul.py (emulator-checked against Python integers before emission) — do not edit.\n")
        f.write("// fq_mul_ptx: %d PTX instructions, result < 2q (caller subtracts q once).  fp_mul_ptx: %d PTX instructions, loose result < 2^256.\n" % (len(pfq.ins) + 8, len(pfp.ins) + 8))
        f.write("#pragma once\n#if defined(__CUDA_ARCH__)\nnamespace sp {\n")
        f.write(emit_c("fq_mul_ptx", pfq, a, b, r, "Montgomery product a*b*2^-256 mod q, unreduced in [0, 2q)") + "\n")
        f.write(emit_c("fp_mul_ptx", pfp, a2, b2, r2, "a*b mod 2^255-19, loose in [0, 2^256)") + "\n")
        f.write("// per-launch constant multiplier of the fold  a0 + r*(a1-a0)  (FqConst, field.cuh):  k[8*j+i] = limb i of (r * 2^(32j) mod q), r out of Montgomery form\n")
        f.write(emit_fold_c("fq_fold_const_ptx", pfo, fa, fd, fK, fr, "a + sum_j b_j * RK[j] folded once through 2^252 = -c: result in (0, 2q), congruent to a0 + r*(a1-a0); b = a1 - a0 + q; %d PTX instructions" % (len(pfo.ins) + 8)) + "\n")
        f.write("}  // namespace sp\n#endif\n")
    print("wrote", out, "fq", len(pfq.ins), "fp", len(pfp.ins), "fold_const", len(pfo.ins))


if __name__ == "__main__":
    main()
