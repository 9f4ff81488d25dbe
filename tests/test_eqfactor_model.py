"""CPU model of the eq-factored sumcheck rounds (spartan_b200/csrc/snark.cpp batched_prove, kernels_sc.cu k_sc_eval_g / k_sc_fold_eval_g / k_eq_suffix).

For sum_x eq(tau, x) A(x) B(x) the round polynomial factors as s_j(t) = prefix_j * eq(tau_j, t) * q_j(t) with q_j(t) = sum_x' E_j[x'] A(t, x') B(t, x'),
E_j = eq(tau[j+1..], .).  The device reduces q_j(0) and the leading coefficient only; the host recovers q_j(1) from the running claim and rebuilds the
evaluations at 0..3 that the reference computes directly (sumcheck.rs:296-355).  This test replays that bookkeeping in Python integers against the
direct evaluation, incl. the suffix tables as sums over the top index bits of E_0 and the table C = prefix * eq(tau[j..], .) used at the switch."""
import random

Q = 2**252 + 27742317777372353535851937790883648493


def eq_table(tau):
    out = [1]
    for x in tau:                      # variable 0 = most significant index bit
        out = [w for v in out for w in (v * ((1 - x) % Q) % Q, v * x % Q)]
    return out


def test_eq_table_order():
    tau = [3, 5]
    t = eq_table(tau)
    assert t[0b10] == 3 * ((1 - 5) % Q) % Q and t[0b01] == ((1 - 3) % Q) * 5 % Q


def test_round_polynomials_and_switch_table():
    rnd = random.Random(11)
    n = 6
    tau = [rnd.randrange(Q) for _ in range(n)]
    A = [rnd.randrange(Q) for _ in range(1 << n)]
    B = [rnd.randrange(Q) for _ in range(1 << n)]
    C = eq_table(tau)
    inv = lambda x: pow(x, Q - 2, Q)
    E0 = eq_table(tau[1:])
    cP, prefix = sum(a * b * c for a, b, c in zip(A, B, C)) % Q, 1
    for j in range(n):
        h = len(A) // 2
        direct = []
        for t in range(4):
            s = 0
            for i in range(h):
                a = (A[i] + t * (A[i + h] - A[i])) % Q
                b = (B[i] + t * (B[i + h] - B[i])) % Q
                c = (C[i] + t * (C[i + h] - C[i])) % Q
                s += a * b * c
            direct.append(s % Q)
        # suffix table of this round = sum over the top j index bits of E0 (k_eq_suffix)
        sz = len(E0) >> j
        E = [sum(E0[b * sz + y] for b in range(1 << j)) % Q for y in range(sz)] if sz else [1]
        assert E == (eq_table(tau[j + 1:]) if j + 1 < n else [1])
        q0 = sum(E[i] * A[i] * B[i] for i in range(h)) % Q
        qinf = sum(E[i] * (A[i + h] - A[i]) * (B[i + h] - B[i]) for i in range(h)) % Q
        q1 = (cP - (1 - tau[j]) * q0) * inv(tau[j]) % Q
        qb = (q1 - q0 - qinf) % Q
        q = lambda t: ((qinf * t + qb) * t + q0) % Q
        rebuilt = [prefix * (((1 - tau[j]) + t * (2 * tau[j] - 1)) % Q) * q(t) % Q for t in range(4)]
        assert rebuilt == direct
        r = rnd.randrange(Q)
        cP, prefix = q(r), prefix * (((1 - tau[j]) + r * (2 * tau[j] - 1)) % Q) % Q
        A = [(A[i] + r * (A[i + h] - A[i])) % Q for i in range(h)]
        B = [(B[i] + r * (B[i + h] - B[i])) % Q for i in range(h)]
        C = [(C[i] + r * (C[i + h] - C[i])) % Q for i in range(h)]
        assert C == [prefix * e % Q for e in (eq_table(tau[j + 1:]) if j + 1 < n else [1])]   # what the standard kernels continue on after the switch
