"""Independent big-integer ("textbook") BFV used only to cross-check the C++ oracle on small N.

TEST INFRASTRUCTURE ONLY.  Pure Python integers: no RNS, no NTT, no fast base conversion -- so an agreement with
oracle/bfv_oracle.cpp is evidence about the algorithms, not about shared code.
"""
from functools import reduce


def crt_compose(residues, moduli):
    """residues[i][n] -> list of integers in [0, prod moduli)."""
    Q = reduce(lambda a, b: a * b, moduli)
    out = [0] * len(residues[0])
    for r, p in zip(residues, moduli):
        qh = Q // p
        c = qh * pow(qh % p, -1, p)
        for n, v in enumerate(r):
            out[n] = (out[n] + int(v) * c) % Q
    return out, Q


def center(v, Q):
    return v - Q if v > Q // 2 else v


def negacyclic_mul(a, b):
    n = len(a)
    out = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k >= n:
                out[k - n] -= x * y
            else:
                out[k] += x * y
    return out


def round_div(a, b):
    """round(a / b) to nearest, b > 0."""
    return (2 * a + b) // (2 * b)


def naive_negacyclic_ntt(poly, psi, p, logn):
    """X[bitrev(j)] = sum_i a_i psi^{(2j+1) i}  -- the transform SEAL's ntt_negacyclic_harvey computes."""
    n = len(poly)

    def bitrev(x):
        r = 0
        for _ in range(logn):
            r = (r << 1) | (x & 1)
            x >>= 1
        return r

    out = [0] * n
    for j in range(n):
        w = pow(psi, 2 * j + 1, p)
        acc, cur = 0, 1
        for a in poly:
            acc = (acc + int(a) * cur) % p
            cur = cur * w % p
        out[bitrev(j)] = acc
    return out


def decrypt_exact(ct_polys, s, t, Q):
    """ct_polys: list of integer polys (c0, c1[, c2]); s: ternary secret; returns m = round(t/Q * [c0 + c1 s + ..]_Q) mod t."""
    n = len(s)
    acc = list(ct_polys[0])
    sp = list(s)
    for part in ct_polys[1:]:
        prod = negacyclic_mul(part, sp)
        acc = [(a + b) % Q for a, b in zip(acc, prod)]
        sp = [center(v % Q, Q) for v in negacyclic_mul(sp, s)]
    return [round_div(t * center(v, Q), Q) % t for v in acc], acc
