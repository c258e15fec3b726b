"""Pins the CPU oracle (oracle/bfv_oracle.cpp) against an independent big-integer BFV and against the constants
SURVEY.md section 8c lists for SEAL 3.2.  CPU only."""
import numpy as np
import pytest

from oracle.oracle_py import Oracle, lib
from oracle import textbook_bfv as tb

SMALL_Q = [1073479681, 1073184769, 1072496641]  # 30-bit primes = 1 mod 2^15 (found by search below)


def _small_primes(count, bits, mod):
    out, c = [], (1 << bits) + 1
    from sympy import isprime
    while len(out) < count:
        c -= mod
        if isprime(c):
            out.append(c)
    return out


def test_default_coeff_moduli_follow_rule():
    # largest b-bit primes = 1 mod 2N, decreasing, with SEAL 3.2's bit splits
    from sympy import isprime
    splits = {4096: [36, 36, 37], 8192: [43, 43, 44, 44, 44], 16384: [48, 48, 48, 49, 49, 49, 49, 49, 49]}
    for N, bits in splits.items():
        o = Oracle({4096: 40961, 8192: 65537, 16384: 65537}[N], N)
        expect, seen = [], {}
        for b in bits:
            c = seen.get(b, (1 << b) + 1)
            while True:
                c -= 2 * N
                if isprime(c):
                    break
            seen[b] = c
            expect.append(c)
        assert o.q == expect
        assert sum(x.bit_length() for x in o.q) == {4096: 109, 8192: 218, 16384: 438}[N]


def test_aux_base_constants():
    o = Oracle(65537, 8192)
    assert o.bsk[:5] == [0x1fffffffffb40001, 0x1fffffffff500001, 0x1fffffffff380001, 0x1fffffffff000001, 0x1ffffffffef00001]
    assert o.bsk[-1] == 0x1fffffffffe00001
    assert lib().orc_gamma(o.h) == 0x1fffffffffc80001


@pytest.mark.parametrize("logn", [3, 5])
def test_ntt_matches_naive_transform(logn):
    N = 1 << logn
    q = _small_primes(2, 30, 1 << 15)
    o = Oracle(65537 if N <= 32768 else 0, N, custom_q=q)
    rng = np.random.default_rng(7)
    for which in list(range(o.k)) + [o.k, 2 * o.k, 2 * o.k + 1]:
        p = o.modulus_of(which)
        psi = int(lib().orc_minimal_primitive_root(2 * N, p))
        # minimal primitive 2N-th root
        roots = [r for r in range(2, min(p, 200000)) if pow(r, N, p) == p - 1][:1]
        if roots:
            assert psi <= roots[0]
        assert pow(psi, N, p) == p - 1
        a = rng.integers(0, p, N, dtype=np.uint64)
        got = o.ntt(which, a)
        assert [int(x) for x in got] == tb.naive_negacyclic_ntt(a, psi, p, logn)
        back = o.ntt(which, got, inverse=True)
        assert np.array_equal(back, a)


def test_minimal_root_is_minimal():
    p, N = 40961, 4096
    psi = int(lib().orc_minimal_primitive_root(2 * N, p))
    cands = [r for r in range(2, p) if pow(r, N, p) == p - 1]
    assert psi == min(cands)


@pytest.mark.parametrize("centered", [0, 1])
def test_multiply_relinearize_against_bigint(centered):
    N, t = 16, 97
    q = _small_primes(3, 30, 1 << 15)
    o = Oracle(t, N, custom_q=q, dbc_relin=10, dbc_galois=20)
    o.set_centered_mtilde(centered)
    o.keygen(11)
    rng = np.random.default_rng(3)
    m1 = rng.integers(0, t, N, dtype=np.uint64)
    m2 = rng.integers(0, t, N, dtype=np.uint64)
    c1, c2 = o.encrypt(m1, 1), o.encrypt(m2, 2)
    # secret key back to coefficient form, as a ternary integer polynomial
    sk = o.secret_key().reshape(o.k, N)
    s0 = o.ntt(0, sk[0], inverse=True)
    s = [tb.center(int(v), o.q[0]) for v in s0]
    assert set(s) <= {-1, 0, 1}

    def polys(ct, size):
        ct = ct.reshape(size, o.k, N)
        out = []
        for part in range(size):
            v, Q = tb.crt_compose([ct[part, i] for i in range(o.k)], o.q)
            out.append(v)
        return out, Q

    p1, Q = polys(c1, 2)
    p2, _ = polys(c2, 2)
    dec, _ = tb.decrypt_exact(p1, s, t, Q)
    assert dec == [int(x) for x in m1]
    assert [int(x) for x in o.decrypt(c1)] == dec

    prod = o.multiply(c1, c2)
    pp, _ = polys(prod, 3)
    # exact tensor over the integers.  The Montgomery lift (mont_rq) leaves the representative of each input in
    # [0, q(1 + k/2^32)) when r_mtilde is taken in [0, m~) (our reading of SEAL 3.2) and in about [-q/2, q/2) when it is
    # centred (SEAL >= 3.3), so the exact product is formed from canonical resp. centred representatives.
    a = [[tb.center(v, Q) if centered else v for v in part] for part in p1]
    b = [[tb.center(v, Q) if centered else v for v in part] for part in p2]
    d0 = tb.negacyclic_mul(a[0], b[0])
    d1 = [x + y for x, y in zip(tb.negacyclic_mul(a[0], b[1]), tb.negacyclic_mul(a[1], b[0]))]
    d2 = tb.negacyclic_mul(a[1], b[1])
    for got, d in zip(pp, (d0, d1, d2)):
        for g, e in zip(got, d):
            exact = tb.round_div(t * e, Q)
            diff = tb.center((g - exact) % Q, Q)
            # BEHZ floor is exact up to the fast-base-conversion overflow alpha in [0,k) (plus rounding vs floor)
            assert abs(diff) <= o.k + 1, diff
    want = [v % t for v in tb.negacyclic_mul([int(x) for x in m1], [int(x) for x in m2])]
    dec3, _ = tb.decrypt_exact(pp, s, t, Q)
    assert dec3 == want
    assert [int(x) for x in o.decrypt(prod)] == want
    rel = o.relinearize(prod)
    pr, _ = polys(rel, 2)
    dec2, _ = tb.decrypt_exact(pr, s, t, Q)
    assert dec2 == want
    assert [int(x) for x in o.decrypt(rel)] == want


def test_galois_and_rotation_semantics():
    N, t = 32, 193
    q = _small_primes(3, 30, 1 << 15)
    o = Oracle(t, N, custom_q=q, dbc_relin=10, dbc_galois=20)
    o.keygen(5)
    vals = np.arange(1, N + 1, dtype=np.uint64)
    ct = o.encrypt(o.encode(vals), 1)
    row = N // 2
    for steps in (1, -1, 3, -5, 7):
        got = o.decode(o.decrypt(o.rotate_rows(ct, steps)))
        want = np.concatenate([np.roll(vals[:row], -steps), np.roll(vals[row:], -steps)])
        assert np.array_equal(got, want), steps
    got = o.decode(o.decrypt(o.rotate_columns(ct)))
    assert np.array_equal(got, np.concatenate([vals[row:], vals[:row]]))
    # NAF decompositions quoted in SURVEY 8c
    assert o.galois_elt_from_step(0) == 2 * N - 1
