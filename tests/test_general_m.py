"""Rows for general (non power-of-two) m: Bluestein transforms and the DoubleCRT operations on top of
them (reference src/bluestein.cpp:77-201, src/CModulus.cpp:148-180,
431-443,555-577).  Checked against the Python restatement (definition-level DFT + remainder mod Phi_m).
For general m the root is pinned: FindPrimitiveRoot is deterministic (src/NumbTh.cpp:435-493)."""
import random

import numpy as np
import pytest

import orc
import pyoracle as po
from helib_b200 import Engine, HbError


def backends():
    return [pytest.param("sim", id="sim"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=backends())
def lib(request):
    return request.getfixturevalue("sim_lib" if request.param == "sim" else "cuda_lib")


CFGS = [(12, 7, 1, 100, 2), (45, 2, 1, 100, 2), (28, 3, 1, 100, 2), (105, 2, 1, 120, 2), (1285, 2, 1, 120, 2)]


def setup(lib, cfg):
    m, p, r, bits, c = cfg
    ch = po.build_mod_chain(m, p, r, bits, c)
    roots = [po.cmod_root(q, m) for q in ch.primes]
    E = Engine(m, ch.primes, None, ch.digits, ch.special, lib=lib)
    assert E.psis == roots, "engine's FindPrimitiveRoot restatement disagrees with the oracle's"
    assert E.N == ch.phim
    return ch, roots, E


def dense(ch, rows):
    out = np.zeros((len(ch.primes), ch.phim), dtype=np.uint64)
    for i, r in rows.items():
        out[i] = np.array(r, dtype=np.uint64)
    return out


@pytest.mark.parametrize("cfg", CFGS)
def test_bluestein_rows_match_definition(lib, cfg):
    ch, roots, E = setup(lib, cfg)
    rnd = random.Random(5)
    idx = ch.ctxt + ch.special
    coef = {i: [rnd.randrange(ch.primes[i]) for _ in range(ch.phim)] for i in idx}
    P = E.poly(dense(ch, coef), idx)
    E.ntt_fwd([P], idx)
    got = P.download(idx)
    small = ch.m <= 200
    for i in idx[:3] if not small else idx:
        ref = po.gen_fft(coef[i], ch.primes[i], ch.m, roots[i]) if small else None
        if small:
            assert list(got[i]) == ref
        else:   # spot-check the definition row[j] = f(zeta^rep(j)) at a few j
            q, zeta, rep = ch.primes[i], roots[i] * roots[i] % ch.primes[i], po.zms_rep(ch.m)
            for j in (0, 1, ch.phim // 2, ch.phim - 1):
                assert int(got[i][j]) == sum(cf * pow(zeta, rep[j] * k, q) for k, cf in enumerate(coef[i])) % q
    E.ntt_inv([P], idx)
    back = P.download(idx)
    for i in idx:
        assert list(back[i]) == coef[i]       # iFFT(FFT(f)) == f (tests/TestHEXL.cpp:189-218)


@pytest.mark.parametrize("cfg", CFGS[:3])
def test_general_m_doublecrt_ops(lib, cfg):
    """toPoly / addPrimes / scaleDownToSet / breakIntoDigits / automorph on Bluestein rows vs big-int Python."""
    ch, roots, E = setup(lib, cfg)
    rnd = random.Random(6)
    m, n = ch.m, ch.phim
    S, Sp = ch.ctxt, sorted(ch.ctxt + ch.special)

    class GenD:   # minimal big-int model of a DoubleCRT over general m
        def __init__(self, rows):
            self.rows = rows

        def to_poly(self, idx):
            Q = ch.product(idx)
            cs = {i: po.gen_ifft(self.rows[i], ch.primes[i], m, roots[i]) for i in idx}
            out = []
            for k in range(n):
                acc = 0
                for i in idx:
                    q = ch.primes[i]; Qi = Q // q
                    acc += Qi * (cs[i][k] * pow(Qi % q, -1, q) % q)
                out.append(po.bal(acc, Q))
            return out

    def rows_of_poly(poly, idx):
        return {i: po.gen_fft([c % ch.primes[i] for c in poly], ch.primes[i], m, roots[i]) for i in idx}

    x = {i: [rnd.randrange(ch.primes[i]) for _ in range(n)] for i in Sp}
    X = GenD(x)
    P = E.poly(dense(ch, x), Sp)
    # toPoly
    assert orc.limbs_to_ints(E.to_poly(P, S)) == X.to_poly(S)
    # scaleDownToSet (drop the special primes), BGV correction with p
    p = ch.p ** ch.r
    Pd = ch.product(ch.special)
    delta = X.to_poly(ch.special)
    pinv = pow(Pd % p, -1, p)
    for k, d in enumerate(delta):
        u = d % p
        if u:
            u = u * pinv % p
            if u > p // 2 or (p % 2 == 0 and u == p // 2 and d < 0):
                u -= p
            delta[k] = d - Pd * u
    drows = rows_of_poly(delta, S)
    want = {i: [((a - b) * pow(Pd % ch.primes[i], -1, ch.primes[i])) % ch.primes[i] for a, b in zip(x[i], drows[i])] for i in S}
    E.scale_down([P], Sp, S, p)
    got = P.download(S)
    for i in S:
        assert list(got[i]) == want[i]
    # addPrimes back to the special primes: new rows are the balanced polynomial's residues
    Y = GenD(want)
    poly = Y.to_poly(S)
    new = rows_of_poly(poly, ch.special)
    E.add_primes([P], S, ch.special)
    got = P.download(ch.special)
    for i in ch.special:
        assert list(got[i]) == new[i]
    # breakIntoDigits: balanced mixed radix
    digs = E.break_into_digits([P], S)[0]
    acc, scale = [0] * n, 1
    for dnum, D in enumerate(digs):
        dset = [i for i in S if i in ch.digits[dnum]]
        drow = D.download(Sp)
        Ei = GenD({i: [int(v) for v in drow[i]] for i in dset}).to_poly(dset)
        for i in Sp:     # every row of the digit is the same small polynomial
            assert list(drow[i]) == rows_of_poly(Ei, [i])[i]
        acc = [a + e * scale for a, e in zip(acc, Ei)]
        scale *= ch.product(ch.digits[dnum])
    assert [po.bal(a, ch.product(S)) for a in acc] == poly
    # noise metadata for general m: basic_embeddingLargestCoeff (src/norms.cpp:129-157) next to the integer
    # results -- breakIntoDigits (ln of each digit's norm), addPrimes and scaleDownToSet (norm of delta / P).  FP64, rel. tol 1e-9.
    import math
    from fractions import Fraction
    P2 = E.poly(P.download(S), S)
    digs2, lognorms = E.break_into_digits_norm([P2], S)
    for dnum, D in enumerate(digs2[0]):
        dset = [i for i in S if i in ch.digits[dnum]]
        drow = D.download(dset)
        Ei = GenD({i: [int(v) for v in drow[i]] for i in dset}).to_poly(dset)
        mant, shift = po.embedding_largest_coeff(Ei, m)
        ref_log = math.log(mant) + shift * math.log(2.0)
        assert abs(lognorms[0, dnum] - ref_log) <= 1e-9 * abs(ref_log) + 1e-9, (dnum, lognorms[0, dnum], ref_log)
    y = {i: [rnd.randrange(ch.primes[i]) for _ in range(n)] for i in Sp}
    Yp = E.poly(dense(ch, y), Sp)
    norms = E.scale_down_norm([Yp], Sp, S, p)
    dl = GenD(y).to_poly(ch.special)
    for k_, d in enumerate(dl):
        u = d % p
        if u:
            u = u * pinv % p
            if u > p // 2 or (p % 2 == 0 and u == p // 2 and d < 0):
                u -= p
            dl[k_] = d - Pd * u
    ff = np.zeros(m)
    ff[:n] = [float(Fraction(d, Pd)) for d in dl]          # fdelta = delta / diffProd (src/Ctxt.cpp:482-485)
    vals = np.fft.fft(ff)
    want_n = float(np.max(np.abs(vals[[i for i in range(1, m // 2 + 1) if math.gcd(i, m) == 1]])))
    assert abs(norms[0] - want_n) <= 1e-9 * want_n, (norms[0], want_n)
    # automorph: F(X) -> F(X^k)
    k = [t for t in range(2, m) if np.gcd(t, m) == 1][1]
    D = E.poly()
    E.automorph([D], [P], S, k)
    rep = po.zms_rep(m)
    got = D.download(S)
    cur = P.download(S)
    for i in S:
        assert [int(v) for v in got[i]] == [int(cur[i][rep.index(rep[j] * k % m)]) for j in range(n)]
    # hoisting for general m (BasicAutomorphPrecon::automorph, src/matmul.cpp:112-184): one breakIntoDigits, then per amount
    # sigma_k on the digits and on c0 + the evk inner product == the same steps done one by one through the engine
    full = sorted(ch.ctxt + ch.special)
    nd = len(digs)
    evk = [E.poly(dense(ch, {i: [rnd.randrange(ch.primes[i]) for _ in range(n)] for i in full}), full) for _ in range(2 * nd)]
    c0rows = {i: [rnd.randrange(ch.primes[i]) for _ in range(n)] for i in S}
    C0 = E.poly(dense(ch, c0rows), S)
    for kk in [t for t in range(2, m) if np.gcd(t, m) == 1][:2] + [m - 1]:
        O0, O1 = E.poly(), E.poly()
        E.automorph_keyswitch_digits([digs], S, [C0], kk, evk[:nd], evk[nd:], [O0], [O1])
        R0, R1 = E.poly(), E.poly()
        E.automorph([R0], [C0], S, kk)
        E.add_primes_and_scale([R0], S, ch.special)
        rot = []
        for D in digs:
            T = E.poly()
            E.automorph([T], [D], full, kk)
            rot.append(T)
        E.keyswitch_digits([rot], full, evk[:nd], evk[nd:], [R0], [R1])
        assert (O0.download(full)[full] == R0.download(full)[full]).all() and (O1.download(full)[full] == R1.download(full)[full]).all(), kk


@pytest.mark.gpu
def test_thinboot_ring_m21845_rows(cuda_lib):
    """BASELINE config 5's ring (m = 21845 = 5*17*257, phi(m) = 16384; tests/GTestThinBootstrapping.cpp:102,
    p=2, bits=580, c=2): Bluestein rows at full size -- definition spot checks, round trip, and the
    relinearisation identity out = P*c0 + sum D_i*b_i checked through toPoly on a few coefficients."""
    m, p, r, bits, c = 21845, 2, 1, 580, 2
    ch = po.build_mod_chain(m, p, r, bits, c)
    assert ch.phim == 16384
    E = Engine(m, ch.primes, None, ch.digits, ch.special, lib=cuda_lib)
    roots = [po.cmod_root(q, m) for q in ch.primes[:2]]
    assert E.psis[:2] == roots
    rnd = random.Random(9)
    idx = ch.ctxt + ch.special
    coef = np.zeros((len(ch.primes), ch.phim), dtype=np.uint64)
    rng = np.random.default_rng(10)
    for i in idx:
        coef[i] = rng.integers(0, ch.primes[i], size=ch.phim, dtype=np.uint64)
    P = E.poly(coef, idx)
    E.ntt_fwd([P], idx)
    got = P.download(idx)
    rep = po.zms_rep(m)
    for i in idx[:2]:
        q, zeta = ch.primes[i], E.psis[i] * E.psis[i] % ch.primes[i]
        f = [int(v) for v in coef[i]]
        for j in (0, 7, ch.phim - 1):
            x = pow(zeta, rep[j], q)
            acc = 0
            for cf in reversed(f):
                acc = (acc * x + cf) % q
            assert int(got[i][j]) == acc
    E.ntt_inv([P], idx)
    assert (P.download(idx)[idx] == coef[idx]).all()
    # scale up by the special primes and back down is the identity (addPrimesAndScale + scaleDownToSet)
    S = ch.ctxt
    E.ntt_fwd([P], S)
    before = P.download(S)
    E.add_primes_and_scale([P], S, ch.special)
    E.scale_down([P], sorted(S + ch.special), S, p)
    assert (P.download(S)[S] == before[S]).all()


def test_thinboot_ring_short_division_plan_on_the_simulator(sim_lib):
    """The same ring on the CPU simulator, two rows: here the division by Phi_m runs at cyclic length L2 = phi(m) = 2^14 (a quarter
    of the chirp length 2^16), so Phi_m's leading coefficient folds onto its constant term and the dividend's coefficients
    k + L2 < m are folded into the remainder -- the case the small rings of CFGS only reach with m = 1285."""
    m = 21845
    ch = po.build_mod_chain(m, 2, 1, 580, 2)
    E = Engine(m, ch.primes, None, ch.digits, ch.special, lib=sim_lib)
    idx = [ch.ctxt[0], ch.special[-1]]
    rng = np.random.default_rng(11)
    coef = np.zeros((len(ch.primes), ch.phim), dtype=np.uint64)
    for i in idx:
        coef[i] = rng.integers(0, ch.primes[i], size=ch.phim, dtype=np.uint64)
    P = E.poly(coef, idx)
    E.ntt_fwd([P], idx)
    got = P.download(idx)
    rep = po.zms_rep(m)
    i = idx[0]
    q, zeta = ch.primes[i], E.psis[i] * E.psis[i] % ch.primes[i]
    f = [int(v) for v in coef[i]]
    for j in (0, 5, ch.phim - 1):
        x, acc = pow(zeta, rep[j], q), 0
        for cf in reversed(f):
            acc = (acc * x + cf) % q
        assert int(got[i][j]) == acc
    E.ntt_inv([P], idx)
    assert (P.download(idx)[idx] == coef[idx]).all()


# ---- powerful basis and the recryption mod-switch ----

@pytest.mark.parametrize("m,mvec,p,bits", [(105, [3, 5, 7], 2, 100), (45, [9, 5], 2, 100), (45, [5, 9], 7, 100), (64, None, 3, 100), (25, None, 2, 80)])
def test_powerful_basis_and_raw_mod_switch(lib, m, mvec, p, bits):
    """PowerfulDCRT::dcrtToPowerful (src/powerful.cpp:393-410) and Ctxt::rawModSwitch (src/Ctxt.cpp:2949-3046) against
    the oracle's big-integer restatement: several prime-power factorisations (both orders), a prime power and a power of
    two (trivial powerful basis), random rows plus coefficients on the rounding boundaries."""
    if m & (m - 1):
        ch, roots, E = setup(lib, (m, p, 1, bits, 2))
    else:
        from common import make
        ch, psis, O, E = make(lib, m, p, 1, bits, 2)
    S = ch.ctxt
    Q = ch.product(S)
    n = ch.phim
    rng = np.random.default_rng(17)
    if mvec is not None:
        E.set_powerful(mvec)
    factors, to_poly = E.powerful_info()
    ix = po.PowerfulIndexes(factors) if len(factors) > 1 else None
    if mvec is not None:
        assert factors == mvec
    if ix is not None:
        assert [int(v) for v in to_poly] == [ix.cube_to_poly[ix.short_to_long[i]] for i in range(n)]
    # a polynomial with chosen balanced coefficients: random, 0, +-1, +-(Q-1)/2, and values next to multiples of Q/q
    q = 2 ** 10 + 1 if p % 2 == 0 else p ** 3 + 1
    p2r = p
    coeffs = [int(rng.integers(-(1 << 62), 1 << 62)) * int(rng.integers(1, 1 << 30)) % Q for _ in range(n)]
    coeffs = [po.bal(c, Q) for c in coeffs]
    special = [0, 1, -1, (Q - 1) // 2, -((Q - 1) // 2), Q // q, Q // q + 1, -(Q // q), (Q // (2 * q)), (Q // (2 * q)) + 1, 3 * Q // (2 * q), 3 * Q // (2 * q) + 1]
    for i, v in enumerate(special[:n]):
        coeffs[i] = po.bal(v, Q)
    X = E.poly()
    L = len(S) + 1
    E.from_limbs([X], S, [orc.ints_to_limbs(coeffs, L)])
    assert orc.limbs_to_ints(E.to_poly(X, S)) == coeffs
    want_pw = [po.bal(c, Q) for c in (po.poly_to_powerful(ix, coeffs) if ix is not None else coeffs)]
    assert orc.limbs_to_ints(E.dcrt_to_powerful(X, S)) == want_pw
    got = [int(v) for v in E.raw_mod_switch(X, S, q, p2r)]
    # the oracle returns the polynomial-basis result; the ABI the powerful-basis one: convert ours with the engine's map
    want_poly = po.raw_mod_switch(ix, [coeffs], Q, q, p2r)[0]
    ours_poly = po.powerful_to_poly(ix, got) if ix is not None else got
    assert ours_poly == want_poly
    # defining property of the switch (src/Ctxt.cpp:3011-3016): x = c*q*Q^-1 (mod p^r) and |c*q/Q - x| <= p^r/2 (+ one q wrap)
    qinv = q * pow(Q, -1, p2r) % p2r
    for c, x in zip(want_pw, got):
        assert (x - c * qinv) % p2r == 0 or (x + q - c * qinv) % p2r == 0 or (x - q - c * qinv) % p2r == 0
    with pytest.raises(HbError):
        E.raw_mod_switch(X, S, p2r * 5, p2r)     # q not coprime to the plaintext space
