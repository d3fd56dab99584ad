"""Pins the C++ oracle (oracle/oracle.cpp) against the Python big-int oracle and against
psi-independent algebraic invariants.  CPU only."""
import json
import os

import numpy as np
import pytest

import orc
import pyoracle as po
from common import chain

CFGS = [(16, 17, 1, 80, 2), (64, 257, 1, 120, 2), (256, 3, 2, 150, 3), (128, -1, 1, 100, 2)]


def setup(cfg):
    ch, psis = chain(*cfg)
    O = orc.Oracle(ch.phim, ch.m, ch.primes, psis, ch.digits, ch.special, nthreads=2)
    return ch, psis, O


def pyd(ch, psis, data, idx):
    return po.PyDCRT(ch, psis, {i: [int(x) for x in data[i]] for i in idx})


@pytest.mark.parametrize("cfg", CFGS)
def test_transform_matches_definition_and_roundtrips(cfg):
    """row[j] = f(psi^(2j+1)) (src/CModulus.cpp:392-426); iFFT(FFT(f)) = f (tests/TestHEXL.cpp:189-218)."""
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(0)
    allp = list(range(len(ch.primes)))
    co = O.random(rng, allp)
    ev = co.copy(); O.ntt_fwd_rows(ev, allp)
    for i in allp[:3]:
        q, psi = ch.primes[i], psis[i]
        f = [int(x) for x in co[i]]
        for j in (0, 1, ch.phim - 1):
            x = pow(psi, 2 * j + 1, q)
            assert int(ev[i][j]) == sum(c * pow(x, k, q) for k, c in enumerate(f)) % q
        assert list(ev[i]) == po.ntt_fwd(f, q, psi)
    back = ev.copy(); O.ntt_inv_rows(back, allp)
    assert (back == co).all()


@pytest.mark.parametrize("cfg", CFGS[:3])
def test_convolution_theorem_vs_schoolbook(cfg):
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(1)
    S = ch.ctxt
    n = ch.phim
    f = [int(x) for x in rng.integers(-50, 50, n)]
    g = [int(x) for x in rng.integers(-50, 50, n)]
    fg = po.negacyclic_mul_schoolbook(f, g)
    F = po.PyDCRT.from_poly(ch, psis, f, S)
    G = po.PyDCRT.from_poly(ch, psis, g, S)
    a, b = O.zeros(), O.zeros()
    for i in S:
        a[i] = np.array(F.rows[i], dtype=np.uint64); b[i] = np.array(G.rows[i], dtype=np.uint64)
    O.pointwise("mul", a, b, S)
    got = orc.limbs_to_ints(O.to_poly(a, S))
    assert got == fg


@pytest.mark.parametrize("cfg", CFGS)
def test_to_poly_add_primes_match_python(cfg):
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(2)
    cur = ch.digits[0]
    add = [i for i in ch.ctxt + ch.special if i not in cur]
    x = O.random(rng, cur)
    P = pyd(ch, psis, x, cur)
    assert orc.limbs_to_ints(O.to_poly(x, cur)) == P.to_poly()
    assert orc.limbs_to_ints(O.to_poly(x, cur, positive=True)) == P.to_poly(positive=True)
    y = x.copy(); O.add_primes(y, cur, add)
    P.add_primes(add)
    for i in cur + add:
        assert list(y[i]) == P.rows[i]
    # addPrimes then removePrimes is the identity on the old rows; new rows == toPoly mod q
    assert (y[cur] == x[cur]).all()


@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("p", [1, 2, 4, 17, 257])
def test_scale_down_postconditions(cfg, p):
    """P*out + delta == in exactly, delta == 0 mod p, |delta| <= P*p/2."""
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(3)
    cur = ch.ctxt + ch.special
    keep = ch.ctxt
    x = O.random(rng, cur)
    before = pyd(ch, psis, x, cur).to_poly()
    y = x.copy()
    delta = orc.limbs_to_ints(O.scale_down(y, cur, keep, p, want_delta=True))
    after = pyd(ch, psis, y, keep).to_poly()
    Pd = ch.product(ch.special)
    Qk = ch.product(keep)
    for b, d, a in zip(before, delta, after):
        assert d % p == 0 and (b - d) % Pd == 0 and abs(d) <= Pd * max(p, 1) // 2 + Pd
        assert po.bal((b - d) // Pd, Qk) == a
    ref = pyd(ch, psis, x, cur); dref = ref.scale_down_to_set(keep, p)
    assert dref == delta
    for i in keep:
        assert list(y[i]) == ref.rows[i]


@pytest.mark.parametrize("cfg", CFGS)
def test_digits_are_balanced_mixed_radix(cfg):
    """sum_i D_i * prod_{j<i} Q_j == bal(x mod Q), |D_i| <= (Q_i-1)/2."""
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(4)
    S = ch.ctxt
    x = O.random(rng, S)
    digs, polys = O.break_into_digits(x, S, want_polys=True)
    xs = pyd(ch, psis, x, S).to_poly()
    Q = ch.product(S)
    acc = [0] * ch.phim
    scale = 1
    for i in range(digs.shape[0]):
        Qi = ch.product(ch.digits[i])
        Ei = orc.limbs_to_ints(polys[i])
        assert all(abs(e) <= (Qi - 1) // 2 for e in Ei)
        acc = [a + e * scale for a, e in zip(acc, Ei)]
        scale *= Qi
    assert [po.bal(a, Q) for a in acc] == xs
    pd, pp = pyd(ch, psis, x, S).break_into_digits()
    for i, d in enumerate(pd):
        assert orc.limbs_to_ints(polys[i]) == pp[i]
        for r in d.index_set:
            assert list(digs[i][r]) == d.rows[r]


@pytest.mark.parametrize("cfg", CFGS[:2])
def test_keyswitch_and_automorph_match_python(cfg):
    ch, psis, O = setup(cfg)
    rng = np.random.default_rng(5)
    S = ch.ctxt
    full = S + ch.special
    x = O.random(rng, S)
    digs = O.break_into_digits(x, S)
    nd = digs.shape[0]
    ea = np.stack([O.random(rng, full) for _ in range(nd)])
    eb = np.stack([O.random(rng, full) for _ in range(nd)])
    o0, o1 = O.zeros(), O.zeros()
    O.keyswitch_digits(digs, full, ea, eb, o0, o1)
    pd, _ = pyd(ch, psis, x, S).break_into_digits()
    r0, r1 = po.key_switch_digits(pd, [pyd(ch, psis, ea[i], full) for i in range(nd)], [pyd(ch, psis, eb[i], full) for i in range(nd)])
    for i in full:
        assert list(o0[i]) == r0.rows[i] and list(o1[i]) == r1.rows[i]
    for k in (3, ch.m - 1):
        y = x.copy(); O.automorph(y, S, k)
        ref = pyd(ch, psis, x, S).automorph(k)
        for i in S:
            assert list(y[i]) == ref.rows[i]
    # automorph is F(X) -> F(X^k): check on coefficients for one row
    f = [int(v) for v in rng.integers(0, 100, ch.phim)]
    F = po.PyDCRT.from_poly(ch, psis, f, S[:1]).automorph(3)
    g = [0] * ch.phim
    for e, c in enumerate(f):
        ee = (3 * e) % ch.m
        if ee < ch.phim:
            g[ee] += c
        else:
            g[ee - ch.phim] -= c
    assert F.to_poly() == g


def test_golden_vectors():
    """tests/golden/*.json were produced by tests/golden/make_golden.py from the Python big-int
    restatement (the reference has no golden vectors for this path, and cannot be built here)."""
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".json"))
    assert files
    for fn in files:
        G = json.load(open(os.path.join(gdir, fn)))
        if "params" not in G:      # other fixtures (e.g. the reference's I/O rows) have their own tests
            continue
        m, p, r, bits, c = G["params"]
        ch, psis = chain(m, p, r, bits, c)
        assert ch.primes == G["primes"] and psis == G["psis"] and ch.digits == G["digits"]
        O = orc.Oracle(ch.phim, ch.m, ch.primes, psis, ch.digits, ch.special, nthreads=1)
        S = ch.ctxt
        x = O.zeros()
        for i in S:
            x[i] = np.array(G["x"][str(i)], dtype=np.uint64)
        assert orc.limbs_to_ints(O.to_poly(x, S)) == G["to_poly"]
        digs = O.break_into_digits(x, S)
        for d, ref in enumerate(G["digits_rows"]):
            for i, row in ref.items():
                assert list(digs[d][int(i)]) == row
        cur = S + ch.special
        y = O.zeros()
        for i in cur:
            y[i] = np.array(G["y"][str(i)], dtype=np.uint64)
        O.scale_down(y, cur, S, G["ptxt_space"])
        for i in S:
            assert list(y[i]) == G["scale_down_rows"][str(i)]


def test_oracle_general_m_conventions_pinned_by_reference_fixture():
    """The reference's own I/O fixtures (tests/test_resources/iotest_ascii*.txt, m=12, p=7) hold evaluation-form rows written
    by a real HElib build: the secret key over five primes and the public encryption key (c0, c1) over three.  With the
    oracle's root (FindPrimitiveRoot restatement) and row order over Z_m^*, the secret-key rows must invert to ONE ternary
    polynomial for every prime, and c0 + c1*s to ONE small multiple of p -- any other root or ordering gives noise."""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "helib_iotest_m12.json")))
    assert len(G["cases"]) == 2
    for case in G["cases"]:
        m, p, primes = case["m"], case["p"], case["primes"]
        assert all((q - 1) % m == 0 and po.is_prime(q) for q in primes)
        sk_coef = None
        for i, q in enumerate(primes):
            root = po.cmod_root(q, m)
            coef = [po.bal(c, q) for c in po.gen_ifft(case["secret_key"][str(i)], q, m, root)]
            assert all(c in (-1, 0, 1) for c in coef), (i, coef)
            assert sk_coef is None or coef == sk_coef
            sk_coef = coef
            # and forward again: the oracle's rows of the recovered polynomial are the fixture's rows
            assert po.gen_fft([c % q for c in coef], q, m, root) == case["secret_key"][str(i)]
        assert any(sk_coef)
        noise = None
        for i in case["pk_prime_set"]:
            q = primes[i]
            root = po.cmod_root(q, m)
            d = [(a + b * s) % q for a, b, s in zip(case["pk_c0"][str(i)], case["pk_c1"][str(i)], case["secret_key"][str(i)])]
            coef = [po.bal(c, q) for c in po.gen_ifft(d, q, m, root)]
            assert all(c % p == 0 and abs(c) < 200 * p for c in coef), (i, coef)      # RLWE1: c0 + c1*s = p*e, e ~ 3.2*sqrt(m)
            assert noise is None or coef == noise
            noise = coef
