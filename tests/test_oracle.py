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


# ---------------------------------------------------------------------------------------------
# Pinning on key-switching matrices written by a real HElib build (reference fixture, m = 12).

def _iotest_cases():
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "helib_iotest_m12.json")))
    return G["cases"]


def _fixture_chain(case):
    m, primes = case["m"], case["primes"]
    special = case["special"]
    ctxt = [i for i in range(len(primes)) if i not in special]
    ch = po.Chain(m=m, p=case["p"], r=case["r"], phim=po.euler_phi(m), primes=list(primes), small=[], ctxt=ctxt, special=list(special),
                  digits=[list(d) for d in case["digits"]])
    roots = [po.cmod_root(q, m) for q in primes]
    return ch, roots


def _regenerate_a(case, ch, W):
    """The a_i of a key-switching matrix: SetSeed(prgSeed), then a[i].randomize() for i = 0..n-1 over ctxt|special primes
    (src/keys.cpp:1199-1204, src/DoubleCRT.cpp:1258-1378) through the restated NTL stream."""
    import ntl_prg
    st = ntl_prg.set_seed(int(W["prg_seed"]))
    full = sorted(ch.ctxt + ch.special)
    return [po.randomize_rows(ch, full, st.get) for _ in range(W["n"])]


def test_ntl_stream_randomize_and_keyswitch_matrix_pinned_by_reference_fixture():
    """SURVEY 8a rows 14 and 20.  The reference's I/O fixtures store four key-switching matrices W[s^r(X^t) -> s] with the b_i rows
    a real HElib wrote and the 256-bit prgSeed of their a_i.  Regenerating a_i with the oracle's restatement of NTL's stream
    (HMAC-SHA256 key derivation + ChaCha20) and of DoubleCRT::randomize, then removing P*prod_{j<i}Q_j*s^r(X^t) and a_i*s,
    must leave ONE small multiple of p on every prime -- a single wrong byte of the stream, a different consumption pattern, a
    different automorphism index map or scaling factor leaves uniform noise instead."""
    for case in _iotest_cases():
        ch, roots = _fixture_chain(case)
        m, p = ch.m, ch.p
        rep = po.zms_rep(m)
        pos = {r: j for j, r in enumerate(rep)}
        P = ch.product(ch.special)
        sk = {int(i): row for i, row in case["secret_key"].items()}
        assert len(case["ksw"]) == 4
        for W in case["ksw"]:
            (spow, xpow, from_id), to_id = W["from"], W["to"]
            assert (from_id, to_id, W["ptxt_space"], W["n"]) == (0, 0, p, len(ch.digits))
            a = _regenerate_a(case, ch, W)
            for d in range(W["n"]):
                noise = None
                for i, q in enumerate(ch.primes):
                    s_from = [sk[i][pos[rep[j] * xpow % m]] for j in range(ch.phim)]          # s(X^t): DoubleCRT::automorph
                    s_from = [pow(x, spow, q) for x in s_from]                                # s^r(X^t): DoubleCRT::Exp
                    fac = P * ch.product([k for j in range(d) for k in ch.digits[j]]) % q    # src/keys.cpp:1239-1242
                    b = W["b"][d][str(i)]
                    rr = [(bb + aa * s - fac * sf) % q for bb, aa, s, sf in zip(b, a[d][i], sk[i], s_from)]
                    coef = [po.bal(c, q) for c in po.gen_ifft(rr, q, m, roots[i])]
                    assert all(c % p == 0 and abs(c) < 200 * p for c in coef), (W["from"], d, i, coef)
                    assert noise is None or coef == noise
                    noise = coef
                assert any(noise)


def test_keyswitch_path_semantics_pinned_by_reference_evk():
    """SURVEY 8a rows 7, 10, 11, 13, 18.  Key-switch with matrices produced by real HElib: a ciphertext of the fixture's key
    (its public encryption key + a plaintext) is squared (tensorProduct), relinearised with the fixture's W[s^2 -> s] through the
    oracle's breakIntoDigits / keySwitchDigits / addPrimesAndScale / scaleDownToSet, and rotated with W[s(X^5) -> s]; decrypting
    with the fixture's secret key (toPoly, balanced) must give mu^2 and mu(X^5).  The oracle's digit decomposition, its factors
    P*prod Q_j, the mod-down rounding and the balanced CRT must therefore be the ones the reference's matrices encode."""
    case = _iotest_cases()[0]
    ch, roots = _fixture_chain(case)
    m, p, n = ch.m, ch.p, ch.phim
    S = list(case["pk_prime_set"])
    assert S == ch.ctxt
    sk = po.PyDCRT(ch, roots, {int(i): list(r) for i, r in case["secret_key"].items()})
    phi = po.cyclotomic_poly(m)

    def polymul_mod_phi(f, g):
        out = [0] * (2 * n)
        for i, a in enumerate(f):
            for j, b in enumerate(g):
                out[i + j] += a * b
        return [c % p for c in po._poly_rem(out, phi)[:n]] if hasattr(po, "_poly_rem") else None

    def decrypt(c0, c1, idxs):
        s = po.PyDCRT(ch, roots, {i: sk.rows[i] for i in idxs})
        t = po.PyDCRT(ch, roots, {i: list(c1.rows[i]) for i in idxs}).mul(s)
        t.add(po.PyDCRT(ch, roots, {i: list(c0.rows[i]) for i in idxs}))
        return [c % p for c in t.to_poly()]

    mu = [3, 1, 0, 5]
    c0 = po.PyDCRT(ch, roots, {int(i): list(r) for i, r in case["pk_c0"].items()})
    c1 = po.PyDCRT(ch, roots, {int(i): list(r) for i, r in case["pk_c1"].items()})
    c0.add(po.PyDCRT.from_poly(ch, roots, mu, S))
    assert decrypt(c0, c1, S) == mu

    def evk(spow, xpow):
        W = [w for w in case["ksw"] if w["from"][:2] == [spow, xpow]][0]
        a = _regenerate_a(case, ch, W)
        A = [po.PyDCRT(ch, roots, a[d]) for d in range(W["n"])]
        B = [po.PyDCRT(ch, roots, {int(i): list(r) for i, r in W["b"][d].items()}) for d in range(W["n"])]
        return A, B

    # --- square + relinearise (src/Ctxt.cpp:1563-1608, 720-786) + drop the special primes (src/Ctxt.cpp:589-593)
    d0 = c0.copy().mul(c0)
    d1 = c0.copy().mul(c1); d1.add(d1.copy())
    d2 = c1.copy().mul(c1)
    A, B = evk(2, 1)
    digits, _ = d2.break_into_digits()
    k0, k1 = po.key_switch_digits(digits, A, B)
    r0 = d0.copy().add_primes_and_scale(ch.special).add(k0)
    r1 = d1.copy().add_primes_and_scale(ch.special).add(k1)
    r0.scale_down_to_set(S, p); r1.scale_down_to_set(S, p)
    mu2 = [0] * (2 * n)
    for i, a_ in enumerate(mu):
        for j, b_ in enumerate(mu):
            mu2[i + j] += a_ * b_
    for k in range(2 * n - 1, n - 1, -1):      # reduce modulo Phi_12 = X^4 - X^2 + 1
        ck = mu2[k]
        for j, pj in enumerate(phi):
            mu2[k - (len(phi) - 1) + j] -= ck * pj
    assert decrypt(r0, r1, S) == [c % p for c in mu2[:n]]

    # --- rotate: sigma_5 on both parts, then switch the s(X^5) part back to s (src/Ctxt.cpp:2437-2515)
    a0, a1 = c0.copy().automorph(5), c1.copy().automorph(5)
    A, B = evk(1, 5)
    digits, _ = a1.break_into_digits()
    k0, k1 = po.key_switch_digits(digits, A, B)
    r0 = a0.copy().add_primes_and_scale(ch.special).add(k0)
    r1 = k1
    r0.scale_down_to_set(S, p); r1.scale_down_to_set(S, p)
    mu5 = [0] * (5 * n)
    for i, a_ in enumerate(mu):
        mu5[5 * i] += a_
    for k in range(len(mu5) - 1, n - 1, -1):
        ck = mu5[k]
        if ck:
            for j, pj in enumerate(phi):
                mu5[k - (len(phi) - 1) + j] -= ck * pj
            mu5[k] = 0
    assert decrypt(r0, r1, S) == [c % p for c in mu5[:n]]


def test_ntl_fft_root_restatement_is_a_valid_deterministic_root():
    """The root derivation of NTL's zz_pContext(INIT_USER_FFT, q) under HElib's fixed seed (src/CModulus.cpp:93-98), restated on top
    of the PINNED stream: deterministic, a primitive m-th root for every chain prime, consistent across m (RootTable entries are
    squares of each other), and usable as the engine/oracle root -- rows computed with it satisfy row[j] = f(psi^(2j+1))."""
    ch = po.build_mod_chain(1 << 17, -1, 1, 230, 2)
    for q in ch.primes:
        psi = po.ntl_fft_root(q, 1 << 17)
        assert psi == po.ntl_fft_root(q, 1 << 17)
        assert pow(psi, 1 << 16, q) == q - 1
        assert po.ntl_fft_root(q, 1 << 16) == psi * psi % q            # RootTable[0][16] = RootTable[0][17]^2
    q = po.build_mod_chain(64, 257, 1, 120, 2).primes[-1]
    psi = po.ntl_fft_root(q, 64)
    f = [3, 1, 4, 1, 5, 9, 2, 6] + [0] * 24
    row = po.ntt_fwd(f, q, psi)
    for j in (0, 1, 7, 31):
        assert row[j] == sum(c * pow(psi, (2 * j + 1) * i, q) for i, c in enumerate(f)) % q
