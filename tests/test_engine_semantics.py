"""Engine through the C ABI: golden fixtures, exact-CRT fallback, reference error behaviour and
an end-to-end BGV semantic check (encrypt -> multiply -> relinearise -> mod-down -> decrypt).
Each test runs on the CPU kernel-logic simulator (not gpu) and on the real CUDA library (-m gpu)."""
import json
import os
import random

import numpy as np
import pytest

import orc
import pyoracle as po
from common import chain, make, rows_equal
from helib_b200 import HbError


def backends():
    return [pytest.param("sim", id="sim"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=backends())
def lib(request):
    return request.getfixturevalue("sim_lib" if request.param == "sim" else "cuda_lib")


def test_golden_vectors_through_engine(lib):
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for fn in sorted(f for f in os.listdir(gdir) if f.endswith(".json")):
        G = json.load(open(os.path.join(gdir, fn)))
        if "params" not in G:      # other fixtures (e.g. the reference's I/O rows) have their own tests
            continue
        ch, psis, O, E = make(lib, *G["params"])
        S = ch.ctxt
        Sp = sorted(S + ch.special)
        x = O.zeros()
        for i in S:
            x[i] = np.array(G["x"][str(i)], dtype=np.uint64)
        P = E.poly(x, S)
        assert orc.limbs_to_ints(E.to_poly(P, S)) == G["to_poly"]
        digs = E.break_into_digits([P], S)[0]
        for d, ref in enumerate(G["digits_rows"]):
            got = digs[d].download(Sp)
            for i, row in ref.items():
                assert list(got[int(i)]) == row
        y = O.zeros()
        for i in Sp:
            y[i] = np.array(G["y"][str(i)], dtype=np.uint64)
        Y = E.poly(y, Sp)
        E.scale_down([Y], Sp, S, G["ptxt_space"])
        got = Y.download(S)
        for i in S:
            assert list(got[i]) == G["scale_down_rows"][str(i)]


@pytest.mark.parametrize("cfg", [(64, 257, 1, 120, 2), (4096, 3, 1, 110, 2), (1 << 17, 257, 1, 230, 2)])
def test_exact_crt_fallback_on_boundary_values(lib, cfg):
    """Coefficients whose CRT value sits on the rounding boundary (+-(Q-1)/2, 0, +-1) force the
    exact multi-limb path; results must still equal the oracle bit for bit."""
    ch, psis, O, E = make(lib, *cfg)
    n = ch.phim
    cur = ch.ctxt + ch.special
    keep = ch.ctxt[:1]
    drop = [i for i in cur if i not in keep]
    Pd = ch.product(drop)
    assert Pd.bit_length() > 70
    half = (Pd - 1) // 2
    rnd = random.Random(7)
    Qall = ch.product(cur)
    specials = [half, -half, 0, 1, -1, half - 1, -(half - 1), half + Pd, -half + 5 * Pd]
    coeffs = [specials[k % len(specials)] + Pd * rnd.randrange(1 << 40) * (k % 3 == 2) if k < 64 else rnd.randrange(Qall) for k in range(n)]
    limbs = orc.ints_to_limbs(coeffs, len(cur) + 1)
    x = O.zeros()
    O.fft_bigpoly(limbs, cur, x)
    for p in (1, 2, ch.p ** ch.r):
        P = E.poly(x, cur)
        E.reset_stats()
        E.scale_down([P], cur, keep, p)
        ref = x.copy(); O.scale_down(ref, cur, keep, p)
        assert rows_equal(P.download(keep), ref, keep), p
        assert E.stats()["exact_fallbacks"] > 0
    # toPoly of the special rows reproduces the planted boundary values exactly
    P = E.poly(x, cur)
    got = orc.limbs_to_ints(E.to_poly(P, drop))
    assert got[:9] == [po.bal(s, Pd) for s in specials]


def test_reference_error_behaviour(lib):
    """Index-set preconditions raise like the reference (RuntimeError -> HB_ERR_INDEX_SET = -2;
    InvalidArgument/LogicError -> HB_ERR_BAD_ARG = -1)."""
    ch, psis, O, E = make(lib, 64, 257, 1, 120, 2)
    S = ch.ctxt
    P, Q = E.poly(), E.poly()
    with pytest.raises(HbError) as ei:          # src/DoubleCRT.cpp:574-575
        E.add_primes([P], S, [S[0]])
    assert ei.value.code == -2
    with pytest.raises(HbError) as ei:          # src/DoubleCRT.cpp:497-498
        E.break_into_digits([P], S + ch.special)
    assert ei.value.code == -2
    with pytest.raises(HbError) as ei:          # src/DoubleCRT.cpp:1165-1167
        E.automorph([Q], [P], S, 2)
    assert ei.value.code == -2
    with pytest.raises(HbError) as ei:          # src/DoubleCRT.cpp:1474-1476
        E.scale_down([P], S, [ch.special[0]], 1)
    assert ei.value.code == -2
    with pytest.raises(HbError) as ei:
        E.ntt_fwd([P], [len(ch.primes)])
    assert ei.value.code == -1
    with pytest.raises(HbError) as ei:          # ptxtSpace >= 1 (src/DoubleCRT.cpp:1472)
        E.scale_down([P], S + ch.special, S, 0)
    assert ei.value.code == -1
    with pytest.raises(HbError) as ei:          # the common set of a multiply is a subset of both operands' sets
        E.mul_relin_moddown([P], [Q], [E.poly()], [E.poly()], S[:-1], S, 1, [E.poly()] * len(ch.digits), [E.poly()] * len(ch.digits))
    assert ei.value.code == -2
    # nothing-to-do cases return quietly (src/DoubleCRT.cpp:569-572, 1468-1470)
    E.add_primes([P], S, [])
    E.scale_down([P], S, S, 1)
    # two contexts do not mix (src/DoubleCRT.cpp:222-223)
    ch2, _, _, E2 = make(lib, 64, 257, 1, 120, 2)
    with pytest.raises(HbError) as ei:
        E.pointwise("add", [P], [E2.poly()], S)
    assert ei.value.code == -2


def sample_small(rng, n, kind):
    if kind == "ternary":
        return [int(x) for x in rng.integers(-1, 2, n)]
    return [int(round(x)) for x in rng.normal(0, 3.2, n)]   # include/helib/Context.h:1080 (stdev 3.2)


def dcrt_of(O, ch, coeffs, idx):
    """Small signed polynomial -> evaluation rows on idx (DoubleCRT(poly, context, s), src/DoubleCRT.cpp:68-85)."""
    d = O.zeros()
    for i in idx:
        d[i] = np.array([c % ch.primes[i] for c in coeffs], dtype=np.uint64)
    O.ntt_fwd_rows(d, idx)
    return d


@pytest.mark.parametrize("cfg", [(128, 257, 1, 150, 2), (4096, 17, 1, 160, 3)])
def test_bgv_multiply_decrypts_to_product(lib, cfg):
    """decrypt(mul + relin + mod-down (E(a), E(b))) == a*b (BGV, exact).
    Key material restated from RLWE1 / GenKeySWmatrix (src/keys.cpp:40-72,1159-1256):
    b_i = p*e_i - a_i*s + P*(prod_{j<i} Q_j)*s^2."""
    ch, psis, O, E = make(lib, *cfg)
    p, n = ch.p, ch.phim
    rng = np.random.default_rng(11)
    full = ch.ctxt + ch.special
    s = sample_small(rng, n, "ternary")
    S_dcrt = dcrt_of(O, ch, s, full)
    s2 = S_dcrt.copy(); O.pointwise("mul", s2, S_dcrt, full)

    def encrypt(msg, idx):
        a = O.random(rng, idx)
        c0 = dcrt_of(O, ch, [p * e + m_ for e, m_ in zip(sample_small(rng, n, "gauss"), msg)], idx)
        t = a.copy(); O.pointwise("mul", t, S_dcrt, idx)
        O.pointwise("sub", c0, t, idx)          # c0 = p*e + m - a*s
        return c0, a

    # key-switching matrix s^2 -> s
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = []
    from_key = s2.copy()
    O.scale_by_primes(from_key, full, ch.special)            # P * s^2
    for i in range(nd):
        b = dcrt_of(O, ch, [p * e for e in sample_small(rng, n, "gauss")], full)
        t = evk_a[i].copy(); O.pointwise("mul", t, S_dcrt, full)
        O.pointwise("sub", b, t, full)
        O.pointwise("add", b, from_key, full)
        evk_b.append(b)
        O.scale_by_primes(from_key, full, ch.digits[i])       # *= prod(digit i)
    evk_b = np.stack(evk_b)

    ma = [int(x) for x in rng.integers(0, p, n)]
    mb = [int(x) for x in rng.integers(0, p, n)]
    S_in = ch.ctxt
    S = ch.ctxt[:-1]
    a0, a1 = encrypt(ma, S_in)
    b0, b1 = encrypt(mb, S_in)
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    A0, A1, B0, B1 = E.poly(a0, S_in), E.poly(a1, S_in), E.poly(b0, S_in), E.poly(b1, S_in)
    E.mul_relin_moddown([A0], [A1], [B0], [B1], S_in, S, p, EA, EB)
    # decrypt: (c0 + c1*s) over S, balanced, mod p  (src/keys.cpp:1327-1400)
    E.pointwise("mul", [A1], [E.poly(S_dcrt, S)], S)
    E.pointwise("add", [A0], [A1], S)
    dec = orc.limbs_to_ints(E.to_poly(A0, S))
    Q = ch.product(S)
    assert max(abs(v) for v in dec) < Q // 4, "noise overflow: parameters too tight for the test"
    # each operand's mod-down by the dropped prime multiplies its plaintext by q_drop^-1 mod p
    qd = ch.primes[S_in[-1]]
    f = pow(pow(qd, -1, p), 2, p)
    expect = [v * f % p for v in po.negacyclic_mul_schoolbook(ma, mb, p)] if n <= 256 else None
    got = [v % p for v in dec]
    if expect is not None:
        assert got == expect
    # and bit-exact agreement with the oracle's own run of the same circuit
    parts = [x.copy() for x in (a0, a1, b0, b1)]
    for x in parts:
        O.scale_down(x, S_in, S, p)
    t0, t1, t2 = O.tensor(*parts, S)
    r0, r1 = O.relinearize(t0, t1, t2, S, evk_a, evk_b)
    Sp = sorted(S + ch.special)
    O.scale_down(r0, Sp, S, p); O.scale_down(r1, Sp, S, p)
    O.pointwise("mul", r1, S_dcrt, S); O.pointwise("add", r0, r1, S)
    assert orc.limbs_to_ints(O.to_poly(r0, S)) == dec


@pytest.mark.parametrize("cfg", [(64, 257, 1, 120, 2), (4096, 17, 1, 160, 3), (1 << 17, 257, 1, 230, 2)])
def test_embedding_norms_match_restated_reference(lib, cfg):
    """Noise metadata: FP64 canonical-embedding norms returned next to the integer
    results.  Floating point: relative tolerance 1e-9 against the numpy restatement of
    embeddingLargestCoeff (src/norms.cpp:204-261,443-485); the integer rows stay bit-exact."""
    import math
    ch, psis, O, E = make(lib, *cfg)
    p = ch.p ** ch.r
    rng = np.random.default_rng(21)
    S = ch.ctxt
    Sp = sorted(S + ch.special)
    x = O.random(rng, S)
    P = E.poly(x, S)
    digs, lognorms = E.break_into_digits_norm([P], S)
    ref_d, polys = O.break_into_digits(x, S, want_polys=True)
    for i, D in enumerate(digs[0]):
        assert rows_equal(D.download(Sp), ref_d[i], Sp)
        mant, shift = po.embedding_largest_coeff(orc.limbs_to_ints(polys[i]), ch.m)
        ref_log = math.log(mant) + shift * math.log(2.0)
        assert abs(lognorms[0, i] - ref_log) <= 1e-9 * abs(ref_log) + 1e-9
    y = O.random(rng, Sp)
    Y = E.poly(y, Sp)
    norms = E.scale_down_norm([Y], Sp, S, p)
    ref = y.copy()
    delta = orc.limbs_to_ints(O.scale_down(ref, Sp, S, p, want_delta=True))
    assert rows_equal(Y.download(S), ref, S)
    Pd = ch.product(ch.special)
    from fractions import Fraction
    fdelta = [float(Fraction(d, Pd)) for d in delta]       # Ctxt::modDownToSet: fdelta = delta / diffProd (src/Ctxt.cpp:482-485)
    n = ch.phim
    k = np.arange(n)
    vals = np.fft.ifft(np.array(fdelta) * np.exp(1j * np.pi * k / n)) * n
    want = float(np.max(np.abs(vals)))
    assert abs(norms[0] - want) <= 1e-9 * want
    assert want <= p / 2.0 * n + 1


def test_wire_format_matches_reference_layout(lib):
    """DoubleCRT::writeTo byte layout (src/DoubleCRT.cpp:1530-1541, src/IndexSet.cpp:288-297,
    src/binio.cpp:103-122) built by hand with struct.pack; round trip; corrupt data is rejected."""
    import struct
    ch, psis, O, E = make(lib, 64, 257, 1, 120, 2)
    rng = np.random.default_rng(41)
    S = ch.ctxt
    x = O.random(rng, S)
    P = E.poly(x, S)
    blob = P.serialize(list(reversed(S)))            # any order in -> ascending index order out
    want = struct.pack("<q", len(S)) + b"".join(struct.pack("<q", i) for i in sorted(S))
    for i in sorted(S):
        want += struct.pack("<ii", ch.phim, 8) + b"".join(struct.pack("<q", int(v)) for v in x[i])
    assert blob == want
    Q = E.poly()
    assert Q.deserialize(blob) == sorted(S)
    assert rows_equal(Q.download(S), x, S)
    bad = bytearray(blob)
    off = 8 + 8 * len(S) + 8
    bad[off:off + 8] = struct.pack("<q", ch.primes[sorted(S)[0]])      # residue == q: out of range
    with pytest.raises(HbError):
        Q.deserialize(bytes(bad))
    with pytest.raises(HbError):
        Q.deserialize(blob[:-5])


# ---- the steps either side of the path (encrypt / decrypt / polynomial -> rows) on the device ----

@pytest.mark.parametrize("cfg", [(64, 257, 1, 120, 2), (4096, 17, 1, 160, 3)])
def test_polynomial_to_rows_on_device(lib, cfg):
    """DoubleCRT(zzX) and DoubleCRT(ZZX) constructors (src/DoubleCRT.cpp:68-105): reduce + NTT on the device from one
    copy of the polynomial; small signed coefficients, full 63-bit ones, and big balanced integers (round trip of
    toPoly), each against the oracle's rows."""
    ch, psis, O, E = make(lib, *cfg)
    n = ch.phim
    full = ch.ctxt + ch.special
    rng = np.random.default_rng(5)
    small = rng.integers(-40, 41, n).astype(np.int64)
    wide = rng.integers(-(1 << 62), 1 << 62, n).astype(np.int64)
    wide[:4] = [-(1 << 63) + 1, (1 << 63) - 1, 0, -1]
    for coeffs in (small, wide):
        P = E.poly()
        E.from_i64([P], full, [coeffs])
        ref = dcrt_of(O, ch, [int(c) for c in coeffs], full)
        assert rows_equal(P.download(full), ref, full)
    # big integers: x -> toPoly -> limbs -> rows must reproduce x on every prime of the set, and extend consistently
    S = ch.ctxt
    x = O.random(rng, S)
    X = E.poly(x, S)
    limbs = E.to_poly(X, S)                      # [N][len(S)+1] two's complement, balanced
    Y = E.poly()
    E.from_limbs([Y], full, [limbs])
    got = Y.download(full)
    assert rows_equal(got, x, S)
    ref = x.copy(); O.add_primes(ref, S, ch.special)     # exact base extension of the same integers
    assert rows_equal(got, ref, ch.special)
    # batched, and a subset of rows only
    A, B = E.poly(), E.poly()
    E.from_i64([A, B], S[:1], [small, wide])
    assert rows_equal(A.download(S[:1]), dcrt_of(O, ch, [int(c) for c in small], S[:1]), S[:1])
    assert rows_equal(B.download(S[:1]), dcrt_of(O, ch, [int(c) for c in wide], S[:1]), S[:1])
    with pytest.raises(HbError):
        E.from_i64([A], [len(ch.primes)], [small])


@pytest.mark.parametrize("cfg", [(128, 257, 1, 150, 2), (64, 2, 1, 120, 2), (256, 3, 2, 150, 2)])
def test_encrypt_decrypt_on_device_match_restated_reference(lib, cfg):
    """PubKey::Encrypt (BGV, src/keys.cpp:381-455) and SecKey::Decrypt (src/keys.cpp:1327-1400) with the sampled
    polynomials given: device rows == oracle rows bit for bit, decrypt(encrypt(m)) == m, and the device's mod-p
    tail equals PolyRed + MulMod of the oracle's big-integer polynomial."""
    ch, psis, O, E = make(lib, *cfg)
    p, n = ch.p ** ch.r, ch.phim
    S = ch.ctxt
    rng = np.random.default_rng(23)
    s = sample_small(rng, n, "ternary")
    sk = po.PyDCRT.from_poly(ch, psis, s, S)
    # public encryption key = RLWE1 sample (src/keys.cpp:40-72): pk1 = a, pk0 = p*e - a*s
    a_rows = {i: [int(x) for x in rng.integers(0, ch.primes[i], n)] for i in S}
    pk1 = po.PyDCRT(ch, psis, a_rows)
    pk0 = po.PyDCRT.from_poly(ch, psis, [p * e for e in sample_small(rng, n, "gauss")], S)
    t = pk1.copy(); t.mul(sk); pk0.sub(t)
    msg = [int(x) for x in rng.integers(0, p, n)]
    r = sample_small(rng, n, "ternary")
    e0, e1 = sample_small(rng, n, "gauss"), sample_small(rng, n, "gauss")
    want = po.encrypt_bgv(ch, psis, pk0, pk1, r, e0, e1, msg, p, S)

    def up(d):
        x = O.zeros()
        for i in S:
            x[i] = np.array(d.rows[i], dtype=np.uint64)
        return E.poly(x, S)
    PK0, PK1, SK = up(pk0), up(pk1), up(sk)
    # device: three polynomials cross the bus as N int64 each; everything else stays in HBM
    fixed = po.balanced_mulmod(msg, ch.product(S) % p, p)
    R, C0, C1 = E.poly(), E.poly(), E.poly()
    E.from_i64([R, C0, C1], S, [r, [p * a + b for a, b in zip(e0, fixed)], [p * a for a in e1]])
    E.muladd([C0, C1], [PK0, PK1], [R, R], S)
    for C, w in ((C0, want[0]), (C1, want[1])):
        got = C.download(S)
        for i in S:
            assert [int(v) for v in got[i]] == w.rows[i]
    # decrypt on the device
    ref_pt, ref_f = po.decrypt_bgv(ch, psis, want, [None, sk], p, 1, S)
    assert ref_pt == msg
    ACC = E.poly()
    E.pointwise("copy", [ACC], [C0], S)
    E.muladd([ACC], [C1], [SK], S)
    assert orc.limbs_to_ints(E.to_poly(ACC, S)) == ref_f
    factor = pow(ch.product(S) % p, -1, p) if p > 2 else 1
    got = E.to_poly_mod_p(ACC, S, p, factor)
    assert [int(v) for v in got] == msg
    # factor 1 == plain PolyRed(abs) of the same integers
    assert [int(v) for v in E.to_poly_mod_p(ACC, S, p, 1)] == [c % p for c in ref_f]
    with pytest.raises(HbError):
        E.to_poly_mod_p(ACC, S, 1, 0)


def test_wire_format_reads_bytes_written_by_the_reference(lib):
    """A DoubleCRT::writeTo record written by a real HElib build (the first ciphertext part inside the reference's
    tests/test_resources/iotest_binLE.bin; tests/golden/make_iotest_fixture.py) deserialises to the rows the ASCII twin
    of that file lists, and serialising them again reproduces the bytes.  (The fixture's ring is m = 12; the record is
    loaded into an N = 4 context over the same three primes -- the wire format does not involve the transform.)"""
    import struct
    from helib_b200.engine import Engine
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "helib_iotest_m12.json")))
    case, rec = G["cases"][0], G["dcrt_record"]
    primes = case["primes"][:3]
    E = Engine(8, primes, None, [[0, 1]], [2], lib=lib)
    assert E.N == 4
    blob = bytes.fromhex(rec["hex"])
    P = E.poly()
    assert P.deserialize(blob) == [0, 1, 2]
    got = P.download([0, 1, 2])
    for i in range(3):
        assert [int(v) for v in got[i]] == case["pk_c0"][str(i)]
    assert P.serialize([0, 1, 2]) == blob
    # the xdouble field next to it: raw double mantissa + int64 exponent (src/binio.cpp:165-171)
    mant, expo = struct.unpack("<dq", bytes.fromhex(rec["noise_bound_field_hex"]))
    assert abs(mant - 2007.04) < 1e-9 and expo == 0
