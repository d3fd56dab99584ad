"""Bit-exact parity of the engine (C ABI -> kernels) against the CPU oracle.

Every test body runs twice: on the CPU kernel-logic simulator build (not gpu; checks the kernel
source's index/modular arithmetic here) and on the real CUDA library on a B200 (-m gpu).
The bar is bit-exact equality of every live row.
"""
import numpy as np
import pytest

import orc
import pyoracle as po
from common import make, rows_equal, ptxt_space
from helib_b200.engine import Engine

SMALL = [  # (m, p, r, bits, c)
    (64, 257, 1, 120, 2),      # N=32, single-phase transform
    (2048, 17, 2, 150, 3),     # N=1024, single-phase, p^r = 289, 3 digits
    (4096, 257, 1, 60, 2),     # BASELINE config 1' (N=2048, two-phase N1=8)
    (8192, -1, 1, 119, 2),     # CKKS, N=4096
]
BIG = [(1 << 17, -1, 1, 1190, 2)]  # BASELINE config 2 (gpu only)


def backends():
    return [pytest.param("sim", id="sim"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=backends())
def lib(request):
    return request.getfixturevalue("sim_lib" if request.param == "sim" else "cuda_lib")


@pytest.mark.parametrize("cfg", SMALL + [(1 << 15, 257, 1, 120, 2)])   # + N = 2^14: the blk kernels with 64 columns, generic cols kernels
def test_ntt_rows_match_oracle(lib, cfg):
    ch, psis, O, E = make(lib, *cfg)
    rng = np.random.default_rng(1)
    allp = list(range(len(ch.primes)))
    data = O.random(rng, allp)
    P = E.poly(data, allp)
    E.ntt_inv([P], allp)
    ref = data.copy(); O.ntt_inv_rows(ref, allp)
    assert rows_equal(P.download(allp), ref, allp)
    E.ntt_fwd([P], allp)
    assert rows_equal(P.download(allp), data, allp)   # iNTT(NTT(x)) == x


def test_rows_with_the_reference_style_roots(lib):
    """psi is an input of the engine: with the roots NTL derives under HElib's fixed seed (pyoracle.ntl_fft_root, restating
    src/CModulus.cpp:93-98 + NTL's IsFFTPrime on the pinned stream) engine and oracle still agree row for row, and the
    psi-independent results (toPoly) equal the ones computed under the default roots."""
    cfg = (8192, 257, 1, 160, 2)
    ch, psis0, O0, E0 = make(lib, *cfg)
    psis = [po.ntl_fft_root(q, ch.m) for q in ch.primes]
    assert psis != psis0
    O = orc.Oracle(ch.phim, ch.m, ch.primes, psis, ch.digits, ch.special, nthreads=4)
    E = Engine(ch.m, ch.primes, psis, ch.digits, ch.special, lib=lib)
    assert E.psis == psis
    rng = np.random.default_rng(9)
    S = ch.ctxt
    f = rng.integers(-50, 50, size=ch.phim).astype(np.int64)
    P = E.poly(); P0 = E0.poly()
    E.from_i64([P], S, f); E0.from_i64([P0], S, f)
    ref = O.zeros()
    for i in S:
        ref[i] = np.array([int(v) % ch.primes[i] for v in f], dtype=np.uint64)
    O.ntt_fwd_rows(ref, S)
    assert rows_equal(P.download(S), ref, S)
    assert not rows_equal(P.download(S), P0.download(S), S)                        # another root, another evaluation order
    assert (E.to_poly(P, S) == E0.to_poly(P0, S)).all()                             # the polynomial itself is the same


@pytest.mark.parametrize("cfg", SMALL)
def test_pointwise_tensor_automorph(lib, cfg):
    ch, psis, O, E = make(lib, *cfg)
    rng = np.random.default_rng(2)
    S = ch.ctxt + ch.special
    a, b = O.random(rng, S), O.random(rng, S)
    for op in ("add", "sub", "mul"):
        A, B = E.poly(a, S), E.poly(b, S)
        E.pointwise(op, [A], [B], S)
        ref = a.copy(); O.pointwise(op, ref, b, S)
        assert rows_equal(A.download(S), ref, S), op
    a1, b1 = O.random(rng, S), O.random(rng, S)
    r0, r1, r2 = O.tensor(a, a1, b, b1, S)
    A0, A1, B0, B1 = E.poly(a, S), E.poly(a1, S), E.poly(b, S), E.poly(b1, S)
    E.tensor([A0], [A1], [B0], [B1], [A0], [A1], [B0], S)     # in place, as the fused path does
    assert rows_equal(A0.download(S), r0, S) and rows_equal(A1.download(S), r1, S) and rows_equal(B0.download(S), r2, S)
    for k in (3, 5, ch.m - 1):
        A, D = E.poly(a, S), E.poly()
        E.automorph([D], [A], S, k)
        ref = a.copy(); O.automorph(ref, S, k)
        assert rows_equal(D.download(S), ref, S), k
    A = E.poly(a, S)
    E.scale_by_primes([A], ch.ctxt, ch.special, inv=True)
    ref = a.copy(); O.scale_by_primes(ref, ch.ctxt, ch.special, inv=True)
    assert rows_equal(A.download(S), ref, ch.ctxt)


@pytest.mark.parametrize("cfg", SMALL)
def test_add_primes_and_to_poly(lib, cfg):
    ch, psis, O, E = make(lib, *cfg)
    rng = np.random.default_rng(3)
    cur = ch.digits[0]
    add = [i for i in ch.ctxt + ch.special if i not in cur]
    x = O.random(rng, cur)
    P = E.poly(x, cur)
    tp = E.to_poly(P, cur)
    assert (tp == O.to_poly(x, cur)).all()
    assert (E.to_poly(P, cur, positive=True) == O.to_poly(x, cur, positive=True)).all()
    E.add_primes([P], cur, add)
    ref = x.copy(); O.add_primes(ref, cur, add)
    allr = cur + add
    assert rows_equal(P.download(allr), ref, allr)
    # addPrimesAndScale
    Q = E.poly(x, cur)
    E.add_primes_and_scale([Q], cur, add)
    ref = x.copy(); O.add_primes_and_scale(ref, cur, add)
    assert rows_equal(Q.download(allr), ref, allr)


@pytest.mark.parametrize("cfg", SMALL)
@pytest.mark.parametrize("pspace", [None, 2, 4])
def test_scale_down(lib, cfg, pspace):
    ch, psis, O, E = make(lib, *cfg)
    p = ptxt_space(ch) if pspace is None else pspace
    rng = np.random.default_rng(4)
    cur = ch.ctxt + ch.special
    for keep in (ch.ctxt, ch.ctxt[:-1] if len(ch.ctxt) > 1 else ch.ctxt):
        x = O.random(rng, cur)
        P = E.poly(x, cur)
        E.scale_down([P], cur, keep, p)
        ref = x.copy(); O.scale_down(ref, cur, keep, p)
        assert rows_equal(P.download(keep), ref, keep)


@pytest.mark.parametrize("cfg", SMALL)
def test_break_into_digits_and_keyswitch(lib, cfg):
    ch, psis, O, E = make(lib, *cfg)
    rng = np.random.default_rng(5)
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    for S in (ch.ctxt, ch.ctxt[:-1] if len(ch.ctxt) > 1 else ch.ctxt):
        Sp = sorted(S + ch.special)
        xs = [O.random(rng, S) for _ in range(2)]       # batch of two items
        Ps = [E.poly(x, S) for x in xs]
        digs = E.break_into_digits(Ps, S)
        for it, x in enumerate(xs):
            ref = O.break_into_digits(x, S)
            assert len(digs[it]) == ref.shape[0]
            for i, D in enumerate(digs[it]):
                assert rows_equal(D.download(Sp), ref[i], Sp), (it, i)
        # full relinearisation of a 3-part ciphertext
        c0s, c1s = [O.random(rng, S) for _ in xs], [O.random(rng, S) for _ in xs]
        C0, C1 = [E.poly(c, S) for c in c0s], [E.poly(c, S) for c in c1s]
        E.relinearize(C0, C1, Ps, S, EA, EB)
        for it, x in enumerate(xs):
            r0, r1 = O.relinearize(c0s[it], c1s[it], x, S, evk_a, evk_b)
            assert rows_equal(C0[it].download(Sp), r0, Sp) and rows_equal(C1[it].download(Sp), r1, Sp)


def oracle_mul_relin_moddown(O, ch, a0, a1, b0, b1, S_in, S, p, evk_a, evk_b):
    """Host orchestration restated from Ctxt::multLowLvl + reLinearize + modDownToSet
    (src/Ctxt.cpp:393-562,720-786,1681-1774) with explicit prime sets."""
    parts = [x.copy() for x in (a0, a1, b0, b1)]
    for x in parts:
        O.scale_down(x, S_in, S, p)
    t0, t1, t2 = O.tensor(parts[0], parts[1], parts[2], parts[3], S)
    r0, r1 = O.relinearize(t0, t1, t2, S, evk_a, evk_b)
    Sp = sorted(S + ch.special)
    O.scale_down(r0, Sp, S, p)
    O.scale_down(r1, Sp, S, p)
    return r0, r1


@pytest.mark.parametrize("cfg", SMALL)
def test_mul_relin_moddown(lib, cfg):
    ch, psis, O, E = make(lib, *cfg)
    p = ptxt_space(ch)
    rng = np.random.default_rng(6)
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    S_in = ch.ctxt
    S = ch.ctxt[:-1] if len(ch.ctxt) > 1 else ch.ctxt
    ops = [[O.random(rng, S_in) for _ in range(4)] for _ in range(2)]
    A0, A1, B0, B1 = ([E.poly(o[k], S_in) for o in ops] for k in range(4))
    E.mul_relin_moddown(A0, A1, B0, B1, S_in, S, p, EA, EB)
    for it, o in enumerate(ops):
        r0, r1 = oracle_mul_relin_moddown(O, ch, *o, S_in, S, p, evk_a, evk_b)
        assert rows_equal(A0[it].download(S), r0, S) and rows_equal(A1[it].download(S), r1, S)


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes (N = 2^16): direct bit-exact parity (the C++ oracle finishes these in
# seconds) plus size-independent properties.  GPU only.

@pytest.mark.gpu
def test_full_size_config2_mul_relin_moddown(cuda_lib):
    cfg = (1 << 17, -1, 1, 1190, 2)
    ch, psis, O, E = make(cuda_lib, *cfg, nthreads=8)
    assert (len(ch.ctxt), len(ch.special)) == (20, 10)
    rng = np.random.default_rng(7)
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    S_in, S = ch.ctxt, ch.ctxt[:-1]
    ops = [[O.random(rng, S_in) for _ in range(4)] for _ in range(2)]
    A0, A1, B0, B1 = ([E.poly(o[k], S_in) for o in ops] for k in range(4))
    E.mul_relin_moddown(A0, A1, B0, B1, S_in, S, 1, EA, EB)
    for it, o in enumerate(ops):
        r0, r1 = oracle_mul_relin_moddown(O, ch, *o, S_in, S, 1, evk_a, evk_b)
        assert rows_equal(A0[it].download(S), r0, S) and rows_equal(A1[it].download(S), r1, S)
    # round trip at full size
    x = O.random(rng, full)
    P = E.poly(x, full)
    E.ntt_inv([P], full); E.ntt_fwd([P], full)
    assert rows_equal(P.download(full), x, full)


@pytest.mark.gpu
def test_full_size_config3_relinearize_bgv(cuda_lib):
    cfg = (1 << 17, 257, 1, 1500, 3)
    ch, psis, O, E = make(cuda_lib, *cfg, nthreads=8)
    assert (len(ch.ctxt), len(ch.special), [len(d) for d in ch.digits]) == (26, 9, [9, 9, 8])
    rng = np.random.default_rng(8)
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    S = ch.ctxt
    Sp = sorted(S + ch.special)
    c0, c1, c2 = (O.random(rng, S) for _ in range(3))
    C0, C1, C2 = E.poly(c0, S), E.poly(c1, S), E.poly(c2, S)
    E.relinearize([C0], [C1], [C2], S, EA, EB)
    r0, r1 = O.relinearize(c0, c1, c2, S, evk_a, evk_b)
    assert rows_equal(C0.download(Sp), r0, Sp) and rows_equal(C1.download(Sp), r1, Sp)
    # mod-down with the BGV correction at full size
    E.scale_down([C0, C1], Sp, S, 257)
    O.scale_down(r0, Sp, S, 257); O.scale_down(r1, Sp, S, 257)
    assert rows_equal(C0.download(S), r0, S) and rows_equal(C1.download(S), r1, S)


def test_sim_full_ring_dimension_small_chain(sim_lib):
    """N = 2^16 with a short chain on the simulator: exercises the register-blocked (v1) cols and
    fused-conversion kernels, which only exist for N = 2^16, without a GPU."""
    cfg = (1 << 17, 257, 1, 230, 2)
    ch, psis, O, E = make(sim_lib, *cfg, nthreads=8)
    rng = np.random.default_rng(9)
    full = ch.ctxt + ch.special
    x = O.random(rng, full)
    P = E.poly(x, full)
    E.ntt_inv([P], full)
    ref = x.copy(); O.ntt_inv_rows(ref, full)
    assert rows_equal(P.download(full), ref, full)
    E.ntt_fwd([P], full)
    assert rows_equal(P.download(full), x, full)
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    S_in = ch.ctxt
    S = ch.ctxt[:-1] if len(ch.ctxt) > 1 else ch.ctxt
    o = [O.random(rng, S_in) for _ in range(4)]
    A0, A1, B0, B1 = ([E.poly(o[k], S_in)] for k in range(4))
    E.mul_relin_moddown(A0, A1, B0, B1, S_in, S, 257, EA, EB)
    r0, r1 = oracle_mul_relin_moddown(O, ch, *o, S_in, S, 257, evk_a, evk_b)
    assert rows_equal(A0[0].download(S), r0, S) and rows_equal(A1[0].download(S), r1, S)


@pytest.mark.parametrize("cfg", [(1 << 17, 257, 1, 230, 3), (1 << 17, -1, 1, 330, 3)])
def test_fused_relinearize_three_digits(lib, cfg):
    """hb_relinearize on the register kernels runs breakIntoDigits fused (the switched part is updated in place, the
    mixed-radix step rides in the forward blk epilogue, the inner product reads a digit's own rows from the part):
    three digits exercise the chained update c2 <- (c2 - E_i)/Q_i over two steps; two items at once; a lower level
    (one ctxt prime dropped) changes the digit sets.  Bit-exact vs the oracle's reLinearize + mod-down."""
    ch, psis, O, E = make(lib, *cfg, nthreads=8)
    p = ptxt_space(ch)
    rng = np.random.default_rng(77)
    full = ch.ctxt + ch.special
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    for S in (ch.ctxt, ch.ctxt[:-1]):
        Sp = sorted(S + ch.special)
        cs = [[O.random(rng, S) for _ in range(3)] for _ in range(2)]
        C0, C1, C2 = ([E.poly(c[k], S) for c in cs] for k in range(3))
        E.relinearize(C0, C1, C2, S, EA, EB)
        E.scale_down(C0 + C1, Sp, S, p)
        for it, c in enumerate(cs):
            r0, r1 = O.relinearize(c[0], c[1], c[2], S, evk_a, evk_b)
            O.scale_down(r0, Sp, S, p); O.scale_down(r1, Sp, S, p)
            assert rows_equal(C0[it].download(S), r0, S) and rows_equal(C1[it].download(S), r1, S), (len(S), it)


def test_sim_generic_modulus_path(sim_lib, monkeypatch):
    """The register kernels have two modulus views: HElib's q = t*2^s+1 (s >= 32) shift form and the
    generic 2^64-q form.  Force the generic one (HB_NO_SPECIAL) on the N = 2^16 circuit."""
    monkeypatch.setenv("HB_NO_SPECIAL", "1")
    test_sim_full_ring_dimension_small_chain(sim_lib)


def _largest_primes(form, count, m):
    """Largest primes below 2^60 with m | q-1: 'sp' = t*2^32+1 (the shift form of the register kernels), 'gen' = k*m+1."""
    out = []
    if form == "sp":
        t = (1 << 28) - 1
        while len(out) < count:
            q = (t << 32) + 1
            if po.is_prime(q):
                out.append(q)
            t -= 1
    else:
        k = ((1 << 60) - 2) // m
        while len(out) < count:
            q = k * m + 1
            if q < (1 << 60) and (q - 1) % (1 << 32) != 0 and po.is_prime(q):
                out.append(q)
            k -= 1
    return out


@pytest.mark.parametrize("form", ["sp", "gen"])
def test_lazy_ranges_at_the_largest_primes(lib, form):
    """The register kernels keep values in [0, 8q + 2^32) (forward) / [0, 4q + 2^48) (inverse) with q < 2^60; the
    margins are smallest for the largest admissible primes and for inputs at the top of the range.  Rows of q-1, of
    alternating 0 / q-1 and random rows through the N = 2^16 transforms and a base extension, against the oracle."""
    from helib_b200.engine import Engine
    m = 1 << 17
    primes = _largest_primes(form, 3, m)
    assert all(q < (1 << 60) and q > (1 << 60) - (1 << 40) for q in primes)
    psis = [po.find_psi(q, m) for q in primes]
    N = m // 2
    O = orc.Oracle(N, m, primes, psis, [[0, 1]], [2], nthreads=8)
    E = Engine(m, primes, psis, [[0, 1]], [2], lib=lib)
    rng = np.random.default_rng(3)
    idx = [0, 1, 2]
    x = O.random(rng, idx)
    for i, q in enumerate(primes):
        x[i][: N // 4] = q - 1
        x[i][N // 4: N // 2: 2] = 0
        x[i][N // 4 + 1: N // 2: 2] = q - 1
    P = E.poly(x, idx)
    E.ntt_inv([P], idx)
    ref = x.copy(); O.ntt_inv_rows(ref, idx)
    assert rows_equal(P.download(idx), ref, idx)
    E.ntt_fwd([P], idx)
    assert rows_equal(P.download(idx), x, idx)
    # forward transform of extreme coefficient vectors
    E.ntt_fwd([P], idx)
    ref = x.copy(); O.ntt_fwd_rows(ref, idx)
    assert rows_equal(P.download(idx), ref, idx)
    # exact base extension {0,1} -> {2} (fused conversion kernel) and mod-down back
    Y = E.poly(x, [0, 1])
    E.add_primes([Y], [0, 1], [2])
    ref = x.copy(); O.add_primes(ref, [0, 1], [2])
    assert rows_equal(Y.download(idx), ref, idx)
    E.scale_down([Y], idx, [0, 1], 1)
    O.scale_down(ref, idx, [0, 1], 1)
    assert rows_equal(Y.download([0, 1]), ref, [0, 1])


@pytest.mark.parametrize("cfg", [(64, 257, 1, 120, 2), (4096, 257, 1, 60, 2), (8192, -1, 1, 119, 2)])
def test_hoisted_automorph_keyswitch(lib, cfg):
    """one breakIntoDigits, many (automorph + keySwitchDigits) -- BasicAutomorphPrecon
    (src/matmul.cpp:60-184).  Checked against the oracle doing automorph on every digit and on c0,
    addPrimesAndScale, keySwitchDigits (src/matmul.cpp:152-170)."""
    ch, psis, O, E = make(lib, *cfg)
    rng = np.random.default_rng(31)
    S = ch.ctxt
    full = S + ch.special
    Sp = sorted(full)
    nd = len(ch.digits)
    c0, c1 = O.random(rng, S), O.random(rng, S)
    C0, C1 = E.poly(c0, S), E.poly(c1, S)
    digs = E.break_into_digits([C1], S)
    ref_digs = O.break_into_digits(c1, S)
    for k in (3, 5, ch.m - 1):
        evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
        evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
        EA = [E.poly(evk_a[i], full) for i in range(nd)]
        EB = [E.poly(evk_b[i], full) for i in range(nd)]
        O0, O1 = E.poly(), E.poly()
        E.automorph_keyswitch_digits(digs, S, [C0], k, EA, EB, [O0], [O1])
        r0 = c0.copy(); O.automorph(r0, S, k); O.add_primes_and_scale(r0, S, ch.special)
        r1 = O.zeros()
        rd = ref_digs.copy()
        for i in range(rd.shape[0]):
            O.automorph(rd[i], Sp, k)
        O.keyswitch_digits(rd, Sp, evk_a, evk_b, r0, r1)
        assert rows_equal(O0.download(Sp), r0, Sp) and rows_equal(O1.download(Sp), r1, Sp), k


@pytest.mark.parametrize("cfg", [(4096, 17, 1, 160, 3), (1 << 17, 257, 1, 230, 2)])
def test_fused_inner_product_and_mixed_radix_step(lib, cfg):
    """hb_keyswitch_digits_fused (addPrimesAndScale folded in, own digit rows read from the switched part) and
    hb_sub_div_by_primes (the mixed-radix step of breakIntoDigits) against the same steps done one by one by the oracle
    (src/Ctxt.cpp:191-230,764-768; src/DoubleCRT.cpp:551-556)."""
    ch, psis, O, E = make(lib, *cfg, nthreads=8)
    rng = np.random.default_rng(41)
    S, full = ch.ctxt, ch.ctxt + ch.special
    Sp = sorted(full)
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    c0, c1, c2 = (O.random(rng, S) for _ in range(3))
    # oracle: digits, scale c0 / c1 up to S | special, inner product
    digs = O.break_into_digits(c2, S)
    r0, r1 = c0.copy(), c1.copy()
    O.add_primes_and_scale(r0, S, ch.special); O.add_primes_and_scale(r1, S, ch.special)
    O.keyswitch_digits(digs, Sp, evk_a, evk_b, r0, r1)
    # engine: digit polynomials hold only the EXTENDED rows, the own rows of digit i are read from `own`
    dsets = [[i for i in S if i in ch.digits[d]] for d in range(nd)]
    own = O.zeros()
    D = []
    for d in range(nd):
        ext = digs[d].copy()
        for i in dsets[d]:
            own[i] = digs[d][i]
            ext[i] = 0          # must not be read
        D.append(E.poly(ext, Sp))
    OWN = E.poly(own, S)
    P = 1
    for i in ch.special:
        P *= ch.primes[i]
    scal = [P % ch.primes[r] if r in S else 0 for r in Sp]
    own_dig = [next(d for d in range(nd) if r in dsets[d]) if r in S else -1 for r in Sp]
    garbage = O.random(rng, ch.special)      # rows with scal == 0 are pure outputs: whatever they held is ignored
    g0 = c0.copy(); g1 = c1.copy()
    for i in ch.special:
        g0[i] = garbage[i]; g1[i] = garbage[i]
    C0, C1 = E.poly(g0, Sp), E.poly(g1, Sp)
    E.keyswitch_digits_fused([D], Sp, EA, EB, [C0], [C1], scal, own=[OWN], own_dig=own_dig)
    assert rows_equal(C0.download(Sp), r0, Sp) and rows_equal(C1.download(Sp), r1, Sp)
    # mixed-radix step: dst = (dst - src) / prod(digit 0) on the rows of digit 1
    if nd > 1:
        a, b = O.random(rng, S), O.random(rng, S)
        A, B = E.poly(a, S), E.poly(b, S)
        E.sub_div_by_primes([A], [B], dsets[1], ch.digits[0])
        ref = a.copy(); O.pointwise("sub", ref, b, dsets[1]); O.scale_by_primes(ref, dsets[1], ch.digits[0], inv=True)
        assert rows_equal(A.download(dsets[1]), ref, dsets[1])


def test_tma_inverse_blk_kernel(lib, monkeypatch):
    """k2_inv_blk (TMA-staged inverse blk phase, HB_INV_V2=1; the cp.async kernel is the default for the inverse direction):
    transform round trip and a full mod-down through it, bit-exact against the oracle."""
    monkeypatch.setenv("HB_INV_V2", "1")
    cfg = (1 << 17, 257, 1, 230, 2)
    ch, psis, O, E = make(lib, *cfg, nthreads=8)
    rng = np.random.default_rng(77)
    S, Sp = ch.ctxt, sorted(ch.ctxt + ch.special)
    x = O.random(rng, Sp)
    P = E.poly(x, Sp)
    E.ntt_inv([P], Sp)
    ref = x.copy(); O.ntt_inv_rows(ref, Sp)
    assert rows_equal(P.download(Sp), ref, Sp)
    E.ntt_fwd([P], Sp)
    assert rows_equal(P.download(Sp), x, Sp)
    E.scale_down([P], Sp, S, ch.p ** ch.r)
    O.scale_down(x, Sp, S, ch.p ** ch.r)
    assert rows_equal(P.download(S), x, S)


def test_single_source_conversion_kernel(lib, monkeypatch):
    """Mod-downs that drop ONE prime (no plaintext correction: the CKKS rescale) run through the dedicated
    kernel k1_conv1 (default; HB_CONV1=0 disables it); results must equal the oracle's scaleDownToSet bit for bit -- dropping a 60-bit ctxt prime, a special
    prime, and (different bit lengths between source and targets) with rows at the extremes."""
    monkeypatch.setenv("HB_CONV1", "1")
    cfg = (1 << 17, -1, 1, 230, 2)
    ch, psis, O, E = make(lib, *cfg, nthreads=8)
    rng = np.random.default_rng(21)
    E.reset_stats() if hasattr(E, "reset_stats") else None
    cases = [(ch.ctxt, ch.ctxt[:-1]), (ch.ctxt + ch.special[:1], ch.ctxt)]
    if ch.small:
        cases.append((sorted(ch.small[:1] + ch.ctxt), ch.ctxt))
        # a 60-bit source with a much smaller target among the kept rows: the only case where y must be reduced modulo q_t
        cases.append((sorted(ch.small[:1] + ch.ctxt), sorted(ch.small[:1] + ch.ctxt[:-1])))
    for cur, keep in cases:
        x = O.random(rng, cur)
        drop = [i for i in cur if i not in keep][0]
        qd = ch.primes[drop]
        x[drop][:8] = [0, 1, qd - 1, (qd - 1) // 2, (qd + 1) // 2, qd // 2 - 1, 2, qd - 2]
        P = E.poly(x, cur)
        E.scale_down([P], cur, keep, 1)
        ref = x.copy(); O.scale_down(ref, cur, keep, 1)
        assert rows_equal(P.download(keep), ref, keep), (cur, keep)
    # batched, two polys at once
    a, b = O.random(rng, ch.ctxt), O.random(rng, ch.ctxt)
    A, B = E.poly(a, ch.ctxt), E.poly(b, ch.ctxt)
    E.scale_down([A, B], ch.ctxt, ch.ctxt[:-1], 1)
    O.scale_down(a, ch.ctxt, ch.ctxt[:-1], 1); O.scale_down(b, ch.ctxt, ch.ctxt[:-1], 1)
    assert rows_equal(A.download(ch.ctxt[:-1]), a, ch.ctxt[:-1]) and rows_equal(B.download(ch.ctxt[:-1]), b, ch.ctxt[:-1])
