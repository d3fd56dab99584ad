"""Prime chain: product C++ (hb_chain_build) vs the Python restatement, plus known answers.

Host-only code: runs against the real libhelib_b200.so (no device needed for the chain ABI).
"""
import ctypes
import math

import pytest

import pyoracle as po
from helib_b200 import Chain, HbError, load_library, library_path

CONFIGS = {
    "cfg1_bgv_m4096": (4096, 257, 1, 60, 2),
    "cfg2_ckks_2^17_1190": (1 << 17, -1, 1, 1190, 2),
    "cfg3_bgv_2^17_1500_c3": (1 << 17, 257, 1, 1500, 3),
    "cfg4_ckks_2^17_1700": (1 << 17, -1, 1, 1700, 2),
    "bgv_m64": (64, 257, 1, 120, 2),
    "bgv_p17r2": (2048, 17, 2, 150, 3),
    "ckks_8192": (8192, -1, 1, 119, 2),
    "thinboot_m21845_chain_only": (21845, 2, 1, 580, 2),   # non power of two: chain logic only
}


@pytest.fixture(scope="module")
def lib():
    return load_library()


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_chain_matches_python_restatement(lib, name):
    m, p, r, bits, c = CONFIGS[name]
    ref = po.build_mod_chain(m, p, r, bits, c)
    ch = Chain(m, p, r, bits, c, lib=lib)
    assert ch.primes == ref.primes
    assert (ch.small, ch.ctxt, ch.special) == (ref.small, ref.ctxt, ref.special)
    assert ch.digits == ref.digits
    assert ch.phim == ref.phim


BOOT = {
    "thinboot_m21845": (21845, 2, 1, 580, 2),     # tests/GTestThinBootstrapping.cpp:102 (BASELINE config 5)
    "boot_m4369": (4369, 2, 1, 400, 3),
    "boot_p17_m105": (105, 17, 1, 300, 2),
    "boot_p17r2_m45": (45, 17, 2, 200, 2),
}


@pytest.mark.parametrize("name", sorted(BOOT))
def test_bootstrappable_chain_matches_python_restatement(lib, name):
    """ContextBuilder::bootstrappable(): default key weight 120, (e, e') from RecryptData::setAE, special primes sized for
    p^(r+e-e') (src/Context.cpp:885-897, src/recryption.cpp:200-256): product C++ == Python restatement, and the branch
    actually changes the chain."""
    m, p, r, bits, c = BOOT[name]
    ref = po.build_mod_chain(m, p, r, bits, c, bootstrappable=True)
    ch = Chain(m, p, r, bits, c, lib=lib, bootstrappable=True)
    assert ch.primes == ref.primes and ch.special == ref.special and ch.digits == ref.digits
    assert (ch.e_param, ch.e_prime_param, ch.sk_hwt) == (ref.e_param, ref.e_prime_param, 120)
    plain = Chain(m, p, r, bits, c, lib=lib)
    assert [ch.primes[i] for i in ch.ctxt] == [plain.primes[i] for i in plain.ctxt]
    logp = lambda chn: sum(math.log2(chn.primes[i]) for i in chn.special)
    assert logp(ch) > logp(plain) + (ch.e_param - ch.e_prime_param) * math.log2(p) - 6   # the special primes absorb p^(e-e')


def test_set_ae_satisfies_the_recryption_inequality():
    """Appendix A of ia.cr/2014/873 as the reference states it (src/recryption.cpp:139-150):
    (f*p^e' + 2*p^r + 2)*B <= p^e/2, with e - e' as small as the search allows."""
    for m, p, r in [(21845, 2, 1), (105, 2, 1), (4369, 2, 1), (105, 17, 1), (45, 17, 2), (57, 7, 1)]:
        e, ep = po.set_ae(m, p, r, 120)
        phim = po.euler_phi(m)
        k = len({f for f in range(2, m + 1) if m % f == 0 and all(f % g for g in range(2, int(f ** 0.5) + 1))})
        B = 0.5 + 10.0 * math.sqrt(phim / m * 120 * (1 << k) / 3.0) * 0.5
        assert e > ep >= 0 and e >= r + 1
        if ep > 0:
            assert p ** e >= (p ** ep * po.compute_fudge(p ** ep, p ** e) + 2 * p ** r + 2) * B * 2
        else:
            assert p ** e >= (2 * p ** r + 2) * B * 2
    ckks = Chain(1 << 13, -1, 1, 119, 2, bootstrappable=True)   # ignored for CKKS (src/Context.cpp:1051-1052)
    assert ckks.e_param == 0 and ckks.primes == Chain(1 << 13, -1, 1, 119, 2).primes


def test_survey_shapes():
    """SURVEY.md section 8 header: the shapes the reference's chain logic yields."""
    ch = po.build_mod_chain(1 << 17, -1, 1, 1190, 2)
    assert (len(ch.small), len(ch.ctxt), len(ch.special)) == (6, 20, 10)
    assert [len(d) for d in ch.digits] == [10, 10]
    assert [q.bit_length() for q in ch.primes[:6]] == [40, 40, 48, 51, 54, 57]
    ch = po.build_mod_chain(1 << 17, 257, 1, 1500, 3)
    assert (len(ch.ctxt), len(ch.special), [len(d) for d in ch.digits]) == (26, 9, [9, 9, 8])
    assert ch.primes[ch.ctxt[0]].bit_length() == 58 and ch.primes[ch.special[0]].bit_length() == 56
    ch = po.build_mod_chain(1 << 17, -1, 1, 1700, 2)
    assert (len(ch.ctxt), len(ch.special)) == (29, 15)
    ch = po.build_mod_chain(4096, 257, 1, 60, 2)
    assert (len(ch.small), len(ch.ctxt), len(ch.special)) == (6, 2, 1)


def test_prime_generator_known_answers():
    """SURVEY.md section 9.12 known answers (restated rule + independent primality test)."""
    assert po.PrimeGenerator(60, 1 << 17).next() == 237 * 2**52 + 1 == 1067353111686807553
    assert po.PrimeGenerator(40, 1 << 17).next() == 113 * 2**33 + 1 == 970662608897
    assert po.PrimeGenerator(58, 1 << 17).next() == 280349076803813377
    assert po.PrimeGenerator(54, 4096).next() == 16044073672507393
    sympy = pytest.importorskip("sympy")
    g = po.PrimeGenerator(60, 1 << 17)
    for _ in range(5):
        q = g.next()
        assert sympy.isprime(q) and (q - 1) % (1 << 17) == 0
        assert (1 << 60) - (1 << 57) <= q < (1 << 60)


def test_ctxt_primes_bits_within_4_percent():
    """reference: tests/TestContext.cpp:205-225 (buildModChain(1016, c=2): total ctxt bits within 4%)."""
    ch = po.build_mod_chain(8192, 3, 1, 1016, 2)
    total = sum(math.log2(ch.primes[i]) for i in ch.ctxt)
    assert 1016 - 0.5 <= total <= 1016 * 1.04


def test_digit_count_clipped(lib):
    """reference: tests/TestContext.cpp:249-306 (number of digits clipped to #ctxt primes)."""
    ch = Chain(64, 257, 1, 50, 5, lib=lib)
    assert len(ch.digits) <= len(ch.ctxt)


def test_bad_arguments(lib):
    with pytest.raises(HbError):
        Chain(4096, 2, 1, 60, 2, lib=lib)      # p divides m (src/PAlgebra.cpp:458)
    with pytest.raises(HbError):
        Chain(4096, 257, 1, 0, 2, lib=lib)     # nBits < 1 (src/Context.cpp:1044-1046)


def test_set4size_prefers_fewest_dropped(lib):
    m, p, r, bits, c = CONFIGS["cfg2_ckks_2^17_1190"]
    ch = Chain(m, p, r, bits, c, lib=lib)
    full = ch.ctxt
    logq = sum(math.log(ch.primes[i]) for i in full)
    lo = logq - 1.5 * math.log(ch.primes[full[-1]])
    got = ch.set4size(lo, lo + 4 * math.log(2), full, full, reverse=True)
    size = sum(math.log(ch.primes[i]) for i in got)
    assert lo <= size <= lo + 4 * math.log(2)
    assert set(i for i in got if i in ch.ctxt) == set(ch.ctxt[:len([i for i in got if i in ch.ctxt])])  # prefix of ctxt primes


def test_c_abi_exports_every_declared_symbol():
    """The C-ABI library loads here (no GPU) and exports every symbol include/*.h declares."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = ctypes.CDLL(library_path())
    names = set()
    for h in ("helib_b200.h", "helib_b200_chain.h"):
        src = open(os.path.join(root, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src))
    assert len(names) > 30
    for n in sorted(names):
        assert hasattr(lib, n), f"missing export {n}"
    assert lib.hb_device_count() == 0 or lib.hb_device_count() > 0


def test_no_device_fails_loudly(lib):
    """No CPU fallback: without a CUDA device the engine refuses to create a context."""
    if lib.hb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from helib_b200 import Engine
    ch = po.build_mod_chain(64, 257, 1, 120, 2)
    with pytest.raises(HbError) as ei:
        Engine(64, ch.primes, lib=lib)
    assert ei.value.code == -3   # HB_ERR_NO_DEVICE


@pytest.mark.parametrize("name", ["cfg1_bgv_m4096", "cfg2_ckks_2^17_1190", "bgv_p17r2", "thinboot_m21845_chain_only"])
def test_set4size_matches_restated_reference(lib, name):
    """hb_chain_set4size (product, C++) vs the Python restatement of ModuliSizes::getSet4Size (src/primeChain.cpp:179-319)
    on random windows: one- and two-operand forms, both search directions, windows that contain candidates and windows
    that fall between them (the one-bit-of-slack fallback), from-sets that are prefixes with and without small primes."""
    import random
    m, p, r, bits, c = CONFIGS[name]
    ref = po.build_mod_chain(m, p, r, bits, c)
    ch = Chain(m, p, r, bits, c, lib=lib)
    MS = po.ModuliSizes(ref)
    rnd = random.Random(hash(name) & 0xffff)
    total = sum(math.log(ref.primes[i]) for i in ref.small + ref.ctxt)
    for trial in range(300):
        def some_set():
            k = rnd.randint(1, len(ref.ctxt))
            s = list(ref.ctxt[:k])
            s += [i for i in ref.small if rnd.random() < 0.3]
            return sorted(s)
        f1 = some_set()
        f2 = some_set() if rnd.random() < 0.6 else None
        low = rnd.uniform(0.0, total * 1.05)
        width = rnd.choice([0.01, 0.5, 2.0, 4 * math.log(2), 40.0])
        rev = rnd.random() < 0.5
        want = MS.get_set4size(low, low + width, f1, f2, rev)
        got = ch.set4size(low, low + width, f1, f2, reverse=rev)
        assert sorted(got) == want, (trial, low, width, f1, f2, rev)
