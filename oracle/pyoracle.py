"""Pure-Python big-integer restatement of HElib's DoubleCRT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under helib_b200/ may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it,
and there only as the checker.

This is the *slow, obviously-correct* oracle (Python ints, O(N^2) or simple
O(N log N) loops).  It pins the fast C++ oracle (oracle/oracle.cpp) and the
CUDA engine on small cases.  Every function cites the reference file:line
whose semantics it restates (paths relative to /root/reference).

Parity status: the reference publishes no golden integer vectors for this
path and cannot be built here (NTL/GMP not vendored).

Pinned by reference output (tests/test_oracle.py over tests/golden/helib_iotest_m12.json,
the rows a real HElib build wrote into tests/test_resources/iotest_*):
  * the general-m conventions (FindPrimitiveRoot root, row order over Z_m^*): secret-key and
    public-key rows invert to one ternary polynomial / one small multiple of p on every prime;
  * NTL's PRG stream (oracle/ntl_prg.py), DoubleCRT::randomize and the key-switching formula
    (src/keys.cpp:1239-1242): the a_i of four stored matrices are regenerated from their prgSeed;
  * breakIntoDigits / keySwitchDigits / addPrimesAndScale / scaleDownToSet / toPoly / automorph:
    a ciphertext relinearised and rotated with the REFERENCE's matrices decrypts correctly.
Still "parity unpinned": the per-prime 2N-th root psi for power-of-two m (NTL derives it from
its root tables, src/CModulus.cpp:93-98,118-119; an input of the engine) and the prime chain of
modern parameter sets (two restatements by the same reader, never compared with HElib's output).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

HELIB_SP_NBITS = 60  # src/macro.h:21 (NTL_SP_NBITS on 64-bit, no HEXL)
PRIMEGEN_B = 3       # src/PrimeGenerator.h:51

# ---------------------------------------------------------------------------
# number theory helpers
# ---------------------------------------------------------------------------

_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin for n < 3.3e24 (replaces NTL::ProbPrime(cand, 60),
    src/PrimeGenerator.h:121 -- a correct test has a deterministic outcome)."""
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def divc(a: int, b: int) -> int:
    """ceil(a/b) for positive ints (helib divc, include/helib/NumbTh.h)."""
    return -((-a) // b)


def bal(x: int, M: int) -> int:
    """Balanced remainder in [-(M-1)/2, (M-1)/2] for odd M; matches
    src/DoubleCRT.cpp:1048-1051,1096-1099 (prod_half=(prod+1)/2; tmp>=prod_half -> tmp-=prod)."""
    x %= M
    if x >= (M + 1) // 2:
        x -= M
    return x


# ---------------------------------------------------------------------------
# prime chain  (src/PrimeGenerator.h:39-127, src/Context.cpp:728-1092)
# ---------------------------------------------------------------------------


class PrimeGenerator:
    """src/PrimeGenerator.h:39-127."""

    def __init__(self, length: int, m: int):
        if not (PRIMEGEN_B <= length <= HELIB_SP_NBITS):
            raise ValueError("PrimeGenerator: len is not in [B, HELIB_SP_NBITS]")
        if not (1 <= m < (1 << HELIB_SP_NBITS)):
            raise ValueError("PrimeGenerator: m is not in [1, NTL_SP_BOUND)")
        self.len, self.m = length, m
        k = 0
        while (m << k) <= (1 << (length - PRIMEGEN_B)):
            k += 1
        self.k = k
        self.t = divc((1 << length) - 1, m << k)

    def next(self) -> int:
        L, m = self.len, self.m
        t_upper = divc((1 << L) - 1, m << self.k)
        while True:
            self.t += 1
            if self.t >= t_upper:
                self.k -= 1
                k_lower = 0 if m % 2 == 0 else 1
                if self.k < k_lower:
                    raise RuntimeError("Prime generator ran out of primes")
                self.t = divc((1 << L) - (1 << (L - PRIMEGEN_B)) - 1, m << self.k)
                t_upper = divc((1 << L) - 1, m << self.k)
            if self.t % 2 == 0:
                continue
            cand = ((self.t * m) << self.k) + 1
            assert (1 << L) - (1 << (L - PRIMEGEN_B)) <= cand < (1 << L)
            if is_prime(cand):
                return cand


def _bit_loss() -> float:
    return -math.log1p(-1.0 / float(1 << PRIMEGEN_B)) / math.log(2.0)


def ctxt_prime_size(nbits: int) -> int:
    """src/Context.cpp:816-843."""
    bit_loss = _bit_loss()
    max_psize = HELIB_SP_NBITS - bit_loss
    nprimes = int(math.ceil(nbits / max_psize))
    target = HELIB_SP_NBITS
    while (10 * (target - 1) >= 9 * HELIB_SP_NBITS and (target - 1) >= 30
           and ((target - 1) - bit_loss) * nprimes >= nbits):
        target -= 1
    return target


def euler_phi(m: int) -> int:
    r, n, p = m, m, 2
    while p * p <= n:
        if n % p == 0:
            while n % p == 0:
                n //= p
            r -= r // p
        p += 1
    if n > 1:
        r -= r // n
    return r


@dataclass
class Chain:
    """The part of helib::Context the hot path needs (include/helib/Context.h:120-180)."""
    m: int
    p: int            # plaintext prime; -1 for CKKS
    r: int
    phim: int
    primes: list = field(default_factory=list)     # q_i in chain (index) order
    small: list = field(default_factory=list)      # indices
    ctxt: list = field(default_factory=list)
    special: list = field(default_factory=list)
    digits: list = field(default_factory=list)     # list of lists of indices
    e_param: int = 0                               # Context::e_param / ePrime_param (bootstrappable chains)
    e_prime_param: int = 0

    @property
    def ckks(self) -> bool:
        return self.p == -1

    @property
    def pow2(self) -> bool:
        return self.m & (self.m - 1) == 0

    def product(self, idxs) -> int:
        out = 1
        for i in idxs:
            out *= self.primes[i]
        return out

    def log_of_product(self, idxs) -> float:
        """Context::logOfProduct -- sum of ln(q) in index order (include/helib/Context.h)."""
        s = 0.0
        for i in sorted(idxs):
            s += math.log(float(self.primes[i]))
        return s


def compute_fudge(p2e_prime: int, p2e: int) -> float:
    """compute_fudge (src/recryption.cpp:154-197)."""
    eps = 0.0
    if p2e_prime > 1:
        eps = 1.0 / (float(p2e_prime) * float(p2e_prime)) if p2e_prime % 2 == 0 else 1.0 / float(p2e)
    return 1 + eps


def set_ae(m: int, p: int, r: int, sk_hwt: int, scale: float = 10.0):
    """RecryptData::setAE (src/recryption.cpp:200-256) -> (e, e'); the bound is Context::boundForRecryption
    (include/helib/Context.h:616-638): 0.5 + scale * sqrt(phi(m)/m * hwt * 2^k / 3) / 2, k = #prime factors of m."""
    phim = euler_phi(m)
    k, mm, f = 0, m, 2
    while f * f <= mm:
        if mm % f == 0:
            k += 1
            while mm % f == 0:
                mm //= f
        f += 1
    if mm > 1:
        k += 1
    coeff_bound = 0.5 + scale * (math.sqrt((phim / m) * sk_hwt * (1 << k) / 3.0) * 0.5)
    p2r = p ** r
    frst = 2 * p2r + 2
    e_bnd, p2e_bnd = 0, 1
    while p2e_bnd <= ((1 << 30) - 2) // p:
        e_bnd += 1
        p2e_bnd *= p
    e_prime, e = 0, r + 1
    while e <= e_bnd and p ** e < frst * coeff_bound * 2:
        e += 1
    if e > e_bnd:
        raise RuntimeError("setAE: cannot find suitable e")
    t = 1
    while t <= e_bnd:
        p2t = p ** t
        e_try = max(r + 1, t + 1)
        while e_try <= e_bnd and e_try - t < e - e_prime:
            if p ** e_try >= (p2t * compute_fudge(p2t, p ** e_try) + frst) * coeff_bound * 2:
                break
            e_try += 1
        if e_try <= e_bnd and e_try - t < e - e_prime:
            e, e_prime = e_try, t
        t += 1
    return e, e_prime


def build_mod_chain(m: int, p: int, r: int, bits: int, c: int, *, sk_hwt: int = 0,
                    resolution: int = 3, bits_in_special: int = 0,
                    bootstrappable: bool = False, stdev: float = 3.2, scale: float = 10.0) -> Chain:
    """Context::buildModChain (src/Context.cpp:1037-1070) =
    addSmallPrimes (:728-790) + addCtxtPrimes (:845-872) + addSpecialPrimes (:874-1035).
    p = -1 selects CKKS (m must then be a power of two)."""
    if bits <= 0:
        raise ValueError("Cannot initialise modulus chain with nBits < 1")
    if p == -1:
        bootstrappable = False                  # src/Context.cpp:1051-1052
    if sk_hwt == 0 and bootstrappable:
        sk_hwt = 120                            # BOOT_DFLT_SK_HWT (include/helib/Context.h:34-35)
    ch = Chain(m=m, p=p, r=r, phim=euler_phi(m))
    ckks = p == -1

    def in_chain(q):
        return q in ch.primes

    # ---- addSmallPrimes (src/Context.cpp:728-790)
    cp = ctxt_prime_size(bits)
    assert cp >= 30 and 9 * HELIB_SP_NBITS <= cp * 10 <= 10 * HELIB_SP_NBITS
    if m <= 0 or m > (1 << 20):
        raise RuntimeError("addSmallPrimes: m undefined or larger than 2^20")
    if resolution < 1 or resolution > 10:
        resolution = 3
    sizes = []
    if cp >= 54:
        smallest = divc(2 * cp, 3)
    elif cp >= 45:
        smallest = divc(7 * cp, 10)
    else:
        smallest = divc(11 * cp, 15)
        sizes.append(smallest)
    sizes += [smallest, smallest]
    delta = resolution
    while cp - delta > smallest:
        sizes.append(cp - delta)
        delta *= 2
    if cp - 3 * resolution > smallest:
        sizes.append(cp - 3 * resolution)
    if resolution == 1 and cp - 11 > smallest:
        sizes.append(cp - 11)
    sizes.sort()
    last, gen = 0, None
    for sz in sizes:
        if sz != last:
            gen = PrimeGenerator(sz, m)
        q = gen.next()
        assert not in_chain(q)
        ch.small.append(len(ch.primes))
        ch.primes.append(q)
        last = sz

    # ---- addCtxtPrimes (src/Context.cpp:845-872)
    gen = PrimeGenerator(cp, m)
    bitlen = 0.0
    while bitlen < bits - 0.5:
        q = gen.next()
        assert not in_chain(q)
        ch.ctxt.append(len(ch.primes))
        ch.primes.append(q)
        bitlen += math.log2(float(q))

    # ---- addSpecialPrimes (src/Context.cpp:874-1035)
    pabs = abs(p)
    phim = ch.phim
    p2r = 1 if ckks else pabs ** r
    p2e = p2r
    if bootstrappable and not ckks:             # bigger p^e for bootstrapping (src/Context.cpp:885-897)
        e, e_prime = set_ae(m, pabs, r, sk_hwt, scale)
        p2e *= pabs ** (e - e_prime)
        ch.e_param, ch.e_prime_param = e, e_prime
    ndg = c
    if ndg > len(ch.ctxt):
        ndg = len(ch.ctxt)
    if ndg <= 0:
        ndg = 1
    digits = [[] for _ in range(ndg)]
    if ndg > 1:
        remaining = list(ch.ctxt)
        for d in range(ndg - 1):
            card = divc(len(remaining), ndg - d)
            for i in remaining:
                digits[d].append(i)
                if len(digits[d]) >= card:
                    break
            remaining = [i for i in remaining if i not in digits[d]]
        if not remaining:
            ndg -= 1
            digits = digits[:ndg]
        else:
            digits[ndg - 1] = remaining
    else:
        digits[0] = list(ch.ctxt)
    ch.digits = digits
    max_digit_log = 0.0
    for dg in digits:
        s = ch.log_of_product(dg)
        if s > max_digit_log:
            max_digit_log = s

    if bits_in_special:
        nbits = float(bits_in_special)
    else:
        h = phim / 2.0 if sk_hwt == 0 else float(sk_hwt)
        log_phim = math.log(phim)
        if log_phim < 1:
            log_phim = 1
        if ckks:
            nbits = (max_digit_log + math.log(stdev) + math.log(ndg) - 0.5 * math.log(h)) / math.log(2.0)
        elif ch.pow2:
            nbits = (max_digit_log + math.log(p2e) + math.log(stdev) + 0.5 * math.log(12.0)
                     + math.log(ndg) - 0.5 * math.log(log_phim) - 2 * math.log(pabs)
                     - math.log(h)) / math.log(2.0)
        else:
            nbits = (max_digit_log + math.log(m) + math.log(p2e) + math.log(stdev)
                     + 0.5 * math.log(12.0) + math.log(ndg) - 0.5 * log_phim
                     - 0.5 * math.log(log_phim) - 2 * math.log(pabs) - math.log(h)) / math.log(2.0)
    if nbits < 1:
        nbits = 1
    bit_loss = _bit_loss()
    max_psize = HELIB_SP_NBITS - bit_loss
    nprimes = int(math.ceil(nbits / max_psize))
    target = HELIB_SP_NBITS
    while ((target - 1) >= 0.55 * HELIB_SP_NBITS and (target - 1) >= 30
           and ((target - 1) - bit_loss) * nprimes >= nbits):
        target -= 1
    gen = PrimeGenerator(target, m)
    while nprimes > 0:
        q = gen.next()
        if in_chain(q):
            continue
        ch.special.append(len(ch.primes))
        ch.primes.append(q)
        nprimes -= 1
    return ch


# ---------------------------------------------------------------------------
# per-prime transform  (src/CModulus.cpp:358-553)
# ---------------------------------------------------------------------------


def find_psi(q: int, two_n: int) -> int:
    """A primitive (two_n)-th root of unity mod q, two_n a power of two.

    The reference takes whatever NTL's zz_pContext(INIT_USER_FFT,q) derives after a
    fixed SetSeed (src/CModulus.cpp:93-98,118-119): *parity unpinned*.  We pin a
    deterministic choice instead: psi = g^((q-1)/two_n) for the smallest g >= 2 that
    is a quadratic non-residue mod q.  The engine takes psi as an input, so a real
    HElib deployment hands over its own (SURVEY.md section 8c calibration trick)."""
    assert (q - 1) % two_n == 0
    g = 2
    while pow(g, (q - 1) // 2, q) != q - 1:
        g += 1
    psi = pow(g, (q - 1) // two_n, q)
    assert pow(psi, two_n // 2, q) == q - 1
    return psi


HELIB_ROOT_SEED = 84547180875373941534287406458029      # src/CModulus.cpp:94
NTL_FFT_MAX_ROOT = 25                                    # NTL_FFTMaxRoot (NTL include/NTL/FFT.h)


def ntl_fft_root(q: int, two_n: int) -> int:
    """The root a real HElib build uses for power-of-two m = two_n: `RootTable[0][log2 m]` of the `zz_pContext(INIT_USER_FFT, q)`
    it constructs right after `SetSeed(84547180875373941534287406458029)` (src/CModulus.cpp:93-98,118-119).  NTL is not in the
    reference tree; this restates its published algorithm (NTL src/FFT.cpp `IsFFTPrime` + `InitFFTPrimeInfo`, src/ZZ.cpp
    `RandomBnd(long)`): write q - 1 = 2^k * t (t odd); draw x = RandomBnd(q) from the seeded stream until x != 0,
    z = x^t != 1 and z has order exactly 2^k; square it down to order 2^min(k, 25) (= w, RootTable[0][mr]); entry j of the table
    is w^(2^(mr-j)).  The stream itself (SetSeed / ChaCha20) and RandomBnd's byte consumption are pinned by the reference's
    key-switching fixtures (oracle/ntl_prg.py, tests/test_oracle.py); the root derivation is restated from memory of NTL's
    sources and NOT pinned by any reference output (no power-of-two-m row exists in the reference tree): the engine therefore
    keeps psi an input, and this function is what a shim without NTL would pass."""
    from ntl_prg import random_bnd, set_seed
    assert two_n & (two_n - 1) == 0 and (q - 1) % two_n == 0
    t, k = q - 1, 0
    while t % 2 == 0:
        t //= 2
        k += 1
    stream = set_seed(HELIB_ROOT_SEED)
    while True:
        x = random_bnd(stream, q)
        if x == 0:
            continue
        z = pow(x, t, q)
        if z == 1:
            continue
        x, j = z, 0
        while True:
            y = z
            z = y * y % q
            j += 1
            if j == k or z == 1:
                break
        if z != 1 or y != q - 1:
            raise ValueError("not an FFT prime")
        if j == k:
            break
    for _ in range(NTL_FFT_MAX_ROOT, k):
        x = x * x % q
    mr = min(k, NTL_FFT_MAX_ROOT)
    lg = two_n.bit_length() - 1
    assert lg <= mr, "Roots count exceeds maximum rootTables size"       # src/CModulus.cpp:108-110
    psi = pow(x, 1 << (mr - lg), q)
    assert pow(psi, two_n // 2, q) == q - 1
    return psi


def ntt_fwd(coeffs, q: int, psi: int):
    """Cmodulus::FFT pow-2 branch (src/CModulus.cpp:362-429): y[i]=x[i]*psi^i, cyclic
    length-N DFT with omega=psi^2, natural order out => row[j] = f(psi^(2j+1)) mod q."""
    n = len(coeffs)
    a = [(int(c) % q) * pow(psi, i, q) % q for i, c in enumerate(coeffs)]
    return _cyclic_dft(a, q, psi * psi % q)


def ntt_inv(row, q: int, psi: int):
    """Cmodulus::iFFT pow-2 branch (src/CModulus.cpp:486-553): inverse cyclic DFT
    (including 1/N) then multiply by psi^-i; coefficients in [0,q)."""
    n = len(row)
    om_inv = pow(psi * psi % q, q - 2, q)
    a = _cyclic_dft([int(x) % q for x in row], q, om_inv)
    ninv = pow(n, q - 2, q)
    ipsi = pow(psi, q - 2, q)
    return [a[i] * ninv % q * pow(ipsi, i, q) % q for i in range(n)]


def _cyclic_dft(a, q, omega):
    """X[k] = sum_n a[n] omega^(nk), natural order in and out (recursive radix-2)."""
    n = len(a)
    if n == 1:
        return list(a)
    ev = _cyclic_dft(a[0::2], q, omega * omega % q)
    od = _cyclic_dft(a[1::2], q, omega * omega % q)
    out = [0] * n
    w = 1
    h = n // 2
    for k in range(h):
        t = w * od[k] % q
        out[k] = (ev[k] + t) % q
        out[k + h] = (ev[k] - t) % q
        w = w * omega % q
    return out


def negacyclic_mul_schoolbook(f, g, mod=None):
    """f*g mod (X^N+1) over the integers (optionally mod `mod`)."""
    n = len(f)
    out = [0] * n
    for i, fi in enumerate(f):
        if fi == 0:
            continue
        for j, gj in enumerate(g):
            k = i + j
            if k < n:
                out[k] += fi * gj
            else:
                out[k - n] -= fi * gj
    if mod is not None:
        out = [x % mod for x in out]
    return out


# ---------------------------------------------------------------------------
# DoubleCRT as {prime index -> row}  (include/helib/DoubleCRT.h:87-94)
# ---------------------------------------------------------------------------


def _row_fwd(chain, psis, poly, i):
    """Cmodulus::FFT: negacyclic transform for power-of-two m (psis[i] = the 2N-th root), Bluestein rows over Z_m^* for
    general m (psis[i] = the root of Cmodulus, cmod_root)."""
    q = chain.primes[i]
    if chain.pow2:
        return ntt_fwd(poly, q, psis[i])
    return gen_fft([int(c) % q for c in poly], q, chain.m, psis[i])


def _row_inv(chain, psis, row, i):
    q = chain.primes[i]
    return ntt_inv(row, q, psis[i]) if chain.pow2 else gen_ifft(row, q, chain.m, psis[i])


class PyDCRT:
    """A DoubleCRT: dict prime-index -> list of N residues (evaluation form)."""

    def __init__(self, chain: Chain, psis, rows=None):
        self.ch, self.psis = chain, psis
        self.rows = dict(rows or {})

    def copy(self):
        return PyDCRT(self.ch, self.psis, {i: list(r) for i, r in self.rows.items()})

    @property
    def index_set(self):
        return sorted(self.rows)

    @classmethod
    def from_poly(cls, chain, psis, poly, idxs):
        """DoubleCRT(poly, context, s) -> FFT (src/DoubleCRT.cpp:68-85,
        src/CModulus.cpp:453-457: coefficients reduced into [0,q) first)."""
        n = chain.phim
        poly = list(poly) + [0] * (n - len(poly))
        return cls(chain, psis, {i: _row_fwd(chain, psis, poly, i) for i in idxs})

    def to_poly(self, idxs=None, positive=False):
        """DoubleCRT::toPoly (src/DoubleCRT.cpp:925-1113)."""
        s1 = [i for i in self.index_set if idxs is None or i in idxs]
        n = self.ch.phim
        if not s1:
            return [0] * n
        Q = self.ch.product(s1)
        coefs = {i: _row_inv(self.ch, self.psis, self.rows[i], i) for i in s1}
        out = []
        for k in range(n):
            acc = 0
            for i in s1:
                q = self.ch.primes[i]
                Qi = Q // q
                t = pow(Qi % q, q - 2, q)
                acc += Qi * (coefs[i][k] * t % q)
            acc %= Q
            out.append(acc if positive else bal(acc, Q))
        return out

    # -- pointwise (src/DoubleCRT.cpp:216-384)
    def _op(self, other, fn):
        if not set(self.rows) <= set(other.rows):
            raise RuntimeError("DoubleCRT::Op: !(map.getIndexSet() <= other.map.getIndexSet())")
        for i in self.rows:
            q = self.ch.primes[i]
            self.rows[i] = [fn(a, b) % q for a, b in zip(self.rows[i], other.rows[i])]
        return self

    def add(self, o):
        return self._op(o, lambda a, b: a + b)

    def sub(self, o):
        return self._op(o, lambda a, b: a - b)

    def mul(self, o):
        return self._op(o, lambda a, b: a * b)

    def mul_scalar(self, c: int):
        for i in self.rows:
            q = self.ch.primes[i]
            cc = c % q
            self.rows[i] = [a * cc % q for a in self.rows[i]]
        return self

    def div_scalar(self, c: int):
        """operator/= (src/DoubleCRT.cpp:1122-1139)."""
        for i in self.rows:
            q = self.ch.primes[i]
            inv = pow(c % q, q - 2, q)
            self.rows[i] = [a * inv % q for a in self.rows[i]]
        return self

    def remove_primes(self, idxs):
        for i in idxs:
            self.rows.pop(i, None)
        return self

    def add_primes(self, idxs):
        """DoubleCRT::addPrimes (src/DoubleCRT.cpp:565-599). Returns the balanced poly."""
        idxs = list(idxs)
        assert not (set(idxs) & set(self.rows))
        poly = self.to_poly()
        for i in idxs:
            self.rows[i] = _row_fwd(self.ch, self.psis, poly, i)
        return poly

    def add_primes_and_scale(self, idxs):
        """DoubleCRT::addPrimesAndScale (src/DoubleCRT.cpp:603-647)."""
        idxs = list(idxs)
        assert not (set(idxs) & set(self.rows))
        f = self.ch.product(idxs)
        self.mul_scalar(f)
        for i in idxs:
            self.rows[i] = [0] * self.ch.phim
        return self

    def scale_down_to_set(self, keep, ptxt_space: int):
        """DoubleCRT::scaleDownToSet (src/DoubleCRT.cpp:1464-1516). Returns delta."""
        diff = [i for i in self.index_set if i not in keep]
        if not diff:
            return None
        assert ptxt_space >= 1 and len(diff) < len(self.rows)
        P = self.ch.product(diff)
        delta = self.to_poly(diff)
        if ptxt_space > 1:
            p = ptxt_space
            p_over_2, p_mod_2 = p // 2, p % 2
            prod_inv = pow(P % p, -1, p)
            for k, d in enumerate(delta):
                u = d % p
                if u != 0:
                    u = u * prod_inv % p
                    if u > p_over_2 or (p_mod_2 == 0 and u == p_over_2 and d < 0):
                        u -= p
                    delta[k] = d - P * u
        self.remove_primes(diff)
        dd = PyDCRT.from_poly(self.ch, self.psis, delta, self.index_set)
        self.sub(dd)
        self.div_scalar(P)
        return delta

    def break_into_digits(self):
        """DoubleCRT::breakIntoDigits (src/DoubleCRT.cpp:479-561). Returns (digits, polys)."""
        ch = self.ch
        assert set(self.rows) <= set(ch.ctxt)
        remaining = set(self.rows)
        n = 0
        while remaining:
            remaining -= set(ch.digits[n])
            n += 1
        all_primes = sorted(set(self.rows) | set(ch.special))
        digits = []
        for i in range(n):
            d = self.copy()
            d.remove_primes([j for j in d.index_set if j not in ch.digits[i]])
            digits.append(d)
        polys = []
        for i in range(n):
            not_in = [j for j in all_primes if j not in digits[i].rows]
            polys.append(digits[i].add_primes(not_in))
            pi = ch.product(ch.digits[i])
            for j in range(i + 1, n):
                digits[j].sub(digits[i])
                digits[j].div_scalar(pi)
        return digits, polys

    def automorph(self, k: int):
        """DoubleCRT::automorph, power-of-two m (src/DoubleCRT.cpp:1160-1202):
        new[j] = old[idx(rep(j)*k mod m)], rep(j) = 2j+1."""
        m = self.ch.m
        if not self.ch.pow2:   # general m: rep(j) = j-th unit of Z_m^* ascending (src/PAlgebra.cpp:535-540)
            rep = zms_rep(m)
            pos = {r: j for j, r in enumerate(rep)}
            assert math.gcd(k, m) == 1
            for i in self.rows:
                old = self.rows[i]
                self.rows[i] = [old[pos[rep[j] * k % m]] for j in range(len(old))]
            return self
        assert k % 2 == 1
        for i in self.rows:
            old = self.rows[i]
            self.rows[i] = [old[(((2 * j + 1) * k) % m - 1) // 2] for j in range(len(old))]
        return self


def key_switch_digits(digits, evk_a, evk_b):
    """Ctxt::keySwitchDigits (src/Ctxt.cpp:191-230): returns (sum D_i*b_i, sum D_i*a_i)
    over the digits' index set (the evk rows cover all ctxt+special primes)."""
    out0 = out1 = None
    for d, a, b in zip(digits, evk_a, evk_b):
        ta = d.copy().mul(a)
        tb = d.copy().mul(b)
        out1 = ta if out1 is None else out1.add(ta)
        out0 = tb if out0 is None else out0.add(tb)
    return out0, out1


# ---------------------------------------------------------------------------
# noise metadata (float64): canonical-embedding L-infinity norm
# ---------------------------------------------------------------------------


def embedding_largest_coeff(coeffs, m: int):
    """embeddingLargestCoeff(ZZX, palg) for power-of-two m (src/norms.cpp:443-485 scale to <= 400 bits,
    :204-261 max over j in Z_m^* of |f(zeta^j)|, zeta = e^(2 pi i/m)).  Returns (mantissa, log2 factor):
    norm = mantissa * 2^shift, as the reference returns an xdouble."""
    import numpy as np
    size = max((abs(int(c)).bit_length() for c in coeffs), default=0)
    shift = max(0, size - 400)
    if m & (m - 1):
        # general m: basic_embeddingLargestCoeff (src/norms.cpp:129-157): length-m DFT of the zero-padded coefficients,
        # max of |.| over i in Z_m^*, 1 <= i <= m/2
        ff = np.zeros(m)
        for i, c in enumerate(coeffs):
            ff[i] = float(int(c) >> shift) if c >= 0 else -float((-int(c)) >> shift)
        vals = np.fft.fft(ff)
        idx = [i for i in range(1, m // 2 + 1) if math.gcd(i, m) == 1]
        return float(np.max(np.abs(vals[idx]))), shift
    n = m // 2
    ff = np.array([float(int(c) >> shift) if c >= 0 else -float((-int(c)) >> shift) for c in coeffs] + [0.0] * (n - len(coeffs)))
    k = np.arange(n)
    tw = np.exp(1j * np.pi * k / n)                 # zeta^k, zeta = e^(i pi / n)
    vals = np.fft.ifft(ff * tw) * n                 # sum_k f_k zeta^k e^(+2 pi i k j / n) = f(zeta^(2j+1))
    return float(np.max(np.abs(vals))), shift


# ---------------------------------------------------------------------------
# general (non power-of-two) m: Bluestein rows  (src/bluestein.cpp, src/CModulus.cpp:148-180,431-443,555-577)
# ---------------------------------------------------------------------------


def prime_factors(n: int):
    out, p = [], 2
    while p * p <= n:
        if n % p == 0:
            out.append(p)
            while n % p == 0:
                n //= p
        p += 1
    if n > 1:
        out.append(n)
    return out


def find_primitive_root(q: int, e: int) -> int:
    """FindPrimitiveRoot (src/NumbTh.cpp:435-493): deterministic.  For each prime p | e take the smallest
    prime g with g^((q-1)/p) != 1 and multiply the elements g^((q-1)/pp) of order pp = p^v || e."""
    assert (q - 1) % e == 0
    root = 1
    for p in prime_factors(e):
        pp = p
        ee = e // p
        while ee % p == 0:
            ee //= p
            pp *= p
        g = 2
        while True:
            if is_prime(g) and pow(g, (q - 1) // p, q) != 1:
                break
            g += 1
        root = root * pow(g, (q - 1) // pp, q) % q
    assert pow(root, e, q) == 1 and all(pow(root, e // p, q) != 1 for p in prime_factors(e))
    return root


def zms_rep(m: int):
    """rep(j): the j-th element of Z_m^* in ascending order (PAlgebra zmsRep, src/PAlgebra.cpp:535-540)."""
    return [j for j in range(1, m) if math.gcd(j, m) == 1] if m > 1 else []


def cyclotomic_poly(m: int):
    """Phi_m(X) over the integers (PAlgebra::getPhimX)."""
    def polydiv_exact(a, b):  # a / b, integer polys as lists (low -> high), exact
        a = list(a)
        out = [0] * (len(a) - len(b) + 1)
        for i in range(len(out) - 1, -1, -1):
            c = a[i + len(b) - 1] // b[-1]
            out[i] = c
            for j, bj in enumerate(b):
                a[i + j] -= c * bj
        return out
    phi = [-1] + [0] * (m - 1) + [1]            # X^m - 1
    for d in range(1, m):
        if m % d == 0:
            phi = polydiv_exact(phi, cyclotomic_poly(d))
    return phi


def cmod_root(q: int, m: int) -> int:
    """The root Cmodulus derives for general m (src/CModulus.cpp:148-165): primitive 2m-th root for even m,
    m-th for odd m."""
    return find_primitive_root(q, 2 * m if m % 2 == 0 else m)


def bluestein_dft(x, n: int, root: int, q: int):
    """What BluesteinFFT computes (src/bluestein.cpp:134-201): X_k = sum_i x_i root^(2ik), k = 0..n-1,
    via X_k = root^(k^2) * sum_i (x_i root^(i^2)) root^(-(k-i)^2); exponents mod 2n (n even) or n (n odd)."""
    e = 2 * n if n % 2 == 0 else n
    pw = [pow(root, i * i % e, q) for i in range(n)]
    rinv = pow(root, q - 2, q)
    ipw = [pow(rinv, i * i % e, q) for i in range(n)]
    y = [(int(x[i]) * pw[i]) % q if i < len(x) else 0 for i in range(n)]
    out = []
    for k in range(n):
        acc = 0
        for i in range(n):
            if y[i]:
                acc += y[i] * ipw[abs(k - i)]
        out.append(acc % q * pw[k] % q)
    return out


def gen_fft(coeffs, q: int, m: int, root: int):
    """Cmodulus::FFT for general m (src/CModulus.cpp:431-443): length-m Bluestein DFT, keep Z_m^* entries:
    row[j] = f(zeta^rep(j)), zeta = root^2."""
    X = bluestein_dft(list(coeffs), m, root, q)
    return [X[r] for r in zms_rep(m)]


def gen_ifft(row, q: int, m: int, root: int):
    """Cmodulus::iFFT for general m (src/CModulus.cpp:555-577): scatter into Z_m^* positions, Bluestein DFT with
    root^-1, reduce mod Phi_m(X), multiply by m^-1.  Coefficients in [0,q)."""
    rep = zms_rep(m)
    x = [0] * m
    for j, r in enumerate(rep):
        x[r] = int(row[j])
    A = bluestein_dft(x, m, pow(root, q - 2, q), q)
    phi = cyclotomic_poly(m)
    A = list(A)
    for k in range(m - 1, len(phi) - 2, -1):          # remainder mod Phi_m (monic)
        c = A[k]
        if c:
            for j, pj in enumerate(phi):
                A[k - (len(phi) - 1) + j] = (A[k - (len(phi) - 1) + j] - c * pj) % q
    minv = pow(m, q - 2, q)
    return [a * minv % q for a in A[:len(phi) - 1]]


# ---------------------------------------------------------------------------------------------
# Encryption / decryption data path.  Randomness is an INPUT here: the reference draws it from
# NTL's PRG (src/sample.cpp), which is not restated -- "parity unpinned" for the sampled values themselves; the
# arithmetic on them is what these functions pin.

def balanced_mulmod(f, a: int, q: int, coin=lambda: 0):
    """balanced_MulMod (src/NumbTh.cpp:876-892): c = f_i*a mod q moved to (-q/2, q/2]; for even q the tie
    c == q/2 is resolved by a coin (NTL::RandomBnd(2)) supplied by the caller."""
    out = []
    for c in f:
        c = (c % q) * a % q
        if c > q // 2 or (q % 2 == 0 and c == q // 2 and coin()):
            c -= q
        out.append(c)
    return out


def encrypt_bgv(chain, psis, pk0: "PyDCRT", pk1: "PyDCRT", r, e0, e1, ptxt, ptxt_space: int, idxs, coin=lambda: 0):
    """PubKey::Encrypt, BGV branch (src/keys.cpp:381-455): ctxt = r*pk + p*(e0,e1) + (ptxt_fixed, 0) with
    ptxt_fixed = balanced(ptxt * (Q mod p) mod p), Q the product of the ciphertext's primes."""
    rr = PyDCRT.from_poly(chain, psis, r, idxs)
    parts = []
    for pk, e in ((pk0, e0), (pk1, e1)):
        part = PyDCRT(chain, psis, {i: list(pk.rows[i]) for i in idxs})
        part.mul(rr)                                                               # parts[i] *= r   (:416)
        part.add(PyDCRT.from_poly(chain, psis, [ptxt_space * x for x in e], idxs))  # e *= p; parts[i] += e (:436-443)
        parts.append(part)
    q_mod_p = chain.product(idxs) % ptxt_space                                     # (:453)
    fixed = balanced_mulmod(list(ptxt) + [0] * (chain.phim - len(ptxt)), q_mod_p, ptxt_space, coin)
    parts[0].add(PyDCRT.from_poly(chain, psis, fixed, idxs))                       # (:454-455)
    return parts


def decrypt_bgv(chain, psis, parts, keys, ptxt_space: int, int_factor: int, idxs):
    """SecKey::Decrypt (src/keys.cpp:1327-1400): ptxt = sum_i part_i * key_i (key None = handle "one"),
    toPoly (balanced), PolyRed to [0,p), times (intFactor*Q)^-1 mod p when p > 2.
    Returns (plaintext mod p, f = the integer polynomial before reduction)."""
    acc = PyDCRT(chain, psis, {i: [0] * chain.phim for i in idxs})
    for part, key in zip(parts, keys):
        t = PyDCRT(chain, psis, {i: list(part.rows[i]) for i in idxs})
        if key is not None:
            t.mul(key)                                                             # key *= part      (:1373)
        acc.add(t)                                                                 # ptxt += key      (:1374)
    f = acc.to_poly(idxs)
    out = [c % ptxt_space for c in f]                                              # PolyRed(..., abs=true) (:1386)
    if ptxt_space > 2:
        factor = chain.product(idxs) % ptxt_space * (int_factor % ptxt_space) % ptxt_space   # (:1389-1392)
        if factor != 1:
            inv = pow(factor, -1, ptxt_space)
            out = [c * inv % ptxt_space for c in out]
    return out, f


# ---------------------------------------------------------------------------------------------
# Powerful basis and Ctxt::rawModSwitch (the mod-switch that opens recryption)

class PowerfulIndexes:
    """PowerfulTranslationIndexes (src/powerful.cpp:151-196): m = prod m_d (prime powers), the CRT index map between
    exponents mod m and the (m_1 x ... x m_k) cube, and the embedding of the (phi(m_1) x ... x phi(m_k)) cube in it."""

    def __init__(self, mvec):
        self.mvec = list(mvec)
        self.m = 1
        for f in self.mvec:
            self.m *= f
        self.phivec = [euler_phi(f) for f in self.mvec]
        self.phim = 1
        for f in self.phivec:
            self.phim *= f
        k = len(self.mvec)
        div = [self.m // f for f in self.mvec]                                   # computeDivVec (:22-32)
        inv = [pow(div[d] % self.mvec[d], -1, self.mvec[d]) if self.mvec[d] > 1 else 0 for d in range(k)]   # computeInvVec (:36-49)
        self.long_prod = [1] * (k + 1)                                           # CubeSignature::getProd(d) = prod of dims d..k-1
        for d in range(k - 1, -1, -1):
            self.long_prod[d] = self.long_prod[d + 1] * self.mvec[d]
        self.short_prod = [1] * (k + 1)
        for d in range(k - 1, -1, -1):
            self.short_prod[d] = self.short_prod[d + 1] * self.phivec[d]
        self.poly_to_cube = [0] * self.m                                         # computePowerToCubeMap (:62-84)
        self.cube_to_poly = [0] * self.m
        for i in range(self.m):
            j = 0
            for d in range(k):
                j += ((i % self.mvec[d]) * inv[d] % self.mvec[d]) * self.long_prod[d + 1]
            self.poly_to_cube[i] = j
            self.cube_to_poly[j] = i
        self.short_to_long = [0] * self.phim                                     # computeShortToLongMap (:92-112)
        for i in range(self.phim):
            j = 0
            for d in range(k):
                j += ((i // self.short_prod[d + 1]) % self.phivec[d]) * self.long_prod[d + 1]
            self.short_to_long[i] = j
        self.cyc = [cyclotomic_poly(f) for f in self.mvec]
        self.phimx = cyclotomic_poly(self.m)


def _poly_rem(a, b):
    """a mod b over Z for monic b (coefficient lists, low -> high)."""
    a = list(a)
    db = len(b) - 1
    for i in range(len(a) - 1, db - 1, -1):
        c = a[i]
        if c:
            for j in range(db + 1):
                a[i - db + j] -= c * b[j]
    return a[:db] + [0] * max(0, db - len(a))


def poly_to_powerful(ix: PowerfulIndexes, poly):
    """PowerfulConversion::polyToPowerful over Z (src/powerful.cpp:203-221 with recursiveReduce :113-148): scatter by
    the CRT map, reduce every hypercolumn of dimension d modulo Phi_{m_d}, read off the short cube."""
    cube = [0] * ix.m
    for i, c in enumerate(poly):
        cube[ix.poly_to_cube[i]] = c
    k = len(ix.mvec)
    for d in range(k):
        stride = ix.long_prod[d + 1]
        for base in range(ix.m):
            if (base // stride) % ix.mvec[d] != 0:
                continue
            col = [cube[base + t * stride] for t in range(ix.mvec[d])]
            r = _poly_rem(col, ix.cyc[d])
            for t in range(ix.mvec[d]):
                cube[base + t * stride] = r[t] if t < len(r) else 0
    return [cube[ix.short_to_long[i]] for i in range(ix.phim)]


def powerful_to_poly(ix: PowerfulIndexes, pw):
    """PowerfulConversion::powerfulToPoly over Z (src/powerful.cpp:223-244)."""
    tmp = [0] * ix.m
    for i in range(ix.phim):
        tmp[ix.cube_to_poly[ix.short_to_long[i]]] = pw[i]
    return _poly_rem(tmp, ix.phimx)[:ix.phim]


def raw_mod_switch(ix, parts_coeffs, Q: int, q: int, p2r: int, coin=lambda: 0):
    """Ctxt::rawModSwitch (src/Ctxt.cpp:2949-3046).  parts_coeffs: per part the balanced coefficients mod Q in the
    polynomial basis (what DoubleCRT::toPoly returns); ix = None when the powerful basis is the polynomial one
    (PowerfulDCRT::triv).  Returns the parts as small integer polynomials."""
    from math import gcd
    assert q > 1 and p2r > 1 and gcd(q, p2r) == 1 and gcd(Q % q, q) == 1
    q_half = Q // 2
    q_inv_p = pow(Q % p2r, -1, p2r)
    out = []
    for coeffs in parts_coeffs:
        pw = [bal(c, Q) for c in poly_to_powerful(ix, coeffs)] if ix is not None else list(coeffs)   # dcrtToPowerful (:393-410)
        res = []
        for c in pw:
            cq = c * q
            X, Y = divmod(cq, Q)                 # floor division, Y in [0,Q)
            if Y > q_half:
                Y -= Q
                X += 1
            delta = (Y % p2r) * q_inv_p % p2r
            if delta > p2r // 2 or (p2r % 2 == 0 and delta == p2r // 2 and (Y < 0 or (Y == 0 and coin()))):
                delta -= p2r
            x = X + delta
            if x > q // 2 or (q % 2 == 0 and x == q // 2 and coin()):
                x -= q
            elif x < -(q // 2) or (q % 2 == 0 and x == -(q // 2) and coin()):
                x += q
            res.append(x)
        out.append(powerful_to_poly(ix, res) if ix is not None else res)
    return out


# ---------------------------------------------------------------------------------------------
# DoubleCRT::randomize (src/DoubleCRT.cpp:1258-1378): rejection sampling of uniform rows from a byte stream.
# The stream itself is NTL's RandomStream (ChaCha20 keyed from SetSeed; NTL 11.4.3, not in the reference tree) and is an
# INPUT here ("parity unpinned" for the bytes; the consumption pattern is the reference's).

def randomize_rows(chain, idxs, get_bytes):
    """get_bytes(n) -> n fresh bytes.  Per row: a fresh 2048-byte buffer at the start of every refill, nb = ceil(k/8)
    little-endian bytes per candidate (k = bits of q-1), masked to k bits, accepted when < q (src/DoubleCRT.cpp:1279-1376)."""
    bufsz = 2048
    rows = {}
    for i in sorted(idxs):
        q = chain.primes[i]
        k = (q - 1).bit_length()
        nb = (k + 7) // 8
        mask = (1 << k) - 1
        row = []
        while len(row) < chain.phim:
            buf = get_bytes(bufsz)
            pos = 0
            while pos <= bufsz - nb and len(row) < chain.phim:
                v = int.from_bytes(buf[pos:pos + nb], "little") & mask
                if v < q:
                    row.append(v)
                pos += nb
        rows[i] = row
    return rows


# ---------------------------------------------------------------------------------------------
# ModuliSizes (src/primeChain.cpp:66-319): the table of candidate prime sets and the two getSet4Size searches that
# Ctxt::multLowLvl / modDownToLevel use to pick the common prime set from noise estimates.

class ModuliSizes:
    def __init__(self, chain):
        """ModuliSizes::init (src/primeChain.cpp:66-121): every subset of the small primes, alone and joined with every
        prefix interval of the ctxt primes; sorted by log-size."""
        self.ifft_cost = 0 if chain.pow2 else 20
        sizes = [(0.0, frozenset())]
        idx = 1
        for i in chain.small:
            sz = math.log(chain.primes[i])
            for j in range(idx, 2 * idx):
                f, s = sizes[j - idx]
                sizes.append((f + sz, s | {i}))
            idx *= 2
        interval = set()
        interval_size = 0.0
        for i in chain.ctxt:
            interval.add(i)
            interval_size += math.log(chain.primes[i])
            for j in range(idx):
                f, s = sizes[j]
                sizes.append((f + interval_size, s | interval))
        self.sizes = sorted(sizes, key=lambda e: (e[0], sorted(e[1])))

    def _cost(self, frm, to):
        """cost_estimate (src/primeChain.cpp:146-157)."""
        add = len(to - frm)
        return 100 * add if self.ifft_cost == 0 else 100 * add + self.ifft_cost * len(frm - to)

    def get_set4size(self, low, high, from1, from2=None, reverse=False):
        """ModuliSizes::getSet4Size, one- and two-operand forms (src/primeChain.cpp:179-243, 250-319)."""
        froms = [frozenset(from1)] + ([frozenset(from2)] if from2 is not None else [])
        n = len(self.sizes)
        idx = 0
        while idx < n and self.sizes[idx][0] < low:
            idx += 1
        best, best_cost = -1, None
        ii = idx
        while ii < n and self.sizes[ii][0] <= high:
            cost = sum(self._cost(f, self.sizes[ii][1]) for f in froms)
            if best_cost is None or cost <= best_cost:
                best, best_cost = ii, cost
            ii += 1
        if best == -1:
            if reverse:
                if ii < n:
                    upper = self.sizes[ii][0] + math.log(2.0)
                    i = ii
                    while i < n and self.sizes[i][0] <= upper:
                        cost = sum(self._cost(f, self.sizes[i][1]) for f in froms)
                        if best_cost is None or cost < best_cost:
                            best, best_cost = i, cost
                        i += 1
            elif idx > 0:
                lower = self.sizes[idx - 1][0] - math.log(2.0)
                i = idx - 1
                while i >= 0 and self.sizes[i][0] >= lower:
                    cost = sum(self._cost(f, self.sizes[i][1]) for f in froms)
                    if best_cost is None or cost < best_cost:
                        best, best_cost = i, cost
                    i -= 1
        return sorted(self.sizes[best][1]) if best != -1 else []
