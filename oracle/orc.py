"""ctypes binding of the C++ CPU oracle (oracle/oracle.cpp -> liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by helib_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_ctx_create.restype = C.c_void_p
    return _LIB


def _u64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _idx(idx):
    arr = np.ascontiguousarray(np.asarray(list(idx), dtype=np.int32))
    return arr, arr.ctypes.data_as(i32p), len(arr)


def limbs_to_ints(a: np.ndarray):
    """[N][L] two's complement little-endian limbs -> python ints."""
    n, L = a.shape
    out = []
    for k in range(n):
        v = 0
        for l in range(L - 1, -1, -1):
            v = (v << 64) | int(a[k, l])
        if int(a[k, L - 1]) >> 63:
            v -= 1 << (64 * L)
        out.append(v)
    return out


def ints_to_limbs(vals, L: int) -> np.ndarray:
    out = np.zeros((len(vals), L), dtype=np.uint64)
    mask = (1 << 64) - 1
    for k, v in enumerate(vals):
        v &= (1 << (64 * L)) - 1
        for l in range(L):
            out[k, l] = (v >> (64 * l)) & mask
    return out


class Oracle:
    """Holds a chain (primes, psi, digit partition) and applies DoubleCRT operations to
    dense [nprimes][N] uint64 matrices in which only the rows named by an index list are live."""

    def __init__(self, N, m, primes, psis, digits=None, special=None, nthreads=1):
        self.N, self.m = int(N), int(m)
        if self.m < 4 or self.m & (self.m - 1) or self.N != self.m // 2:
            raise ValueError("the C++ oracle restates the power-of-two path only (m = 2N); general m is checked against pyoracle")
        self.primes = [int(q) for q in primes]
        self.psis = [int(p) for p in psis]
        self.np = len(self.primes)
        q = np.array(self.primes, dtype=np.uint64)
        ps = np.array(self.psis, dtype=np.uint64)
        self.h = C.c_void_p(lib().orc_ctx_create(C.c_long(self.N), C.c_long(self.m), self.np, _u64(q), _u64(ps), nthreads))
        self.digits = [list(d) for d in (digits or [])]
        self.special = list(special or [])
        digit_of = np.full(self.np, -1, dtype=np.int32)
        for d, lst in enumerate(self.digits):
            for i in lst:
                digit_of[i] = d
        sp = np.ascontiguousarray(np.array(self.special, dtype=np.int32))
        lib().orc_ctx_set_chain(self.h, digit_of.ctypes.data_as(i32p), len(self.digits), sp.ctypes.data_as(i32p), len(sp))

    def __del__(self):
        try:
            lib().orc_ctx_destroy(self.h)
        except Exception:
            pass

    def set_threads(self, n):
        lib().orc_set_threads(self.h, int(n))

    def zeros(self):
        return np.zeros((self.np, self.N), dtype=np.uint64)

    def random(self, rng, idx):
        """Uniform residues in [0,q_i) on rows idx (as DoubleCRT::randomize produces,
        src/DoubleCRT.cpp:1365-1371)."""
        out = self.zeros()
        for i in idx:
            out[i] = rng.integers(0, self.primes[i], size=self.N, dtype=np.uint64)
        return out

    def ntt_fwd_rows(self, data, idx):
        a, p, n = _idx(idx)
        lib().orc_ntt_fwd_rows(self.h, _u64(data), p, n)

    def ntt_inv_rows(self, data, idx):
        a, p, n = _idx(idx)
        lib().orc_ntt_inv_rows(self.h, _u64(data), p, n)

    def to_poly(self, data, idx, positive=False, L=None):
        a, p, n = _idx(idx)
        L = L or (n + 1)
        out = np.zeros((self.N, L), dtype=np.uint64)
        lib().orc_to_poly(self.h, _u64(data), p, n, int(positive), _u64(out), L)
        return out

    def fft_bigpoly(self, limbs, idx, data):
        a, p, n = _idx(idx)
        lib().orc_fft_bigpoly(self.h, _u64(limbs), limbs.shape[1], p, n, _u64(data))

    def pointwise(self, op, dst, src, idx):
        a, p, n = _idx(idx)
        lib().orc_pointwise(self.h, {"add": 0, "sub": 1, "mul": 2}[op], _u64(dst), _u64(src), p, n)

    def scale_by_primes(self, data, idx, fidx, inv=False):
        a, p, n = _idx(idx)
        b, pf, nf = _idx(fidx)
        lib().orc_scale_by_primes(self.h, _u64(data), p, n, pf, nf, int(inv))

    def scale_by_word(self, data, idx, scalar):
        a, p, n = _idx(idx)
        lib().orc_scale_by_word(self.h, _u64(data), p, n, C.c_uint64(int(scalar)))

    def add_primes(self, data, cur, add, want_poly=False):
        a, pc, nc = _idx(cur)
        b, pa, na = _idx(add)
        if want_poly:
            L = nc + 1
            poly = np.zeros((self.N, L), dtype=np.uint64)
            lib().orc_add_primes(self.h, _u64(data), pc, nc, pa, na, _u64(poly), L)
            return poly
        lib().orc_add_primes(self.h, _u64(data), pc, nc, pa, na, None, 0)

    def add_primes_and_scale(self, data, cur, add):
        a, pc, nc = _idx(cur)
        b, pa, na = _idx(add)
        lib().orc_add_primes_and_scale(self.h, _u64(data), pc, nc, pa, na)

    def scale_down(self, data, cur, keep, ptxt_space=1, want_delta=False):
        a, pc, nc = _idx(cur)
        b, pk, nk = _idx(keep)
        if want_delta:
            L = (nc - nk) + 2
            delta = np.zeros((self.N, L), dtype=np.uint64)
            lib().orc_scale_down(self.h, _u64(data), pc, nc, pk, nk, C.c_long(ptxt_space), _u64(delta), L)
            return delta
        lib().orc_scale_down(self.h, _u64(data), pc, nc, pk, nk, C.c_long(ptxt_space), None, 0)

    def break_into_digits(self, data, cur, want_polys=False):
        a, pc, nc = _idx(cur)
        nd = len(self.digits)
        out = np.zeros((nd, self.np, self.N), dtype=np.uint64)
        L = max(len(d) for d in self.digits) + 1
        polys = np.zeros((nd, self.N, L), dtype=np.uint64) if want_polys else None
        n = lib().orc_break_into_digits(self.h, _u64(data), pc, nc, _u64(out), _u64(polys) if want_polys else None, L)
        return (out[:n], polys[:n]) if want_polys else out[:n]

    def keyswitch_digits(self, digits, idx, evk_a, evk_b, out0, out1):
        a, p, n = _idx(idx)
        digits = np.ascontiguousarray(digits)
        lib().orc_keyswitch_digits(self.h, _u64(digits), digits.shape[0], p, n, _u64(evk_a), _u64(evk_b), _u64(out0), _u64(out1))

    def automorph(self, data, idx, k):
        a, p, n = _idx(idx)
        lib().orc_automorph(self.h, _u64(data), p, n, C.c_long(k))

    def tensor(self, a0, a1, b0, b1, idx):
        a, p, n = _idx(idx)
        o0, o1, o2 = self.zeros(), self.zeros(), self.zeros()
        lib().orc_tensor(self.h, _u64(a0), _u64(a1), _u64(b0), _u64(b1), _u64(o0), _u64(o1), _u64(o2), p, n)
        return o0, o1, o2

    # ---- composites used by tests and the CPU baseline (host orchestration restated from
    #      Ctxt::reLinearize / keySwitchPart, src/Ctxt.cpp:720-842)
    def relinearize(self, c0, c1, c2, S, evk_a, evk_b):
        """3-part (1, s, s^2) ciphertext over ctxt primes S -> 2-part over S | special."""
        Sp = sorted(set(S) | set(self.special))
        c0 = c0.copy(); c1 = c1.copy()
        self.add_primes_and_scale(c0, S, self.special)
        self.add_primes_and_scale(c1, S, self.special)
        digits = self.break_into_digits(c2, S)
        self.keyswitch_digits(digits, Sp, evk_a, evk_b, c0, c1)
        return c0, c1
