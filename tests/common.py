"""Shared helpers: build (chain, oracle, engine) triples for the parity tests."""
import numpy as np

import orc
import pyoracle as po
from helib_b200.engine import Engine

_CHAINS = {}


def chain(m, p, r, bits, c):
    key = (m, p, r, bits, c)
    if key not in _CHAINS:
        ch = po.build_mod_chain(m, p, r, bits, c)
        psis = [po.find_psi(q, m) for q in ch.primes]
        _CHAINS[key] = (ch, psis)
    return _CHAINS[key]


def make(lib, m, p, r, bits, c, nthreads=4):
    ch, psis = chain(m, p, r, bits, c)
    O = orc.Oracle(ch.phim, ch.m, ch.primes, psis, ch.digits, ch.special, nthreads=nthreads)
    E = Engine(m, ch.primes, psis, ch.digits, ch.special, lib=lib)
    return ch, psis, O, E


def rows_equal(a, b, idx):
    idx = list(idx)
    return bool((a[idx] == b[idx]).all())


def ptxt_space(ch):
    return 1 if ch.p == -1 else ch.p ** ch.r
