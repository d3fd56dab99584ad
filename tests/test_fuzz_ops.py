"""Differential fuzz of the DoubleCRT operations: random index sets and random sequences of addPrimes / addPrimesAndScale /
scaleDownToSet (several plaintext spaces) / automorph / scale-by-primes / pointwise / transform round trips / breakIntoDigits on
the CPU kernel-logic simulator, every intermediate state compared row for row with the C++ oracle.  The index sets the parity
tests use are the ones the multiply path produces; this walks through arbitrary subsets of the chain (one source prime, more
sources than targets, small primes mixed with special ones), which is what DoubleCRT's interface allows
(src/DoubleCRT.cpp:565-647,1464-1516)."""
import random
from math import gcd

import numpy as np
import pytest

from common import make

SMALL = [(64, 257, 1, 200, 3), (2048, 17, 2, 250, 3), (4096, 257, 1, 300, 2), (8192, -1, 1, 300, 3), (8192, 2, 3, 200, 2)]
BIG = [(1 << 17, 257, 1, 330, 3), (1 << 17, -1, 1, 400, 2)]


def walk(lib, cfg, rnd, nops):
    m, p, r, bits, c = cfg
    ch, psis, O, E = make(lib, *cfg, nthreads=8)
    rng = np.random.default_rng(rnd.randrange(1 << 30))
    allp = list(range(len(ch.primes)))
    p2r = 1 if p == -1 else p ** r
    cur = sorted(rnd.sample(allp, rnd.randint(1, len(allp))))
    x = O.random(rng, cur)
    P = E.poly(x, cur)
    log = [("init", cfg, cur)]
    done = 0
    while done < nops:
        op = rnd.choice(["add_primes", "add_primes_and_scale", "scale_down", "scale_down", "automorph", "scale_by_primes", "pointwise", "roundtrip", "digits"])
        if op in ("add_primes", "add_primes_and_scale"):
            rest = [i for i in allp if i not in cur]
            if not rest:
                continue
            add = sorted(rnd.sample(rest, rnd.randint(1, len(rest))))
            log.append((op, list(cur), add))
            getattr(E, op)([P], cur, add)
            getattr(O, op)(x, cur, add)
            cur = sorted(cur + add)
        elif op == "scale_down":
            if len(cur) < 2:
                continue
            keep = sorted(rnd.sample(cur, rnd.randint(1, len(cur) - 1)))
            ps = 1 if p == -1 else rnd.choice([1, p2r, p2r, 2, 3, 4, 65537])
            if any(gcd(ps, ch.primes[i]) != 1 for i in cur if i not in keep):
                continue
            log.append((op, list(cur), keep, ps))
            E.scale_down([P], cur, keep, ps)
            O.scale_down(x, cur, keep, ps)
            cur = keep
        elif op == "automorph":
            k = rnd.randrange(1, m, 2)
            log.append((op, list(cur), k))
            Pd = E.poly()
            E.automorph([Pd], [P], cur, k)
            O.automorph(x, cur, k)
            P = Pd
        elif op == "scale_by_primes":
            f = sorted(rnd.sample(allp, rnd.randint(1, min(4, len(allp)))))
            inv = rnd.random() < 0.5
            if inv and any(i in cur for i in f):
                continue                                   # q_i has no inverse modulo itself
            log.append((op, list(cur), f, inv))
            E.scale_by_primes([P], cur, f, inv)
            O.scale_by_primes(x, cur, f, inv)
        elif op == "pointwise":
            y = O.random(rng, cur)
            o = rnd.choice(["add", "sub", "mul"])
            log.append((op, list(cur), o))
            E.pointwise(o, [P], [E.poly(y, cur)], cur)
            O.pointwise(o, x, y, cur)
        elif op == "roundtrip":
            sub = sorted(rnd.sample(cur, rnd.randint(1, len(cur))))
            log.append((op, sub))
            E.ntt_inv([P], sub)
            E.ntt_fwd([P], sub)
        else:
            if not cur or any(i not in ch.ctxt for i in cur):
                continue
            log.append((op, list(cur)))
            ref = O.break_into_digits(x, cur)
            full = sorted(set(cur) | set(ch.special))
            D = E.break_into_digits([P], cur)[0]
            assert len(D) == len(ref), log
            for di in range(len(ref)):
                assert (D[di].download(full)[full] == ref[di][full]).all(), (di, log)
            done += 1
            continue
        done += 1
        got = P.download(cur)
        bad = [i for i in cur if not (got[i] == x[i]).all()]
        assert not bad, (bad, log)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_walks_small_rings(sim_lib, seed):
    rnd = random.Random(seed)
    for _ in range(6):
        walk(sim_lib, rnd.choice(SMALL), rnd, 25)


def test_random_walk_full_size_ring(sim_lib):
    """N = 2^16: the register-blocked and TMA-staged kernels and the fused conversions with arbitrary source/target splits."""
    rnd = random.Random(11)
    for cfg in BIG:
        walk(sim_lib, cfg, rnd, 14)


GENERAL = [(12, 7, 1, 150, 2), (45, 2, 1, 150, 2), (28, 3, 1, 150, 3), (105, 2, 1, 180, 2), (45, 7, 2, 150, 2), (63, 2, 1, 150, 2), (85, 2, 1, 150, 2),
           (30, 7, 1, 150, 2)]   # m = 30: the division by Phi_m needs the full chirp length (2(m - phi(m)) - 1 > m)


@pytest.mark.parametrize("seed", [1, 2])
def test_random_walks_general_m(sim_lib, seed):
    """The same for Bluestein rows (general m, incl. rings no other test uses: m = 63, 85), against a big-integer model:
    toPoly by CRT over the current set, addPrimes = residues of the balanced polynomial, scaleDownToSet with the reference's
    delta correction and tie rule, automorph as the permutation of Z_m^*."""
    import pyoracle as po
    from helib_b200.engine import Engine
    rnd = random.Random(seed)
    for _ in range(8):
        m, p, r, bits, c = rnd.choice(GENERAL)
        ch = po.build_mod_chain(m, p, r, bits, c)
        roots = [po.cmod_root(q, m) for q in ch.primes]
        E = Engine(m, ch.primes, None, ch.digits, ch.special, lib=sim_lib)
        n, allp, rep = ch.phim, list(range(len(ch.primes))), po.zms_rep(m)

        def to_poly(rows, idx):
            Q = ch.product(idx)
            cs = {i: po.gen_ifft(rows[i], ch.primes[i], m, roots[i]) for i in idx}
            out = []
            for k in range(n):
                acc = 0
                for i in idx:
                    q = ch.primes[i]
                    Qi = Q // q
                    acc += Qi * (cs[i][k] * pow(Qi % q, -1, q) % q)
                out.append(po.bal(acc, Q))
            return out

        def rows_of_poly(poly, idx):
            return {i: po.gen_fft([cc % ch.primes[i] for cc in poly], ch.primes[i], m, roots[i]) for i in idx}

        def dense(rows):
            out = np.zeros((len(ch.primes), n), dtype=np.uint64)
            for i, rr in rows.items():
                out[i] = np.array(rr, dtype=np.uint64)
            return out

        cur = sorted(rnd.sample(allp, rnd.randint(1, len(allp))))
        x = {i: [rnd.randrange(ch.primes[i]) for _ in range(n)] for i in cur}
        P = E.poly(dense(x), cur)
        log = [("init", (m, p, r, bits, c), cur)]
        done = 0
        while done < 6:
            op = rnd.choice(["add_primes", "scale_down", "scale_down", "automorph", "roundtrip"])
            if op == "add_primes":
                rest = [i for i in allp if i not in cur]
                if not rest:
                    continue
                add = sorted(rnd.sample(rest, rnd.randint(1, len(rest))))
                log.append((op, list(cur), add))
                x.update(rows_of_poly(to_poly(x, cur), add))
                E.add_primes([P], cur, add)
                cur = sorted(cur + add)
            elif op == "scale_down":
                if len(cur) < 2:
                    continue
                keep = sorted(rnd.sample(cur, rnd.randint(1, len(cur) - 1)))
                ps = rnd.choice([1, p ** r, p ** r, 2, 3, 4, 65537])
                diff = [i for i in cur if i not in keep]
                if any(gcd(ps, ch.primes[i]) != 1 for i in diff):
                    continue
                log.append((op, list(cur), keep, ps))
                Pd = ch.product(diff)
                delta = to_poly(x, diff)
                if ps > 1:
                    pinv = pow(Pd % ps, -1, ps)
                    for k, d in enumerate(delta):
                        u = d % ps
                        if u:
                            u = u * pinv % ps
                            if u > ps // 2 or (ps % 2 == 0 and u == ps // 2 and d < 0):
                                u -= ps
                            delta[k] = d - Pd * u
                drows = rows_of_poly(delta, keep)
                x = {i: [((a - b) * pow(Pd % ch.primes[i], -1, ch.primes[i])) % ch.primes[i] for a, b in zip(x[i], drows[i])] for i in keep}
                E.scale_down([P], cur, keep, ps)
                cur = keep
            elif op == "automorph":
                k = rnd.choice([t for t in range(1, m) if gcd(t, m) == 1])
                log.append((op, list(cur), k))
                x = {i: [x[i][rep.index(rep[j] * k % m)] for j in range(n)] for i in cur}
                D = E.poly()
                E.automorph([D], [P], cur, k)
                P = D
            else:
                sub = sorted(rnd.sample(cur, rnd.randint(1, len(cur))))
                log.append((op, sub))
                E.ntt_inv([P], sub)
                E.ntt_fwd([P], sub)
            done += 1
            got = P.download(cur)
            bad = [i for i in cur if [int(v) for v in got[i]] != [int(v) for v in x[i]]]
            assert not bad, (bad, log)


def oracle_mul_relin_moddown(O, ch, a0, a1, b0, b1, S_in, S, p, evk_a, evk_b):
    parts = [x.copy() for x in (a0, a1, b0, b1)]
    for x in parts:
        O.scale_down(x, S_in, S, p)
    t0, t1, t2 = O.tensor(parts[0], parts[1], parts[2], parts[3], S)
    r0, r1 = O.relinearize(t0, t1, t2, S, evk_a, evk_b)
    Sp = sorted(S + ch.special)
    O.scale_down(r0, Sp, S, p)
    O.scale_down(r1, Sp, S, p)
    return r0, r1


def test_multiply_on_level_sets_with_holes(sim_lib):
    """Ctxt x Ctxt on arbitrary subsets of the ctxt primes.  A set that skips every prime of a middle digit makes breakIntoDigits
    carry a zero digit whose full product still divides the later ones (src/DoubleCRT.cpp:488-493,509-561); the fused
    relinearisation converts from a digit's own rows and used to run off the end of an empty one (found by this walk at
    m = 2^17, bits = 500, c = 4, S = {7, 12, 13})."""
    from common import ptxt_space
    rnd = random.Random(5)
    for cfg, fixed in (((1 << 17, 257, 1, 500, 4), [[7, 12, 13], [6, 13]]), ((4096, 257, 1, 300, 3), [])):
        ch, psis, O, E = make(sim_lib, *cfg, nthreads=8)
        p = ptxt_space(ch)
        rng = np.random.default_rng(rnd.randrange(1 << 30))
        full = ch.ctxt + ch.special
        nd = len(ch.digits)
        evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
        evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
        EA = [E.poly(evk_a[i], full) for i in range(nd)]
        EB = [E.poly(evk_b[i], full) for i in range(nd)]
        cases = [(s, s) for s in fixed if all(i in ch.ctxt for i in s)]
        for _ in range(1 if cfg[0] > 4096 else 3):
            S_in = sorted(rnd.sample(ch.ctxt, rnd.randint(2, len(ch.ctxt))))
            cases.append((S_in, sorted(rnd.sample(S_in, len(S_in) - rnd.randint(0, min(2, len(S_in) - 1))))))
        for S_in, S in cases:
            o = [O.random(rng, S_in) for _ in range(4)]
            A0, A1, B0, B1 = ([E.poly(o[k], S_in)] for k in range(4))
            E.mul_relin_moddown(A0, A1, B0, B1, S_in, S, p, EA, EB)
            r0, r1 = oracle_mul_relin_moddown(O, ch, *o, S_in, S, p, evk_a, evk_b)
            assert (A0[0].download(S)[S] == r0[S]).all() and (A1[0].download(S)[S] == r1[S]).all(), (cfg, S_in, S)


def test_relinearize_with_six_digits(sim_lib):
    """c = 6 at N = 2^16: more key-switching columns than the fused inner product is instantiated for (four); the call used to
    answer HB_ERR_UNSUPPORTED on this ring while smaller rings accepted the same chain shape."""
    cfg = (1 << 17, 257, 1, 600, 6)
    ch, psis, O, E = make(sim_lib, *cfg, nthreads=8)
    assert len(ch.digits) == 6
    rng = np.random.default_rng(8)
    full, S = ch.ctxt + ch.special, ch.ctxt
    Sp = sorted(full)
    nd = len(ch.digits)
    evk_a = np.stack([O.random(rng, full) for _ in range(nd)])
    evk_b = np.stack([O.random(rng, full) for _ in range(nd)])
    EA = [E.poly(evk_a[i], full) for i in range(nd)]
    EB = [E.poly(evk_b[i], full) for i in range(nd)]
    c = [O.random(rng, S) for _ in range(3)]
    C0, C1, C2 = ([E.poly(c[k], S)] for k in range(3))
    E.relinearize(C0, C1, C2, S, EA, EB)
    r0, r1 = O.relinearize(c[0], c[1], c[2], S, evk_a, evk_b)
    assert (C0[0].download(Sp)[Sp] == r0[Sp]).all() and (C1[0].download(Sp)[Sp] == r1[Sp]).all()
