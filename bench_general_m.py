#!/usr/bin/env python
"""bench_general_m.py -- row-level throughput on BASELINE config 5's ring (m = 21845, phi(m) = 16384, Bluestein
rows): forward/inverse transforms and a full relinearise + mod-down at the thin-bootstrapping chain
(p=2, bits=580, c=2; tests/GTestThinBootstrapping.cpp:102).  End-to-end recryption is out of scope;
this reports what the accelerated DoubleCRT layer delivers underneath it.  1 GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from helib_b200 import Chain, Engine
    m, p, r, bits, c = 21845, 2, 1, 580, 2
    ch = Chain(m, p, r, bits, c)
    E = Engine(m, ch.primes, None, ch.digits, ch.special, device=0)
    E.set_stream(torch.cuda.current_stream().cuda_stream)
    N, npr = E.N, E.np
    S, full = ch.ctxt, ch.ctxt + ch.special
    nd = len(ch.digits)
    rng = np.random.default_rng(5)
    B = 8

    def rand(idx):
        out = np.zeros((npr, N), dtype=np.uint64)
        for i in idx:
            out[i] = rng.integers(0, ch.primes[i], size=N, dtype=np.uint64)
        return out

    EA = [E.poly(rand(full), full) for _ in range(nd)]
    EB = [E.poly(rand(full), full) for _ in range(nd)]
    C = [[E.poly(rand(S), S) for _ in range(3)] for _ in range(B)]
    C0, C1, C2 = ([x[k] for x in C] for k in range(3))
    Sp = sorted(full)

    def timed(fn, steps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    t_f = timed(lambda: E.ntt_fwd(C0, S))
    t_i = timed(lambda: E.ntt_inv(C0, S))

    def ks():
        E.relinearize(C0, C1, C2, S, EA, EB)
        E.scale_down(C0 + C1, Sp, S, p)
    t_ks = timed(ks, 3)
    rows = B * len(S)
    print(json.dumps({"metric": "general_m_rows", "m": m, "phim": N, "bluestein_length": 65536, "division_length": 1 << (max(N, 2 * (m - N) - 1) - 1).bit_length(), "primes": {"ctxt": len(S), "special": len(ch.special), "digits": nd},
                      "fwd_rows_per_s": rows / (t_f / 1e3), "inv_rows_per_s": rows / (t_i / 1e3),
                      "relin_moddown_per_s": B / (t_ks / 1e3), "ms": {"fwd": t_f, "inv": t_i, "relin_moddown": t_ks}, "batch": B, "n_gpus": 1}))


if __name__ == "__main__":
    main()
