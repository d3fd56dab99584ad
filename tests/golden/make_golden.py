#!/usr/bin/env python
"""Generates tests/golden/*.json from oracle/pyoracle.py (pure-Python big-int restatement).

The reference (HElib) holds no golden integer vectors for the DoubleCRT path and cannot be built
in this environment (NTL/GMP are fetched at configure time), so these fixtures pin the *restated*
semantics: any later change to the C++ oracle or the CUDA engine must keep reproducing them.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle as po  # noqa: E402

CASES = {"bgv_m16_p17": (16, 17, 1, 80, 2), "bgv_m32_p257_c3": (32, 257, 1, 150, 3), "ckks_m32": (32, -1, 1, 100, 2)}

for name, (m, p, r, bits, c) in CASES.items():
    ch = po.build_mod_chain(m, p, r, bits, c)
    psis = [po.find_psi(q, m) for q in ch.primes]
    rnd = random.Random(name)
    S = ch.ctxt
    x = {i: [rnd.randrange(ch.primes[i]) for _ in range(ch.phim)] for i in S}
    X = po.PyDCRT(ch, psis, x)
    digs, _ = X.break_into_digits()
    cur = S + ch.special
    y = {i: [rnd.randrange(ch.primes[i]) for _ in range(ch.phim)] for i in cur}
    Y = po.PyDCRT(ch, psis, y)
    ps = 1 if p == -1 else p ** r
    Y.scale_down_to_set(S, ps)
    out = {
        "params": [m, p, r, bits, c], "primes": ch.primes, "psis": psis, "digits": ch.digits, "special": ch.special,
        "x": {str(i): v for i, v in x.items()}, "to_poly": po.PyDCRT(ch, psis, x).to_poly(),
        "digits_rows": [{str(i): d.rows[i] for i in d.index_set} for d in digs],
        "y": {str(i): v for i, v in y.items()}, "ptxt_space": ps,
        "scale_down_rows": {str(i): Y.rows[i] for i in S},
    }
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f)
    print("wrote", name)
