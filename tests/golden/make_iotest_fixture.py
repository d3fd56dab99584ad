#!/usr/bin/env python
"""Extract the DoubleCRT rows of the reference's own I/O fixtures (tests/test_resources/iotest_ascii{LE,BE}.txt:
a context for m=12, p=7 with five chain primes, the public encryption key and the secret key written by a real
HElib build) into tests/golden/helib_iotest_m12.json.  Run in the build container only (reads /root/reference);
the tests read the JSON.  These are the only evaluation-form rows produced by the reference itself in its tree,
so they pin the oracle's general-m conventions: the root chosen by FindPrimitiveRoot and the row order over Z_m^*."""
import json
import os
import re
import sys

SRC = "/root/reference/tests/test_resources"


def parse_file(path):
    toks = re.findall(r"\[|\]|-?\d+\.\d+|-?\d+", open(path).read())

    def parse(pos):
        out = []
        while pos < len(toks):
            t = toks[pos]
            if t == "[":
                sub, pos = parse(pos + 1)
                out.append(sub)
            elif t == "]":
                return out, pos + 1
            else:
                out.append(float(t) if "." in t else int(t))
                pos += 1
        return out, pos
    return parse(0)[0]


def rows_of(dcrt):
    return {str(i): dcrt[1 + k] for k, i in enumerate(dcrt[0])}


def main():
    out = {"source": "reference tests/test_resources/iotest_asciiLE.txt, iotest_asciiBE.txt (legacy ASCII I/O fixtures)", "cases": []}
    for name in ("iotest_asciiLE.txt", "iotest_asciiBE.txt"):
        tree = parse_file(os.path.join(SRC, name))
        m, p, r = tree[0][0], tree[0][1], tree[0][2]
        ctx = tree[1]
        nprimes = ctx[2]
        primes = ctx[3:3 + nprimes]
        special = ctx[1]
        enc = tree[2][1]                       # [ptxtSpace, noise, primeSet, nparts, part, part]
        parts = enc[4:]
        blocks = []

        def rec(x):
            if isinstance(x, list):
                if len(x) == nprimes + 1 and isinstance(x[0], list) and x[0] == list(range(nprimes)) and all(isinstance(v, list) for v in x[1:]):
                    blocks.append(x)
                for y in x:
                    rec(y)
        rec(tree)
        out["cases"].append({
            "file": name, "m": m, "p": p, "r": r, "primes": primes, "special": special,
            "pk_ptxt_space": enc[0], "pk_prime_set": enc[2],
            "pk_c0": rows_of(parts[0][0]), "pk_c0_handle": parts[0][1],
            "pk_c1": rows_of(parts[1][0]), "pk_c1_handle": parts[1][1],
            "secret_key": rows_of(blocks[-1]),
            # chain layout of the legacy context record: stdev, special primes, #primes, primes, #digits, digits
            "digits": [d for d in ctx[3 + nprimes + 1:3 + nprimes + 1 + ctx[3 + nprimes]]],
            # key-switching matrices W[s^r(X^t) -> s]: [[r, t, fromID], toID, ptxtSpace, n, b_0 .. b_{n-1}, prgSeed]
            # (KeySwitch::writeTo, src/keySwitching.cpp): the b_i rows as written by real HElib, the 256-bit seed of the a_i as a decimal string
            "ksw": [{"from": mt[0], "to": mt[1], "ptxt_space": mt[2], "n": mt[3],
                     "b": [rows_of(mt[4 + d]) for d in range(mt[3])], "prg_seed": str(mt[4 + mt[3]])}
                    for mt in tree[2][4:4 + tree[2][3]]],
        })
    # the binary twin of the LE file holds the same key; its first ciphertext part is a DoubleCRT::writeTo record
    # (IndexSet, then per row int32 length, int32 intSize, little-endian int64 values) written by the real library
    import struct
    raw = open(os.path.join(SRC, "iotest_binLE.bin"), "rb").read()
    cx = raw.find(b"|CX[")
    off = cx + 4 + 8 + 16            # eye catcher, ptxtSpace, noise bound (xdouble = double + int64)
    card = struct.unpack("<q", raw[off:off + 8])[0]
    off += 8 + 8 * card + 8          # the ciphertext's prime set, number of parts
    card = struct.unpack("<q", raw[off:off + 8])[0]
    end = off + 8 + 8 * card
    for _ in range(card):
        ln, isz = struct.unpack("<ii", raw[end:end + 8])
        end += 8 + ln * isz
    out["dcrt_record"] = {"file": "iotest_binLE.bin", "offset": off, "hex": raw[off:end].hex(),
                          "noise_bound_field_hex": raw[cx + 12:cx + 28].hex(),
                          "note": "same rows as cases[0].pk_c0 (iotest_asciiLE.txt); the 16-byte field is the xdouble noise bound 2007.04 = (double mantissa, int64 exponent 0)"}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helib_iotest_m12.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    sys.exit(main())
