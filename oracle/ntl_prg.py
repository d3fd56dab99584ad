"""Restatement of NTL's pseudo-random stream (TEST INFRASTRUCTURE, part of the oracle).

HElib draws the pseudo-random halves a_i of a key-switching matrix from NTL's global RandomStream after
NTL::SetSeed(prgSeed) (reference call sites: src/keys.cpp:1199-1204, src/Ctxt.cpp:199-206,
src/DoubleCRT.cpp:1258-1378).  NTL (pinned: 11.4.3+, CMakeLists.txt:77; not in the reference tree) is an
un-vendored dependency, so its published algorithm is restated here from the NTL sources' documented behaviour:

  SetSeed(ZZ s)           : data = BytesFromZZ(|s|) (little-endian, NumBytes(s) bytes); SetSeed(data)
  SetSeed(bytes)          : key = DeriveKey(32, data); stream = RandomStream(key)
  DeriveKey               : K = HMAC-SHA256(key = "", data); out = HMAC-SHA256(K, counter_le64(0)) || HMAC-SHA256(K, counter(1)) ...
  RandomStream(key)       : ChaCha20 (20 rounds, "expand 32-byte k"), 64-bit block counter starting at 0, zero nonce;
                            get(n) returns the next n key-stream bytes.

PINNING: tests/test_oracle.py regenerates the a_i of the key-switching matrices stored (with their prgSeed) in the
reference's own fixture tests/test_resources/iotest_ascii*.txt and checks  b_i + a_i*s - P*prod_{j<i}Q_j*s'  against
a small multiple of p on every row -- which only holds if every byte of this stream, the consumption pattern of
DoubleCRT::randomize and the key-switching formula (src/keys.cpp:1239-1242) are the reference's.
"""
import hashlib
import hmac
import struct

MASK32 = 0xFFFFFFFF


def derive_key(data: bytes, klen: int = 32) -> bytes:
    K = hmac.new(b"", data, hashlib.sha256).digest()
    out = b""
    counter = 0
    while len(out) < klen:
        out += hmac.new(K, struct.pack("<Q", counter), hashlib.sha256).digest()
        counter += 1
    return out[:klen]


def _rotl(x, n):
    return ((x << n) & MASK32) | (x >> (32 - n))


def _qr(x, a, b, c, d):
    x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 16)
    x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 12)
    x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 8)
    x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 7)


def chacha20_block(key: bytes, counter: int) -> bytes:
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(struct.unpack("<8I", key)) + [counter & MASK32, (counter >> 32) & MASK32, 0, 0]
    x = st[:]
    for _ in range(10):
        _qr(x, 0, 4, 8, 12); _qr(x, 1, 5, 9, 13); _qr(x, 2, 6, 10, 14); _qr(x, 3, 7, 11, 15)
        _qr(x, 0, 5, 10, 15); _qr(x, 1, 6, 11, 12); _qr(x, 2, 7, 8, 13); _qr(x, 3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & MASK32 for a, b in zip(x, st)])


class RandomStream:
    def __init__(self, key: bytes):
        self.key, self.counter, self.buf = key, 0, b""

    def get(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += chacha20_block(self.key, self.counter)
            self.counter += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out


def zz_bytes(seed: int) -> bytes:
    seed = abs(seed)
    return seed.to_bytes((seed.bit_length() + 7) // 8, "little")


def set_seed(seed: int) -> RandomStream:
    """NTL::SetSeed(const ZZ&)."""
    return RandomStream(derive_key(zz_bytes(seed)))


def random_bnd(stream: RandomStream, n: int) -> int:
    """NTL::RandomBnd(long n): l = NumBits(n-1), ceil(l/8) little-endian bytes per trial, masked to l bits, rejection."""
    if n <= 1:
        return 0
    l = (n - 1).bit_length()
    nb = (l + 7) // 8
    mask = (1 << l) - 1
    while True:
        v = int.from_bytes(stream.get(nb), "little") & mask
        if v < n:
            return v
