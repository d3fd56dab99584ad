// oracle.cpp -- CPU restatement of HElib's DoubleCRT / key-switch hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under helib_b200/ may link, load or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and there only as the checker / the timed CPU baseline.
//
// Parity status: the reference (homenc/HElib v2.2.0) cannot be built here (NTL >= 11.4.3
// and GMP >= 6.2.0 are fetched by URL at configure time: CMakeLists.txt:75-77,241) and its
// tests hold no golden integer vectors for this path.  This file
// restates the reference's algorithms function by function (citations are paths relative
// to /root/reference) and is pinned three ways (tests/test_oracle.py):
//   * against oracle/pyoracle.py (Python big-int arithmetic) and psi-independent invariants;
//   * against REFERENCE OUTPUT: the rows a real HElib build wrote into its own I/O fixtures
//     tests/test_resources/iotest_* (tests/golden/helib_iotest_m12.json) -- secret key, public
//     key and four key-switching matrices whose a_i are regenerated from the stored prgSeed
//     (oracle/ntl_prg.py + randomize), and the key-switch path (breakIntoDigits,
//     keySwitchDigits, addPrimesAndScale, scaleDownToSet, toPoly) run with the reference's own
//     matrices decrypts correctly under the fixture's secret key.
// Still "parity unpinned": the choice of the 2N-th root psi for power-of-two m (NTL's root
// tables; an argument of the engine and of this oracle) and the prime chain of modern
// parameter sets (restated, never compared with primes printed by HElib).
//
// Threading mirrors the reference: across primes for the transforms
// (src/DoubleCRT.cpp:79-84) and across coefficients for the CRT (src/DoubleCRT.cpp:1062-1102).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef __int128 i128;

namespace {

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
inline u64 powmod(u64 a, u64 e, u64 q) {
  u64 r = 1 % q; a %= q;
  while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
  return r;
}
inline u64 invmod(u64 a, u64 q) { return powmod(a, q - 2, q); }  // q prime

// NTL::PrepMulModPrecon / MulModPrecon equivalent (Shoup): wp = floor(w*2^64/q)
inline u64 shoup_prep(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
inline u64 shoup_mul(u64 a, u64 w, u64 wp, u64 q) {
  u64 hi = (u64)(((u128)a * wp) >> 64);
  u64 r = a * w - hi * q;
  return r >= q ? r - q : r;
}

// ---------------------------------------------------------------- fixed-width signed bigints
// L-limb little-endian two's complement.  L <= MAXL.
const int MAXL = 72;

struct Big {
  int L;
  u64 w[MAXL];
};
inline void big_zero(Big& a, int L) { a.L = L; memset(a.w, 0, sizeof(u64) * L); }
inline bool big_neg(const Big& a) { return (a.w[a.L - 1] >> 63) != 0; }
inline void big_add(Big& a, const Big& b) {  // a += b
  u128 c = 0;
  for (int i = 0; i < a.L; i++) { c += (u128)a.w[i] + b.w[i]; a.w[i] = (u64)c; c >>= 64; }
}
inline void big_sub(Big& a, const Big& b) {  // a -= b
  u64 borrow = 0;
  for (int i = 0; i < a.L; i++) {
    u128 d = (u128)a.w[i] - b.w[i] - borrow;
    a.w[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
  }
}
inline void big_muladd_small(Big& a, const Big& b, u64 s) {  // a += b*s  (b >= 0)
  u128 c = 0;
  for (int i = 0; i < a.L; i++) { c += (u128)b.w[i] * s + a.w[i]; a.w[i] = (u64)c; c >>= 64; }
}
inline void big_mulsub_small(Big& a, const Big& b, u64 s) {  // a -= b*s (b >= 0)
  u128 c = 0; u64 borrow = 0;
  for (int i = 0; i < a.L; i++) {
    c += (u128)b.w[i] * s; u64 lo = (u64)c; c >>= 64;
    u128 d = (u128)a.w[i] - lo - borrow; a.w[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
  }
}
inline int big_cmp(const Big& a, const Big& b) {  // signed compare
  bool na = big_neg(a), nb = big_neg(b);
  if (na != nb) return na ? -1 : 1;
  for (int i = a.L - 1; i >= 0; i--) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  return 0;
}
inline void big_mul_small_inplace(Big& a, u64 s) {  // a *= s (a >= 0)
  u128 c = 0;
  for (int i = 0; i < a.L; i++) { c += (u128)a.w[i] * s; a.w[i] = (u64)c; c >>= 64; }
}
inline u64 big_mod_small_unsigned(const u64* w, int L, u64 q) {
  u128 r = 0;
  for (int i = L - 1; i >= 0; i--) r = ((r << 64) | w[i]) % q;
  return (u64)r;
}
// value mod q in [0,q) for a signed two's complement value (NTL conv(zz_p, ZZ) semantics,
// src/CModulus.cpp:453-457)
inline u64 big_mod_small(const u64* w, int L, u64 q, u64 two64L_mod_q) {
  u64 r = big_mod_small_unsigned(w, L, q);
  if (w[L - 1] >> 63) r = submod(r, two64L_mod_q, q);
  return r;
}
inline void big_half_ceil(Big& a) {  // a = (a+1)/2 , a >= 0
  Big one; big_zero(one, a.L); one.w[0] = 1; big_add(a, one);
  for (int i = 0; i < a.L; i++) a.w[i] = (a.w[i] >> 1) | (i + 1 < a.L ? a.w[i + 1] << 63 : 0);
}

// ---------------------------------------------------------------- context

struct PrimeTab {
  u64 q, psi, ipsi;
  std::vector<u64> powers, powers_aux;    // psi^i,  src/CModulus.cpp:121-127
  std::vector<u64> ipowers, ipowers_aux;  // psi^-i, src/CModulus.cpp:129-135
  std::vector<u64> wtab, wtab_aux;        // omega^k (k < N/2), omega = psi^2
  std::vector<u64> iwtab, iwtab_aux;      // omega^-k
  u64 ninv, ninv_aux;
};

struct Ctx {
  long N, m;
  int logN;
  int nprimes;
  int nthreads;
  std::vector<PrimeTab> pt;
  std::vector<int> digit_of;  // digit number of prime i, or -1
  int ndigits;
  std::vector<int> special;   // indices of special primes
  std::vector<u64> brev;      // bit reversal table
};

void parallel_for(int nthreads, long n, const std::function<void(long, long)>& fn) {
  if (nthreads <= 1 || n <= 1) { fn(0, n); return; }
  int nt = (int)std::min<long>(nthreads, n);
  std::vector<std::thread> th;
  long chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    long a = t * chunk, b = std::min(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto& x : th) x.join();
}

// Cmodulus::FFT_aux pow-2 branch (src/CModulus.cpp:362-429): twist by psi^i, cyclic DIF FFT
// (output bit-reversed, as NTL::FFTFwd), then BitReverseCopy to natural order.
void ntt_fwd_row(const Ctx& c, int pi, const u64* coef, u64* row) {
  const PrimeTab& p = c.pt[pi];
  const long N = c.N; const u64 q = p.q;
  std::vector<u64> y(N);
  for (long i = 0; i < N; i++) y[i] = shoup_mul(coef[i], p.powers[i], p.powers_aux[i], q);
  // Gentleman-Sande DIF: natural in, bit-reversed out
  for (long len = N / 2, step = 1; len >= 1; len >>= 1, step <<= 1) {
    for (long start = 0; start < N; start += 2 * len) {
      for (long j = 0; j < len; j++) {
        u64 u = y[start + j], v = y[start + j + len];
        y[start + j] = addmod(u, v, q);
        u64 d = submod(u, v, q);
        long tw = j * step;
        y[start + j + len] = shoup_mul(d, p.wtab[tw], p.wtab_aux[tw], q);
      }
    }
  }
  for (long i = 0; i < N; i++) row[c.brev[i]] = y[i];
}

// Cmodulus::iFFT pow-2 branch (src/CModulus.cpp:486-553): BitReverseCopy, inverse cyclic
// transform incl. 1/N (NTL::FFTRev1), multiply by psi^-i; output in [0,q).
void ntt_inv_row(const Ctx& c, int pi, const u64* row, u64* coef) {
  const PrimeTab& p = c.pt[pi];
  const long N = c.N; const u64 q = p.q;
  std::vector<u64> y(N);
  for (long i = 0; i < N; i++) y[i] = row[c.brev[i]];
  // Cooley-Tukey DIT: bit-reversed in, natural out, inverse twiddles
  for (long len = 1, step = N / 2; len < N; len <<= 1, step >>= 1) {
    for (long start = 0; start < N; start += 2 * len) {
      for (long j = 0; j < len; j++) {
        long tw = j * step;
        u64 u = y[start + j];
        u64 v = shoup_mul(y[start + j + len], p.iwtab[tw], p.iwtab_aux[tw], q);
        y[start + j] = addmod(u, v, q);
        y[start + j + len] = submod(u, v, q);
      }
    }
  }
  for (long i = 0; i < N; i++) {
    u64 t = shoup_mul(y[i], p.ninv, p.ninv_aux, q);
    coef[i] = shoup_mul(t, p.ipowers[i], p.ipowers_aux[i], q);
  }
}

struct SetInfo {  // per prime-set constants used by toPoly
  int n; int L;
  std::vector<u64> q, t, t_aux;  // t = (prod/q)^-1 mod q   (src/DoubleCRT.cpp:1033-1041)
  std::vector<double> qrecip;
  std::vector<Big> prod1;        // prod / q_j
  Big prod, prod_half;
};

void make_setinfo(const Ctx& c, const int* idx, int n, SetInfo& s) {
  s.n = n; s.L = n + 3;
  if (s.L > MAXL) abort();
  s.q.resize(n); s.t.resize(n); s.t_aux.resize(n); s.qrecip.resize(n); s.prod1.resize(n);
  big_zero(s.prod, s.L); s.prod.w[0] = 1;
  for (int j = 0; j < n; j++) {
    s.q[j] = c.pt[idx[j]].q; s.qrecip[j] = 1.0 / double(s.q[j]);
    big_mul_small_inplace(s.prod, s.q[j]);
  }
  for (int j = 0; j < n; j++) {
    big_zero(s.prod1[j], s.L); s.prod1[j].w[0] = 1;
    for (int k = 0; k < n; k++) if (k != j) big_mul_small_inplace(s.prod1[j], s.q[k]);
    u64 r = big_mod_small_unsigned(s.prod1[j].w, s.L, s.q[j]);
    s.t[j] = invmod(r, s.q[j]); s.t_aux[j] = shoup_prep(s.t[j], s.q[j]);
  }
  s.prod_half = s.prod; big_half_ceil(s.prod_half);
}

// One coefficient of DoubleCRT::toPoly's CRT loop (src/DoubleCRT.cpp:1076-1100).
inline void crt_one(const SetInfo& s, const u64* rem, bool positive, Big& tmp) {
  big_zero(tmp, s.L);
  double quotient = 0;
  for (int j = 0; j < s.n; j++) {
    u64 r = shoup_mul(rem[j], s.t[j], s.t_aux[j], s.q[j]);
    big_muladd_small(tmp, s.prod1[j], r);
    quotient += double(r) * s.qrecip[j];
  }
  big_mulsub_small(tmp, s.prod, (u64)(long)quotient);
  while (big_neg(tmp)) big_add(tmp, s.prod);
  while (big_cmp(tmp, s.prod) >= 0) big_sub(tmp, s.prod);
  if (!positive && big_cmp(tmp, s.prod_half) >= 0) big_sub(tmp, s.prod);
}

// DoubleCRT::toPoly (src/DoubleCRT.cpp:925-1113).  out: N x Lout limbs, two's complement.
void to_poly(const Ctx& c, const u64* data, const int* idx, int n, bool positive, u64* out, int Lout) {
  const long N = c.N;
  if (n == 0) { memset(out, 0, sizeof(u64) * N * Lout); return; }
  std::vector<u64> remtab((size_t)N * n);
  parallel_for(c.nthreads, n, [&](long a, long b) {
    std::vector<u64> tmp(N);
    for (long j = a; j < b; j++) {
      ntt_inv_row(c, idx[j], data + (size_t)idx[j] * N, tmp.data());
      for (long h = 0; h < N; h++) remtab[(size_t)h * n + j] = tmp[h];
    }
  });
  SetInfo s; make_setinfo(c, idx, n, s);
  parallel_for(c.nthreads, N, [&](long a, long b) {
    Big tmp;
    for (long h = a; h < b; h++) {
      crt_one(s, &remtab[(size_t)h * n], positive, tmp);
      u64 ext = big_neg(tmp) ? ~0ULL : 0ULL;
      for (int l = 0; l < Lout; l++) out[(size_t)h * Lout + l] = l < tmp.L ? tmp.w[l] : ext;
    }
  });
}

// DoubleCRT::FFT(const ZZX&, const IndexSet&) (src/DoubleCRT.cpp:68-85) with
// Cmodulus::FFT(vec_long&, const ZZX&) (src/CModulus.cpp:446-460).
void fft_bigpoly(const Ctx& c, const u64* poly, int L, const int* idx, int n, u64* data) {
  const long N = c.N;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    std::vector<u64> tmp(N);
    for (long j = a; j < b; j++) {
      u64 q = c.pt[idx[j]].q;
      u64 two64L = powmod(powmod(2, 64, q), L, q);
      for (long h = 0; h < N; h++) tmp[h] = big_mod_small(poly + (size_t)h * L, L, q, two64L);
      ntt_fwd_row(c, idx[j], tmp.data(), data + (size_t)idx[j] * N);
    }
  });
}

u64 product_mod(const Ctx& c, const int* idx, int n, u64 q) {
  u64 r = 1 % q;
  for (int j = 0; j < n; j++) r = mulmod(r, c.pt[idx[j]].q % q, q);
  return r;
}

void scale_rows(const Ctx& c, u64* data, const int* idx, int n, const std::function<u64(u64)>& factor_mod) {
  const long N = c.N;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    for (long j = a; j < b; j++) {
      u64 q = c.pt[idx[j]].q; u64 f = factor_mod(q); u64 fa = shoup_prep(f, q);
      u64* row = data + (size_t)idx[j] * N;
      for (long h = 0; h < N; h++) row[h] = shoup_mul(row[h], f, fa, q);
    }
  });
}

}  // namespace

// =====================================================================================
extern "C" {

void* orc_ctx_create(long N, long m, int nprimes, const u64* q, const u64* psi, int nthreads) {
  Ctx* c = new Ctx;
  c->N = N; c->m = m; c->nprimes = nprimes; c->nthreads = nthreads > 0 ? nthreads : 1;
  c->logN = 0; while ((1L << c->logN) < N) c->logN++;
  c->brev.resize(N);
  for (long i = 0; i < N; i++) {
    u64 r = 0; for (int b = 0; b < c->logN; b++) if (i >> b & 1) r |= 1ULL << (c->logN - 1 - b);
    c->brev[i] = r;
  }
  c->pt.resize(nprimes);
  c->digit_of.assign(nprimes, -1); c->ndigits = 0;
  parallel_for(c->nthreads, nprimes, [&](long a, long b) {
    for (long i = a; i < b; i++) {
      PrimeTab& p = c->pt[i];
      p.q = q[i]; p.psi = psi[i]; p.ipsi = invmod(psi[i], q[i]);
      p.powers.resize(N); p.powers_aux.resize(N); p.ipowers.resize(N); p.ipowers_aux.resize(N);
      u64 w = 1, iw = 1;
      for (long k = 0; k < N; k++) {
        p.powers[k] = w; p.powers_aux[k] = shoup_prep(w, p.q);
        p.ipowers[k] = iw; p.ipowers_aux[k] = shoup_prep(iw, p.q);
        w = mulmod(w, p.psi, p.q); iw = mulmod(iw, p.ipsi, p.q);
      }
      long H = N / 2 > 0 ? N / 2 : 1;
      p.wtab.resize(H); p.wtab_aux.resize(H); p.iwtab.resize(H); p.iwtab_aux.resize(H);
      u64 om = mulmod(p.psi, p.psi, p.q), iom = mulmod(p.ipsi, p.ipsi, p.q);
      w = 1; iw = 1;
      for (long k = 0; k < H; k++) {
        p.wtab[k] = w; p.wtab_aux[k] = shoup_prep(w, p.q);
        p.iwtab[k] = iw; p.iwtab_aux[k] = shoup_prep(iw, p.q);
        w = mulmod(w, om, p.q); iw = mulmod(iw, iom, p.q);
      }
      p.ninv = invmod((u64)N % p.q, p.q); p.ninv_aux = shoup_prep(p.ninv, p.q);
    }
  });
  return c;
}

void orc_ctx_destroy(void* h) { delete (Ctx*)h; }

void orc_ctx_set_chain(void* h, const int* digit_of, int ndigits, const int* special, int nspecial) {
  Ctx* c = (Ctx*)h;
  c->digit_of.assign(digit_of, digit_of + c->nprimes); c->ndigits = ndigits;
  c->special.assign(special, special + nspecial);
}

void orc_set_threads(void* h, int nthreads) { ((Ctx*)h)->nthreads = nthreads > 0 ? nthreads : 1; }

void orc_ntt_fwd(void* h, int pi, const u64* coef, u64* row) { ntt_fwd_row(*(Ctx*)h, pi, coef, row); }
void orc_ntt_inv(void* h, int pi, const u64* row, u64* coef) { ntt_inv_row(*(Ctx*)h, pi, row, coef); }

// rows `idx` of data <- transforms of the (small, already reduced or not) coefficient rows
void orc_ntt_fwd_rows(void* h, u64* data, const int* idx, int n) {
  Ctx& c = *(Ctx*)h;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    std::vector<u64> tmp(c.N);
    for (long j = a; j < b; j++) {
      u64* row = data + (size_t)idx[j] * c.N;
      memcpy(tmp.data(), row, sizeof(u64) * c.N);
      ntt_fwd_row(c, idx[j], tmp.data(), row);
    }
  });
}
void orc_ntt_inv_rows(void* h, u64* data, const int* idx, int n) {
  Ctx& c = *(Ctx*)h;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    std::vector<u64> tmp(c.N);
    for (long j = a; j < b; j++) {
      u64* row = data + (size_t)idx[j] * c.N;
      memcpy(tmp.data(), row, sizeof(u64) * c.N);
      ntt_inv_row(c, idx[j], tmp.data(), row);
    }
  });
}

void orc_to_poly(void* h, const u64* data, const int* idx, int n, int positive, u64* out, int Lout) {
  to_poly(*(Ctx*)h, data, idx, n, positive != 0, out, Lout);
}

void orc_fft_bigpoly(void* h, const u64* poly, int L, const int* idx, int n, u64* data) {
  fft_bigpoly(*(Ctx*)h, poly, L, idx, n, data);
}

// op: 0 add, 1 sub, 2 mul   -- DoubleCRT::Op / do_mul (src/DoubleCRT.cpp:216-337)
void orc_pointwise(void* h, int op, u64* dst, const u64* src, const int* idx, int n) {
  Ctx& c = *(Ctx*)h; const long N = c.N;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    for (long j = a; j < b; j++) {
      u64 q = c.pt[idx[j]].q;
      u64* d = dst + (size_t)idx[j] * N; const u64* s = src + (size_t)idx[j] * N;
      if (op == 0) for (long k = 0; k < N; k++) d[k] = addmod(d[k], s[k], q);
      else if (op == 1) for (long k = 0; k < N; k++) d[k] = submod(d[k], s[k], q);
      else for (long k = 0; k < N; k++) d[k] = mulmod(d[k], s[k], q);
    }
  });
}

// DoubleCRT::Op(ZZ, MulFun) with the scalar given as a product of chain primes
// (what addPrimesAndScale / operator/= need): rows *= prod(q_k, k in fidx)  [inverse if inv]
void orc_scale_by_primes(void* h, u64* data, const int* idx, int n, const int* fidx, int nf, int inv) {
  Ctx& c = *(Ctx*)h;
  scale_rows(c, data, idx, n, [&](u64 q) {
    u64 f = product_mod(c, fidx, nf, q);
    return inv ? invmod(f, q) : f;
  });
}

// rows *= (scalar mod q)   (DoubleCRT::Op(ZZ) for a word-sized scalar, src/DoubleCRT.cpp:339-361)
void orc_scale_by_word(void* h, u64* data, const int* idx, int n, u64 scalar) {
  Ctx& c = *(Ctx*)h;
  scale_rows(c, data, idx, n, [&](u64 q) { return scalar % q; });
}

// DoubleCRT::addPrimes (src/DoubleCRT.cpp:565-599).  poly_out may be null (N x L limbs otherwise).
void orc_add_primes(void* h, u64* data, const int* cur, int ncur, const int* add, int nadd, u64* poly_out, int L) {
  Ctx& c = *(Ctx*)h;
  if (nadd == 0) return;
  int Lw = ncur + 1;
  std::vector<u64> poly;
  u64* pp = poly_out;
  if (!pp) { poly.resize((size_t)c.N * Lw); pp = poly.data(); L = Lw; }
  to_poly(c, data, cur, ncur, false, pp, L);
  fft_bigpoly(c, pp, L, add, nadd, data);
}

// DoubleCRT::addPrimesAndScale (src/DoubleCRT.cpp:603-647)
void orc_add_primes_and_scale(void* h, u64* data, const int* cur, int ncur, const int* add, int nadd) {
  Ctx& c = *(Ctx*)h;
  if (nadd == 0) return;
  orc_scale_by_primes(h, data, cur, ncur, add, nadd, 0);
  for (int j = 0; j < nadd; j++) memset(data + (size_t)add[j] * c.N, 0, sizeof(u64) * c.N);
}

// DoubleCRT::scaleDownToSet (src/DoubleCRT.cpp:1464-1516).  cur = current index set,
// keep = s & cur (non-empty), delta_out optional (N x L limbs).
void orc_scale_down(void* h, u64* data, const int* cur, int ncur, const int* keep, int nkeep,
                    long ptxtSpace, u64* delta_out, int L) {
  Ctx& c = *(Ctx*)h; const long N = c.N;
  std::vector<int> diff;
  for (int j = 0; j < ncur; j++) if (std::find(keep, keep + nkeep, cur[j]) == keep + nkeep) diff.push_back(cur[j]);
  if (diff.empty()) return;
  int nd = (int)diff.size();
  int Lw = nd + 2;  // room for |delta| <= P*(1+p)/2 with p < 2^62
  std::vector<u64> dbuf;
  u64* delta = delta_out;
  if (!delta) { dbuf.resize((size_t)N * Lw); delta = dbuf.data(); L = Lw; }
  to_poly(c, data, diff.data(), nd, false, delta, L);
  if (ptxtSpace > 1) {
    u64 p = (u64)ptxtSpace;
    u64 p_over_2 = p / 2, p_mod_2 = p % 2;
    // prodInv = InvMod(rem(diffProd, ptxtSpace), ptxtSpace): p need not be prime -> ext. Euclid
    u64 Pm = product_mod(c, diff.data(), nd, p);
    i128 t0 = 0, t1 = 1; i128 r0 = p, r1 = Pm;
    while (r1 != 0) { i128 qq = r0 / r1; i128 t2 = t0 - qq * t1; t0 = t1; t1 = t2; i128 r2 = r0 - qq * r1; r0 = r1; r1 = r2; }
    if (r0 != 1) abort();
    u64 prodInv = (u64)((t0 % (i128)p + (i128)p) % (i128)p);
    Big P; big_zero(P, L); P.w[0] = 1;
    for (int j = 0; j < nd; j++) big_mul_small_inplace(P, c.pt[diff[j]].q);
    u64 two64L = powmod(powmod(2, 64, p), L, p);
    parallel_for(c.nthreads, N, [&](long a, long b) {
      Big d; d.L = L;
      for (long k = a; k < b; k++) {
        u64* w = delta + (size_t)k * L;
        u64 u = big_mod_small(w, L, p, two64L);
        if (u != 0) {
          bool neg = (w[L - 1] >> 63) != 0;
          u = mulmod(u, prodInv, p);
          bool minus = u > p_over_2 || (p_mod_2 == 0 && u == p_over_2 && neg);
          memcpy(d.w, w, sizeof(u64) * L);
          if (minus) big_muladd_small(d, P, p - u);   // delta -= P*(u-p)
          else big_mulsub_small(d, P, u);             // delta -= P*u
          memcpy(w, d.w, sizeof(u64) * L);
        }
      }
    });
  }
  // removePrimes(diff); *this -= delta; *this /= diffProd
  std::vector<u64> dd((size_t)c.nprimes * N);
  fft_bigpoly(c, delta, L, keep, nkeep, dd.data());
  orc_pointwise(h, 1, data, dd.data(), keep, nkeep);
  orc_scale_by_primes(h, data, keep, nkeep, diff.data(), nd, 1);
}

// DoubleCRT::breakIntoDigits (src/DoubleCRT.cpp:479-561).  cur must be ctxt primes only.
// digits_out: [maxdig][nprimes][N]; returns the number of digits n.  Each digit is defined
// over cur | special.  polys_out (optional): [n][N][L] balanced digit polynomials.
int orc_break_into_digits(void* h, const u64* data, const int* cur, int ncur, u64* digits_out,
                          u64* polys_out, int L) {
  Ctx& c = *(Ctx*)h; const long N = c.N;
  const size_t PS = (size_t)c.nprimes * N;
  std::vector<char> remaining(c.nprimes, 0);
  for (int j = 0; j < ncur; j++) remaining[cur[j]] = 1;
  int n = 0; int left = ncur;
  for (; left > 0; n++) for (int i = 0; i < c.nprimes; i++) if (remaining[i] && c.digit_of[i] == n) { remaining[i] = 0; left--; }
  std::vector<int> all(cur, cur + ncur);
  all.insert(all.end(), c.special.begin(), c.special.end());
  std::sort(all.begin(), all.end());
  std::vector<std::vector<int>> dset(n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < ncur; j++) if (c.digit_of[cur[j]] == i) {
      dset[i].push_back(cur[j]);
      memcpy(digits_out + i * PS + (size_t)cur[j] * N, data + (size_t)cur[j] * N, sizeof(u64) * N);
    }
  }
  for (int i = 0; i < n; i++) {
    std::vector<int> notIn;
    for (int a : all) if (std::find(dset[i].begin(), dset[i].end(), a) == dset[i].end()) notIn.push_back(a);
    orc_add_primes(h, digits_out + i * PS, dset[i].data(), (int)dset[i].size(), notIn.data(), (int)notIn.size(),
                   polys_out ? polys_out + (size_t)i * N * L : nullptr, L);
    // pi = product of the FULL context digit i (src/DoubleCRT.cpp:551)
    std::vector<int> full;
    for (int k = 0; k < c.nprimes; k++) if (c.digit_of[k] == i) full.push_back(k);
    for (int j = i + 1; j < n; j++) {
      orc_pointwise(h, 1, digits_out + j * PS, digits_out + i * PS, dset[j].data(), (int)dset[j].size());
      orc_scale_by_primes(h, digits_out + j * PS, dset[j].data(), (int)dset[j].size(), full.data(), (int)full.size(), 1);
    }
  }
  return n;
}

// Ctxt::keySwitchDigits (src/Ctxt.cpp:191-230): out1 += digit_i * a_i ; out0 += digit_i * b_i
// over idx (= primeSet of the accumulating ciphertext).  evk_a/evk_b: [ndig][nprimes][N].
void orc_keyswitch_digits(void* h, const u64* digits, int ndig, const int* idx, int n,
                          const u64* evk_a, const u64* evk_b, u64* out0, u64* out1) {
  Ctx& c = *(Ctx*)h; const long N = c.N;
  const size_t PS = (size_t)c.nprimes * N;
  std::vector<u64> tmp(PS);
  for (int i = 0; i < ndig; i++) {
    for (int j = 0; j < n; j++) memcpy(&tmp[(size_t)idx[j] * N], digits + i * PS + (size_t)idx[j] * N, sizeof(u64) * N);
    orc_pointwise(h, 2, tmp.data(), evk_a + i * PS, idx, n);
    orc_pointwise(h, 0, out1, tmp.data(), idx, n);
    for (int j = 0; j < n; j++) memcpy(&tmp[(size_t)idx[j] * N], digits + i * PS + (size_t)idx[j] * N, sizeof(u64) * N);
    orc_pointwise(h, 2, tmp.data(), evk_b + i * PS, idx, n);
    orc_pointwise(h, 0, out0, tmp.data(), idx, n);
  }
}

// DoubleCRT::automorph for power-of-two m (src/DoubleCRT.cpp:1160-1202): new[j]=old[idx(rep(j)*k mod m)]
void orc_automorph(void* h, u64* data, const int* idx, int n, long k) {
  Ctx& c = *(Ctx*)h; const long N = c.N, m = c.m;
  parallel_for(c.nthreads, n, [&](long a, long b) {
    std::vector<u64> tmp(m);
    for (long r = a; r < b; r++) {
      u64* row = data + (size_t)idx[r] * N;
      for (long j = 0; j < N; j++) tmp[2 * j + 1] = row[j];
      for (long j = 0; j < N; j++) row[j] = tmp[(u64)((u128)(2 * j + 1) * (u64)k % (u64)m)];
    }
  });
}

// Ctxt::tensorProduct for two canonical 2-part ciphertexts (src/Ctxt.cpp:1563-1608):
// out = [a0*b0, a0*b1 + a1*b0, a1*b1]
void orc_tensor(void* h, const u64* a0, const u64* a1, const u64* b0, const u64* b1,
                u64* o0, u64* o1, u64* o2, const int* idx, int n) {
  Ctx& c = *(Ctx*)h; const long N = c.N;
  parallel_for(c.nthreads, n, [&](long ra, long rb) {
    for (long r = ra; r < rb; r++) {
      u64 q = c.pt[idx[r]].q; size_t off = (size_t)idx[r] * N;
      for (long k = 0; k < N; k++) {
        u64 x0 = a0[off + k], x1 = a1[off + k], y0 = b0[off + k], y1 = b1[off + k];
        o0[off + k] = mulmod(x0, y0, q);
        o1[off + k] = addmod(mulmod(x0, y1, q), mulmod(x1, y0, q), q);
        o2[off + k] = mulmod(x1, y1, q);
      }
    }
  });
}

}  // extern "C"
