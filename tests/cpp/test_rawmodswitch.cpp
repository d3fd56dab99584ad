// Ctxt::rawModSwitch through the mirror (include/helib_b200_ctxt.hpp) for a general m with a three-factor powerful basis
// and for a power-of-two m: after the switch to an external modulus q the ciphertext still decrypts -- modulo q, with the
// small integer polynomials the switch returns -- to the same plaintext (times q*Q^-1 mod p^r), which is what recryption
// relies on (src/recryption.cpp:1015-1046).  Exit codes: 0 ok, 3 no CUDA device, 1 failure.
#include <cstdio>
#include <random>

#include "helib_b200_ctxt.hpp"

using namespace hb;

static std::vector<long> sample_ternary(std::mt19937_64& g, long n) { std::vector<long> v(n); for (auto& x : v) x = (long)(g() % 3) - 1; return v; }
static std::vector<long> sample_gauss(std::mt19937_64& g, long n, double sigma) { std::normal_distribution<double> d(0, sigma); std::vector<long> v(n); for (auto& x : v) x = std::lround(d(g)); return v; }
static DoubleCRT random_rows(const Context& ctx, const IndexSet& s, std::mt19937_64& g) {
  const long N = ctx.getPhiM();
  std::vector<uint64_t> dense((size_t)ctx.numPrimes() * N, 0);
  for (long i : s) for (long k = 0; k < N; k++) dense[(size_t)i * N + k] = g() % (uint64_t)ctx.ithPrime(i);
  return DoubleCRT::fromRows(ctx, s, dense);
}
// a*b mod (Phi_m, integers), inputs of length phi(m)
static std::vector<long> mul_mod_phi(const std::vector<long>& a, const std::vector<long>& b, const std::vector<long>& phimx, long m) {
  const long n = (long)phimx.size() - 1;
  std::vector<long> t((size_t)(2 * n), 0);
  for (long i = 0; i < n; i++) if (a[i]) for (long j = 0; j < n; j++) t[i + j] += a[i] * b[j];
  if ((m & (m - 1)) == 0) { for (long k = 2 * n - 1; k >= n; k--) { t[k - n] -= t[k]; t[k] = 0; } }   // X^n = -1
  else for (long k = 2 * n - 1; k >= n; k--) { const long c = t[k]; if (c) for (long j = 0; j <= n; j++) t[k - n + j] -= c * phimx[j]; }
  t.resize((size_t)n);
  return t;
}

static int run(long m, long p, long q) {
  Context ctx(m, p, 1, /*bits=*/100, /*c=*/2);
  const long N = ctx.getPhiM();
  std::mt19937_64 gen(99 + m);
  KeyInfo pk; pk.context = &ctx; pk.ckks = false; pk.scale = 10.0;
  pk.skBound = pk.scale * std::sqrt(double(N) * 2.0 / 3.0);
  std::vector<long> s = sample_ternary(gen, N);
  Ctxt c(pk, p);
  c.primeSet = ctx.getCtxtPrimes();
  DoubleCRT S(s, ctx, c.primeSet);
  std::vector<long> msg(N), pt(N), e = sample_gauss(gen, N, 3.2);
  unsigned long Qp = 1;
  for (long i : c.primeSet) Qp = Qp * (unsigned long)(ctx.ithPrime(i) % p) % (unsigned long)p;
  for (long k = 0; k < N; k++) { msg[k] = (long)(gen() % (unsigned long)p); pt[k] = p * e[k] + (long)(Qp * (unsigned long)msg[k] % (unsigned long)p); }
  DoubleCRT c1 = random_rows(ctx, c.primeSet, gen);
  DoubleCRT c0(pt, ctx, c.primeSet);
  { DoubleCRT t(c1); t *= S; c0 -= t; }
  c.parts.emplace_back(c0, SKHandle());
  c.parts.emplace_back(c1, SKHandle(1, 1, 0));
  c.noiseBound = XD(double(p) * pk.noiseBoundForGaussian(3.2, N) + pk.noiseBoundForMod(p, N));
  std::vector<std::vector<long>> zz;
  const double scaled = c.rawModSwitch(zz, q);
  if (zz.size() != 2 || (long)zz[0].size() != N) { std::printf("m=%ld: wrong shape\n", m); return 1; }
  long mx = 0;
  for (auto& part : zz) for (long v : part) mx = std::max(mx, std::labs(v));
  // (the polynomial-basis coefficients of a non-trivial powerful basis may exceed q/2: the reference does not re-reduce them)
  if ((m & (m - 1)) == 0 && mx > q / 2) { std::printf("m=%ld: coefficient %ld outside [-q/2, q/2]\n", m, mx); return 1; }
  // decrypt modulo q with the switched parts: w = z0 + z1*s  (symmetric mod q), then mod p; expect msg * (q * Q^-1 * Q) = msg * q ... mod p
  std::vector<long> phimx = Ctxt::cyclotomic(m);
  std::vector<long> w = mul_mod_phi(zz[1], s, phimx, m);
  long wmax = 0;
  for (long k = 0; k < N; k++) { long v = (w[k] + zz[0][k]) % q; if (v > q / 2) v -= q; if (v < -(q / 2)) v += q; w[k] = v; wmax = std::max(wmax, std::labs(v)); }
  if (wmax >= q / 2 - 1) { std::printf("m=%ld: switched noise %ld too close to q/2=%ld for the check\n", m, wmax, q / 2); return 1; }
  // u = p*e + Qp*msg decrypts mod Q; after the switch w = u * q * Q^-1 (mod p): msg' = w * (q * Q^-1 * Qp)^-1 ... with u = Qp*msg mod p
  const long qmodp = q % p;
  for (long k = 0; k < N; k++) {
    // w = (p e + Qp msg) q Q^-1 = msg q (mod p)   since Qp = Q mod p
    long lhs = ((w[k] % p) + p) % p, rhs = (long)((unsigned long)msg[k] * (unsigned long)qmodp % (unsigned long)p);
    if (lhs != rhs) { std::printf("m=%ld: plaintext lost by rawModSwitch at %ld (%ld vs %ld)\n", m, k, lhs, rhs); return 1; }
  }
  std::printf("rawModSwitch OK: m=%ld phi=%ld q=%ld, |w|max=%ld, scaled noise estimate %.3g\n", m, N, q, wmax, scaled);
  return 0;
}

int main() {
  if (hb_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
  try {
    if (run(105, 2, (1L << 12) + 1)) return 1;      // three prime factors: powerful basis differs from the polynomial basis
    if (run(45, 2, (1L << 12) + 1)) return 1;       // 9 * 5: a prime-power factor
    if (run(128, 3, 3 * 3 * 3 * 3 * 3 * 3 * 3 + 1)) return 1;   // power of two: trivial powerful basis, even q
    std::printf("rawmodswitch OK\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
