// Random walk over the CKKS branches of the helib::Ctxt mirror: a pool of ciphertexts with scaling factors and their
// plaintext polynomials over the reals (long double mirror); random sequences of +=, -=, multiplyBy, negate and
// dropSmallAndSpecialPrimes.  After every step: decode (toPoly / ratFactor) must equal the mirror within the TRACKED noise
// bound -- the reference's approximate-number tests compare with a tolerance the same way (tests/GTestApproxNums.cpp:180-235).
// usage: test_ckks_walk [seed] [steps].  Exit codes: 0 ok, 3 no CUDA device, 1 failure.
#include <cstdio>
#include <cstdlib>
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
static long double limbs_to_ld(const uint64_t* w, int L) {
  const bool neg = w[L - 1] >> 63;
  long double mag = 0;
  for (int l = L - 1; l >= 0; l--) mag = mag * 18446744073709551616.0L + (long double)(neg ? ~w[l] : w[l]);
  return neg ? -(mag + 1) : mag;
}
typedef std::vector<long double> RPoly;

int main(int argc, char** argv) {
  if (hb_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
  const unsigned long seed = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 1;
  const int steps = argc > 2 ? std::atoi(argv[2]) : 30;
  try {
    const long m = 2048;
    Context ctx(m, /*p=*/-1, /*r=*/20, /*bits=*/360, /*c=*/3);
    const long N = ctx.getPhiM();
    std::mt19937_64 gen(777ULL + seed);
    const double sigma = 3.2;
    const IndexSet allq = ctx.getCtxtPrimes() | ctx.getSpecialPrimes();
    KeyInfo pk; pk.context = &ctx; pk.ckks = true; pk.scale = 10.0; pk.hwt = 0;
    pk.skBound = pk.scale * std::sqrt(double(N) * 2.0 / 3.0);
    std::vector<long> s = sample_ternary(gen, N);
    DoubleCRT S(s, ctx, allq);
    {
      KeySwitch W; W.fromKey = SKHandle(2, 1, 0); W.toKeyID = 0; W.ptxtSpace = 1;
      DoubleCRT fromKey(S); fromKey *= S;
      fromKey.multiplyByPrimes(ctx.getSpecialPrimes());
      for (size_t i = 0; i < ctx.getDigits().size(); i++) {
        W.a.push_back(random_rows(ctx, allq, gen));
        DoubleCRT b(sample_gauss(gen, N, sigma), ctx, allq);
        DoubleCRT t(W.a.back()); t *= S; b -= t;
        b += fromKey;
        W.b.push_back(b);
        fromKey.multiplyByPrimes(ctx.getDigit(i));
      }
      W.noiseBound = XD(pk.noiseBoundForGaussian(sigma, N));
      pk.keySwitching.push_back(W);
    }
    const double Delta = std::ldexp(1.0, 30);
    auto rand_msg = [&]() { std::vector<long> v(N); for (auto& x : v) x = (long)(gen() % 5) - 2; return v; };
    auto encrypt = [&](const std::vector<long>& msg) {
      Ctxt c(pk, 1);
      c.primeSet = ctx.getCtxtPrimes();
      std::vector<long> e = sample_gauss(gen, N, sigma), pt(N);
      for (long k = 0; k < N; k++) pt[k] = (long)(Delta * msg[k]) + e[k];
      DoubleCRT c1 = random_rows(ctx, c.primeSet, gen);
      DoubleCRT c0(pt, ctx, c.primeSet);
      DoubleCRT t(c1); t.Mul(S, false); c0 -= t;
      c.parts.emplace_back(c0, SKHandle());
      c.parts.emplace_back(c1, SKHandle(1, 1, 0));
      c.noiseBound = XD(pk.noiseBoundForGaussian(sigma, N));
      c.ratFactor = XD(Delta);
      c.ptxtMag = XD(embeddingLargestCoeff(msg, m));
      return c;
    };
    std::vector<DoubleCRT> sKeys; sKeys.push_back(S);
    auto decode = [&](const Ctxt& c, double* bound) {
      std::vector<long> dummy; std::vector<uint64_t> limbs; int L = 0;
      hb::Decrypt(dummy, c, sKeys, &limbs, &L);
      RPoly out(N);
      const long double rf = std::ldexp((long double)c.ratFactor.m, (int)c.ratFactor.e);
      for (long k = 0; k < N; k++) out[k] = limbs_to_ld(&limbs[(size_t)k * L], L) / rf;
      *bound = std::exp2(c.noiseBound.ln() / std::log(2.0) - (std::log2((double)c.ratFactor.m) + (double)c.ratFactor.e));
      return out;
    };
    auto mul = [&](const RPoly& x, const RPoly& y) {
      RPoly r(N, 0.0L);
      for (long i = 0; i < N; i++) { if (x[i] == 0) continue; for (long j = 0; j < N; j++) { long k = i + j; if (k >= N) r[k - N] -= x[i] * y[j]; else r[k] += x[i] * y[j]; } }
      return r;
    };
    const double ln2 = std::log(2.0);
    const int POOL = 3;
    std::vector<Ctxt> ct; std::vector<RPoly> pt;
    auto fresh = [&](int i) { std::vector<long> mm = rand_msg(); RPoly r(mm.begin(), mm.end()); if ((int)ct.size() <= i) { ct.push_back(encrypt(mm)); pt.push_back(r); } else { ct[i] = encrypt(mm); pt[i] = r; } };
    for (int i = 0; i < POOL; i++) fresh(i);
    auto capacity_bits = [&](const Ctxt& c) { return (pk.logOfProduct(c.primeSet & ctx.getCtxtPrimes()) - c.noiseBound.ln()) / ln2; };
    auto magnitude = [&](const RPoly& r) { long double mx = 0; for (auto v : r) mx = std::max(mx, std::fabs(v)); return mx; };
    int nmul = 0, nrefresh = 0; double worst_ratio = 0;
    for (int st = 0; st < steps; st++) {
      const int a = (int)(gen() % POOL), b = (int)(gen() % POOL);
      const int op = (int)(gen() % 6);
      const char* name = "";
      if (op == 2 && (capacity_bits(ct[a]) < 160 || capacity_bits(ct[b]) < 160 || magnitude(pt[a]) * magnitude(pt[b]) * N > 1e12L)) { fresh(a); nrefresh++; name = "refresh"; }
      else switch (op) {
        case 0: name = "+="; { Ctxt o = ct[b]; ct[a] += o; } { RPoly y = pt[b]; for (long k = 0; k < N; k++) pt[a][k] += y[k]; } break;
        case 1: name = "-="; { Ctxt o = ct[b]; ct[a] -= o; } { RPoly y = pt[b]; for (long k = 0; k < N; k++) pt[a][k] -= y[k]; } break;
        case 2: name = "multiplyBy"; { Ctxt o = ct[b]; ct[a].multiplyBy(o); } pt[a] = mul(pt[a], RPoly(pt[b])); nmul++; break;
        case 3: name = "negate"; ct[a].negate(); for (auto& v : pt[a]) v = -v; break;
        case 4: name = "dropSmallAndSpecialPrimes"; ct[a].dropSmallAndSpecialPrimes(); break;
        default: name = "copy"; ct[a] = ct[b]; pt[a] = pt[b]; break;
      }
      double bound = 0;
      RPoly got = decode(ct[a], &bound);
      long double err = 0;
      for (long k = 0; k < N; k++) err = std::max(err, std::fabs(got[k] - pt[a][k]));
      const long double slack = magnitude(pt[a]) * 1e-15L;    // the mirror's own rounding
      if (err > bound + slack) { std::printf("seed %lu step %d (%s): error %.3Lg exceeds the tracked bound %.3g\n", seed, st, name, err, bound); return 1; }
      // the bound must stay small against the TRACKED plaintext magnitude (the mirror polynomial itself may have cancelled, a - a)
      const double mag = std::max(ct[a].ptxtMag.to_double(), 1.0);
      if ((double)magnitude(pt[a]) > mag * (1 + 1e-9)) { std::printf("seed %lu step %d (%s): plaintext size %.3Lg above the tracked magnitude %.3g\n", seed, st, name, magnitude(pt[a]), mag); return 1; }
      if (bound > 1e-2 * mag) { std::printf("seed %lu step %d (%s): tracked bound %.3g is useless against a tracked magnitude of %.3g\n", seed, st, name, bound, mag); return 1; }
      if (bound > 0) worst_ratio = std::max(worst_ratio, (double)(err / bound));
    }
    ctx.sync();
    std::printf("ckks walk OK: seed %lu, %d steps, %d products, %d refreshes, largest error/bound %.3g\n", seed, steps, nmul, nrefresh, worst_ratio);
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
