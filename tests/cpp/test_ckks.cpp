// CKKS branch of the helib::Ctxt mirror (include/helib_b200_ctxt.hpp): fresh ciphertexts with a scaling factor,
// multiplyBy (computeIntervalForMul's CKKS interval, mod-switch, tensor, relin_CKKS_adjust, relinearise, device norms),
// decrypt = toPoly, divide by ratFactor, compare with the plaintext product within the tracked noise bound.
// Reference test style: tests/GTestApproxNums.cpp:180-235 (encrypt, multiply, decrypt, compare with a tolerance).
// Exit codes: 0 ok, 3 no CUDA device, 1 failure.
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
// two's-complement limbs -> long double (|value| far below 2^16000)
static long double limbs_to_ld(const uint64_t* w, int L) {
  const bool neg = w[L - 1] >> 63;
  long double mag = 0;
  for (int l = L - 1; l >= 0; l--) mag = mag * 18446744073709551616.0L + (long double)(neg ? ~w[l] : w[l]);
  return neg ? -(mag + 1) : mag;
}

int main() {
  if (hb_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
  try {
    const long m = 8192;
    Context ctx(m, /*p=*/-1, /*r=*/20, /*bits=*/300, /*c=*/2);
    const long N = ctx.getPhiM();
    std::mt19937_64 gen(7);
    const double sigma = 3.2;
    const IndexSet allq = ctx.getCtxtPrimes() | ctx.getSpecialPrimes();
    KeyInfo pk; pk.context = &ctx; pk.ckks = true; pk.scale = 10.0; pk.hwt = 0;
    pk.skBound = pk.scale * std::sqrt(double(N) * 2.0 / 3.0);
    std::vector<long> s = sample_ternary(gen, N);
    DoubleCRT S(s, ctx, allq);
    {   // s^2 -> s  (GenKeySWmatrix with ptxtSpace 1, src/keys.cpp:1159-1256)
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
    auto encrypt = [&](const std::vector<long>& msg) {   // symmetric CKKS encryption of Delta*msg (SecKey::Encrypt, CKKS branch)
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
    auto decode = [&](const Ctxt& c, double* noise_log2) {   // Decrypt (CKKS: the integer polynomial), then / ratFactor
      std::vector<long> dummy; std::vector<uint64_t> limbs; int L = 0;
      hb::Decrypt(dummy, c, sKeys, &limbs, &L);
      std::vector<double> out(N);
      const long double rf = std::ldexp((long double)c.ratFactor.m, (int)c.ratFactor.e);
      for (long k = 0; k < N; k++) out[k] = (double)(limbs_to_ld(&limbs[(size_t)k * L], L) / rf);
      if (noise_log2) *noise_log2 = c.noiseBound.ln() / std::log(2.0) - (std::log2((double)c.ratFactor.m) + (double)c.ratFactor.e);
      return out;
    };
    std::vector<long> ma(N), mb(N);
    for (long k = 0; k < N; k++) { ma[k] = (long)(gen() % 7) - 3; mb[k] = (long)(gen() % 7) - 3; }
    Ctxt ca = encrypt(ma), cb = encrypt(mb);
    {
      double nl; std::vector<double> d = decode(ca, &nl);
      for (long k = 0; k < N; k++) if (std::fabs(d[k] - ma[k]) > 1e-4) { std::printf("fresh decode mismatch at %ld: %g vs %ld\n", k, d[k], ma[k]); return 1; }
    }
    const double logq0 = pk.logOfProduct(ca.primeSet);
    ca.multiplyBy(cb);
    if (!ca.inCanonicalForm()) { std::printf("not canonical after multiplyBy\n"); return 1; }
    if (!(ca.primeSet >= ctx.getSpecialPrimes())) { std::printf("special primes missing after relinearisation\n"); return 1; }
    double nl = 0;
    std::vector<double> prod = decode(ca, &nl);
    // tolerance: the tracked noise bound relative to the scaling factor is a bound on the canonical-embedding norm of the
    // error; coefficients are bounded by it as well (power-of-two m: |coeff| <= ||.||_canon)
    const double tol = std::exp2(nl);
    double worst = 0;
    for (long t = 0; t < 64; t++) {
      long k = (t * 521 + 3) % N; long acc = 0;
      for (long i = 0; i < N; i++) { long j = k - i; acc += j >= 0 ? ma[i] * mb[j] : -(ma[i] * mb[j + N]); }
      worst = std::max(worst, std::fabs(prod[k] - (double)acc));
    }
    if (worst > tol) { std::printf("CKKS product error %.3g exceeds the tracked bound %.3g\n", worst, tol); return 1; }
    if (tol > 1.0) { std::printf("tracked bound %.3g is useless for integers in [-3,3] products\n", tol); return 1; }
    // mod-down to the ctxt primes and decode again: same values
    ca.dropSmallAndSpecialPrimes();
    std::vector<double> prod2 = decode(ca, &nl);
    for (long t = 0; t < 64; t++) { long k = (t * 521 + 3) % N; if (std::fabs(prod2[k] - prod[k]) > std::exp2(nl) + tol) { std::printf("mod-down changed the value at %ld\n", k); return 1; } }
    // addCtxt across different scaling factors (equalizeRationalFactors, src/Ctxt.cpp:1199-1351): product + fresh
    std::vector<long> mc(N);
    for (long k = 0; k < N; k++) mc[k] = (long)(gen() % 7) - 3;
    Ctxt cc = encrypt(mc);
    const double rf_before = std::log2((double)ca.ratFactor.m) + (double)ca.ratFactor.e;
    Ctxt sum = ca; sum += cc;
    if (std::fabs((sum.ratFactor / cc.ratFactor).to_double() - std::round((sum.ratFactor / cc.ratFactor).to_double())) > 1e-6 && sum.ratFactor.e < 60) { std::printf("common factor is not an integer multiple\n"); return 1; }
    std::vector<double> ds = decode(sum, &nl);
    const double tol2 = std::exp2(nl);
    for (long t = 0; t < 64; t++) {
      long k = (t * 521 + 3) % N;
      if (std::fabs(ds[k] - (prod2[k] + (double)mc[k])) > tol2 + tol) { std::printf("CKKS sum mismatch at %ld: %g vs %g (bound %g)\n", k, ds[k], prod2[k] + mc[k], tol2); return 1; }
    }
    if (tol2 > 0.05) { std::printf("tracked bound after the sum %.3g is useless\n", tol2); return 1; }
    Ctxt diff = cc; diff -= ca;
    std::vector<double> dd = decode(diff, &nl);
    for (long t = 0; t < 64; t++) { long k = (t * 521 + 3) % N; if (std::fabs(dd[k] - ((double)mc[k] - prod2[k])) > std::exp2(nl) + tol) { std::printf("CKKS difference mismatch at %ld\n", k); return 1; } }
    // constants: multByConstantCKKS and addConstantCKKS with an integer-coefficient constant encoded at factor 2^20
    {
      const double dc = std::ldexp(1.0, 20);
      std::vector<long> kc(N, 0), kcs(N, 0);
      kc[0] = 3; kc[1] = -2;                                   // constant polynomial 3 - 2X
      for (long k = 0; k < N; k++) kcs[k] = (long)(dc * kc[k]);
      Ctxt cm = encrypt(mc);
      DoubleCRT K(kcs, ctx, cm.primeSet);
      cm.multByConstantCKKS(K, XD(embeddingLargestCoeff(kc, m)), XD(dc), 0.0);
      std::vector<double> dm = decode(cm, &nl);
      for (long t = 0; t < 64; t++) {
        long k = (t * 521 + 3) % N;
        const long prev = k == 0 ? -mc[N - 1] : mc[k - 1];    // X * f wraps with a sign
        const double want = 3.0 * mc[k] - 2.0 * prev;
        if (std::fabs(dm[k] - want) > std::exp2(nl) + 1e-6) { std::printf("multByConstantCKKS mismatch at %ld: %g vs %g\n", k, dm[k], want); return 1; }
      }
      Ctxt cadd = encrypt(mc);
      DoubleCRT K2(kcs, ctx, cadd.primeSet);
      cadd.addConstantCKKS(K2, XD(embeddingLargestCoeff(kc, m)), XD(dc));   // ratFactor 2^30 / 2^20: the constant is scaled by 1024
      std::vector<double> da = decode(cadd, &nl);
      for (long k = 0; k < 4; k++) if (std::fabs(da[k] - (mc[k] + kc[k])) > std::exp2(nl) + 1e-6) { std::printf("addConstantCKKS mismatch at %ld: %g vs %ld\n", k, da[k], mc[k] + kc[k]); return 1; }
    }
    ctx.sync();
    std::printf("ckks add OK: log2 ratFactor %.1f + 30.0 -> %.1f, bound %.3g\n", rf_before, std::log2((double)sum.ratFactor.m) + (double)sum.ratFactor.e, tol2);
    std::printf("ckks OK: %.0f -> %.0f bits after multiplyBy, log2 ratFactor %.1f, error %.3g <= bound %.3g, KS-noise-ratio %.3g\n",
                logq0 / std::log(2.0), pk.logOfProduct(ca.primeSet) / std::log(2.0), std::log2((double)ca.ratFactor.m) + (double)ca.ratFactor.e, worst, tol, ca.lastKSNoiseRatio);
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
