// C++ test of the helib::DoubleCRT mirror (include/helib_b200_doublecrt.hpp) over the C ABI.
// Exercises the reference's own unit-level properties for this path:
//   FFT o iFFT = identity on polynomials          (tests/TestHEXL.cpp:189-218)
//   addPrimes / removePrimes round trip            
//   index-set precondition failures raise          (src/DoubleCRT.cpp:227-253)
//   (a*b) computed in evaluation form == schoolbook negacyclic product (toPoly)
// Exit codes: 0 ok, 3 no CUDA device (expected on the CPU box), 1 failure.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "helib_b200_doublecrt.hpp"

static long coeff_of(const std::vector<uint64_t>& limbs, int L, long k) {  // small values only
  return (long)limbs[(size_t)k * L];
}

int main() {
  if (hb_device_count() <= 0) { std::printf("no CUDA device: %s\n", "engine has no CPU path"); return 3; }
  try {
    hb::Context ctx(/*m=*/4096, /*p=*/257, /*r=*/1, /*bits=*/120, /*c=*/2);
    const long N = ctx.getPhiM();
    const hb::IndexSet S = ctx.getCtxtPrimes();
    std::vector<long> f(N), g(N);
    unsigned long long st = 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (long)((st >> 33) % 21) - 10; };
    for (long i = 0; i < N; i++) { f[i] = rnd(); g[i] = rnd(); }
    hb::DoubleCRT F(f, ctx, S), G(g, ctx, S);
    // product in evaluation form vs schoolbook
    hb::DoubleCRT H(F);
    H *= G;
    int L = 0;
    std::vector<uint64_t> hp = H.toPoly(S, false, L);
    for (long k : {0L, 1L, N / 2, N - 1}) {
      long acc = 0;
      for (long i = 0; i < N; i++) { long j = k - i; if (j >= 0) acc += f[i] * g[j]; else acc -= f[i] * g[j + N]; }
      if (coeff_of(hp, L, k) != acc) { std::printf("product mismatch at %ld\n", k); return 1; }
    }
    // Cmodulus surface (include/helib/CModulus.h:104-157): the monomial X evaluates to the root itself in y[0] (the calibration
    // SURVEY 8c describes), y[j] = psi^(2j+1), and iFFT inverts FFT
    {
      hb::Cmodulus cm(ctx, S.first());
      const long q = cm.getQ(), psi = cm.getRoot();
      if ((long)cm.getM() != 4096 || (long)cm.getPhiM() != N) { std::printf("Cmodulus: m / phi(m)\n"); return 1; }
      std::vector<long> mono(2, 0), y, back; mono[1] = 1;
      cm.FFT(y, mono);
      if (y[0] != psi) { std::printf("Cmodulus::FFT(X)[0] = %ld, root = %ld\n", y[0], psi); return 1; }
      unsigned __int128 w = (unsigned __int128)psi * psi % q, cur = psi;
      for (long j = 0; j < 8; j++) { if ((long)cur != y[j]) { std::printf("Cmodulus::FFT(X)[%ld] != psi^(2j+1)\n", j); return 1; } cur = cur * w % q; }
      cm.FFT(y, f);
      cm.iFFT(back, y);
      for (long k = 0; k < N; k++) { long want = f[k] < 0 ? f[k] + q : f[k]; if (back[k] != want) { std::printf("Cmodulus: iFFT(FFT(f)) != f at %ld\n", k); return 1; } }
      if (y != F.getOneRow(S.first())) { std::printf("Cmodulus::FFT row differs from the DoubleCRT row\n"); return 1; }
    }
    // addPrimes to the special primes and back
    hb::DoubleCRT E(F);
    E.addPrimes(ctx.getSpecialPrimes());
    std::vector<long> row_new = E.getOneRow(ctx.getSpecialPrimes().first());
    hb::DoubleCRT F2(f, ctx, S | ctx.getSpecialPrimes());
    if (row_new != F2.getOneRow(ctx.getSpecialPrimes().first())) { std::printf("addPrimes row mismatch\n"); return 1; }
    E.removePrimes(ctx.getSpecialPrimes());
    if (E.getOneRow(S.first()) != F.getOneRow(S.first())) { std::printf("removePrimes changed a row\n"); return 1; }
    // scaleDownToSet keeps the value / P up to the small rounding term: (P*x) scaled down by P == x
    hb::DoubleCRT X(F);
    X.addPrimesAndScale(ctx.getSpecialPrimes());
    X.scaleDownToSet(S, 1);
    if (X.getOneRow(S.first()) != F.getOneRow(S.first())) { std::printf("scale up/down is not the identity\n"); return 1; }
    // digits: two digits, each over S | special
    std::vector<hb::DoubleCRT> digits;
    F.breakIntoDigits(digits);
    if (digits.size() != ctx.getDigits().size()) { std::printf("digit count\n"); return 1; }
    // ZZX-style construction: toPoly limbs -> rows reproduces the object (DoubleCRT(ZZX) o toPoly = identity)
    {
      hb::DoubleCRT R = hb::DoubleCRT::fromLimbs(ctx, S, hp, L);
      if (R.getOneRow(S.first()) != H.getOneRow(S.first()) || R.getOneRow(S.last()) != H.getOneRow(S.last())) { std::printf("fromLimbs round trip\n"); return 1; }
      // mulAdd: Z = G; Z += F*G  ==  H + G
      hb::DoubleCRT Z(G); Z.mulAdd(F, G);
      hb::DoubleCRT W(H); W += G;
      if (Z.getOneRow(S.first()) != W.getOneRow(S.first())) { std::printf("mulAdd mismatch\n"); return 1; }
      // toPolyModP == the balanced coefficients reduced into [0,p)
      std::vector<long> mp = H.toPolyModP(S, 257, 1);
      for (long k : {0L, 1L, N / 2, N - 1}) { long v = coeff_of(hp, L, k) % 257; if (v < 0) v += 257; if (mp[k] != v) { std::printf("toPolyModP mismatch at %ld\n", k); return 1; } }
    }
    // randomize from a byte stream (splitmix64 bytes; tests/test_cpp_shim.py recomputes the rows with the oracle)
    {
      uint64_t st2 = 0x1234567ULL;
      auto get = [&](unsigned char* b, long n) {
        for (long i = 0; i < n; i += 8) {
          st2 += 0x9E3779B97F4A7C15ULL; uint64_t z = st2;
          z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
          for (int k = 0; k < 8 && i + k < n; k++) b[i + k] = (unsigned char)(z >> (8 * k));
        }
      };
      hb::DoubleCRT R(ctx, S);
      R.randomize(get);
      std::printf("randomize:");
      for (long i : S) { std::vector<long> row = R.getOneRow(i); unsigned long h = 1469598103934665603UL; for (long v : row) h = (h ^ (unsigned long)v) * 1099511628211UL; std::printf(" %ld:%lu", i, h); }
      std::printf("\n");
    }
    // error behaviour
    bool threw = false;
    try { hb::DoubleCRT A(ctx, S), B(ctx, hb::IndexSet(S.first())); A += B; } catch (const hb::RuntimeError&) { threw = true; }
    if (!threw) { std::printf("missing RuntimeError for index-set violation\n"); return 1; }
    threw = false;
    try { hb::DoubleCRT A(F); A.addPrimes(hb::IndexSet(S.first())); } catch (const hb::RuntimeError&) { threw = true; }
    if (!threw) { std::printf("missing RuntimeError for non-disjoint addPrimes\n"); return 1; }
    ctx.sync();
    std::printf("shim OK: N=%ld primes=%ld digits=%zu\n", N, ctx.numPrimes(), digits.size());
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
