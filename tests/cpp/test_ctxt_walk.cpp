// Random walk over the helib::Ctxt mirror (include/helib_b200_ctxt.hpp): a pool of BGV ciphertexts and their plaintext
// polynomials in Z_p[X]/(X^N+1); random sequences of +=, -=, multiplyBy, multByConstant, addConstant, negate, smartAutomorph,
// dropSmallAndSpecialPrimes; after every step the touched ciphertext is decrypted and compared coefficient for coefficient
// with the plaintext mirror, and the tracked noise bound must dominate the measured noise.  The reference tests its Ctxt the
// same way on fixed sequences (tests/TestCtxt.cpp:110-142, tests/GTestGeneral.cpp:298-353); this one walks.
// usage: test_ctxt_walk [seed] [steps].  Exit codes: 0 ok, 3 no CUDA device, 1 failure.
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

int main(int argc, char** argv) {
  if (hb_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
  const unsigned long seed = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 1;
  const int steps = argc > 2 ? std::atoi(argv[2]) : 40;
  try {
    const long m = 2048, p = 257;
    Context ctx(m, p, 1, /*bits=*/360, /*c=*/3);
    const long N = ctx.getPhiM();
    std::mt19937_64 gen(20260923ULL + seed);
    const double sigma = 3.2;
    const IndexSet allq = ctx.getCtxtPrimes() | ctx.getSpecialPrimes();
    KeyInfo pk; pk.context = &ctx; pk.ckks = false; pk.scale = 10.0; pk.hwt = 0;
    pk.skBound = pk.scale * std::sqrt(double(N) * 2.0 / 3.0);
    std::vector<long> s = sample_ternary(gen, N);
    DoubleCRT S(s, ctx, allq);
    auto bytes = [&](unsigned char* b, long n) { for (long i = 0; i < n; i++) b[i] = (unsigned char)gen(); };
    auto genKeySW = [&](const DoubleCRT& fromKey, const SKHandle& h) {
      pk.keySwitching.push_back(hb::genKeySWmatrix(ctx, fromKey, h, 0, S, p, false, sigma, gen, [&](DoubleCRT& a) { a.randomize(bytes); }));
    };
    { DoubleCRT s2(S); s2 *= S; genKeySW(s2, SKHandle(2, 1, 0)); }
    const long amts[2] = {3, m - 1};
    for (long amt : amts) { DoubleCRT sk(S); sk.automorph(amt); genKeySW(sk, SKHandle(1, amt, 0)); }
    pk.setKeySwitchMap(0);
    Ctxt pubEncrKey(pk, p);
    {
      pubEncrKey.primeSet = ctx.getCtxtPrimes();
      std::vector<long> e = sample_gauss(gen, N, sigma);
      DoubleCRT c1 = random_rows(ctx, pubEncrKey.primeSet, gen);
      DoubleCRT c0(e, ctx, pubEncrKey.primeSet); c0 *= p;
      DoubleCRT t(c1); t.Mul(S, false); c0 -= t;
      pubEncrKey.parts.emplace_back(c0, SKHandle());
      pubEncrKey.parts.emplace_back(c1, SKHandle(1, 1, 0));
      pubEncrKey.noiseBound = XD(double(p) * pk.noiseBoundForGaussian(sigma, N));
    }
    std::vector<DoubleCRT> sKeys; sKeys.push_back(S);
    auto encrypt = [&](const std::vector<long>& msg) {
      Ctxt c(pk, p);
      hb::EncryptionSample smp = hb::drawEncryptionSample(ctx, sigma, gen);
      if (hb::Encrypt(c, pubEncrKey, msg, p, smp) != p) throw hb::LogicError("Encrypt changed the plaintext space");
      return c;
    };
    auto decrypt = [&](const Ctxt& c, double* maxabs) {
      std::vector<long> out; std::vector<uint64_t> limbs; int L = 0;
      hb::Decrypt(out, c, sKeys, &limbs, &L);
      double mx = 0;
      for (long k = 0; k < N; k++) {
        bool neg = limbs[(size_t)k * L + L - 1] >> 63; double mag = 0;
        for (int l = L - 1; l >= 0; l--) { uint64_t w = limbs[(size_t)k * L + l]; if (neg) w = ~w; mag = mag * 18446744073709551616.0 + (double)w; }
        mx = std::max(mx, mag + (neg ? 1 : 0));
      }
      *maxabs = mx;
      return out;
    };
    auto rand_msg = [&]() { std::vector<long> v(N); for (auto& x : v) x = (long)(gen() % p); return v; };
    auto mul = [&](const std::vector<long>& x, const std::vector<long>& y) {   // negacyclic product mod p
      std::vector<long> r(N, 0);
      for (long i = 0; i < N; i++) { if (!x[i]) continue; for (long j = 0; j < N; j++) { long k = i + j; long t = x[i] * y[j] % p; if (k >= N) { k -= N; t = (p - t) % p; } r[k] = (r[k] + t) % p; } }
      return r;
    };
    auto autom = [&](const std::vector<long>& f, long k) {
      std::vector<long> g(N, 0);
      for (long i = 0; i < N; i++) { long e = (long)(((unsigned __int128)(unsigned long)i * (unsigned long)k) % (unsigned long)m); long v = f[i]; if (e >= N) { e -= N; v = (p - v) % p; } g[e] = (g[e] + v) % p; }
      return g;
    };
    const double ln2 = std::log(2.0);
    const int POOL = 4;
    std::vector<Ctxt> ct; std::vector<std::vector<long>> pt;
    for (int i = 0; i < POOL; i++) { pt.push_back(rand_msg()); ct.push_back(encrypt(pt.back())); }
    auto capacity_bits = [&](const Ctxt& c) { return (pk.logOfProduct(c.primeSet & ctx.getCtxtPrimes()) - c.noiseBound.ln()) / ln2; };
    int nmul = 0, nrefresh = 0; double worst_margin = 1e9;
    for (int st = 0; st < steps; st++) {
      const int a = (int)(gen() % POOL); int b = (int)(gen() % POOL);
      const int op = (int)(gen() % 9);
      const char* name = "";
      if ((op == 2 || op == 6) && (capacity_bits(ct[a]) < 150 || capacity_bits(ct[b]) < 150)) {   // no room for a product / rotation: recrypt by hand
        pt[a] = rand_msg(); ct[a] = encrypt(pt[a]); nrefresh++; name = "refresh";
      } else switch (op) {
        case 0: name = "+="; { Ctxt o = ct[b]; ct[a] += o; } for (long k = 0; k < N; k++) pt[a][k] = (pt[a][k] + pt[b][k]) % p; break;
        case 1: name = "-="; { Ctxt o = ct[b]; ct[a] -= o; } for (long k = 0; k < N; k++) pt[a][k] = ((pt[a][k] - pt[b][k]) % p + p) % p; break;
        case 2: name = "multiplyBy"; { Ctxt o = ct[b]; ct[a].multiplyBy(o); } pt[a] = mul(pt[a], std::vector<long>(pt[b])); nmul++; break;
        case 3: { name = "multByConstant"; std::vector<long> cst = rand_msg(); DoubleCRT d(cst, ctx, ct[a].primeSet); ct[a].multByConstant(d); pt[a] = mul(pt[a], cst); } break;
        case 4: { name = "addConstant"; std::vector<long> cst = rand_msg(); DoubleCRT d(cst, ctx, ct[a].primeSet); ct[a].addConstant(d); for (long k = 0; k < N; k++) pt[a][k] = (pt[a][k] + cst[k]) % p; } break;
        case 5: name = "negate"; ct[a].negate(); for (long k = 0; k < N; k++) pt[a][k] = (p - pt[a][k]) % p; break;
        case 6: { const long ks[4] = {3, 9, m - 1, 27}; const long k = ks[gen() % 4]; name = "smartAutomorph"; ct[a].smartAutomorph(k); pt[a] = autom(pt[a], k); } break;
        case 7: name = "dropSmallAndSpecialPrimes"; ct[a].dropSmallAndSpecialPrimes(); break;
        default: name = "copy"; ct[a] = ct[b]; pt[a] = pt[b]; break;
      }
      double maxabs = 0;
      std::vector<long> got = decrypt(ct[a], &maxabs);
      if (got != pt[a]) { long bad = 0; while (got[bad] == pt[a][bad]) bad++; std::printf("seed %lu step %d (%s): plaintext mismatch at coefficient %ld (%ld != %ld)\n", seed, st, name, bad, got[bad], pt[a][bad]); return 1; }
      const double est = ct[a].noiseBound.ln() / ln2, act = std::log2(std::max(maxabs, 1.0));
      if (act > est) { std::printf("seed %lu step %d (%s): noise estimate 2^%.1f below the actual 2^%.1f\n", seed, st, name, est, act); return 1; }
      worst_margin = std::min(worst_margin, est - act);
    }
    ctx.sync();
    std::printf("walk OK: seed %lu, %d steps, %d products, %d refreshes, smallest estimate-actual margin %.1f bits\n", seed, steps, nmul, nrefresh, worst_margin);
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
