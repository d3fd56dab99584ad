// End-to-end C++ test of the helib::Ctxt mirror (include/helib_b200_ctxt.hpp): BGV encrypt -> multiplyBy
// (noise-driven prime-set choice through getSet4Size, mod-switch, tensor, relinearise with device-computed
// norms) -> decrypt == plaintext product.  This is the reference's own test style (encrypt, operate, decrypt,
// compare with a plaintext mirror: tests/TestCtxt.cpp:110-142, tests/GTestGeneral.cpp:298-353).
// Exit codes: 0 ok, 3 no CUDA device, 1 failure.
#include <cstdio>
#include <random>
#include <sstream>
#include <cstring>

#include "helib_b200_ctxt.hpp"

using namespace hb;
typedef unsigned __int128 u128;

static std::vector<long> sample_ternary(std::mt19937_64& g, long n) { std::vector<long> v(n); for (auto& x : v) x = (long)(g() % 3) - 1; return v; }
static std::vector<long> sample_gauss(std::mt19937_64& g, long n, double sigma) { std::normal_distribution<double> d(0, sigma); std::vector<long> v(n); for (auto& x : v) x = std::lround(d(g)); return v; }
static DoubleCRT random_rows(const Context& ctx, const IndexSet& s, std::mt19937_64& g) {
  const long N = ctx.getPhiM();
  std::vector<uint64_t> dense((size_t)ctx.numPrimes() * N, 0);
  for (long i : s) for (long k = 0; k < N; k++) dense[(size_t)i * N + k] = g() % (uint64_t)ctx.ithPrime(i);
  return DoubleCRT::fromRows(ctx, s, dense);
}

int main() {
  if (hb_device_count() <= 0) { std::printf("no CUDA device\n"); return 3; }
  try {
    const long m = 8192, p = 257;
    Context ctx(m, p, 1, /*bits=*/300, /*c=*/2);
    const long N = ctx.getPhiM();
    std::mt19937_64 gen(20260922);
    const double sigma = 3.2;
    const IndexSet allq = ctx.getCtxtPrimes() | ctx.getSpecialPrimes();
    KeyInfo pk; pk.context = &ctx; pk.ckks = false; pk.scale = 10.0; pk.hwt = 0;
    pk.skBound = pk.scale * std::sqrt(double(N) * 2.0 / 3.0);    // high-probability bound on ||s||_canon for ternary s
    // secret key and the s^2 -> s key-switching matrix (SecKey::GenKeySWmatrix, src/keys.cpp:1159-1256)
    std::vector<long> s = sample_ternary(gen, N);
    DoubleCRT S(s, ctx, allq);
    // key-switching matrices through the mirror's GenKeySWmatrix / RLWE1, a_i by DoubleCRT::randomize over a byte stream
    auto bytes = [&](unsigned char* b, long n) { for (long i = 0; i < n; i++) b[i] = (unsigned char)gen(); };
    auto genKeySW = [&](const DoubleCRT& fromKey, const SKHandle& h) {
      pk.keySwitching.push_back(hb::genKeySWmatrix(ctx, fromKey, h, 0, S, p, false, sigma, gen, [&](DoubleCRT& a) { a.randomize(bytes); }));
    };
    { DoubleCRT s2(S); s2 *= S; genKeySW(s2, SKHandle(2, 1, 0)); }                 // s^2 -> s
    for (long amt : {3L, m - 1}) { DoubleCRT sk(S); sk.automorph(amt); genKeySW(sk, SKHandle(1, amt, 0)); }   // s(X^amt) -> s
    pk.setKeySwitchMap(0);

    // public encryption key: an RLWE1 encryption of zero over the ctxt primes (SecKey::GenSecKey, src/keys.cpp:1129-1157)
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
    auto encrypt = [&](const std::vector<long>& msg) {     // PubKey::Encrypt through the mirror (device path)
      Ctxt c(pk, p);
      hb::EncryptionSample smp = hb::drawEncryptionSample(ctx, sigma, gen);
      if (hb::Encrypt(c, pubEncrKey, msg, p, smp) != p) throw hb::LogicError("Encrypt changed the plaintext space");
      return c;
    };
    auto decrypt = [&](const Ctxt& c, double* maxabs) {   // SecKey::Decrypt through the mirror (device path)
      std::vector<long> out; std::vector<uint64_t> limbs; int L = 0;
      hb::Decrypt(out, c, sKeys, maxabs ? &limbs : nullptr, &L);
      if (maxabs) {
        double mx = 0;
        for (long k = 0; k < N; k++) {
          bool neg = limbs[(size_t)k * L + L - 1] >> 63; double mag = 0;
          for (int l = L - 1; l >= 0; l--) { uint64_t w = limbs[(size_t)k * L + l]; if (neg) w = ~w; mag = mag * 18446744073709551616.0 + (double)w; }
          mx = std::max(mx, mag + (neg ? 1 : 0));
        }
        *maxabs = mx;
      }
      return out;
    };

    std::vector<long> ma(N), mb(N), mc(N);
    for (long k = 0; k < N; k++) { ma[k] = (long)(gen() % p); mb[k] = (long)(gen() % p); mc[k] = (long)(gen() % p); }
    Ctxt ca = encrypt(ma), cb = encrypt(mb), cc = encrypt(mc);
    { std::vector<long> chk = decrypt(ca, nullptr); if (chk != ma) { std::printf("fresh decrypt mismatch\n"); return 1; } }
    auto negacyclic_at = [&](const std::vector<long>& x, const std::vector<long>& y, long k) {
      long acc = 0;
      for (long i = 0; i < N; i++) { long j = k - i; long term = j >= 0 ? x[i] * y[j] : -(x[i] * y[j + N]); acc = (acc + term) % p; }
      return ((acc % p) + p) % p;
    };
    const double ln2 = std::log(2.0);
    double logq0 = pk.logOfProduct(ca.primeSet);
    // ---- first product: fresh operands, getSet4Size picks a set inside [lo, hi] (may even add small primes)
    hb::setTimersOn(); hb::fhe_stats() = true;     // the reference's timers / statistics under their own names (src/timing.cpp)
    ca.multiplyBy(cb);
    hb::setTimersOff();
    {
      const hb::FHEtimer* tm = hb::getTimerByName("multiplyBy");
      const hb::FHEtimer* tk = hb::getTimerByName("KS_loop");
      const hb::FHEtimer* tr = hb::getTimerByName("reLinearize");
      if (!tm || !tk || !tr || tm->getNumCalls() != 1 || tk->getNumCalls() < 1 || tm->getTime() <= 0 || tm->getTime() < tr->getTime()) { std::printf("timers: multiplyBy / reLinearize / KS_loop not recorded\n"); return 1; }
      if (hb::fhe_stats_map().count("KS-noise-ratio") != 1 || hb::fhe_stats_map()["KS-noise-ratio"].count < 1) { std::printf("stats: KS-noise-ratio not recorded\n"); return 1; }
    }
    if (!ca.inCanonicalForm()) { std::printf("result not canonical\n"); return 1; }
    if (!(ctx.getSpecialPrimes() <= ca.primeSet)) { std::printf("special primes missing after reLinearize\n"); return 1; }
    double logq1 = pk.logOfProduct(ca.lastCommonPrimeSet);
    if (logq1 > ca.lastHi + 1e-9 || logq1 < ca.lastLo - ln2 - 1e-9) { std::printf("common set size %.1f outside [%.1f, %.1f] bits\n", logq1 / ln2, ca.lastLo / ln2, ca.lastHi / ln2); return 1; }
    if (!(ca.lastCommonPrimeSet & ctx.getCtxtPrimes()).isInterval()) { std::printf("common ctxt primes are not an interval\n"); return 1; }
    double maxabs = 0;
    std::vector<long> ab_full = decrypt(ca, &maxabs);
    std::vector<long> ab(N);
    for (long k = 0; k < N; k++) ab[k] = ab_full[k];
    for (long t = 0; t < 48; t++) {
      long k = (t * 131) % N;
      if (ab[k] != negacyclic_at(ma, mb, k)) { std::printf("product mismatch at %ld\n", k); return 1; }
    }
    double est = ca.noiseBound.ln() / ln2, act = std::log2(std::max(maxabs, 1.0));
    if (act > est) { std::printf("noise estimate too small: 2^%.1f < actual 2^%.1f\n", est, act); return 1; }
    // ---- second product: the noisy product times a fresh ciphertext must shrink the modulus
    ca.multiplyBy(cc);
    double logq2 = pk.logOfProduct(ca.lastCommonPrimeSet);
    if (logq2 >= logq1) { std::printf("second product did not shrink the modulus (%.1f -> %.1f bits)\n", logq1 / ln2, logq2 / ln2); return 1; }
    std::vector<long> abc = decrypt(ca, &maxabs);
    for (long t = 0; t < 32; t++) {
      long k = (t * 257 + 5) % N;
      if (abc[k] != negacyclic_at(ab, mc, k)) { std::printf("second product mismatch at %ld\n", k); return 1; }
    }
    est = ca.noiseBound.ln() / ln2; act = std::log2(std::max(maxabs, 1.0));
    if (act > est) { std::printf("noise estimate too small after 2 products: 2^%.1f < actual 2^%.1f\n", est, act); return 1; }
    // ---- operands with different plaintext spaces (p^2 x p): multLowLvl must equalise to gcd = p on BOTH operands
    //      (src/Ctxt.cpp:1717-1725) and reLinearize must reduce to the matrix's space (src/Ctxt.cpp:771-775)
    {
      const long p2 = p * p;
      Ctxt pubEncrKey2(pk, p2);
      pubEncrKey2.primeSet = ctx.getCtxtPrimes();
      std::vector<long> e2 = sample_gauss(gen, N, sigma);
      DoubleCRT k1 = random_rows(ctx, pubEncrKey2.primeSet, gen);
      DoubleCRT k0(e2, ctx, pubEncrKey2.primeSet); k0 *= p2;
      { DoubleCRT t(k1); t.Mul(S, false); k0 -= t; }
      pubEncrKey2.parts.emplace_back(k0, SKHandle());
      pubEncrKey2.parts.emplace_back(k1, SKHandle(1, 1, 0));
      pubEncrKey2.noiseBound = XD(double(p2) * pk.noiseBoundForGaussian(sigma, N));
      std::vector<long> mw(N);
      for (long k = 0; k < N; k++) mw[k] = (long)(gen() % p2);
      Ctxt cw(pk, p2);
      hb::EncryptionSample smp = hb::drawEncryptionSample(ctx, sigma, gen);
      if (hb::Encrypt(cw, pubEncrKey2, mw, p2, smp) != p2) { std::printf("Encrypt at p^2 changed the plaintext space\n"); return 1; }
      if (decrypt(cw, nullptr) != mw) { std::printf("p^2 fresh decrypt mismatch\n"); return 1; }
      Ctxt cn = encrypt(mb);
      cw.multiplyBy(cn);
      if (cw.ptxtSpace != p) { std::printf("mixed plaintext spaces: product space %ld, expected %ld\n", cw.ptxtSpace, p); return 1; }
      std::vector<long> mwp(N); for (long k = 0; k < N; k++) mwp[k] = mw[k] % p;
      std::vector<long> got = decrypt(cw, nullptr);
      for (long t = 0; t < 32; t++) {
        long k = (t * 193 + 7) % N;
        if (got[k] != negacyclic_at(mwp, mb, k)) { std::printf("mixed plaintext spaces: product mismatch at %ld\n", k); return 1; }
      }
      bool threw = false;
      try { Ctxt bad = encrypt(ma); bad.ptxtSpace = 3; bad.intFactor = 1; Ctxt o = encrypt(mb); bad.multLowLvl(o); } catch (const hb::LogicError&) { threw = true; }
      if (!threw) { std::printf("co-prime plaintext spaces must throw\n"); return 1; }
    }
    // drop the special primes again (cleanUp path) and decrypt once more
    ca.dropSmallAndSpecialPrimes();
    if (decrypt(ca, nullptr) != abc) { std::printf("mod-down changed the plaintext\n"); return 1; }
    // rotations: smartAutomorph in map-driven steps, and the hoisted form sharing one digit decomposition
    {
      auto apply = [&](const std::vector<long>& f, long k) {   // f(X^k) mod (X^N + 1, p)
        std::vector<long> g(N, 0);
        for (long i = 0; i < N; i++) { long e = (long)(((unsigned __int128)(unsigned long)i * (unsigned long)k) % (unsigned long)m); long v = f[i]; if (e >= N) { e -= N; v = (p - v) % p; } g[e] = (g[e] + v) % p; }
        return g;
      };
      Ctxt cr = encrypt(mb);
      Ctxt c9 = cr; c9.smartAutomorph(9);                        // two steps of 3
      if (!c9.inCanonicalForm() || decrypt(c9, nullptr) != apply(mb, 9)) { std::printf("smartAutomorph(9) mismatch\n"); return 1; }
      Ctxt cc2 = cr; cc2.smartAutomorph(m - 1);                  // complex conjugation matrix
      if (decrypt(cc2, nullptr) != apply(mb, m - 1)) { std::printf("smartAutomorph(m-1) mismatch\n"); return 1; }
      hb::BasicAutomorphPrecon pre(cr);
      for (long k : {3L, 27L, (3 * (m - 1)) % m}) {
        auto h = pre.automorph(k);
        double mx = 0;
        if (decrypt(*h, &mx) != apply(mb, k)) { std::printf("hoisted automorph(%ld) mismatch\n", k); return 1; }
        if (std::log2(std::max(mx, 1.0)) > h->noiseBound.ln() / ln2) { std::printf("hoisted noise estimate too small\n"); return 1; }
      }
      // hoisting must agree with the plain path bit for bit when one step suffices (sigma_k commutes with the digits)
      Ctxt c3 = cr; c3.smartAutomorph(3);
      auto h3 = pre.automorph(3);
      for (size_t i = 0; i < 2; i++)
        if (c3.parts[i].dcrt.getOneRow(c3.primeSet.first()) != h3->parts[i].dcrt.getOneRow(c3.primeSet.first())) { std::printf("hoisted rows differ from smartAutomorph rows\n"); return 1; }
      bool threw = false;
      try { Ctxt bad(pk, p); bad = cr; bad.parts[1].skHandle.secretKeyID = 1; bad.smartAutomorph(3); } catch (const hb::LogicError&) { threw = true; }
      if (!threw) { std::printf("missing LogicError for unreachable automorphism\n"); return 1; }
    }
    // linear operations: addCtxt with unequal prime sets and integer factors, addConstant, multByConstant
    {
      Ctxt x = encrypt(ma), y = encrypt(mb), z = encrypt(mc);
      Ctxt sum = x; sum += y;
      std::vector<long> d = decrypt(sum, nullptr);
      for (long k = 0; k < N; k++) if (d[k] != (ma[k] + mb[k]) % p) { std::printf("addCtxt mismatch at %ld\n", k); return 1; }
      Ctxt xy = x; xy.multiplyBy(y);                 // fewer primes, intFactor != 1 in general
      Ctxt t = xy; t += z;                           // z is mod-UPped and the factors harmonised
      Ctxt u = z; u -= xy;
      std::vector<long> dt = decrypt(t, nullptr), du = decrypt(u, nullptr);
      for (long i = 0; i < 32; i++) {
        long k = (i * 131 + 7) % N, prod = negacyclic_at(ma, mb, k);
        if (dt[k] != (prod + mc[k]) % p) { std::printf("product + fresh mismatch at %ld\n", k); return 1; }
        if (du[k] != ((mc[k] - prod) % p + p) % p) { std::printf("fresh - product mismatch at %ld\n", k); return 1; }
      }
      Ctxt v = xy;
      DoubleCRT cst(mc, ctx, v.primeSet);
      v.addConstant(cst);
      std::vector<long> dv = decrypt(v, nullptr);
      for (long i = 0; i < 32; i++) { long k = (i * 131 + 7) % N; if (dv[k] != (negacyclic_at(ma, mb, k) + mc[k]) % p) { std::printf("addConstant mismatch at %ld\n", k); return 1; } }
      Ctxt w = x;
      DoubleCRT cst2(mb, ctx, w.primeSet);
      w.multByConstant(cst2);
      std::vector<long> dw = decrypt(w, nullptr);
      for (long i = 0; i < 32; i++) { long k = (i * 131 + 7) % N; if (dw[k] != negacyclic_at(ma, mb, k)) { std::printf("multByConstant mismatch at %ld\n", k); return 1; } }
    }
    // binary wire format of Ctxt::writeTo / read (src/Ctxt.cpp:2584-2641): layout, size and round trip
    {
      std::stringstream ss;
      ca.writeTo(ss);
      const std::string bytes = ss.str();
      const size_t card = (size_t)ca.primeSet.card(), np = ca.parts.size();
      const size_t expect = 24 + 4 + 16 + 48 + 8 + 8 * card + 8 + np * ((8 + 8 * card) + card * (8 + 8 * (size_t)N) + 24) + 4;
      if (bytes.size() != expect) { std::printf("serialized size %zu != %zu\n", bytes.size(), expect); return 1; }
      if (bytes.compare(0, 4, "|HE[") != 0 || bytes[6] != 1 || bytes[12] != 20 || bytes.compare(20, 8, "]HE||CX[") != 0 || bytes.compare(bytes.size() - 4, 4, "]CX|") != 0) { std::printf("header / eye catchers\n"); return 1; }
      int64_t ps = 0; std::memcpy(&ps, bytes.data() + 28, 8);
      if (ps != ca.ptxtSpace) { std::printf("ptxtSpace field\n"); return 1; }
      Ctxt back(pk, 0);
      back.read(ss);
      if (back.ptxtSpace != ca.ptxtSpace || back.intFactor != ca.intFactor || back.primeSet != ca.primeSet || back.parts.size() != np) { std::printf("round trip: metadata\n"); return 1; }
      if (std::fabs(back.noiseBound.ln() - ca.noiseBound.ln()) > 1e-12) { std::printf("round trip: noiseBound\n"); return 1; }
      for (size_t i = 0; i < np; i++)
        if (!(back.parts[i].skHandle == ca.parts[i].skHandle) || back.parts[i].dcrt.getOneRow(ca.primeSet.first()) != ca.parts[i].dcrt.getOneRow(ca.primeSet.first())) { std::printf("round trip: part %zu\n", i); return 1; }
      if (decrypt(back, nullptr) != decrypt(ca, nullptr)) { std::printf("round trip: decryption differs\n"); return 1; }
      std::string bad = bytes; bad[25] = 'Q';
      std::stringstream sb(bad);
      bool threw = false;
      try { Ctxt x(pk, 0); x.read(sb); } catch (const hb::RuntimeError&) { threw = true; }
      if (!threw) { std::printf("missing error for a damaged eye catcher\n"); return 1; }
    }
    const long before = ctx.getCtxtPrimes().card(), common = ca.lastCommonPrimeSet.card();
    const double logq_before = logq0, logq_common = logq2;
    ctx.sync();
    std::printf("ctxt OK: %ld primes (%.0f bits) -> common %ld primes (%.0f bits), log2 noise est %.1f >= actual %.1f, KS-noise-ratio %.3g, mod-switch ratio %.3g\n",
                before, logq_before / std::log(2.0), common, logq_common / std::log(2.0), est, act, ca.lastKSNoiseRatio, ca.lastModSwitchRatio);
    return 0;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
}
