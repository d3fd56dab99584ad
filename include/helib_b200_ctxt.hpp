// helib_b200_ctxt.hpp -- header-only C++17 mirror of the hot subset of helib::Ctxt over hb::DoubleCRT.
//
// Host orchestration stays on the host exactly as in the reference:
// noise estimates in extended-range floating point, the choice of the prime set for a product
// (computeIntervalForMul + ModuliSizes::getSet4Size), when to mod-switch and when to key-switch.  Every data
// operation goes to the engine through hb::DoubleCRT.  Method names, argument meaning and failure behaviour
// follow the reference (citations: paths in the HElib tree):
//   SKHandle::mul                  include/helib/Ctxt.h:155-185
//   modUpToSet / bringToSet        src/Ctxt.cpp:346-389
//   modDownToSet                   src/Ctxt.cpp:393-562      (added noise from the device-computed ||delta/P||)
//   dropSmallAndSpecialPrimes      src/Ctxt.cpp:589-662
//   relin_CKKS_adjust              src/Ctxt.cpp:664-716
//   reLinearize / keySwitchPart    src/Ctxt.cpp:720-842
//   keySwitchDigits                src/Ctxt.cpp:191-230
//   tensorProduct                  src/Ctxt.cpp:1563-1608
//   computeIntervalForMul          src/Ctxt.cpp:1610-1656
//   multLowLvl / multiplyBy        src/Ctxt.cpp:1681-1774
//   modSwitchAddedNoiseBound       src/Ctxt.cpp:2560-2582
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <complex>
#include <functional>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <ostream>
#include <istream>
#include <cstring>

#include "helib_b200_doublecrt.hpp"

namespace hb {

// NTL::xdouble stand-in: value = m * 2^e, enough range for noise bounds of 2^2000 and beyond.
struct XD {
  double m = 0; long e = 0;
  XD() = default;
  XD(double v) { int ex = 0; m = std::frexp(v, &ex); e = ex; }
  static XD make(double m_, long e_) { XD r; int ex = 0; r.m = std::frexp(m_, &ex); r.e = e_ + ex; if (r.m == 0) r.e = 0; return r; }
  static XD exp(double lnv) {
    if (!(lnv > -1e300)) return XD();   // ln 0 = -inf (the engine's answer for a zero polynomial) or ln() of a zero XD
    double l2 = lnv / std::log(2.0); long fl = (long)std::floor(l2); return make(std::exp2(l2 - fl), fl);
  }
  double ln() const { return m <= 0 ? -DBL_MAX : std::log(m) + e * std::log(2.0); }
  double to_double() const { return std::ldexp(m, (int)std::max(-2000L, std::min(2000L, e))); }
  XD operator*(const XD& o) const { return make(m * o.m, e + o.e); }
  XD operator/(const XD& o) const { return make(m / o.m, e - o.e); }
  XD operator+(const XD& o) const {
    if (m == 0) return o;
    if (o.m == 0) return *this;
    if (e >= o.e) { long d = e - o.e; return d > 1100 ? *this : make(m + std::ldexp(o.m, (int)-d), e); }
    return o + *this;
  }
  XD operator-() const { XD r = *this; r.m = -r.m; return r; }
  XD operator-(const XD& o) const { return *this + (-o); }
  XD abs() const { XD r = *this; r.m = std::fabs(r.m); return r; }
  // exact integer value floor(v) of a non-negative XD as mant * 2^shift (mant < 2^63)
  void floorParts(uint64_t& mant, long& shift) const {
    if (m <= 0 || e <= 0) { mant = 0; shift = 0; return; }
    if (e <= 62) { mant = (uint64_t)std::floor(std::ldexp(m, (int)e)); shift = 0; return; }
    mant = (uint64_t)std::ldexp(m, 53); shift = e - 53;      // 53-bit mantissa: already an integer
  }
  bool operator<(const XD& o) const { if (m <= 0 || o.m <= 0) return m < o.m; return e != o.e ? e < o.e : m < o.m; }
  bool operator>(const XD& o) const { return o < *this; }
  bool operator<=(const XD& o) const { return !(o < *this); }
};

// include/helib/Ctxt.h:82-260
struct SKHandle {
  long powerOfS = 0, powerOfX = 1, secretKeyID = 0;
  SKHandle() = default;
  SKHandle(long s, long x, long id) : powerOfS(s), powerOfX(x), secretKeyID(id) {}
  bool isOne() const { return powerOfS == 0; }
  bool isBase(long id = 0) const { return powerOfS == 1 && powerOfX == 1 && (id < 0 || secretKeyID == id); }
  bool operator==(const SKHandle& o) const { return powerOfS == o.powerOfS && powerOfX == o.powerOfX && secretKeyID == o.secretKeyID; }
  bool mul(const SKHandle& a, const SKHandle& b) {   // include/helib/Ctxt.h:155-185
    if (a.isOne()) { *this = b; return b.secretKeyID >= 0; }
    if (b.isOne()) { *this = a; return a.secretKeyID >= 0; }
    if (a.secretKeyID == -1 || b.secretKeyID == -1 || a.secretKeyID != b.secretKeyID || a.powerOfX != b.powerOfX) { secretKeyID = -1; return false; }
    secretKeyID = a.secretKeyID; powerOfX = a.powerOfX; powerOfS = a.powerOfS + b.powerOfS;
    return true;
  }
};

struct CtxtPart {
  DoubleCRT dcrt;
  SKHandle skHandle;
  CtxtPart(const DoubleCRT& d, const SKHandle& h) : dcrt(d), skHandle(h) {}
};

// helib::KeySwitch (include/helib/keySwitching.h:86-100) with the pseudo-random a_i expanded once
struct KeySwitch {
  SKHandle fromKey; long toKeyID = 0; long ptxtSpace = 0;
  std::vector<DoubleCRT> a, b;
  XD noiseBound;
};

// the part of helib::PubKey / Context the ciphertext logic consults
struct KeyInfo {
  const Context* context;
  bool ckks = false;
  double scale = 10.0;           // Context::scale (include/helib/Context.h:151)
  long hwt = 0;                  // Context::getHwt()
  double skBound = 0;            // PubKey::getSKeyBound (src/keys.cpp:280)
  std::vector<KeySwitch> keySwitching;
  const KeySwitch* getKeySWmatrix(const SKHandle& from, long toID) const {
    for (auto& w : keySwitching) if (w.fromKey == from && w.toKeyID == toID) return &w;
    return nullptr;
  }
  // PubKey::setKeySwitchMap / getNextKSWmatrix / isReachable (src/keys.cpp:122-172,310-319): BFS from 1 over the
  // automorphism matrices W[s(X^n) -> s(X)]; map[k] = matrix of the first step towards k
  std::vector<std::vector<long>> keySwitchMap;
  void setKeySwitchMap(long keyId = 0) {
    const long m = context->getM();
    std::vector<std::pair<long, long>> edges;
    for (size_t i = 0; i < keySwitching.size(); i++) {
      const KeySwitch& mat = keySwitching[i];
      if (mat.toKeyID == keyId && mat.fromKey.powerOfS == 1 && mat.fromKey.secretKeyID == keyId) edges.emplace_back(mat.fromKey.powerOfX, (long)i);
    }
    if (keyId >= (long)keySwitchMap.size()) keySwitchMap.resize(keyId + 1);
    keySwitchMap[keyId].assign((size_t)m, -1);
    std::vector<long> queue{1};
    for (size_t h = 0; h < queue.size(); h++)
      for (auto& e : edges) {
        long next = (long)(((unsigned __int128)(unsigned long)queue[h] * (unsigned long)e.first) % (unsigned long)m);
        if (keySwitchMap[keyId][next] == -1) { keySwitchMap[keyId][next] = e.second; queue.push_back(next); }
      }
  }
  bool isReachable(long k, long keyID) const { return keyID < (long)keySwitchMap.size() && keySwitchMap[keyID].at((size_t)k) >= 0; }
  const KeySwitch* getNextKSWmatrix(long fromXPower, long fromID) const {
    long i = keySwitchMap.at((size_t)fromID).at((size_t)fromXPower);
    return i >= 0 ? &keySwitching[(size_t)i] : nullptr;
  }
  double noiseBoundForUniform(double mag, long deg) const { return scale * std::sqrt(double(deg) / 3.0) * mag; }   // include/helib/Context.h:475-478
  double noiseBoundForMod(long modulus, long deg) const {   // include/helib/Context.h:517-524
    double var = double(modulus) * double(modulus) / 12.0; if (modulus % 2 == 0) var += 1.0 / 6.0;
    return scale * std::sqrt(deg * var);
  }
  double noiseBoundForGaussian(double sigma, long deg) const { return scale * std::sqrt(double(deg)) * sigma; }   // include/helib/Context.h:541-544
  double logOfProduct(const IndexSet& s) const { double x = 0; for (long i : s) x += std::log((double)context->ithPrime(i)); return x; }
};

class Ctxt {
 public:
  const KeyInfo& pubKey;
  const Context& context;
  std::vector<CtxtPart> parts;
  IndexSet primeSet;
  long ptxtSpace;
  XD noiseBound;
  long intFactor = 1;
  XD ratFactor = XD(1.0), ptxtMag = XD(1.0);
  static constexpr double safety = 0.6931471805599453;   // log 2, top of src/Ctxt.cpp

  Ctxt(const KeyInfo& pk, long ptxtSp) : pubKey(pk), context(*pk.context), ptxtSpace(ptxtSp), noiseBound(0.0) {}
  Ctxt& operator=(const Ctxt& o) {
    parts = o.parts; primeSet = o.primeSet; ptxtSpace = o.ptxtSpace; noiseBound = o.noiseBound;
    intFactor = o.intFactor; ratFactor = o.ratFactor; ptxtMag = o.ptxtMag;
    if (o.lastKSNoiseRatio != 0) lastKSNoiseRatio = o.lastKSNoiseRatio;
    if (o.lastModSwitchRatio != 0) lastModSwitchRatio = o.lastModSwitchRatio;
    return *this;
  }
  Ctxt(const Ctxt&) = default;
  bool isCKKS() const { return pubKey.ckks; }
  bool isEmpty() const { return parts.empty(); }
  double logOfPrimeSet() const { return pubKey.logOfProduct(primeSet); }
  long getPartIndexByHandle(const SKHandle& h) const { for (size_t i = 0; i < parts.size(); i++) if (parts[i].skHandle == h) return (long)i; return -1; }
  bool inCanonicalForm(long keyID = 0) const {
    if (parts.size() > 2) return false;
    if (parts.size() > 0 && !parts[0].skHandle.isOne()) return false;
    if (parts.size() > 1 && !parts[1].skHandle.isBase(keyID)) return false;
    return true;
  }
  XD totalNoiseBound() const { return isCKKS() ? ptxtMag * ratFactor + noiseBound : noiseBound; }   // include/helib/Ctxt.h:1358-1364
  // src/Ctxt.cpp:116-127; polyNormBnd = 1 for power-of-two m (PAlgebra::getPolyNormBnd), else supplied by the caller
  bool isCorrect(double polyNormBnd = 1.0) const { return (totalNoiseBound() * XD(polyNormBnd)).ln() <= std::log(0.48) + logOfPrimeSet(); }
  bool verifyPrimeSet() const {   // src/Ctxt.cpp:177-186
    IndexSet s = primeSet & context.getSpecialPrimes();
    if (!empty(s) && s != context.getSpecialPrimes()) return false;
    return (primeSet & context.getCtxtPrimes()).isInterval();
  }
  XD modSwitchAddedNoiseBound() const {   // src/Ctxt.cpp:2560-2582
    XD added(0.0);
    for (auto& part : parts) {
      if (part.skHandle.isOne()) added = added + XD(1.0);
      else added = added + XD::exp(part.skHandle.powerOfS * std::log(pubKey.skBound));
    }
    return added * XD(pubKey.noiseBoundForUniform(double(ptxtSpace) / 2.0, context.getPhiM()));
  }

  void modUpToSet(const IndexSet& s) {   // src/Ctxt.cpp:346-371
    IndexSet setDiff = s / primeSet;
    if (empty(setDiff)) return;
    double f = 0;
    for (auto& part : parts) f = part.dcrt.addPrimesAndScale(setDiff);
    noiseBound = noiseBound * XD::exp(f);
    ratFactor = ratFactor * XD::exp(f);
    primeSet.insert(setDiff);
    if (!verifyPrimeSet()) throw LogicError("primeSet is no longer valid");
  }
  void modDownToSet(const IndexSet& s) {   // src/Ctxt.cpp:393-562 ("real mod switching" branch)
    HB_TIMER_START(context.handle());
    IndexSet intersection = primeSet & s;
    if (empty(intersection)) throw RuntimeError("modDownToSet called with a disjoint set");
    IndexSet setDiff = primeSet / intersection;
    if (empty(setDiff)) return;
    XD addedNoiseBound = modSwitchAddedNoiseBound();
    XD addedNoise(0.0);
    for (auto& part : parts) {
      const double norm = part.dcrt.scaleDownToSetNorm(intersection, ptxtSpace);   // ||delta/P||_canon, computed on the device
      if (part.skHandle.isOne()) addedNoise = addedNoise + XD(norm);
      else addedNoise = addedNoise + XD(norm) * XD::exp(part.skHandle.powerOfS * std::log(pubKey.skBound));
    }
    XD f = XD::exp(pubKey.logOfProduct(setDiff));
    ratFactor = ratFactor / f;
    noiseBound = noiseBound / f;
    noiseBound = noiseBound + addedNoise;
    lastModSwitchRatio = (addedNoise / addedNoiseBound).to_double();   // the reference's "mod-switch-added-noise" statistic
    HB_STATS_UPDATE("mod-switch-added-noise", lastModSwitchRatio);      // src/Ctxt.cpp:537
    primeSet.remove(setDiff);
    if (!verifyPrimeSet()) throw LogicError("primeSet is no longer valid");
  }
  void bringToSet(const IndexSet& s) {   // src/Ctxt.cpp:373-389
    if (empty(s)) { IndexSet tmp(context.getCtxtPrimes().first()); modUpToSet(tmp); modDownToSet(tmp); }
    else { modUpToSet(s); modDownToSet(s); }
  }
  void dropSmallAndSpecialPrimes() {   // src/Ctxt.cpp:589-662
    if (primeSet.disjointFrom(context.getSmallPrimes())) { modDownToSet(context.getCtxtPrimes()); return; }
    IndexSet target = primeSet & context.getCtxtPrimes();
    IndexSet dropping = primeSet / target;
    double log_dropping = pubKey.logOfProduct(dropping);
    double log_modswitch_noise = modSwitchAddedNoiseBound().ln();
    double log_noise = noiseBound.m <= 0 ? -DBL_MAX : noiseBound.ln();
    double log_compensation = 0;
    log_modswitch_noise += 3 * std::log(2.0);
    if (log_noise - log_dropping + log_compensation < log_modswitch_noise) {
      IndexSet candidates = context.getCtxtPrimes() / target;
      for (long i : candidates) {
        target.insert(i);
        log_compensation += std::log((double)context.ithPrime(i));
        if (log_noise - log_dropping + log_compensation >= log_modswitch_noise) break;
      }
    }
    bringToSet(target);
  }
  void relin_CKKS_adjust() {   // src/Ctxt.cpp:664-716
    if (!isCKKS()) return;
    long phim = context.getPhiM();
    double h = pubKey.hwt == 0 ? phim / 2.0 : (double)pubKey.hwt;
    double log_phim = std::max(1.0, std::log((double)phim));
    double beta = pubKey.scale * std::sqrt(phim * log_phim * h / 12.0);
    double gamma = beta * 8;
    if (XD(gamma) > noiseBound) {
      long xf = (long)std::ceil(gamma / noiseBound.to_double());
      for (auto& part : parts) part.dcrt *= xf;
      noiseBound = noiseBound * XD((double)xf);
      ratFactor = ratFactor * XD((double)xf);
    }
  }
  void addPart(const DoubleCRT& part, const SKHandle& handle) {   // src/Ctxt.cpp:851-893 (matchPrimeSet, non-negative)
    if (!(primeSet <= part.getIndexSet())) throw RuntimeError("Ctxt::addPart: ctxt has primes not in part");
    long j = getPartIndexByHandle(handle);
    if (j >= 0) parts[j].dcrt.Add(part, /*matchIndexSets=*/false);
    else {
      parts.emplace_back(part, handle);
      if (part.getIndexSet() != primeSet) parts.back().dcrt.removePrimes(part.getIndexSet() / primeSet);
    }
  }
  // src/Ctxt.cpp:191-230 -- the two products per digit and their accumulation run as ONE engine launch
  void keySwitchDigits(const KeySwitch& W, std::vector<DoubleCRT>& digits) {
    HB_NTIMER_START(KS_loop, context.handle());   // src/Ctxt.cpp:205
    long j0 = getPartIndexByHandle(SKHandle()), j1 = getPartIndexByHandle(SKHandle(1, 1, W.toKeyID));
    if (j0 < 0) { parts.emplace_back(DoubleCRT(context, primeSet), SKHandle()); j0 = (long)parts.size() - 1; }
    if (j1 < 0) { parts.emplace_back(DoubleCRT(context, primeSet), SKHandle(1, 1, W.toKeyID)); j1 = (long)parts.size() - 1; }
    std::vector<hb_poly*> dg, ea, eb;
    for (size_t i = 0; i < digits.size(); i++) { dg.push_back(digits[i].handle()); ea.push_back(W.a[i].handle()); eb.push_back(W.b[i].handle()); }
    hb_poly* o0[1] = {parts[j0].dcrt.handle()}; hb_poly* o1[1] = {parts[j1].dcrt.handle()};
    auto idx = primeSet.vec();
    check(hb_keyswitch_digits(dg.data(), (int)digits.size(), (int)digits.size(), 1, idx.data(), (int)idx.size(), ea.data(), eb.data(), o0, o1));
  }
  void keySwitchPart(const CtxtPart& p, const KeySwitch& W) {   // src/Ctxt.cpp:805-842
    HB_TIMER_START(context.handle());
    if (!context.getSpecialPrimes().disjointFrom(p.dcrt.getIndexSet())) throw LogicError("Special primes and CtxtPart's index set have non-empty intersection");
    if (p.skHandle.isOne() || p.skHandle.isBase(W.toKeyID)) {
      CtxtPart pp = p;
      pp.dcrt.addPrimesAndScale(context.getSpecialPrimes());
      addPart(pp.dcrt, pp.skHandle);
      return;
    }
    if (!(W.fromKey == p.skHandle)) throw LogicError("Secret key handles do not match");
    std::vector<DoubleCRT> polyDigits;
    XD addedNoise(0.0);
    for (double ln : p.dcrt.breakIntoDigitsLogNorms(polyDigits)) addedNoise = addedNoise + XD::exp(ln);   // sum of ||E_i||, computed on the device
    addedNoise = addedNoise * W.noiseBound;
    keySwitchDigits(W, polyDigits);
    lastKSNoiseRatio = (addedNoise / noiseBound).to_double();   // "KS-noise-ratio"
    HB_STATS_UPDATE("KS-noise-ratio", lastKSNoiseRatio);         // src/Ctxt.cpp:835
    noiseBound = noiseBound + addedNoise;
  }
  void reLinearize(long keyID = 0) {   // src/Ctxt.cpp:720-786
    HB_TIMER_START(context.handle());
    if (isEmpty() || inCanonicalForm(keyID)) return;
    dropSmallAndSpecialPrimes();
    relin_CKKS_adjust();
    long g = ptxtSpace;
    double logProd = pubKey.logOfProduct(context.getSpecialPrimes());
    Ctxt tmp(pubKey, ptxtSpace);
    tmp.intFactor = intFactor; tmp.ptxtMag = ptxtMag;
    tmp.noiseBound = noiseBound * XD::exp(logProd);
    tmp.primeSet = primeSet | context.getSpecialPrimes();
    tmp.ratFactor = ratFactor * XD::exp(logProd);
    for (CtxtPart& part : parts) {
      if (part.skHandle.isOne() || part.skHandle.isBase(keyID)) {
        part.dcrt.addPrimesAndScale(context.getSpecialPrimes());
        tmp.addPart(part.dcrt, part.skHandle);
        continue;
      }
      const KeySwitch* W = pubKey.getKeySWmatrix(part.skHandle, keyID);
      if (!W) throw LogicError("No key-switching matrix exists");
      if (g > 1) { tmp.reducePtxtSpace(W->ptxtSpace); g = tmp.ptxtSpace; }   // g == 1 for CKKS (src/Ctxt.cpp:771-775)
      tmp.keySwitchPart(part, *W);
    }
    *this = tmp;
  }
  void tensorProduct(const Ctxt& c1, const Ctxt& c2) {   // src/Ctxt.cpp:1563-1608
    HB_TIMER_START(context.handle());
    parts.clear();
    primeSet = c1.primeSet;
    long ptxtSp = c1.ptxtSpace;
    if (ptxtSp > 2) {
      unsigned long q = 1;
      for (long i : c1.primeSet) q = (unsigned long)(((unsigned __int128)q * (unsigned long)(context.ithPrime(i) % ptxtSp)) % (unsigned long)ptxtSp);
      intFactor = (long)(((unsigned __int128)c1.intFactor * c2.intFactor) % ptxtSp);
      intFactor = (long)(((unsigned __int128)intFactor * q) % ptxtSp);
    }
    for (auto& p1 : c1.parts)
      for (auto& p2 : c2.parts) {
        CtxtPart tmpPart = p2;
        if (!tmpPart.skHandle.mul(p1.skHandle, tmpPart.skHandle)) throw LogicError("Ctxt::tensorProduct: cannot multiply secret-key handles");
        tmpPart.dcrt *= p1.dcrt;
        long k = getPartIndexByHandle(tmpPart.skHandle);
        if (k >= 0) parts[k].dcrt += tmpPart.dcrt;
        else parts.push_back(tmpPart);
      }
    if (isCKKS()) {
      noiseBound = c1.noiseBound * c2.ptxtMag * c2.ratFactor + c2.noiseBound * c1.ptxtMag * c1.ratFactor + c1.noiseBound * c2.noiseBound;
      ratFactor = c1.ratFactor * c2.ratFactor;
      ptxtMag = c1.ptxtMag * c2.ptxtMag;
    } else noiseBound = c1.noiseBound * c2.noiseBound;
  }
  static void computeIntervalForMul(double& lo, double& hi, const Ctxt& c1, const Ctxt& c2) {   // src/Ctxt.cpp:1610-1656
    const double slack = 4 * std::log(2.0);
    double cap1 = c1.logOfPrimeSet() - std::max(c1.noiseBound, XD(1.0)).ln();
    double cap2 = c2.logOfPrimeSet() - std::max(c2.noiseBound, XD(1.0)).ln();
    double adn1 = c1.modSwitchAddedNoiseBound().ln(), adn2 = c2.modSwitchAddedNoiseBound().ln();
    if (c1.isCKKS()) { lo = std::max(cap1 + adn1, cap2 + adn2) + safety; hi = lo + slack; }
    else { hi = std::min(cap1 + adn1, cap2 + adn2) - safety; lo = hi - slack; }
  }
  void multLowLvl(const Ctxt& other_orig) {   // src/Ctxt.cpp:1681-1753 (non-destructive, distinct operands)
    HB_TIMER_START(context.handle());
    if (isEmpty()) return;
    if (other_orig.isEmpty()) { *this = other_orig; return; }
    if (isCKKS() != other_orig.isCKKS()) throw LogicError("Scheme mismatch");
    if (&context != &other_orig.context) throw LogicError("Context mismatch");
    if (&pubKey != &other_orig.pubKey) throw LogicError("Public key mismatch");
    if (isCKKS() && (ptxtSpace != 1 || other_orig.ptxtSpace != 1)) throw LogicError("Plaintext spaces incompatible");
    Ctxt other = other_orig;
    if (!isCKKS()) {   // equalize plaintext spaces (src/Ctxt.cpp:1717-1725); reducePtxtSpace also reduces intFactor
      long g = std::gcd(ptxtSpace, other.ptxtSpace);
      if (g <= 1) throw LogicError("Plaintext spaces are co-prime");
      reducePtxtSpace(g);
      other.reducePtxtSpace(g);
    }
    double lo, hi;
    computeIntervalForMul(lo, hi, *this, other);
    auto f1 = primeSet.vec(), f2 = other.primeSet.vec();
    std::vector<int32_t> out(context.numPrimes()); int nout = 0;
    check(hb_chain_set4size(context.chain(), lo, hi, f1.data(), (int)f1.size(), f2.data(), (int)f2.size(), isCKKS() ? 1 : 0, out.data(), &nout));
    IndexSet common(out.begin(), out.begin() + nout);
    lastCommonPrimeSet = common; lastLo = lo; lastHi = hi;
    bringToSet(common);
    other.bringToSet(common);
    Ctxt tmp(pubKey, ptxtSpace);
    tmp.tensorProduct(*this, other);
    *this = tmp;
  }
  // ---- linear operations (addConstant / multByConstant: BGV branches only)
  void negate() { for (auto& part : parts) part.dcrt.Negate(); }   // src/Ctxt.cpp:1190-1194
  void reducePtxtSpace(long newPtxtSpace) {   // src/Ctxt.cpp:576-584
    long g = std::gcd(ptxtSpace, newPtxtSpace);
    if (g <= 1) throw LogicError("New and old plaintext spaces are coprime");
    ptxtSpace = g; intFactor %= g;
  }
  static long balRem(long a, long q) { return a > q / 2 ? a - q : a; }   // include/helib/NumbTh.h:140-146
  void mulIntFactor(long e) {   // src/Ctxt.cpp:331-340
    if (e == 1) return;
    intFactor = (long)(((unsigned __int128)(unsigned long)intFactor * (unsigned long)e) % (unsigned long)ptxtSpace);
    long bal_e = balRem(e, ptxtSpace);
    for (auto& part : parts) part.dcrt *= bal_e;
    noiseBound = noiseBound * XD((double)std::labs(bal_e));
  }
  void addCtxt(const Ctxt& other, bool negative = false) {   // src/Ctxt.cpp:1406-1556
    if (&context != &other.context) throw LogicError("Context mismatch");
    if (&pubKey != &other.pubKey) throw LogicError("Public key mismatch");
    if (other.isEmpty()) return;
    if (isEmpty()) { *this = other; if (negative) negate(); return; }
    if (isCKKS()) { if (ptxtSpace != 1 || other.ptxtSpace != 1) throw LogicError("Plaintext spaces incompatible"); }
    else reducePtxtSpace(other.ptxtSpace);
    Ctxt tmp(pubKey, other.ptxtSpace);
    const Ctxt* other_pt = &other;
    if (ptxtSpace != other_pt->ptxtSpace) { tmp = other; tmp.reducePtxtSpace(ptxtSpace); other_pt = &tmp; }
    IndexSet s = other_pt->primeSet / primeSet;
    if (!empty(s)) modUpToSet(s);
    s = primeSet / other_pt->primeSet;
    if (!empty(s)) { if (other_pt != &tmp) { tmp = other; other_pt = &tmp; } tmp.modUpToSet(s); }
    if (isCKKS()) { if (other_pt != &tmp) { tmp = other; other_pt = &tmp; } equalizeRationalFactors(*this, tmp, context.getR()); }
    long e1 = 1, e2 = 1;
    if (!isCKKS() && intFactor != other_pt->intFactor) {   // harmonise: e1*f1 == e2*f2 (mod ptxtSpace) with the least noise growth (:1475-1527)
      const long f1 = intFactor, f2 = other_pt->intFactor;
      const long ratio = (long)(((unsigned __int128)(unsigned long)f2 * (unsigned long)invMod(f1, ptxtSpace)) % (unsigned long)ptxtSpace);
      auto noiseNorm = [&](long a, long b) { return noiseBound * XD((double)std::labs(balRem(a, ptxtSpace))) + other_pt->noiseBound * XD((double)std::labs(balRem(b, ptxtSpace))); };
      auto mc = [&](long a) { a %= ptxtSpace; return a < 0 ? a + ptxtSpace : a; };
      long r0 = ptxtSpace, t0 = 0, r1 = ratio, t1 = 1;
      long e1_best = r1, e2_best = t1;
      XD noise_best = noiseNorm(e1_best, e2_best);
      const long pp = context.getP();
      while (r1 != 0) {
        long q = r0 / r1, r2 = r0 % r1, t2 = t0 - t1 * q;
        r0 = r1; r1 = r2; t0 = t1; t1 = t2;
        long e1_try = mc(r1), e2_try = mc(t1);
        if (e1_try % pp != 0) { XD n = noiseNorm(e1_try, e2_try); if (n < noise_best) { e1_best = e1_try; e2_best = e2_try; noise_best = n; } }
      }
      e1 = e1_best; e2 = e2_best;
    }
    if (e2 != 1) { if (other_pt != &tmp) { tmp = other; other_pt = &tmp; } tmp.mulIntFactor(e2); }
    if (e1 != 1) mulIntFactor(e1);
    for (const CtxtPart& part : other_pt->parts) {
      long j = getPartIndexByHandle(part.skHandle);
      if (j >= 0) { if (negative) parts[j].dcrt -= part.dcrt; else parts[j].dcrt += part.dcrt; }
      else { parts.push_back(part); if (negative) parts.back().dcrt.Negate(); }
    }
    ptxtMag = ptxtMag + other_pt->ptxtMag;
    noiseBound = noiseBound + other_pt->noiseBound;
  }
  // src/Ctxt.cpp:1199-1351 ("NEW VERSION"): bring two CKKS ciphertexts to a common scaling factor by multiplying them by
  // the numerator / denominator of a continued-fraction approximation of the ratio, stopping as soon as the
  // discretisation error is within sqrt(2) of the unavoidable one.  r = Context::getPrecision().
  static void equalizeRationalFactors(Ctxt& c1, Ctxt& c2, long r) {
    Ctxt& big = (c1.ratFactor > c2.ratFactor) ? c1 : c2;
    Ctxt& small = (c1.ratFactor > c2.ratFactor) ? c2 : c1;
    const XD x = big.ratFactor / small.ratFactor;
    const double denomBound = std::ldexp(1.0, (int)r + 1);
    const double epsilon = 0.125 / denomBound;
    auto calc_err = [](const XD& f, const XD& m1, const XD& f1, const XD& e1, const XD& m2, const XD& f2, const XD& e2) {
      return m1 * (f1 / f - XD(1.0)).abs() + m2 * (f2 / f - XD(1.0)).abs() + (e1 + e2) / f;
    };
    auto floorXD = [](const XD& v) { uint64_t mant; long sh; v.floorParts(mant, sh); return XD::make((double)mant, sh); };
    XD xi = x - floorXD(x + XD(epsilon));
    double prevDenom = 0, denom = 1;
    XD numer = floorXD(XD(denom) * x + XD(0.5));
    const XD m1 = big.ptxtMag, of1 = big.ratFactor, oe1 = big.noiseBound;
    const XD m2 = small.ptxtMag, of2 = small.ratFactor, oe2 = small.noiseBound;
    const XD target_error = oe1 / of1 + oe2 / of2;
    XD f, fe1, fe2;
    for (;;) {
      const XD xdenom(denom);
      const XD f1 = of1 * xdenom, e1 = oe1 * xdenom, f2 = of2 * numer, e2 = oe2 * numer;
      const XD err1 = calc_err(f1, m1, f1, e1, m2, f2, e2), err2 = calc_err(f2, m1, f1, e1, m2, f2, e2);
      XD err;
      if (err1 < err2) { f = f1; fe1 = e1; fe2 = e2 + m2 * (f2 - f1).abs(); err = err1; }
      else { f = f2; fe1 = e1 + m1 * (f2 - f1).abs(); fe2 = e2; err = err2; }
      if (err < XD(std::sqrt(2.0)) * target_error) break;
      if (xi.m <= 0) break;
      xi = XD(1.0) / xi;
      const XD ai = floorXD(xi + XD(epsilon));
      xi = xi - ai;
      const double tmpDenom = denom * ai.to_double() + prevDenom;
      if (tmpDenom > denomBound) break;
      prevDenom = denom; denom = tmpDenom;
      numer = floorXD(XD(denom) * x + XD(0.5));
    }
    if (denom != 1) for (auto& part : big.parts) part.dcrt *= (long)denom;
    big.ratFactor = f; big.noiseBound = fe1;
    uint64_t nm; long nsh; numer.floorParts(nm, nsh);
    if (!(nm == 1 && nsh == 0)) for (auto& part : small.parts) part.dcrt.mulByPow2Scaled(nm, nsh);
    small.ratFactor = f; small.noiseBound = fe2;
  }
  Ctxt& operator+=(const Ctxt& o) { addCtxt(o); return *this; }
  Ctxt& operator-=(const Ctxt& o) { addCtxt(o, true); return *this; }
  // src/Ctxt.cpp:896-935 (BGV): the constant is scaled by intFactor*Q mod p so that it decrypts unscaled
  void addConstant(const DoubleCRT& dcrt, double size = -1.0) {
    if (isCKKS()) throw LogicError("Ctxt::addConstant: use addConstantCKKS (explicit size and factor)");
    if (size < 0.0) size = pubKey.noiseBoundForMod(ptxtSpace, context.getPhiM());
    long f = 1;
    if (ptxtSpace > 2) {
      unsigned long q = 1;
      for (long i : primeSet) q = (unsigned long)(((unsigned __int128)q * (unsigned long)(context.ithPrime(i) % ptxtSpace)) % (unsigned long)ptxtSpace);
      f = balRem((long)(((unsigned __int128)(unsigned long)intFactor * q) % (unsigned long)ptxtSpace), ptxtSpace);
    }
    noiseBound = noiseBound + XD(size * (double)std::labs(f));
    if (f == 1) addPart(dcrt, SKHandle(0, 1, 0));
    else { DoubleCRT tmp = dcrt; tmp *= f; addPart(tmp, SKHandle(0, 1, 0)); }
  }
  // multByConstantCKKS (src/Ctxt.cpp:1905-1938): dcrt encodes slots of magnitude <= size at scaling factor `factor`;
  // roundingErr = the encoding's rounding error.  The reference's defaults come from EncryptedArrayCx (the encoding layer,
  // out of scope here), so the three values are explicit arguments.
  void multByConstantCKKS(const DoubleCRT& dcrt, const XD& size, const XD& factor, double roundingErr) {
    if (isEmpty()) return;
    if (!isCKKS()) throw LogicError("multByConstantCKKS on a BGV ciphertext");
    noiseBound = noiseBound * factor * size + XD(roundingErr) * ratFactor * ptxtMag + noiseBound * XD(roundingErr);   // must come first
    ptxtMag = ptxtMag * size;
    ratFactor = ratFactor * factor;
    for (auto& part : parts) part.dcrt.Mul(dcrt, /*matchIndexSets=*/false);
  }
  // addConstantCKKS (src/Ctxt.cpp:941-1052): the constant (scaling factor `factor`) is multiplied by round(ratFactor/factor)
  // so that it matches the ciphertext's factor.  The reference adds primes (addSomePrimes) when that rounding alone would
  // cost more than 2^-precision of accuracy; the mirror reports that case instead.
  void addConstantCKKS(const DoubleCRT& dcrt, const XD& size_in, const XD& factor) {
    if (!isCKKS()) throw LogicError("addConstantCKKS on a BGV ciphertext");
    const XD size = size_in.m <= 0 ? XD(1.0) : size_in;
    if (factor.m <= 0) throw InvalidArgument("addConstantCKKS: the scaling factor of the constant must be given");
    XD ratio = ratFactor / factor + XD(0.5);
    uint64_t mant; long sh; ratio.floorParts(mant, sh);
    const XD r = XD::make((double)mant, sh);
    const double inaccuracy = std::fabs((r * factor / ratFactor).to_double() - 1.0);
    if (inaccuracy * std::ldexp(1.0, (int)context.getR()) > 1.0) throw LogicError("addConstantCKKS: scaling factors too far apart (the reference calls addSomePrimes here)");
    ptxtMag = ptxtMag + size;
    noiseBound = noiseBound + XD(0.5);
    IndexSet delta = primeSet / dcrt.getIndexSet();
    if (mant == 1 && sh == 0 && empty(delta)) { addPart(dcrt, SKHandle(0, 1, 0)); return; }
    DoubleCRT tmp = dcrt;
    if (!empty(delta)) tmp.addPrimes(delta);
    if (!(mant == 1 && sh == 0)) tmp.mulByPow2Scaled(mant, sh);
    addPart(tmp, SKHandle(0, 1, 0));
  }
  // src/Ctxt.cpp:1832-1856 (BGV)
  void multByConstant(const DoubleCRT& dcrt, double size = -1.0) {
    if (isEmpty()) return;
    if (isCKKS()) throw LogicError("Ctxt::multByConstant: use multByConstantCKKS (explicit size, factor and rounding error)");
    if (size < 0.0) size = pubKey.noiseBoundForMod(ptxtSpace, context.getPhiM());
    for (auto& part : parts) part.dcrt.Mul(dcrt, /*matchIndexSets=*/false);
    noiseBound = noiseBound * XD(size);
  }
  long getKeyID() const { for (auto& part : parts) if (!part.skHandle.isOne()) return part.skHandle.secretKeyID; return 0; }   // src/Ctxt.cpp:2550-2557
  void cleanUp() {   // src/Ctxt.cpp:788-797
    reLinearize();
    if (!primeSet.disjointFrom(context.getSpecialPrimes()) || !primeSet.disjointFrom(context.getSmallPrimes())) dropSmallAndSpecialPrimes();
  }
  static long invMod(long a, long m) {
    long b = m, x0 = 1, x1 = 0; a %= m; if (a < 0) a += m;
    while (b) { long q = a / b, t = a - q * b; a = b; b = t; t = x0 - q * x1; x0 = x1; x1 = t; }
    if (a != 1) throw InvalidArgument("InvMod: not invertible");
    x0 %= m; return x0 < 0 ? x0 + m : x0;
  }
  void automorph(long k) {   // src/Ctxt.cpp:2437-2457: F(X) -> F(X^k), no change in the noise bound
    if (isEmpty()) return;
    const long m = context.getM();
    if (k <= 0 || k >= m || std::gcd(k, m) != 1) throw LogicError("k must be in Zm*");
    for (auto& part : parts) {
      part.dcrt.automorph(k);
      if (!part.skHandle.isOne()) part.skHandle.powerOfX = (long)(((unsigned __int128)(unsigned long)part.skHandle.powerOfX * (unsigned long)k) % (unsigned long)m);
    }
  }
  void complexConj() { automorph(context.getM() - 1); }   // src/Ctxt.cpp:2517-2523
  void smartAutomorph(long k) {   // src/Ctxt.cpp:2462-2515: automorphism then re-linearisation, in the steps the key-switching map allows
    const long m = context.getM();
    k %= m; if (k < 0) k += m;
    if (isEmpty() || k == 1) return;
    if (std::gcd(k, m) != 1) throw LogicError("k must be in Zm*");
    const long keyID = getKeyID();
    if (!pubKey.isReachable(k, keyID)) throw LogicError("no key-switching matrices for k=" + std::to_string(k) + ", keyID=" + std::to_string(keyID));
    if (!inCanonicalForm(keyID)) { reLinearize(keyID); if (!inCanonicalForm(keyID)) throw LogicError("Re-linearization failed: not in canonical form"); }
    while (k != 1) {
      const KeySwitch* matrix = pubKey.getNextKSWmatrix(k, keyID);
      const long amt = matrix->fromKey.powerOfX;
      automorph(amt);
      reLinearize(keyID);
      k = (long)(((unsigned __int128)(unsigned long)k * (unsigned long)invMod(amt, m)) % (unsigned long)m);
    }
  }
  // Phi_m(X) over Z (Cyclotomic(m)): Phi_1 = X - 1; Phi_{np}(X) = Phi_n(X^p) / Phi_n(X) for p not dividing n, Phi_n(X^p) otherwise
  static std::vector<long> cyclotomic(long m) {
    std::vector<long> phi{-1, 1};
    long n = 1;
    for (long p = 2; m > 1; p++) {
      bool first = true;
      while (m % p == 0) {
        m /= p;
        std::vector<long> up((phi.size() - 1) * p + 1, 0);
        for (size_t i = 0; i < phi.size(); i++) up[i * p] = phi[i];
        if (first && n % p != 0) {   // exact division of up by the monic phi
          std::vector<long> quo(up.size() - phi.size() + 1, 0);
          for (long i = (long)up.size() - 1; i >= (long)phi.size() - 1; i--) {
            const long c = up[i]; quo[i - (phi.size() - 1)] = c;
            if (c) for (size_t j = 0; j < phi.size(); j++) up[i - (phi.size() - 1) + j] -= c * phi[j];
          }
          phi = quo;
        } else phi = up;
        n *= p; first = false;
      }
    }
    return phi;
  }
  // Ctxt::rawModSwitch (src/Ctxt.cpp:2949-3046): mod-switch to an external modulus q for bootstrapping.  The scaling and
  // rounding run on the device in the powerful basis (hb_raw_mod_switch); the small result is brought back to the
  // polynomial basis here (PowerfulDCRT::powerfulToZZX, src/powerful.cpp:355-391).  Returns the scaled noise estimate.
  double rawModSwitch(std::vector<std::vector<long>>& zzParts, long q) const {
    if (q <= 1) throw InvalidArgument("q must be greater than 1");
    if (ptxtSpace <= 1) throw LogicError("Plaintext space must be greater than 1 for mod switching");
    if (std::gcd(q, ptxtSpace) != 1) throw LogicError("New modulus and current plaintext space must be co-prime");
    const long phim = context.getPhiM(), m = context.getM();
    int32_t nf = 0; std::vector<int32_t> toPoly((size_t)phim);
    check(hb_ctx_powerful_info(context.handle(), &nf, nullptr, toPoly.data()));
    std::vector<long> phimx;
    if (nf > 1) phimx = cyclotomic(m);
    zzParts.assign(parts.size(), std::vector<long>());
    auto idx = primeSet.vec();
    for (size_t i = 0; i < parts.size(); i++) {
      std::vector<int64_t> pw((size_t)phim);
      check(hb_raw_mod_switch(parts[i].dcrt.handle(), idx.data(), (int)idx.size(), (uint64_t)q, (uint64_t)ptxtSpace, pw.data()));
      if (nf <= 1) { zzParts[i].assign(pw.begin(), pw.end()); continue; }
      std::vector<long> tmp((size_t)m, 0);
      for (long k = 0; k < phim; k++) tmp[toPoly[k]] = pw[k];
      for (long k = m - 1; k >= phim; k--) {   // rem(tmp, Phi_m): Phi_m is monic of degree phi(m)
        const long c = tmp[k];
        if (c) for (long j = 0; j <= phim; j++) tmp[k - phim + j] -= c * phimx[j];
      }
      zzParts[i].assign(tmp.begin(), tmp.begin() + phim);
    }
    return (noiseBound * XD::exp(std::log((double)q) - logOfPrimeSet())).to_double();
  }
  // ---- binary wire format, Ctxt::writeTo / read (src/Ctxt.cpp:2584-2641, src/binio.h:91-137, src/binio.cpp:75-178):
  //   24-byte SerializeHeader<Ctxt> ("|HE[", format version 0.0.1.0, library version, struct id 20, 7 reserved, "]HE|"),
  //   "|CX[", ptxtSpace, intFactor, ptxtMag, ratFactor, noiseBound (xdouble = raw double mantissa + int64 exponent),
  //   primeSet, vector<CtxtPart> (count, then DoubleCRT + SKHandle each), "]CX|".  All integers little-endian int64.
  // NTL's xdouble is x * (2^114)^e with 2^-57 <= |x| <= 2^57 (NTL 11.4.3 xdouble.h, NTL_XD_BOUND); NTL is not in the
  // reference tree, so this split is restated from its documentation: parity unpinned for those 16-byte fields.
  static void writeXD(std::ostream& str, const XD& v) {
    double x = 0; int64_t e = 0;
    if (v.m != 0) {
      long E = v.e;                      // v = m * 2^E, 0.5 <= |m| < 1
      e = E > 57 ? (E - 57 + 113) / 114 : (E < -56 ? -((-56 - E + 113) / 114) : 0);
      x = std::ldexp(v.m, (int)(E - 114 * e));
    }
    str.write(reinterpret_cast<const char*>(&x), 8); str.write(reinterpret_cast<const char*>(&e), 8);
  }
  static XD readXD(std::istream& str) {
    double x = 0; int64_t e = 0;
    str.read(reinterpret_cast<char*>(&x), 8); str.read(reinterpret_cast<char*>(&e), 8);
    return XD::make(x, 114 * (long)e);
  }
  static void writeInt(std::ostream& str, int64_t v) { str.write(reinterpret_cast<const char*>(&v), 8); }
  static int64_t readInt(std::istream& str) { int64_t v = 0; str.read(reinterpret_cast<char*>(&v), 8); return v; }
  void writeTo(std::ostream& str) const {
    const char header[24] = {'|', 'H', 'E', '[', 0, 0, 1, 0, 2, 2, 0, 0, 20, 0, 0, 0, 0, 0, 0, 0, ']', 'H', 'E', '|'};
    str.write(header, 24);
    str.write("|CX[", 4);
    writeInt(str, ptxtSpace); writeInt(str, intFactor);
    writeXD(str, ptxtMag); writeXD(str, ratFactor); writeXD(str, noiseBound);
    writeInt(str, primeSet.card());
    for (long i : primeSet) writeInt(str, i);
    writeInt(str, (int64_t)parts.size());
    for (const CtxtPart& part : parts) {
      part.dcrt.writeTo(str);
      writeInt(str, part.skHandle.powerOfS); writeInt(str, part.skHandle.powerOfX); writeInt(str, part.skHandle.secretKeyID);
    }
    str.write("]CX|", 4);
  }
  void read(std::istream& str) {
    char header[24];
    str.read(header, 24);
    if (!str || std::memcmp(header, "|HE[", 4) != 0 || std::memcmp(header + 20, "]HE|", 4) != 0) throw RuntimeError("Eye catchers for header mismatch");
    const char ver[4] = {0, 0, 1, 0};
    if (std::memcmp(header + 4, ver, 4) != 0) throw RuntimeError("Header: version not supported");
    char eye[4];
    str.read(eye, 4);
    if (std::memcmp(eye, "|CX[", 4) != 0) throw RuntimeError("Could not find pre-ciphertext eye catcher");
    ptxtSpace = readInt(str); intFactor = readInt(str);
    ptxtMag = readXD(str); ratFactor = readXD(str); noiseBound = readXD(str);
    const int64_t card = readInt(str);
    if (!str || card < 0 || card > context.numPrimes()) throw RuntimeError("Ctxt::read: bad prime set");
    primeSet = IndexSet();
    for (int64_t i = 0; i < card; i++) primeSet.insert(readInt(str));
    const int64_t np = readInt(str);
    if (!str || np < 0 || np > 64) throw RuntimeError("Ctxt::read: bad part count");
    parts.clear();
    for (int64_t i = 0; i < np; i++) {
      DoubleCRT d(context, IndexSet::emptySet());
      d.read(str);
      SKHandle h; h.powerOfS = readInt(str); h.powerOfX = readInt(str); h.secretKeyID = readInt(str);
      parts.emplace_back(d, h);
    }
    str.read(eye, 4);
    if (!str || std::memcmp(eye, "]CX|", 4) != 0) throw RuntimeError("Could not find post-ciphertext eye catcher");
  }
  void multiplyBy(const Ctxt& other) {   // src/Ctxt.cpp:1757-1774
    HB_TIMER_START(context.handle());
    if (isEmpty()) return;
    if (other.isEmpty()) { *this = other; return; }
    multLowLvl(other);
    reLinearize();
  }
  // statistics the reference records through HELIB_STATS_UPDATE (src/Ctxt.cpp:537,835)
  double lastModSwitchRatio = 0, lastKSNoiseRatio = 0, lastLo = 0, lastHi = 0;
  IndexSet lastCommonPrimeSet;
};

// BasicAutomorphPrecon (src/matmul.cpp:60-184): hoisting -- one breakIntoDigits shared by many automorphisms of the
// same ciphertext.  Each automorph(k) is ONE engine launch for power-of-two m (sigma_k applied in the load stage of
// the evaluation-key inner product, hb_automorph_keyswitch_digits); general m permutes the digits first.
class BasicAutomorphPrecon {
  Ctxt ctxt;
  XD noise;
  std::vector<DoubleCRT> polyDigits;
 public:
  double lastKSNoiseRatioHoist = 0;   // "KS-noise-ratio-hoist" (src/matmul.cpp:101)
  explicit BasicAutomorphPrecon(const Ctxt& c) : ctxt(c), noise(1.0) {
    if (ctxt.parts.size() >= 1 && !ctxt.parts[0].skHandle.isOne()) throw LogicError("Invalid ciphertext (secret key handle for part 0 is not one)");
    if (ctxt.parts.size() <= 1) return;
    ctxt.cleanUp();
    if (!ctxt.inCanonicalForm(ctxt.getKeyID())) throw LogicError("Ciphertext is not in canonical form");
    ctxt.relin_CKKS_adjust();
    XD addedNoise(0.0);
    for (double ln : ctxt.parts[1].dcrt.breakIntoDigitsLogNorms(polyDigits)) addedNoise = addedNoise + XD::exp(ln);
    XD max_ks_noise(0.0);
    for (const KeySwitch& ks : ctxt.pubKey.keySwitching) if (max_ks_noise < ks.noiseBound) max_ks_noise = ks.noiseBound;
    addedNoise = addedNoise * max_ks_noise;
    noise = ctxt.noiseBound * XD::exp(ctxt.pubKey.logOfProduct(ctxt.context.getSpecialPrimes()));
    lastKSNoiseRatioHoist = (addedNoise / noise).to_double();
    noise = noise + addedNoise;
  }
  std::shared_ptr<Ctxt> automorph(long k) const {
    if (k == 1 || ctxt.isEmpty()) return std::make_shared<Ctxt>(ctxt);
    const Context& context = ctxt.context;
    const KeyInfo& pubKey = ctxt.pubKey;
    const long m = context.getM();
    auto result = std::make_shared<Ctxt>(pubKey, ctxt.ptxtSpace);
    result->noiseBound = noise;
    result->intFactor = ctxt.intFactor;
    result->primeSet = ctxt.primeSet | context.getSpecialPrimes();
    if (ctxt.isCKKS()) {
      result->ptxtMag = ctxt.ptxtMag;
      result->ratFactor = ctxt.ratFactor * XD::exp(pubKey.logOfProduct(context.getSpecialPrimes()));
    }
    if (ctxt.parts.size() == 1) {   // only the constant part: no key switch
      DoubleCRT tmp = ctxt.parts[0].dcrt;
      tmp.automorph(k);
      tmp.addPrimesAndScale(context.getSpecialPrimes());
      result->addPart(tmp, ctxt.parts[0].skHandle);
      return result;
    }
    const long keyID = ctxt.getKeyID();
    if (!pubKey.isReachable(k, keyID)) throw LogicError("no key-switching matrices for k=" + std::to_string(k) + ", keyID=" + std::to_string(keyID));
    const KeySwitch& W = *pubKey.getNextKSWmatrix(k, keyID);
    const long amt = W.fromKey.powerOfX;
    const bool pow2 = (m & (m - 1)) == 0;
    if (pow2) {
      DoubleCRT o0(context, result->primeSet), o1(context, result->primeSet);
      std::vector<hb_poly*> dg, ea, eb;
      for (size_t i = 0; i < polyDigits.size(); i++) { dg.push_back(polyDigits[i].handle()); ea.push_back(W.a[i].handle()); eb.push_back(W.b[i].handle()); }
      hb_poly* c0[1] = {ctxt.parts[0].dcrt.handle()}; hb_poly* p0[1] = {o0.handle()}; hb_poly* p1[1] = {o1.handle()};
      auto S = ctxt.primeSet.vec();
      check(hb_automorph_keyswitch_digits(dg.data(), (int)dg.size(), (int)dg.size(), 1, S.data(), (int)S.size(), c0, (uint64_t)amt, ea.data(), eb.data(), p0, p1));
      result->parts.emplace_back(o0, SKHandle());
      result->parts.emplace_back(o1, SKHandle(1, 1, W.toKeyID));
    } else {
      DoubleCRT tmp = ctxt.parts[0].dcrt;
      tmp.automorph(amt);
      tmp.addPrimesAndScale(context.getSpecialPrimes());
      result->addPart(tmp, ctxt.parts[0].skHandle);
      std::vector<DoubleCRT> tmpDigits = polyDigits;
      for (auto& d : tmpDigits) d.automorph(amt);
      result->keySwitchDigits(W, tmpDigits);
    }
    if ((amt - k) % m != 0) result->smartAutomorph((long)(((unsigned __int128)(unsigned long)k * (unsigned long)Ctxt::invMod(amt, m)) % (unsigned long)m));
    return result;
  }
};

// ---- the steps either side of the path ------------------------------------------------------------
// Sampling follows the reference's DISTRIBUTIONS (src/sample.cpp); its bit stream (NTL's PRG) is not restated, so
// the sampled values are an input of Encrypt below and "parity unpinned" is confined to them.
struct EncryptionSample {
  std::vector<long> r, e0, e1;
  double r_bound = 0, e0_bound = 0, e1_bound = 0;   // what sampleSmallBounded / sampleGaussianBounded return
};
// max_j |f(zeta^(2j+1))| for power-of-two m (embeddingLargestCoeff, src/norms.cpp:204-261): twist + N-point FFT on the host
inline double embeddingLargestCoeff(const std::vector<long>& f, long m) {
  const long N = m / 2;
  if (m < 4 || (m & (m - 1)) != 0) throw LogicError("embeddingLargestCoeff: host version is for power-of-two m");
  std::vector<std::complex<double>> z((size_t)N);
  const double pi = 3.14159265358979323846;
  for (long k = 0; k < N; k++) z[k] = (k < (long)f.size() ? double(f[k]) : 0.0) * std::polar(1.0, pi * double(k) / double(N));
  for (long i = 1, j = 0; i < N; i++) { long bit = N >> 1; for (; j & bit; bit >>= 1) j ^= bit; j ^= bit; if (i < j) std::swap(z[i], z[j]); }
  for (long len = 2; len <= N; len <<= 1) {
    const std::complex<double> wl = std::polar(1.0, 2 * pi / double(len));
    for (long i = 0; i < N; i += len) { std::complex<double> w(1.0, 0.0);
      for (long k = 0; k < len / 2; k++) { auto u = z[i + k], v = z[i + k + len / 2] * w; z[i + k] = u + v; z[i + k + len / 2] = u - v; w *= wl; } }
  }
  double mx = 0; for (auto& c : z) mx = std::max(mx, std::abs(c));
  return mx;
}
template <class Gen> void sampleSmall(std::vector<long>& poly, long n, Gen& g) {   // src/sample.cpp:67-103 (prob = 1/2)
  poly.assign((size_t)n, 0);
  for (long i = 0; i < n; i++) { unsigned u = (unsigned)(g() & 3u); poly[i] = (u & 1u) ? long(u & 2u) - 1 : 0; }
}
template <class Gen> void sampleGaussian(std::vector<long>& poly, long n, double stdev, Gen& g) {   // src/sample.cpp:140-187
  std::normal_distribution<double> d(0.0, stdev);
  poly.assign((size_t)n, 0);
  for (long i = 0; i < n; i++) poly[i] = std::lround(d(g));
}
template <class Gen> double sampleSmallBounded(std::vector<long>& poly, const Context& ctx, Gen& g) {   // src/sample.cpp:342-396
  const long phim = ctx.getPhiM();
  const double bound = std::sqrt(phim * std::log(double(phim)) / 2.0);
  double val; long count = 0;
  do { sampleSmall(poly, phim, g); val = embeddingLargestCoeff(poly, ctx.getM()); } while (++count < 1000 && val > bound);
  if (val > bound) throw RuntimeError("Error: sampleSmallBounded, after 1000 trials, still val > bound");
  return bound;
}
template <class Gen> double sampleGaussianBounded(std::vector<long>& poly, const Context& ctx, double stdev, Gen& g) {   // src/sample.cpp:459-512
  const long phim = ctx.getPhiM();
  const double bound = stdev * std::sqrt(phim * std::log(double(phim)));
  double val; long count = 0;
  do { sampleGaussian(poly, phim, stdev, g); val = embeddingLargestCoeff(poly, ctx.getM()); } while (++count < 1000 && val > bound);
  if (val > bound) throw RuntimeError("Error: sampleGaussianBounded, after 1000 trials, still val > bound");
  return bound;
}
template <class Gen> EncryptionSample drawEncryptionSample(const Context& ctx, double stdev, Gen& g) {
  EncryptionSample s;
  s.r_bound = sampleSmallBounded(s.r, ctx, g);
  s.e0_bound = sampleGaussianBounded(s.e0, ctx, stdev, g);
  s.e1_bound = sampleGaussianBounded(s.e1, ctx, stdev, g);
  return s;
}

// PubKey::Encrypt, BGV branch (src/keys.cpp:358-488): ctxt = r*pk + p*(e0,e1) + (ptxt_fixed, 0).  Three polynomials
// cross the bus as phi(m) words each; the products, sums and transforms run on the device.
// tieCoin stands for NTL::RandomBnd(2) in balanced_MulMod (src/NumbTh.cpp:876-892, even ptxtSpace only).
inline long Encrypt(Ctxt& ctxt, const Ctxt& pubEncrKey, const std::vector<long>& ptxt, long ptxtSpace,
                    const EncryptionSample& smp, const std::function<bool()>& tieCoin = [] { return false; }) {
  const Context& context = pubEncrKey.context;
  if (&ctxt.pubKey != &pubEncrKey.pubKey) throw LogicError("Public key and context public key mismatch");
  if (pubEncrKey.isCKKS()) throw LogicError("Encrypt: BGV only (CKKSencrypt is separate in the reference)");
  if (pubEncrKey.parts.size() != 2) throw LogicError("Encrypt: public encryption key must have two parts");
  if (ptxtSpace != pubEncrKey.ptxtSpace) {
    ptxtSpace = std::gcd(ptxtSpace, pubEncrKey.ptxtSpace);
    if (ptxtSpace <= 1) throw RuntimeError("Plaintext-space mismatch on encryption");
  }
  const long phim = context.getPhiM();
  if ((long)ptxt.size() > phim) throw InvalidArgument("plaintext degree >= phi(m)");
  ctxt = pubEncrKey;
  const IndexSet& S = ctxt.primeSet;
  DoubleCRT r(smp.r, context, S);
  ctxt.noiseBound = XD(smp.r_bound) * pubEncrKey.noiseBound;
  unsigned long QmodP = 1;
  for (long i : S) QmodP = (unsigned long)(((unsigned __int128)QmodP * (unsigned long)(context.ithPrime(i) % ptxtSpace)) % (unsigned long)ptxtSpace);
  for (size_t i = 0; i < ctxt.parts.size(); i++) {
    const std::vector<long>& ei = i == 0 ? smp.e0 : smp.e1;
    std::vector<long> c((size_t)phim, 0);
    for (long k = 0; k < phim && k < (long)ei.size(); k++) {
      const __int128 v = (__int128)ei[k] * ptxtSpace;
      if (v > ((__int128)1 << 61) || v < -((__int128)1 << 61)) throw InvalidArgument("Encrypt: ptxtSpace * e does not fit a word");
      c[k] = (long)v;
    }
    if (i == 0)
      for (long k = 0; k < (long)ptxt.size(); k++) {   // ptxt_fixed = balanced(ptxt * (Q mod p) mod p)  (:453-455)
        long t = ptxt[k] % ptxtSpace; if (t < 0) t += ptxtSpace;
        long f = (long)(((unsigned __int128)(unsigned long)t * QmodP) % (unsigned long)ptxtSpace);
        if (f > ptxtSpace / 2 || (ptxtSpace % 2 == 0 && f == ptxtSpace / 2 && tieCoin())) f -= ptxtSpace;
        c[k] += f;
      }
    DoubleCRT e(c, context, S);                    // p*e_i (+ ptxt_fixed)
    e.mulAdd(ctxt.parts[i].dcrt, r);               // + pk_i * r      (:416,443)
    ctxt.parts[i].dcrt = e;
    XD e_bound = XD((i == 0 ? smp.e0_bound : smp.e1_bound) * double(ptxtSpace));
    if (i == 1) e_bound = e_bound * XD(pubEncrKey.pubKey.skBound);
    ctxt.noiseBound = ctxt.noiseBound + e_bound;
  }
  ctxt.noiseBound = ctxt.noiseBound + XD(pubEncrKey.pubKey.noiseBoundForMod(ptxtSpace, phim));   // (:462,476)
  ctxt.ptxtSpace = ptxtSpace;
  ctxt.intFactor = 1;
  return ptxtSpace;
}

// SecKey::Decrypt (src/keys.cpp:1327-1400).  sKeys[id] = the secret key polynomials in DoubleCRT form.
// BGV: plaintxt in [0, ptxtSpace).  CKKS (or f_limbs != nullptr): the integer polynomial before reduction is
// returned as phi(m) x L two's-complement limbs.
inline void Decrypt(std::vector<long>& plaintxt, const Ctxt& c, const std::vector<DoubleCRT>& sKeys,
                    std::vector<uint64_t>* f_limbs = nullptr, int* L = nullptr, double polyNormBnd = 1.0) {
  if (!c.isCorrect(polyNormBnd)) throw LogicError("Decrypting with too much noise");
  const Context& context = c.context;
  const IndexSet& P = c.primeSet;
  DoubleCRT ptxt(context, P);
  for (const CtxtPart& part : c.parts) {
    if (part.skHandle.isOne()) { ptxt.Add(part.dcrt, false); continue; }
    DoubleCRT key = sKeys.at((size_t)part.skHandle.secretKeyID);
    key.addPrimes(P / key.getIndexSet());              // key.setPrimes(ptxtPrimes)  (include/helib/DoubleCRT.h:275-279)
    key.removePrimes(key.getIndexSet() / P);
    if (part.skHandle.powerOfX > 1) key.automorph(part.skHandle.powerOfX);
    if (part.skHandle.powerOfS > 1) { DoubleCRT base(key); for (long e = 1; e < part.skHandle.powerOfS; e++) key *= base; }   // Exp (src/DoubleCRT.cpp:1142-1156)
    ptxt.mulAdd(key, part.dcrt);
  }
  if (c.isCKKS() || f_limbs) {
    int l = 0; std::vector<uint64_t> limbs = ptxt.toPoly(P, false, l);
    if (f_limbs) *f_limbs = std::move(limbs);
    if (L) *L = l;
    if (c.isCKKS()) return;
  }
  const long p = c.ptxtSpace;
  long factor = 1;
  if (p > 2) {   // multiply by (intFactor * Q)^-1 mod p  (:1388-1398)
    unsigned long f = 1;
    for (long i : P) f = (unsigned long)(((unsigned __int128)f * (unsigned long)(context.ithPrime(i) % p)) % (unsigned long)p);
    long jf = c.intFactor % p; if (jf < 0) jf += p;
    f = (unsigned long)(((unsigned __int128)f * (unsigned long)jf) % (unsigned long)p);
    if (f != 1) {   // InvMod by extended Euclid
      long a = (long)f, b = p, x0 = 1, x1 = 0;
      while (b) { long q = a / b, t = a - q * b; a = b; b = t; t = x0 - q * x1; x0 = x1; x1 = t; }
      if (a != 1) throw LogicError("Decrypt: intFactor*Q not invertible mod ptxtSpace");
      factor = x0 % p; if (factor < 0) factor += p;
    }
  }
  plaintxt = ptxt.toPolyModP(P, p, factor);
}

// RLWE1 (src/keys.cpp:39-72): c0 = p*e - c1*s for a short e; returns the high-probability bound on the canonical
// embedding of the decryption.  e is drawn by sampleGaussianBounded above (the reference's distribution and rejection bound).
template <class Gen> double RLWE1(DoubleCRT& c0, const DoubleCRT& c1, const DoubleCRT& s, long p, double stdev, Gen& g) {
  if (p <= 0) throw InvalidArgument("Cannot generate RLWE instance with nonpositive p");
  const Context& context = s.getContext();
  if ((context.getM() & (context.getM() - 1)) != 0) stdev *= std::sqrt((double)context.getM());
  std::vector<long> e;
  double bound = sampleGaussianBounded(e, context, stdev, g);
  c0 = DoubleCRT(e, context, c0.getIndexSet());
  if (p > 1) { c0 *= p; bound *= p; }
  DoubleCRT tmp(c1);
  tmp.Mul(s, /*matchIndexSets=*/false);
  c0 -= tmp;
  return bound;
}
// SecKey::GenKeySWmatrix (src/keys.cpp:1159-1256): W[fromKey -> toKey] over ctxt | special primes,
// b_i = p*e_i - a_i*s + P*(prod_{j<i} Q_j)*fromKey.  fromKey = s^r(X^t) is passed in already transformed;
// drawA(a_i) fills a_i with uniform rows (the reference: a[i].randomize() under SetSeed(prgSeed); DoubleCRT::randomize here).
template <class Gen, class DrawA>
KeySwitch genKeySWmatrix(const Context& context, DoubleCRT fromKey, const SKHandle& fromHandle, long toKeyID,
                         const DoubleCRT& toKey, long p, bool ckks, double stdev, Gen& g, DrawA&& drawA) {
  KeySwitch W; W.fromKey = fromHandle; W.toKeyID = toKeyID;
  if (ckks) p = 1;
  else if (p < 2) throw LogicError("Invalid p value found generating BGV key-switching matrix");
  W.ptxtSpace = p;
  const IndexSet all = context.getCtxtPrimes() | context.getSpecialPrimes();
  const size_t n = context.getDigits().size();
  for (size_t i = 0; i < n; i++) { W.a.emplace_back(context, all); drawA(W.a.back()); }
  for (size_t i = 0; i < n; i++) { W.b.emplace_back(context, all); W.noiseBound = XD(RLWE1(W.b[i], W.a[i], toKey, p, stdev, g)); }
  fromKey.addPrimes(all / fromKey.getIndexSet());
  fromKey.multiplyByPrimes(context.getSpecialPrimes());
  for (size_t i = 0; i < n; i++) { W.b[i] += fromKey; fromKey.multiplyByPrimes(context.getDigit((long)i)); }
  return W;
}

}  // namespace hb
