// hb_chain.cpp -- host C++17 owner of the prime chain (north_star: "host C++17 owns
// Context/primeChain").  Produces, for identical (m, p, r, bits, c), the same primes in the same
// index order as helib::Context::buildModChain, so device rows line up with the reference's.
//
// Behaviour follows (paths relative to the HElib tree):
//   PrimeGenerator            src/PrimeGenerator.h:39-127
//   ctxtPrimeSize             src/Context.cpp:816-843
//   addSmallPrimes            src/Context.cpp:728-790
//   addCtxtPrimes             src/Context.cpp:845-872
//   addSpecialPrimes + digits src/Context.cpp:874-1035
//   ModuliSizes               src/primeChain.cpp:68-335
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/helib_b200.h"
#include "../../include/helib_b200_chain.h"

namespace {

typedef unsigned __int128 u128;
const long SP_NBITS = 60;   // HELIB_SP_NBITS without HEXL (src/macro.h:16-23)
const long GEN_B = 3;       // PrimeGenerator::B

uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }
uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
  uint64_t r = 1 % q; a %= q;
  while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
  return r;
}
// deterministic Miller-Rabin, exact for n < 3.3e24; stands in for NTL::ProbPrime(cand, 60)
bool is_prime(uint64_t n) {
  static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return false;
  for (uint64_t p : bases) if (n % p == 0) return n == p;
  uint64_t d = n - 1; int s = 0;
  while ((d & 1) == 0) { d >>= 1; s++; }
  for (uint64_t a : bases) {
    uint64_t x = powmod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < s; i++) { x = mulmod(x, x, n); if (x == n - 1) { comp = false; break; } }
    if (comp) return false;
  }
  return true;
}
long divc(long a, long b) { return (a + b - 1) / b; }

class PrimeGen {
  long len, m, k, t;
 public:
  PrimeGen(long len_, long m_) : len(len_), m(m_) {
    if (len < GEN_B || len > SP_NBITS) throw std::invalid_argument("PrimeGenerator: len is not in [B, HELIB_SP_NBITS]");
    if (m < 1 || m >= (1L << SP_NBITS)) throw std::invalid_argument("PrimeGenerator: m is not in [1, NTL_SP_BOUND)");
    k = 0;
    while ((m << k) <= (1L << (len - GEN_B))) k++;
    t = divc((1L << len) - 1, m << k);
  }
  long next() {
    long t_upper = divc((1L << len) - 1, m << k);
    for (;;) {
      t++;
      if (t >= t_upper) {
        k--;
        long k_lower = (m % 2 == 0) ? 0 : 1;
        if (k < k_lower) throw std::runtime_error("Prime generator ran out of primes");
        t = divc((1L << len) - (1L << (len - GEN_B)) - 1, m << k);
        t_upper = divc((1L << len) - 1, m << k);
      }
      if (t % 2 == 0) continue;
      long cand = ((t * m) << k) + 1;
      if (is_prime((uint64_t)cand)) return cand;
    }
  }
};

double bit_loss() { return -std::log1p(-1.0 / double(1L << GEN_B)) / std::log(2.0); }

long ctxt_prime_size(long nBits) {
  double bl = bit_loss();
  double maxPsize = SP_NBITS - bl;
  long nPrimes = long(std::ceil(nBits / maxPsize));
  long target = SP_NBITS;
  while (10 * (target - 1) >= 9 * SP_NBITS && (target - 1) >= 30 && ((target - 1) - bl) * nPrimes >= nBits) target--;
  return target;
}

long euler_phi(long m) {
  long r = m, n = m;
  for (long p = 2; p * p <= n; p++) if (n % p == 0) { while (n % p == 0) n /= p; r -= r / p; }
  if (n > 1) r -= r / n;
  return r;
}

}  // namespace

struct hb_chain {
  long m, p, r, phim;
  bool ckks, pow2;
  std::vector<long> primes;
  std::vector<int> kind;       // 0 small, 1 ctxt, 2 special
  std::vector<int> digit_of;   // digit number or -1
  int ndigits;
  std::vector<std::pair<double, std::vector<int>>> sizes;  // ModuliSizes table
  long iFFT_cost;
  std::string err;

  bool in_chain(long q) const { return std::find(primes.begin(), primes.end(), q) != primes.end(); }
  void add(long q, int k) { primes.push_back(q); kind.push_back(k); digit_of.push_back(-1); }
  std::vector<int> of_kind(int k) const { std::vector<int> v; for (size_t i = 0; i < kind.size(); i++) if (kind[i] == k) v.push_back((int)i); return v; }
  double log_of_product(const std::vector<int>& s) const { double x = 0; for (int i : s) x += std::log((double)primes[i]); return x; }

  long e_param = 0, ePrime_param = 0, hwt_param = 0;   // Context::e_param / ePrime_param / hwt_param

  // compute_fudge (src/recryption.cpp:154-197): the v-coefficients of the recryption are not quite uniform
  static double compute_fudge(long p2ePrime, long p2e) {
    double eps = 0;
    if (p2ePrime > 1) {
      if (p2ePrime % 2 == 0) eps = 1.0 / ((double)p2ePrime * (double)p2ePrime);
      else eps = 1.0 / (double)p2e;
    }
    return 1 + eps;
  }
  static long ipow(long b, long e) { long r = 1; while (e-- > 0) r *= b; return r; }
  // RecryptData::setAE (src/recryption.cpp:200-256) with Context::boundForRecryption (include/helib/Context.h:616-638)
  void set_ae(long& e, long& ePrime, long skHwt, double scale) const {
    long k = 0; { long mm = m; for (long f = 2; f * f <= mm; f++) if (mm % f == 0) { k++; while (mm % f == 0) mm /= f; } if (mm > 1) k++; }
    const double mrat = (double)phim / (double)m;
    const double stddev = std::sqrt(mrat * (double)skHwt * (double)(1L << k) / 3.0) * 0.5;
    const double coeff_bound = 0.5 + scale * stddev;
    long p2r = ipow(p, r);
    const long frstTerm = 2 * p2r + 2;
    long e_bnd = 0, p2e_bnd = 1;
    while (p2e_bnd <= ((1L << 30) - 2) / p) { e_bnd++; p2e_bnd *= p; }   // largest e with p^e + 1 < 2^30
    ePrime = 0;
    e = r + 1;
    while (e <= e_bnd && (double)ipow(p, e) < frstTerm * coeff_bound * 2) e++;
    if (e > e_bnd) throw std::runtime_error("setAE: cannot find suitable e");
    long ePrimeTry = 1;
    while (ePrimeTry <= e_bnd) {
      const long p2ePrimeTry = ipow(p, ePrimeTry);
      long eTry = std::max(r + 1, ePrimeTry + 1);
      while (eTry <= e_bnd && eTry - ePrimeTry < e - ePrime) {
        const long p2eTry = ipow(p, eTry);
        const double fudge = compute_fudge(p2ePrimeTry, p2eTry);
        if ((double)p2eTry >= ((double)p2ePrimeTry * fudge + frstTerm) * coeff_bound * 2) break;
        eTry++;
      }
      if (eTry <= e_bnd && eTry - ePrimeTry < e - ePrime) { e = eTry; ePrime = ePrimeTry; }
      ePrimeTry++;
    }
  }

  void build(long bits, long nDgts, long skHwt, long resolution, long bitsInSpecial, double stdev, bool bootstrappable = false, double scale = 10.0) {
    if (bits <= 0) throw std::invalid_argument("Cannot initialise modulus chain with nBits < 1");
    if (skHwt < 0) throw std::invalid_argument("invalid skHwt parameter");
    if (ckks) bootstrappable = false;                       // src/Context.cpp:1051-1052
    if (skHwt == 0 && bootstrappable) skHwt = 120;          // BOOT_DFLT_SK_HWT = MIN_SK_HWT (include/helib/Context.h:34-35)
    hwt_param = skHwt;
    // ---- addSmallPrimes
    long cp = ctxt_prime_size(bits);
    if (m <= 0 || m > (1 << 20)) throw std::runtime_error("addSmallPrimes: m undefined or larger than 2^20");
    if (resolution < 1 || resolution > 10) resolution = 3;
    std::vector<long> sz;
    long smallest;
    if (cp >= 54) smallest = divc(2 * cp, 3);
    else if (cp >= 45) smallest = divc(7 * cp, 10);
    else { smallest = divc(11 * cp, 15); sz.push_back(smallest); }
    sz.push_back(smallest); sz.push_back(smallest);
    for (long delta = resolution; cp - delta > smallest; delta *= 2) sz.push_back(cp - delta);
    if (cp - 3 * resolution > smallest) sz.push_back(cp - 3 * resolution);
    if (resolution == 1 && cp - 11 > smallest) sz.push_back(cp - 11);
    std::sort(sz.begin(), sz.end());
    long last = 0; std::unique_ptr<PrimeGen> gen;
    for (long s : sz) {
      if (s != last) gen.reset(new PrimeGen(s, m));
      add(gen->next(), 0);
      last = s;
    }
    // ---- addCtxtPrimes
    {
      PrimeGen g(cp, m);
      double bitlen = 0;
      while (bitlen < bits - 0.5) { long q = g.next(); add(q, 1); bitlen += std::log2((double)q); }
    }
    // ---- digits + addSpecialPrimes
    long pabs = std::labs(p);
    long p2r = 1;
    if (!ckks) for (long i = 0; i < r; i++) p2r *= pabs;
    long p2e = p2r;
    if (bootstrappable && !ckks) {   // bigger p^e for bootstrapping (src/Context.cpp:885-897)
      long e, ePrime;
      set_ae(e, ePrime, skHwt, scale);
      p2e *= ipow(pabs, e - ePrime);
      e_param = e; ePrime_param = ePrime;
    }
    std::vector<int> ctxt = of_kind(1);
    long nCtxt = (long)ctxt.size();
    if (nDgts > nCtxt) nDgts = nCtxt;
    if (nDgts <= 0) nDgts = 1;
    std::vector<std::vector<int>> digits(nDgts);
    if (nDgts > 1) {
      std::vector<int> remaining = ctxt;
      for (long d = 0; d < nDgts - 1; d++) {
        long card = divc((long)remaining.size(), nDgts - d);
        for (int i : remaining) { digits[d].push_back(i); if ((long)digits[d].size() >= card) break; }
        std::vector<int> rest;
        for (int i : remaining) if (std::find(digits[d].begin(), digits[d].end(), i) == digits[d].end()) rest.push_back(i);
        remaining.swap(rest);
      }
      if (remaining.empty()) { nDgts--; digits.resize(nDgts); }
      else digits[nDgts - 1] = remaining;
    } else digits[0] = ctxt;
    ndigits = (int)nDgts;
    for (int d = 0; d < ndigits; d++) for (int i : digits[d]) digit_of[i] = d;
    double maxDigitLog = 0;
    for (auto& d : digits) maxDigitLog = std::max(maxDigitLog, log_of_product(d));
    double nBits;
    if (bitsInSpecial) nBits = (double)bitsInSpecial;
    else {
      double h = skHwt == 0 ? phim / 2.0 : (double)skHwt;
      double log_phim = std::log((double)phim);
      if (log_phim < 1) log_phim = 1;
      if (ckks)
        nBits = (maxDigitLog + std::log(stdev) + std::log((double)nDgts) - 0.5 * std::log(h)) / std::log(2.0);
      else if (pow2)
        nBits = (maxDigitLog + std::log((double)p2e) + std::log(stdev) + 0.5 * std::log(12.0) + std::log((double)nDgts) -
                 0.5 * std::log(log_phim) - 2 * std::log((double)pabs) - std::log(h)) / std::log(2.0);
      else
        nBits = (maxDigitLog + std::log((double)m) + std::log((double)p2e) + std::log(stdev) + 0.5 * std::log(12.0) +
                 std::log((double)nDgts) - 0.5 * log_phim - 0.5 * std::log(log_phim) - 2 * std::log((double)pabs) - std::log(h)) / std::log(2.0);
    }
    if (nBits < 1) nBits = 1;
    double bl = bit_loss();
    double maxPsize = SP_NBITS - bl;
    long nPrimes = long(std::ceil(nBits / maxPsize));
    long target = SP_NBITS;
    while ((target - 1) >= 0.55 * SP_NBITS && (target - 1) >= 30 && ((target - 1) - bl) * nPrimes >= nBits) target--;
    PrimeGen g(target, m);
    while (nPrimes > 0) { long q = g.next(); if (in_chain(q)) continue; add(q, 2); nPrimes--; }
    init_sizes();
  }

  // ModuliSizes::init (src/primeChain.cpp:68-128)
  void init_sizes() {
    iFFT_cost = pow2 ? 0 : 20;
    sizes.clear();
    sizes.push_back({0.0, {}});
    size_t idx = 1;
    for (int i : of_kind(0)) {
      double sz = std::log((double)primes[i]);
      for (size_t j = idx; j < 2 * idx; j++) { auto e = sizes[j - idx]; e.first += sz; e.second.push_back(i); sizes.push_back(e); }
      idx *= 2;
    }
    std::vector<int> interval; double isz = 0;
    for (int i : of_kind(1)) {
      interval.push_back(i); isz += std::log((double)primes[i]);
      for (size_t j = 0; j < idx; j++) { auto e = sizes[j]; e.first += isz; e.second.insert(e.second.end(), interval.begin(), interval.end()); std::sort(e.second.begin(), e.second.end()); sizes.push_back(e); }
    }
    // std::sort on pair<double,IndexSet>: IndexSet ordering only breaks exact size ties
    std::stable_sort(sizes.begin(), sizes.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  }
  static long card_minus(const std::vector<int>& a, const std::vector<int>& b) {  // |a \ b|
    long n = 0; for (int x : a) if (std::find(b.begin(), b.end(), x) == b.end()) n++; return n;
  }
  long cost(const std::vector<int>& from, const std::vector<int>& to) const {  // src/primeChain.cpp:150-160
    if (iFFT_cost == 0) return 100 * card_minus(to, from);
    return 100 * card_minus(to, from) + iFFT_cost * card_minus(from, to);
  }
  // ModuliSizes::getSet4Size, two-operand form (src/primeChain.cpp:257-335); from2 may be empty (one-operand :179-250)
  std::vector<int> set4size(double low, double high, const std::vector<int>& f1, const std::vector<int>* f2, bool reverse) const {
    long n = (long)sizes.size();
    long idx = std::lower_bound(sizes.begin(), sizes.end(), low, [](const auto& e, double v) { return e.first < v; }) - sizes.begin();
    long best = -1, bestCost = LONG_MAX, ii = idx;
    auto c2 = [&](long i) { return cost(f1, sizes[i].second) + (f2 ? cost(*f2, sizes[i].second) : 0); };
    for (; ii < n && sizes[ii].first <= high; ii++) { long c = c2(ii); if (c <= bestCost) { best = ii; bestCost = c; } }
    if (best == -1) {
      if (reverse) {
        if (ii < n) { double ub = sizes[ii].first + std::log(2.0); for (long i = ii; i < n && sizes[i].first <= ub; ++i) { long c = c2(i); if (c < bestCost) { best = i; bestCost = c; } } }
      } else if (idx > 0) {
        double lb = sizes[idx - 1].first - std::log(2.0);
        for (long i = idx - 1; i >= 0 && sizes[i].first >= lb; --i) { long c = c2(i); if (c < bestCost) { best = i; bestCost = c; } }
      }
    }
    if (best == -1) return {};
    return sizes[best].second;
  }
};

static thread_local std::string g_chain_err;

extern "C" const char* hb_chain_last_error(void) { return g_chain_err.c_str(); }

extern "C" int hb_chain_build(hb_chain** out, uint64_t m, int64_t p, int r, int bits, int c, int sk_hwt, int resolution,
                              int bits_in_special, double stdev) {
  return hb_chain_build_ex(out, m, p, r, bits, c, sk_hwt, resolution, bits_in_special, stdev, 0, 10.0);
}
extern "C" int hb_chain_recrypt_params(const hb_chain* ch, int64_t* e, int64_t* e_prime, int64_t* sk_hwt) {
  if (!ch) return HB_ERR_BAD_ARG;
  if (e) *e = ch->e_param;
  if (e_prime) *e_prime = ch->ePrime_param;
  if (sk_hwt) *sk_hwt = ch->hwt_param;
  return HB_OK;
}
extern "C" int hb_chain_build_ex(hb_chain** out, uint64_t m, int64_t p, int r, int bits, int c, int sk_hwt, int resolution,
                                 int bits_in_special, double stdev, int will_be_bootstrappable, double scale) {
  if (!out) return HB_ERR_BAD_ARG;
  std::unique_ptr<hb_chain> ch(new hb_chain());
  try {
    if (m < 2) throw std::invalid_argument("Bad Z_m^* modulus m (must be greater than 1)");
    ch->m = (long)m; ch->p = (long)p; ch->r = r; ch->ckks = p == -1;
    ch->pow2 = (m & (m - 1)) == 0;
    if (!ch->ckks && (p < 2 || m % (uint64_t)p == 0)) throw std::invalid_argument("Modulus pp divides mm");  // src/PAlgebra.cpp:458
    if (ch->ckks && !ch->pow2) throw std::invalid_argument("CKKS requires m to be a power of two");
    ch->phim = euler_phi((long)m);
    ch->build(bits, c, sk_hwt, resolution, bits_in_special, stdev > 0 ? stdev : 3.2, will_be_bootstrappable != 0, scale > 0 ? scale : 10.0);
  } catch (const std::exception& e) {
    g_chain_err = e.what();
    return HB_ERR_BAD_ARG;
  }
  *out = ch.release();
  return HB_OK;
}
extern "C" void hb_chain_destroy(hb_chain* ch) { delete ch; }
extern "C" int hb_chain_info(const hb_chain* ch, int* nprimes, int* nsmall, int* nctxt, int* nspecial, int* ndigits, int64_t* phim) {
  if (!ch) return HB_ERR_BAD_ARG;
  if (nprimes) *nprimes = (int)ch->primes.size();
  if (nsmall) *nsmall = (int)ch->of_kind(0).size();
  if (nctxt) *nctxt = (int)ch->of_kind(1).size();
  if (nspecial) *nspecial = (int)ch->of_kind(2).size();
  if (ndigits) *ndigits = ch->ndigits;
  if (phim) *phim = ch->phim;
  return HB_OK;
}
extern "C" int hb_chain_get(const hb_chain* ch, uint64_t* primes, int32_t* kind, int32_t* digit_of) {
  if (!ch) return HB_ERR_BAD_ARG;
  for (size_t i = 0; i < ch->primes.size(); i++) {
    if (primes) primes[i] = (uint64_t)ch->primes[i];
    if (kind) kind[i] = ch->kind[i];
    if (digit_of) digit_of[i] = ch->digit_of[i];
  }
  return HB_OK;
}
extern "C" int hb_chain_set4size(const hb_chain* ch, double low, double high, const int32_t* from1, int n1,
                                 const int32_t* from2, int n2, int reverse, int32_t* out, int* nout) {
  if (!ch || !out || !nout) return HB_ERR_BAD_ARG;
  std::vector<int> f1(from1, from1 + n1), f2;
  if (from2) f2.assign(from2, from2 + n2);
  std::vector<int> s = ch->set4size(low, high, f1, from2 ? &f2 : nullptr, reverse != 0);
  *nout = (int)s.size();
  for (size_t i = 0; i < s.size(); i++) out[i] = s[i];
  return HB_OK;
}
