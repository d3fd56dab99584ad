/* helib_b200_chain.h -- C ABI of the host-side prime-chain builder (pure host C++17, no CUDA).
 *
 * Reproduces helib::Context::buildModChain for identical (m, p, r, bits, c): same primes, same
 * index order (small, then ctxt, then special), same digit partition, same ModuliSizes table.
 * reference: src/PrimeGenerator.h:39-127, src/Context.cpp:728-1092, src/primeChain.cpp:68-335.
 */
#ifndef HELIB_B200_CHAIN_H
#define HELIB_B200_CHAIN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_chain hb_chain;

/* p = -1 selects CKKS.  c = number of key-switching digits (ContextBuilder::c, default 3);
 * sk_hwt, resolution, bits_in_special, stdev as ContextBuilder (0 / 3 / 0 / 3.2 by default;
 * include/helib/Context.h:1067-1087).  Bootstrappable chains are not supported. */
int hb_chain_build(hb_chain** out, uint64_t m, int64_t p, int r, int bits, int c, int sk_hwt, int resolution,
                   int bits_in_special, double stdev);
/* The same with ContextBuilder::bootstrappable(): will_be_bootstrappable != 0 (ignored for CKKS) sets the default secret-key
 * weight to BOOT_DFLT_SK_HWT = 120 when sk_hwt == 0 and sizes the special primes for p^(r + e - e') with (e, e') from
 * RecryptData::setAE (src/Context.cpp:885-897, src/recryption.cpp:200-256); scale = Context::scale (10 by default). */
int hb_chain_build_ex(hb_chain** out, uint64_t m, int64_t p, int r, int bits, int c, int sk_hwt, int resolution,
                      int bits_in_special, double stdev, int will_be_bootstrappable, double scale);
/* Context::e_param / ePrime_param (0 unless bootstrappable) and hwt_param. */
int hb_chain_recrypt_params(const hb_chain* ch, int64_t* e, int64_t* e_prime, int64_t* sk_hwt);
void hb_chain_destroy(hb_chain* ch);
const char* hb_chain_last_error(void);
int hb_chain_info(const hb_chain* ch, int* nprimes, int* nsmall, int* nctxt, int* nspecial, int* ndigits, int64_t* phim);
/* kind[i]: 0 small, 1 ctxt, 2 special (Context::smallPrimes/ctxtPrimes/specialPrimes);
 * digit_of[i]: key-switching digit of ctxt prime i or -1 (Context::digits). */
int hb_chain_get(const hb_chain* ch, uint64_t* primes, int32_t* kind, int32_t* digit_of);
/* ModuliSizes::getSet4Size (src/primeChain.cpp:179-335); from2 == NULL selects the one-set form. */
int hb_chain_set4size(const hb_chain* ch, double low, double high, const int32_t* from1, int n1,
                      const int32_t* from2, int n2, int reverse, int32_t* out, int* nout);

#ifdef __cplusplus
}
#endif
#endif
