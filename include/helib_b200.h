/* helib_b200.h -- C ABI of the B200-native DoubleCRT / NTT / key-switch engine.
 *
 * This is the drop-in boundary for HElib's hot path.  HElib itself has no
 * plugin/FFI seam above row granularity (src/intelExt.h:22-58 is per-row and CPU-only), so the
 * boundary sits at the helib::DoubleCRT method level: each entry point below replaces the body of
 * one DoubleCRT / Cmodulus / Ctxt method, cited as `reference: path:line` (paths relative to the
 * HElib source tree, v2.2.0).
 *
 * Conventions
 *  - Opaque handles.  The engine owns device memory; the caller owns host buffers.
 *  - Every function returns 0 on success or a negative HB_ERR_* code; nothing throws across the
 *    ABI.  hb_last_error() returns a thread-local message for the last failure.
 *  - A device polynomial (hb_poly) is a dense matrix uint64[nprimes][N]: the row of chain prime i
 *    holds canonical residues in [0, q_i) in HElib's evaluation order row[j] = f(psi_i^(2j+1))
 *    (reference: src/CModulus.cpp:392-426, src/PAlgebra.cpp:535-540).  Which rows are live is
 *    the caller's metadata (helib::IndexSet), passed to every call as an index list.
 *  - Host-side dense matrices use the same [nprimes][N] layout; only the rows named by the index
 *    list are read or written.
 *  - Calls are stream-ordered on the context's CUDA stream and asynchronous unless stated;
 *    hb_ctx_sync() or any download synchronises.  A context may be used by one host thread at a
 *    time (the reference's DoubleCRT has the same value-semantic rule).
 *  - There is no CPU fallback: without a CUDA device hb_ctx_create fails with HB_ERR_NO_DEVICE.
 */
#ifndef HELIB_B200_H
#define HELIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_BAD_ARG (-1)      /* helib::InvalidArgument / LogicError */
#define HB_ERR_INDEX_SET (-2)    /* helib::RuntimeError: index-set precondition violated */
#define HB_ERR_NO_DEVICE (-3)
#define HB_ERR_CUDA (-4)
#define HB_ERR_OOM (-5)
#define HB_ERR_UNSUPPORTED (-6)

#define HB_OP_ADD 0
#define HB_OP_SUB 1
#define HB_OP_MUL 2
#define HB_OP_NEG 3
#define HB_OP_COPY 7

typedef struct hb_ctx hb_ctx;
typedef struct hb_poly hb_poly;

int hb_version(void);
const char* hb_last_error(void);
int hb_device_count(void);

/* ---- context: the device image of helib::Context's prime chain ---------------------------
 * reference: include/helib/Context.h:120-180 (moduli, smallPrimes/ctxtPrimes/specialPrimes,
 * digits), src/CModulus.cpp:62-135 (per-prime tables).  m must be a power of two (phi(m)=m/2);
 * q[i] are the chain primes in index order; psi[i] a primitive m-th root of unity mod q[i]
 * (i.e. 2N-th root, N = m/2), or psi == NULL to let the engine derive one deterministically
 * (smallest quadratic non-residue g, psi = g^((q-1)/m)). */
int hb_ctx_create(hb_ctx** out, int device, uint64_t m, int nprimes, const uint64_t* q, const uint64_t* psi);
void hb_ctx_destroy(hb_ctx* ctx);
/* digit_of[i] = digit number of ctxt prime i or -1; special = indices of the special primes.
 * reference: src/Context.cpp:902-928 (digits), :1014-1028 (special primes). */
int hb_ctx_set_chain(hb_ctx* ctx, const int32_t* digit_of, int ndigits, const int32_t* special, int nspecial);
int hb_ctx_get_psi(hb_ctx* ctx, uint64_t* psi_out);
int hb_ctx_sync(hb_ctx* ctx);
/* out[0] = exact-CRT fallback evaluations, out[1] = kernels launched, out[2] = bytes of device memory held */
int hb_ctx_stats(hb_ctx* ctx, uint64_t* out3);
int hb_ctx_reset_stats(hb_ctx* ctx);
/* Launch-timing hooks for bench.py: elapsed GPU milliseconds on the context's stream between the
 * two marks (CUDA events on the launching stream). */
int hb_ctx_mark_begin(hb_ctx* ctx);
int hb_ctx_mark_end(hb_ctx* ctx, float* ms_out);
/* Per-kernel profile for the roofline report: while enabled, every launch is bracketed by CUDA
 * events on the launching stream.  hb_ctx_profile_get(i, ...) returns HB_ERR_BAD_ARG past the end.
 * bytes = algorithmic HBM bytes (each row read once / written once) summed over the launches. */
int hb_ctx_profile(hb_ctx* ctx, int enable);
int hb_ctx_profile_get(hb_ctx* ctx, int i, char* name, int namelen, uint64_t* launches, double* ms, uint64_t* bytes);

/* ---- device polynomials (helib::DoubleCRT storage, include/helib/DoubleCRT.h:87-94) ---- */
int hb_poly_create(hb_ctx* ctx, hb_poly** out);          /* zero-filled [nprimes][N] */
void hb_poly_destroy(hb_poly* p);
int hb_poly_upload(hb_poly* p, const int32_t* idx, int n, const uint64_t* host_dense);
int hb_poly_download(hb_poly* p, const int32_t* idx, int n, uint64_t* host_dense);  /* synchronises */
int hb_poly_download_async(hb_poly* p, const int32_t* idx, int n, uint64_t* host_dense); /* stream-ordered; hb_ctx_sync() before reading */

/* Wire format: byte-compatible with DoubleCRT::writeTo / read (src/DoubleCRT.cpp:1530-1561):
 * IndexSet (int64 card, int64 indices) then per row int32 length, int32 intSize(=8), little-endian int64 values
 * (src/binio.cpp:103-122).  hb_poly_deserialize also accepts intSize=4 rows and rejects residues outside [0,q). */
int hb_poly_serialized_size(hb_poly* p, int n, uint64_t* bytes);
int hb_poly_serialize(hb_poly* p, const int32_t* idx, int n, void* buf, uint64_t buflen);
int hb_poly_deserialize(hb_poly* p, const void* buf, uint64_t buflen, int32_t* idx_out, int* n_out);

/* ---- per-prime transforms: Cmodulus::FFT / iFFT (src/CModulus.cpp:362-429, 486-553) -----
 * In place on rows idx of each poly: coefficient rows (values in [0,q)) <-> evaluation rows. */
int hb_ntt_fwd(hb_poly* const* polys, int nitems, const int32_t* idx, int n);
int hb_ntt_inv(hb_poly* const* polys, int nitems, const int32_t* idx, int n);

/* ---- row-wise ring operations: DoubleCRT::Op<Add|Sub|Mul>, Negate, operator=
 * (src/DoubleCRT.cpp:216-384).  dst op= src on rows idx. */
int hb_pointwise(int op, hb_poly* const* dst, hb_poly* const* src, int nitems, const int32_t* idx, int n);
/* rows idx *= scalars[r] (already reduced mod q): DoubleCRT::Op(ZZ, MulFun) (src/DoubleCRT.cpp:339-361) */
int hb_scale_rows(hb_poly* const* polys, int nitems, const int32_t* idx, int n, const uint64_t* scalars);
/* rows idx *= prod(q_k : k in fidx) or its inverse: DoubleCRT::operator/= (src/DoubleCRT.cpp:1122-1139) */
int hb_scale_by_primes(hb_poly* const* polys, int nitems, const int32_t* idx, int n, const int32_t* fidx, int nf, int inverse);
int hb_zero_rows(hb_poly* const* polys, int nitems, const int32_t* idx, int n);

/* DoubleCRT::addPrimesAndScale (src/DoubleCRT.cpp:603-647) */
int hb_add_primes_and_scale(hb_poly* const* polys, int nitems, const int32_t* cur, int ncur, const int32_t* add, int nadd);
/* DoubleCRT::addPrimes (src/DoubleCRT.cpp:565-599): exact base extension of rows cur to rows add */
int hb_add_primes(hb_poly* const* polys, int nitems, const int32_t* cur, int ncur, const int32_t* add, int nadd);
/* DoubleCRT::scaleDownToSet (src/DoubleCRT.cpp:1464-1516): rows keep <- (x - delta)/P, rows cur\keep dropped
 * (left as garbage; the caller's index set shrinks).  ptxt_space = 1 for CKKS. */
int hb_scale_down(hb_poly* const* polys, int nitems, const int32_t* cur, int ncur, const int32_t* keep, int nkeep, uint64_t ptxt_space);
/* DoubleCRT::toPoly (src/DoubleCRT.cpp:925-1113): balanced (or positive) big integers,
 * out[N][Lout] little-endian two's-complement limbs.  Synchronises. */
int hb_to_poly(hb_poly* p, const int32_t* idx, int n, int positive, uint64_t* out_limbs, int Lout);
/* Tail of SecKey::Decrypt (src/keys.cpp:1381-1399): PolyRed(toPoly(ptxt), ptxt_space) times factor
 * (= (intFactor*Q)^-1 mod ptxt_space, or 1), out[N] in [0, ptxt_space).  The big integers stay on the device:
 * N words come back instead of N*(n+1).  ptxt_space >= 2, coprime to the primes in idx.  Synchronises. */
int hb_to_poly_mod_p(hb_poly* p, const int32_t* idx, int n, uint64_t ptxt_space, uint64_t factor, int64_t* out);
/* DoubleCRT(const zzX&, context, s) / FFT(const zzX&, s) (src/DoubleCRT.cpp:87-105, src/CModulus.cpp:339-356):
 * coeffs[nitems][N] signed 64-bit coefficients (|c| < 2^63) are copied once, reduced modulo every prime in idx
 * and transformed on the device.  Synchronises (the host buffer may be reused on return). */
int hb_poly_from_i64(hb_poly* const* polys, int nitems, const int32_t* idx, int n, const int64_t* coeffs);
/* DoubleCRT(const ZZX&, context, s) / FFT(const ZZX&, s) (src/DoubleCRT.cpp:68-85; the per-prime `convert`,
 * src/CModulus.cpp:453-457, timer FFT_remainder): limbs[nitems][N][L] little-endian two's-complement big
 * integers (the layout hb_to_poly produces).  Synchronises. */
int hb_poly_from_limbs(hb_poly* const* polys, int nitems, const int32_t* idx, int n, const uint64_t* limbs, int L);
/* dst += a * b row-wise: `key *= part; ptxt += key` of SecKey::Decrypt (src/keys.cpp:1373-1374) and
 * `parts[i] *= r; parts[i] += e` of PubKey::Encrypt (src/keys.cpp:416,443) in one pass. */
int hb_muladd(hb_poly* const* dst, hb_poly* const* a, hb_poly* const* b, int nitems, const int32_t* idx, int n);
/* ---- powerful basis and the recryption mod-switch ----
 * hb_ctx_set_powerful: PowerfulDCRT(context, mvec) (src/powerful.cpp:246-318): the factorisation of m into prime powers
 *   (what Context::buildRecryptData passes); optional -- without it the engine factors m itself on first use.
 * hb_ctx_powerful_info: number of factors, the factors, and to_poly[i] = cubeToPolyMap[shortToLongMap[i]], the exponent
 *   of X that powerful coefficient i is written to before the reduction modulo Phi_m (powerfulToPoly, src/powerful.cpp:223-244).
 * hb_dcrt_to_powerful: PowerfulDCRT::dcrtToPowerful (src/powerful.cpp:393-410): the balanced integers modulo Q of the
 *   powerful-basis coefficients, out[N][Lout] two's-complement limbs.
 * hb_raw_mod_switch: the per-part body of Ctxt::rawModSwitch (src/Ctxt.cpp:2976-3037): out[N] = the powerful-basis
 *   coefficients scaled by q/Q, rounded with the correction that keeps them = c*q*Q^-1 modulo ptxt_space, reduced
 *   symmetrically mod q (a tie of an even q is left at +-q/2; the reference flips a coin there). */
int hb_ctx_set_powerful(hb_ctx* ctx, const int64_t* mvec, int k);
int hb_ctx_powerful_info(hb_ctx* ctx, int32_t* nfactors, int64_t* mvec, int32_t* to_poly);
int hb_dcrt_to_powerful(hb_poly* p, const int32_t* idx, int n, uint64_t* out_limbs, int Lout);
int hb_raw_mod_switch(hb_poly* p, const int32_t* idx, int n, uint64_t q, uint64_t ptxt_space, int64_t* out);
/* DoubleCRT::breakIntoDigits (src/DoubleCRT.cpp:479-561).  src rows cur (ctxt primes only);
 * digits[item*maxdig + i] receives digit i over cur | special.  *ndig_out = number of digits. */
int hb_break_into_digits(hb_poly* const* src, int nitems, const int32_t* cur, int ncur, hb_poly* const* digits, int maxdig, int* ndig_out);
/* Ctxt::keySwitchDigits (src/Ctxt.cpp:191-230): out0 += sum_i D_i*b_i, out1 += sum_i D_i*a_i on rows idx.
 * evk_a[i]: the expanded pseudo-random a_i rows (the reference regenerates them from prgSeed on
 * every call, src/Ctxt.cpp:196-206; here they are expanded once by the host and cached). */
int hb_keyswitch_digits(hb_poly* const* digits, int maxdig, int ndig, int nitems, const int32_t* idx, int n,
                        hb_poly* const* evk_a, hb_poly* const* evk_b, hb_poly* const* out0, hb_poly* const* out1);
/* The same inner product with the two passes around it folded in (what Ctxt::keySwitchPart + reLinearize do per part,
 * src/Ctxt.cpp:764-768,805-842):
 *   out0[r] = scal[r]*out0[r] + sum_i D_i[r]*b_i[r]   (same for out1 / a_i);  scal[r] == 0 => the row is a pure output
 *   (addPrimesAndScale: scal[r] = prod(special) mod q_r on the rows the part already has, 0 on the special rows);
 *   own / own_dig (optional): rows with own_dig[r] = i >= 0 take digit i from own[item] (the part being switched, whose rows
 *   ARE digit i's own rows) instead of digits[item*maxdig + i] -- no copies of the part into the digit polynomials.
 * Power-of-two m, at most 4 digits. */
int hb_keyswitch_digits_fused(hb_poly* const* digits, int maxdig, int ndig, int nitems, const int32_t* idx, int n,
                              hb_poly* const* evk_a, hb_poly* const* evk_b, hb_poly* const* out0, hb_poly* const* out1,
                              const uint64_t* scal, hb_poly* const* own, const int32_t* own_dig);
/* The mixed-radix step of DoubleCRT::breakIntoDigits (src/DoubleCRT.cpp:551-556) in one pass:
 * dst = (dst - src) / prod(q_f, f in fidx) on rows idx. */
int hb_sub_div_by_primes(hb_poly* const* dst, hb_poly* const* src, int nitems, const int32_t* idx, int n, const int32_t* fidx, int nf);
/* Hoisted automorphism + key switch: BasicAutomorphPrecon::automorph (src/matmul.cpp:112-184),
 * the rotation path of Ctxt::smartAutomorph (src/Ctxt.cpp:2462-2515).  The digits of the s-part are computed once
 * with hb_break_into_digits; for each amount k (odd, < m) one launch applies sigma_k in the load stage:
 *   out0 = P*sigma_k(c0) + sum_i sigma_k(D_i)*b_i ,  out1 = sum_i sigma_k(D_i)*a_i     over S | special
 * (evk_a/evk_b: the matrix for s(X^k) -> s).  Outputs must not alias inputs.  Power-of-two m. */
int hb_automorph_keyswitch_digits(hb_poly* const* digits, int maxdig, int ndig, int nitems, const int32_t* S, int nS,
                                  hb_poly* const* c0, uint64_t k, hb_poly* const* evk_a, hb_poly* const* evk_b,
                                  hb_poly* const* out0, hb_poly* const* out1);
/* Ctxt::tensorProduct of two canonical 2-part ciphertexts (src/Ctxt.cpp:1563-1608) */
int hb_tensor(hb_poly* const* a0, hb_poly* const* a1, hb_poly* const* b0, hb_poly* const* b1,
              hb_poly* const* o0, hb_poly* const* o1, hb_poly* const* o2, int nitems, const int32_t* idx, int n);
/* DoubleCRT::automorph (src/DoubleCRT.cpp:1160-1202): dst[j] = src[idx(rep(j)*k mod m)], dst != src */
int hb_automorph(hb_poly* const* dst, hb_poly* const* src, int nitems, const int32_t* idx, int n, uint64_t k);

/* ---- noise metadata: canonical-embedding norms (embeddingLargestCoeff, src/norms.cpp:204-261,443-485),
 * computed in FP64 on the device from the coefficient data the conversion kernels already hold.
 * Same operations as above plus the norm outputs; these variants synchronise (they return host values).
 *  hb_add_primes_norm        : log_norms[item] = ln max_j |f(zeta^j)| of the balanced polynomial being extended
 *  hb_break_into_digits_norm : log_norms[item*maxdig+i] = ln ||E_i||  (breakIntoDigits returns their sum, src/DoubleCRT.cpp:542-545)
 *  hb_scale_down_norm        : norms[item] = ||delta/P||  (the fdelta norms of Ctxt::modDownToSet, src/Ctxt.cpp:476-505) */
int hb_add_primes_norm(hb_poly* const* polys, int nitems, const int32_t* cur, int ncur, const int32_t* add, int nadd, double* log_norms);
int hb_break_into_digits_norm(hb_poly* const* src, int nitems, const int32_t* cur, int ncur, hb_poly* const* digits, int maxdig, int* ndig_out, double* log_norms);
int hb_scale_down_norm(hb_poly* const* polys, int nitems, const int32_t* cur, int ncur, const int32_t* keep, int nkeep, uint64_t ptxt_space, double* norms);

/* ---- prime-sharded base conversion (one rank per GPU, rows sharded by prime index).
 * The exact conversion of hb_add_primes / hb_scale_down split where residues must cross shards:
 *   hb_conv_make_y : for the owned rows of the source set D: y_j = iNTT(row_j) * (Q_D/q_j)^-1 mod q_j,
 *                    written to ypolys rows `owned` in coefficient order (local work, no communication);
 *   [caller: all-gather of the y rows over NCCL/NVLink so that every rank holds all rows of D]
 *   hb_conv_from_y : exact CRT of the gathered y rows, reduction mod the target primes this rank owns,
 *                    forward transform into dst rows tgt.  mode 0: dst = x (addPrimes, src/DoubleCRT.cpp:565-599);
 *                    mode 1: dst = (dst - x) / Q_D with the BGV correction (scaleDownToSet, :1464-1516). */
int hb_conv_make_y(hb_poly* const* polys, int nitems, const int32_t* D, int nD, const int32_t* owned, int nOwned, hb_poly* const* ypolys);
int hb_conv_from_y(hb_poly* const* ypolys, int nitems, const int32_t* D, int nD, const int32_t* tgt, int nT,
                   uint64_t ptxt_space, hb_poly* const* dst, int mode);
/* Fused "make y + all-gather": like hb_conv_make_y, but the final kernel also stores the y rows into the
 * y buffers of up to 8 peer GPUs (peer_ypolys[p*nitems + item], obtained with hb_poly_ipc_open), so the rows
 * cross NVLink once, straight into place.  The caller orders a cross-rank barrier (e.g. a 1-element NCCL
 * all-reduce on the same stream) before hb_conv_from_y reads the buffers. */
int hb_conv_make_y_bcast(hb_poly* const* polys, int nitems, const int32_t* D, int nD, const int32_t* owned, int nOwned,
                         hb_poly* const* ypolys, hb_poly* const* peer_ypolys, int npeers);
/* CUDA IPC export / import of a polynomial's device buffer (one process per GPU; 64-byte handle). */
int hb_poly_ipc_export(hb_poly* p, void* handle64);
int hb_poly_ipc_open(hb_ctx* ctx, const void* handle64, hb_poly** out);
/* Alias caller-owned device memory (uint64[nprimes][N]) as a polynomial; hb_poly_destroy does not free it. */
int hb_poly_wrap(hb_ctx* ctx, void* device_ptr, hb_poly** out);
/* Issue the context's work on a caller-provided CUDA stream (cudaStream_t), e.g. the stream the caller's
 * NCCL collectives are ordered on. */
int hb_ctx_set_stream(hb_ctx* ctx, void* cuda_stream);

/* ---- fused ciphertext-level paths (host orchestration of Ctxt::reLinearize / keySwitchPart,
 * src/Ctxt.cpp:720-842, and Ctxt::multLowLvl + reLinearize + modDownToSet, src/Ctxt.cpp:393-562,
 * 1681-1774) with explicit prime sets (the noise-driven choice stays in the host Ctxt layer).
 * hb_relinearize: (c0,c1,c2) over ctxt primes S  ->  (c0,c1) over S | special   (c2 is consumed). */
int hb_relinearize(hb_poly* const* c0, hb_poly* const* c1, hb_poly* const* c2, int nitems,
                   const int32_t* S, int nS, hb_poly* const* evk_a, hb_poly* const* evk_b, int ndig_evk);
/* hb_mul_relin_moddown: operands (a0,a1),(b0,b1) over S_in; mod-down both to S (ptxt_space),
 * tensor, relinearise over S | special, mod-down the result to S.  Result in (a0,a1) rows S.
 * S is the common set of Ctxt::multiplyBy / multLowLvl (src/Ctxt.cpp:1700-1712) and must be a subset of S_in
 * (HB_ERR_INDEX_SET otherwise). */
int hb_mul_relin_moddown(hb_poly* const* a0, hb_poly* const* a1, hb_poly* const* b0, hb_poly* const* b1, int nitems,
                         const int32_t* S_in, int nS_in, const int32_t* S, int nS, uint64_t ptxt_space,
                         hb_poly* const* evk_a, hb_poly* const* evk_b, int ndig_evk);

#ifdef __cplusplus
}
#endif
#endif /* HELIB_B200_H */
