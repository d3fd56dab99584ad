// hb_device_gen.cuh -- rows for general (non power-of-two) m: Bluestein transforms and the unfused
// exact base conversion on plain coefficient rows.
//
// Reference: BluesteinInit / BluesteinFFT (src/bluestein.cpp:77-201) and the general-m branches of
// Cmodulus::FFT_aux / iFFT (src/CModulus.cpp:148-180,431-443,555-577).
//   forward : X_k = root^(k^2) * sum_i (x_i root^(i^2)) * root^(-(k-i)^2), k in Z_m^*  (row[j] = X_rep(j))
//   inverse : scatter the row into Z_m^* positions, the same DFT with root^-1, reduce mod Phi_m(X), times m^-1.
// The two chirp convolutions are cyclic convolutions of length L = 2^ceil(log2(2m-1)) done with the power-of-two
// transform kernels on *cyclic* twiddle tables (same butterfly network, tables omega^brev instead of psi^brev).
// The remainder mod Phi_m uses rev(Phi)^-1 = rev((X^m-1)/Phi_m) mod X^d, d = m - phi(m), so no per-prime power-series
// inversion is needed, and both of its products run at the shorter cyclic length L2 = 2^ceil(log2 max(phi(m), 2d-1)):
// the quotient is a product of two length-d polynomials; q*Phi_m (degree < m) is only needed in its low phi(m)
// coefficients, and everything that wraps around modulo X^L2 - 1 lands on coefficients >= phi(m) of q*Phi_m, which equal
// those of the dividend A because the remainder has degree < phi(m) -- so (A mod Phi_m)[k] = A[k] + sum_{j>=1} A[k + j*L2]
// - (q*Phi_m mod X^L2 - 1)[k].  For m = 21845 this is L2 = 2^14 against L = 2^16.
#pragma once
#include "hb_device.cuh"

struct HbGenPrime {
  const ulonglong2* pw;    // [m]  root^(i^2)      (+Shoup)
  const ulonglong2* ipw;   // [m]  root^(-i^2)     (+Shoup)
  const u64* RbHat;        // [L]  cyclic transform of b[j] = root^(-(j-(m-1))^2), j = 0..2m-2
  const u64* iRbHat;       // [L]  same for root^-1
  const u64* invHat;       // [L2] cyclic transform (length L2) of rev((X^m-1)/Phi_m) mod X^d
  const u64* phiHat;       // [L2] cyclic transform (length L2) of Phi_m mod X^L2 - 1
  u64 minv, minv_s;        // m^-1 mod q
};

struct HbGenJob {
  u64 m, phim, L, L2, d;   // d = m - phi(m); L2 = cyclic length of the division by Phi_m
  const int* rep;          // [phim] j-th unit of Z_m^*
  const int* irep;         // [m]    index of unit i, or -1
  HbRows rows;
  int nitems;
  const u64* src[HB_MAXB];  // polynomial rows, stride phim
  u64* dst[HB_MAXB];
  u64* w0[HB_MAXB];         // work buffers, row stride L (chirp convolutions) or L2 (division by Phi_m)
  u64* w1[HB_MAXB];
  int which;                // selects the fixed vector in k_gen_mulvec: 0 RbHat, 1 iRbHat, 2 invHat, 3 phiHat
};

enum { HB_GEN_PRE_FWD = 0, HB_GEN_POST_FWD, HB_GEN_PRE_INV, HB_GEN_POST_INV, HB_GEN_QREV, HB_GEN_FIN, HB_GEN_MULVEC };

// grid = (blocks over max(L, m), nrows, nitems)
__global__ void __launch_bounds__(HB_THREADS) k_gen(const HbPrimeDev* __restrict__ primes, const HbGenPrime* __restrict__ gp, HbGenJob J, int op) {
  const int pi = J.rows.prime[blockIdx.y];
  const HbPrimeDev P = primes[pi];
  const HbGenPrime G = gp[pi];
  const u64 q = P.q;
  const int it = blockIdx.z;
  const size_t prow = (size_t)pi * J.phim, wrow = (size_t)pi * J.L;
  const u64* src = J.src[it] ? J.src[it] + prow : nullptr;
  u64* dst = J.dst[it] ? J.dst[it] + prow : nullptr;
  u64* w0 = J.w0[it] + wrow;
  const size_t m = J.m, phim = J.phim, L = J.L, L2 = J.L2, d = J.d;
  u64* w0s = J.w0[it] + (size_t)pi * L2;                          // the same buffers addressed as rows of the short plan
  u64* w1s = J.w1[it] ? J.w1[it] + (size_t)pi * L2 : nullptr;
  const bool short_op = op == HB_GEN_QREV || op == HB_GEN_FIN || (op == HB_GEN_MULVEC && J.which >= 2);
  const size_t lim = short_op ? L2 : L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lim; i += (size_t)gridDim.x * blockDim.x) {
    switch (op) {
      case HB_GEN_PRE_FWD:   // y_i = x_i * root^(i^2), zero padded (src/bluestein.cpp:151-155)
        w0[i] = i < phim ? hb_mul_shoup(src[i], G.pw[i].x, G.pw[i].y, q) : 0;
        break;
      case HB_GEN_POST_FWD:  // row[j] = z[rep(j) + m-1] * root^(rep(j)^2) (src/bluestein.cpp:190-196, src/CModulus.cpp:435-443)
        if (i < phim) { const size_t k = (size_t)J.rep[i]; dst[i] = hb_mul_shoup(w0[k + m - 1], G.pw[k].x, G.pw[k].y, q); }
        break;
      case HB_GEN_PRE_INV: { // scatter into Z_m^* positions (src/CModulus.cpp:557-562) and chirp with root^-1
        u64 v = 0;
        if (i < m) { const int j = J.irep[i]; if (j >= 0) v = hb_mul_shoup(src[j], G.ipw[i].x, G.ipw[i].y, q); }
        w0[i] = v;
      } break;
      case HB_GEN_POST_INV:  // A_k = z[k+m-1] * root^(-k^2), k < m; dst_k = A_k + A_(k+L2) + ... (k < phim), w1 = first d coefficients of rev(A)
        if (i < phim) {
          u64 s = hb_mul_shoup(w0[i + m - 1], G.ipw[i].x, G.ipw[i].y, q);
          for (size_t k = i + L2; k < m; k += L2) s = hb_addmod(s, hb_mul_shoup(w0[k + m - 1], G.ipw[k].x, G.ipw[k].y, q), q);
          dst[i] = s;
        }
        if (i < L2) {
          u64 v = 0;
          if (i < d) { const size_t k = m - 1 - i; v = hb_mul_shoup(w0[k + m - 1], G.ipw[k].x, G.ipw[k].y, q); }
          w1s[i] = v;
        }
        break;
      case HB_GEN_QREV:      // quotient by Phi_m: q_k = (rev(A) * rev(Phi)^-1 mod X^d)[d-1-k]
        w0s[i] = i < d ? w1s[d - 1 - i] : 0;
        break;
      case HB_GEN_FIN:       // coefficients = (A - q*Phi) * m^-1 (src/CModulus.cpp:566-577)
        if (i < phim) dst[i] = hb_mul_shoup(hb_submod(dst[i], w0s[i], q), G.minv, G.minv_s, q);
        break;
      case HB_GEN_MULVEC: {
        if (J.which < 2) { const u64* v = J.which == 0 ? G.RbHat : G.iRbHat; w0[i] = hb_mulmod(w0[i], v[i], P); }
        else if (J.which == 2) w1s[i] = hb_mulmod(w1s[i], G.invHat[i], P);   // the quotient product runs in w1
        else w0s[i] = hb_mulmod(w0s[i], G.phiHat[i], P);
      } break;
    }
  }
}

// Unfused exact base conversion on coefficient rows (general m): per coefficient y_j = r_j*(Q/q_j)^-1,
// v = round(sum y_j/q_j) (+ BGV correction), then x mod q_t for every target.  Same arithmetic as the
// fused kernels (hb_conv_v); src/dst are coefficient rows of stride N.
struct HbPlainConvJob {
  const HbConvDev* cv;
  const u64* t; const u64* t_s;   // (Q/q_j)^-1 mod q_j (+Shoup), without N^-1
  u64 N;
  int nitems;
  const u64* src[HB_MAXB];
  u64* dst[HB_MAXB];
  double* frac[HB_MAXB];
  u64* stats;
};
__global__ void __launch_bounds__(HB_THREADS) k_conv_plain(const HbPrimeDev* __restrict__ primes, HbPlainConvJob J) {
  const HbConvDev* cv = J.cv;
  const int n = cv->n, nt = cv->nt;
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= J.N) return;
  const u64* src = J.src[blockIdx.y];
  u64* dst = J.dst[blockIdx.y];
  u64 y[HB_MAXROWS];
  for (int j = 0; j < n; j++) {
    const int pi = cv->src_prime[j];
    y[j] = hb_mul_shoup(src[(size_t)pi * J.N + k], J.t[j], J.t_s[j], primes[pi].q);
  }
  double f;
  double* fr = J.frac[blockIdx.y];
  const i64 v = hb_conv_v(cv, y, 1, J.stats, fr ? &f : nullptr);
  if (fr) fr[k] = f;
  for (int t = 0; t < nt; t++) {
    const int pi = cv->tgt_prime[t];
    const HbPrimeDev P = primes[pi];
    const u64* ct = cv->c + (size_t)t * n;
    u64 hi = 0, lo = 0;
    for (int j = 0; j < n; j++) hb_mac128(hi, lo, y[j], ct[j]);
    if (v >= 0) hb_mac128(hi, lo, (u64)v, cv->negQ[t]);
    else hb_mac128(hi, lo, (u64)(-v), cv->Qmod[t]);
    dst[(size_t)pi * J.N + k] = hb_reduce128(hi, lo, P);
  }
}

// Canonical-embedding norm for general m: basic_embeddingLargestCoeff (src/norms.cpp:129-157) -- max over i in Z_m^*, i <= m/2,
// of |sum_k f_k W^(ik)|, W = e^(2 pi I/m) -- evaluated directly in FP64 (phi(m)^2/2 multiply-adds per polynomial: ~1.3e8 for
// m = 21845, tens of microseconds of the B200's FP64 rate; no length-m FFT plan needed).  W is a host-built table of m entries.
struct HbGenNormJob { u64 m, phim; const double* frac; const double2* W; const int* rep; unsigned long long* maxbits; };
__global__ void __launch_bounds__(HB_THREADS) k_gen_norm(HbGenNormJob J) {
  HB_SMEM_DECL
  double* F = (double*)HB_SMEM;   // [1024] tile of the coefficients
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double* f = J.frac + (size_t)blockIdx.y * J.phim;
  const u64 i = t < J.phim ? (u64)J.rep[t] : 0;
  const bool live = t < J.phim && 2 * i <= J.m;   // i and m - i give conjugate values
  double re = 0, im = 0;
  u64 idx = 0;   // i*k mod m
  for (size_t k0 = 0; k0 < J.phim; k0 += 1024) {
    __syncthreads();
    for (size_t k = threadIdx.x; k < 1024; k += blockDim.x) F[k] = k0 + k < J.phim ? f[k0 + k] : 0.0;
    __syncthreads();
    if (live) {
      const size_t kn = J.phim - k0 < 1024 ? J.phim - k0 : 1024;
      for (size_t k = 0; k < kn; k++) {
        const double2 w = J.W[idx];
        re += F[k] * w.x; im += F[k] * w.y;
        idx += i; if (idx >= J.m) idx -= J.m;
      }
    }
  }
  double mx = live ? re * re + im * im : 0.0;
#ifdef HB_SIM
  unsigned long long bits; memcpy(&bits, &mx, 8);
  if (bits > J.maxbits[blockIdx.y]) J.maxbits[blockIdx.y] = bits;
#else
  atomicMax(J.maxbits + blockIdx.y, (unsigned long long)__double_as_longlong(mx));  // non-negative doubles order like integers
#endif
}

// automorphism for general m: new[j] = old[idx(rep(j)*k mod m)] (src/DoubleCRT.cpp:1160-1202)
struct HbGenAutoJob { u64 m, phim, k; const int* rep; const int* irep; HbRows rows; int nitems; const u64* src[HB_MAXB]; u64* dst[HB_MAXB]; };
__global__ void __launch_bounds__(HB_THREADS) k_gen_automorph(HbGenAutoJob J) {
  const int pi = J.rows.prime[blockIdx.y];
  const size_t off = (size_t)pi * J.phim;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < J.phim; j += (size_t)gridDim.x * blockDim.x) {
    const u64 r = ((u64)J.rep[j] * J.k) % J.m;   // m <= 2^20, so the product fits
    J.dst[blockIdx.z][off + j] = J.src[blockIdx.z][off + J.irep[r]];
  }
}

// ------------------------------------------------------------------------------------------
// Powerful basis (src/powerful.cpp) and Ctxt::rawModSwitch (src/Ctxt.cpp:2949-3046), SURVEY 8f-4.
// polyToPowerful is a Z-linear map with coefficients in {0, +-1}: scatter by the CRT index map into the
// (m_1 x ... x m_k) cube, reduce every hypercolumn of dimension d modulo Phi_{m_d}, read off the
// (phi(m_1) x ... x phi(m_k)) sub-cube.  Applied to the coefficient rows of every prime, then the exact CRT of the
// engine gives the balanced integers modulo Q -- PowerfulDCRT::dcrtToPowerful without leaving the device.
struct HbPwJob2 {
  u64 m, phim;
  HbRows rows;
  const int* cube_to_poly;    // [m]    exponent held by a cube cell
  const int* short_to_long;   // [phim] cube cell of a powerful coefficient
  const u64* src;             // coefficient rows [nprimes][phim]
  u64* cube;                  // scratch [nprimes][m]
  u64* dst;                   // powerful rows [nprimes][phim]
  // reduction pass of one dimension (m_d = p^e): stride = cells between consecutive coordinates, b = p^(e-1)
  u64 stride, md, p, b;
};
__global__ void __launch_bounds__(HB_THREADS) k_pw_scatter(const HbPwJob2 J) {
  const int pi = J.rows.prime[blockIdx.y];
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < J.m; j += (size_t)gridDim.x * blockDim.x) {
    const int i = J.cube_to_poly[j];
    J.cube[(size_t)pi * J.m + j] = (u64)i < J.phim ? J.src[(size_t)pi * J.phim + i] : 0;
  }
}
// column mod Phi_{p^e}(X) = sum_{j<p} X^(j*b): c[j*b + t] -= c[(p-1)*b + t] for j < p-1   (the top block is then dead)
__global__ void __launch_bounds__(HB_THREADS) k_pw_reduce(const HbPrimeDev* __restrict__ primes, const HbPwJob2 J) {
  const int pi = J.rows.prime[blockIdx.y];
  const u64 q = primes[pi].q;
  u64* cube = J.cube + (size_t)pi * J.m;
  const u64 ncol = J.m / J.md, work = ncol * J.b;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < work; w += (u64)gridDim.x * blockDim.x) {
    const u64 col = w / J.b, t = w - col * J.b;
    const u64 outer = col / J.stride, inner = col - outer * J.stride;    // column = all cells sharing the other coordinates
    u64* c = cube + outer * J.stride * J.md + inner;
    const u64 top = c[((J.p - 1) * J.b + t) * J.stride];
    for (u64 j = 0; j + 1 < J.p; j++) { u64* x = c + (j * J.b + t) * J.stride; *x = hb_submod(*x, top, q); }
  }
}
__global__ void __launch_bounds__(HB_THREADS) k_pw_gather(const HbPwJob2 J) {
  const int pi = J.rows.prime[blockIdx.y];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < J.phim; i += (size_t)gridDim.x * blockDim.x)
    J.dst[(size_t)pi * J.phim + i] = J.cube[(size_t)pi * J.m + J.short_to_long[i]];
}

// rawModSwitch of one coefficient vector (powerful basis): c balanced mod Q  ->  x = round(c*q/Q) + delta, reduced
// symmetrically mod q, where delta = bal(Y * Q^-1 mod p^r) and Y = c*q - round(c*q/Q)*Q (src/Ctxt.cpp:2990-3033).
// With y_j = c_j*(Q/q_j)^-1 mod q_j (so c = sum y_j Q_j - V0*Q) and y'_j = y_j*q mod q_j = y_j*q - k_j*q_j (so
// Y = sum y'_j Q_j - vY*Q):   round(c*q/Q) = sum_j k_j + vY - q*V0,   and hb_conv_v with the plaintext modulus
// returns vY + delta with exactly the reference's tie rule.  All quantities are exact integers.
struct HbRawMsJob {
  const HbConvDev* cv;
  const u64* t; const u64* t_s;      // (Q/q_j)^-1 mod q_j
  u64 N, q;
  const u64* src;                    // [nprimes][N] coefficient-like rows
  i64* out;                          // [N]
  u64* stats;
};
__global__ void __launch_bounds__(HB_THREADS) k_raw_mod_switch(const HbPrimeDev* __restrict__ primes, HbRawMsJob J) {
  const HbConvDev* cv = J.cv;
  const int n = cv->n;
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= J.N) return;
  u64 y[HB_MAXROWS];
  for (int j = 0; j < n; j++) {
    const int pi = cv->src_prime[j];
    y[j] = hb_mul_shoup(J.src[(size_t)pi * J.N + k], J.t[j], J.t_s[j], primes[pi].q);
  }
  const i64 v0 = hb_conv_v(cv, y, 1, J.stats, nullptr, false);
  i64 ksum = 0;
  for (int j = 0; j < n; j++) {
    const HbPrimeDev P = primes[cv->src_prime[j]];
    const u64 lo = y[j] * J.q, hi = __umul64hi(y[j], J.q);
    const u64 r = hb_reduce128(hi, lo, P);
    u64 inv = P.q;                                        // q_j^-1 mod 2^64 (Newton; q_j odd)
    for (int it = 0; it < 5; it++) inv *= 2 - P.q * inv;
    ksum += (i64)((lo - r) * inv);                        // exact quotient (y_j*q - r)/q_j < q
    y[j] = r;
  }
  const i64 v1 = hb_conv_v(cv, y, 1, J.stats, nullptr, true);
  i64 x = ksum + v1 - (i64)J.q * v0;
  const i64 qh = (i64)(J.q >> 1);
  if (x > qh) x -= (i64)J.q;                              // ties of an even q stay (the reference flips a coin there)
  else if (x < -qh) x += (i64)J.q;
  J.out[k] = x;
}
