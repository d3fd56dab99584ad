// Micro-benchmark: issue throughput of the integer / FP instructions the engine's kernels are built from,
// alone and in the mixes that matter (IMAD-class next to IADD3-class), on one B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
// Output: warp-instructions per clock per SM for each stream (4.0 = one per cycle per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned u32;

#define CH 8   // independent chains per thread
enum { OP_IMAD, OP_IMAD_WIDE, OP_IMAD_HI, OP_IADD3, OP_IADDX, OP_LOP3, OP_SHF, OP_CSUB, OP_FFMA, OP_DFMA,
       OP_MIX_IMAD_IADD, OP_MIX_WIDE_IADD, OP_MIX_WIDE_2IADD, OP_MIX_WIDE_FFMA, OP_MIX_WIDE_DFMA, OP_MIX_HI_IADD,
       OP_MUL64LO, OP_MUL64HI, OP_IMADMOV, OP_MIX_IMAD_WIDE, OP_NOPS };
static const char* NAMES[] = {"IMAD (mad.lo.u32)", "IMAD.WIDE.U32 (mad.wide.u32)", "IMAD.HI.U32 (mad.hi.u32)", "IADD3 (add.u32 x2 fused)", "IADD3+IADD3.X (add.cc/addc)",
  "LOP3 (xor/and)", "SHF (funnel shift)", "ISETP + 2 predicated IADD3 (csub)", "FFMA", "DFMA",
  "mix IMAD : IADD3 = 1:1", "mix IMAD.WIDE : IADD3 = 1:1", "mix IMAD.WIDE : IADD3 = 1:2", "mix IMAD.WIDE : FFMA = 1:1", "mix IMAD.WIDE : DFMA = 1:1", "mix IMAD.HI : IADD3 = 1:1",
  "mul.lo.u64 (compound)", "mul.hi.u64 (compound)", "IMAD.MOV-like (mad.lo x,1,0)", "mix IMAD : IMAD.WIDE = 1:1"};
// instructions counted per inner step and chain for each stream
static const int PER[] = {1, 1, 1, 1, 2, 1, 1, 3, 1, 1, 2, 2, 3, 2, 2, 2, 1, 1, 1, 2};

template <int OP>
__global__ void __launch_bounds__(256) k(u64* out, u32 seed, int rounds) {
  u32 a[CH], b[CH], c[CH]; u64 w[CH]; float f[CH]; double d[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { a[i] = seed * (threadIdx.x + 1) + i; b[i] = a[i] * 2654435761u + 1; c[i] = b[i] ^ 0x9e3779b9u; w[i] = ((u64)a[i] << 32) | b[i]; f[i] = (float)(a[i] & 1023) * 1e-3f; d[i] = (double)(b[i] & 1023) * 1e-3; }
  for (int r = 0; r < rounds; r++) {
#pragma unroll
    for (int s = 0; s < 8; s++) {
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
        if (OP == OP_IMAD_WIDE) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i]));
        if (OP == OP_IMAD_HI) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i]));
        if (OP == OP_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
        if (OP == OP_IADDX) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(seed));
        if (OP == OP_LOP3) asm volatile("{.reg .u32 t; xor.b32 t, %0, %1; and.b32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
        if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
        if (OP == OP_CSUB) asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %3; @p sub.cc.u32 %0, %0, %2; @p subc.u32 %1, %1, %3;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(seed));
        if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % CH]));
        if (OP == OP_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
        if (OP == OP_MIX_IMAD_IADD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(b[i]), "r"(seed)); }
        if (OP == OP_MIX_WIDE_IADD) { asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i])); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(b[i]), "r"(seed)); }
        if (OP == OP_MIX_WIDE_2IADD) { asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i])); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(b[i]), "r"(seed)); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(seed)); }
        if (OP == OP_MIX_WIDE_FFMA) { asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i])); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % CH])); }
        if (OP == OP_MIX_WIDE_DFMA) { asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i])); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH])); }
        if (OP == OP_MIX_HI_IADD) { asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i])); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(b[i]), "r"(seed)); }
        if (OP == OP_MUL64LO) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CH] | 1));
        if (OP == OP_MUL64HI) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CH] | 0x8000000000000001ull));
        if (OP == OP_IMADMOV) asm volatile("mad.lo.u32 %0, %1, 1, 0;" : "=r"(a[i]) : "r"(a[(i + 1) % CH]));
        if (OP == OP_MIX_IMAD_WIDE) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w[i]) : "r"(b[i])); }
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= a[i] ^ ((u64)b[i] << 7) ^ c[i] ^ w[i] ^ (u64)__float_as_uint(f[i]) ^ (u64)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> static void run(u64* out, int sms, double clk_hz) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int rounds = 4000;
  for (int wps : {8, 16}) {   // warps per SM sub-partition = blocks/SM * 8 / 4
    int blocks = sms * wps / 2;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); k<OP><<<blocks, 256>>>(out, 12345u + rep, rounds); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double winstr = (double)blocks * 8 * rounds * 8 * CH * PER[OP];
    printf("%-44s warps/SMSP %2d: %8.3f ms  %6.3f warp-instr/clk/SM (counted), %6.2f Gop-steps/s\n", NAMES[OP], wps / 4 * 2, best,
           winstr / (best * 1e-3) / clk_hz / sms, (double)blocks * 256 * rounds * 8 * CH / (best * 1e-3) * 1e-9);
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double clk = clk_khz * 1e3;
  printf("%s, %d SMs, clock %.0f MHz (attribute; results assume it)\n", p.name, p.multiProcessorCount, clk * 1e-6);
  u64* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 16 * 256 * 8);
  const int sms = p.multiProcessorCount;
  run<OP_IMAD>(out, sms, clk); run<OP_IMAD_WIDE>(out, sms, clk); run<OP_IMAD_HI>(out, sms, clk); run<OP_IADD3>(out, sms, clk); run<OP_IADDX>(out, sms, clk);
  run<OP_LOP3>(out, sms, clk); run<OP_SHF>(out, sms, clk); run<OP_CSUB>(out, sms, clk); run<OP_FFMA>(out, sms, clk); run<OP_DFMA>(out, sms, clk);
  run<OP_MIX_IMAD_IADD>(out, sms, clk); run<OP_MIX_WIDE_IADD>(out, sms, clk); run<OP_MIX_WIDE_2IADD>(out, sms, clk); run<OP_MIX_WIDE_FFMA>(out, sms, clk); run<OP_MIX_WIDE_DFMA>(out, sms, clk);
  run<OP_MIX_HI_IADD>(out, sms, clk); run<OP_MUL64LO>(out, sms, clk); run<OP_MUL64HI>(out, sms, clk); run<OP_IMADMOV>(out, sms, clk); run<OP_MIX_IMAD_WIDE>(out, sms, clk);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
