// cusim.h -- a tiny CUDA *kernel-logic simulator* for unit tests.
//
// TEST INFRASTRUCTURE ONLY.  It lets the engine's .cu sources be compiled with g++
// (-DHB_SIM) so the kernels' index arithmetic and modular arithmetic can be checked against
// the oracle on a machine without a GPU.  Each CTA's threads run as ucontext fibers on one OS
// thread; __syncthreads() yields to the fiber scheduler.  It is never built into, loaded by or
// linked with the product library (helib_b200/libhelib_b200.so), which requires a real device.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 r; r.x = a; r.y = b; return r; }

namespace cusim {
extern dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
extern char* smem_;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void bar_sync(int id, int count);  // named barrier (PTX bar.sync id, count)
void syncwarp();                   // full-warp __syncwarp()
void yield();                      // let the other threads of the CTA run (spin waits)
}  // namespace cusim
#define __syncwarp() cusim::syncwarp()

#define threadIdx cusim::threadIdx_
#define blockIdx cusim::blockIdx_
#define blockDim cusim::blockDim_
#define gridDim cusim::gridDim_
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __syncthreads() cusim::syncthreads()
#define __ldg(p) (*(p))

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; *p += v; return o;
}

// ---- runtime shims -------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
#define cudaSuccess 0
#define cudaMemcpyHostToDevice 1
#define cudaMemcpyDeviceToHost 2
#define cudaMemcpyDeviceToDevice 3
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
static inline const char* cudaGetErrorString(cudaError_t) { return "sim"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define HB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cusim::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })

#ifdef CUSIM_IMPLEMENTATION
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
namespace cusim {
static void segv_handler(int sig) {
  void* bt[64];
  int n = backtrace(bt, 64);
  const char msg[] = "cusim: fatal signal, backtrace:\n";
  if (write(2, msg, sizeof(msg) - 1)) {}
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
__attribute__((constructor)) static void install_handler() {
  if (getenv("CUSIM_BACKTRACE")) { signal(SIGSEGV, segv_handler); signal(SIGBUS, segv_handler); }
}
dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
char* smem_ = nullptr;

struct Fiber {
  ucontext_t ctx;
  char* stack;
  bool done;
};
static ucontext_t sched_ctx;
static std::vector<Fiber> fibers;
static int cur_fiber = -1;
static const std::function<void()>* cur_body = nullptr;
static const size_t STACK = 256 * 1024;

static void fiber_entry() {
  (*cur_body)();
  fibers[cur_fiber].done = true;
  swapcontext(&fibers[cur_fiber].ctx, &sched_ctx);
}

static int bar_arrived[80], bar_gen[80];   // 0..15: CTA-level named barriers; 16 + w: __syncwarp of warp w
static unsigned cur_nthr = 0;

// counter/generation barriers: a fiber arrives, then yields until the generation advances
void bar_sync(int id, int count) {
  int me = cur_fiber;
  int gen = bar_gen[id];
  if (++bar_arrived[id] == count) { bar_arrived[id] = 0; bar_gen[id]++; return; }
  while (bar_gen[id] == gen) swapcontext(&fibers[me].ctx, &sched_ctx);
}
void syncthreads() { bar_sync(0, (int)cur_nthr); }
// warp-level barrier (all 32 lanes; the kernels only use full-warp __syncwarp)
void syncwarp() { bar_sync(16 + cur_fiber / 32, 32); }
// give the other fibers of the CTA a turn (spin-wait loops: mbarrier waits)
void yield() { int me = cur_fiber; swapcontext(&fibers[me].ctx, &sched_ctx); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  unsigned nthr = block.x * block.y * block.z;
  std::vector<char> shared(smem + 64);
  if (fibers.size() < nthr) {
    size_t old = fibers.size();
    fibers.resize(nthr);
    for (size_t i = old; i < nthr; i++) fibers[i].stack = (char*)malloc(STACK);
  }
  gridDim_ = grid; blockDim_ = block;
  cur_body = &body;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        smem_ = shared.data();
        cur_nthr = nthr;
        for (int b = 0; b < 80; b++) { bar_arrived[b] = 0; bar_gen[b] = 0; }
        for (unsigned t = 0; t < nthr; t++) {
          getcontext(&fibers[t].ctx);
          fibers[t].ctx.uc_stack.ss_sp = fibers[t].stack;
          fibers[t].ctx.uc_stack.ss_size = STACK;
          fibers[t].ctx.uc_link = &sched_ctx;
          fibers[t].done = false;
          makecontext(&fibers[t].ctx, fiber_entry, 0);
        }
        unsigned remaining = nthr;
        while (remaining) {
          remaining = 0;
          for (unsigned t = 0; t < nthr; t++) {
            if (fibers[t].done) continue;
            cur_fiber = (int)t;
            blockIdx_ = dim3(bx, by, bz);
            threadIdx_ = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            swapcontext(&sched_ctx, &fibers[t].ctx);
            if (!fibers[t].done) remaining++;
          }
        }
      }
  cur_body = nullptr;
}
}  // namespace cusim
#endif
