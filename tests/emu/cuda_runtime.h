// TEST INFRASTRUCTURE ONLY -- a single-header CUDA execution-model emulator for the host CPU.
//
// It exists so that the *unmodified* kernels of ramscb_b200/csrc can be compiled with g++ and run,
// thread for thread, in a GPU-less container: tests/emu/build_emu.py rewrites the `<<<...>>>` launch
// syntax and `extern __shared__` declarations of a scratch copy of the sources and compiles them
// against this header (which shadows <cuda_runtime.h>).  Nothing under ramscb_b200/ references this
// directory; the product library is built by nvcc only and has no CPU path.
//
// Model: a kernel launch runs its blocks one after the other; the threads of a block are fibers
// (hand-written x86-64 context switch) scheduled round-robin on one OS thread.  __syncthreads()
// yields until every live thread of the block has arrived; the warp shuffles exchange values
// through a per-warp buffer with the same arrive-then-release protocol, so warp-synchronous code
// behaves as on the device.  `__shared__` variables are function-local statics (one block is
// resident at a time).  Streams are synchronous; stream capture records closures and a graph launch
// replays them, with kernel arguments frozen at capture time like a CUDA graph.
// fma() is the correctly rounded std::fma and the build uses -ffp-contract=off, so the arithmetic
// is the device's IEEE arithmetic operation for operation (exp/pow come from libm, as in the oracle).
#pragma once
#if !defined(__x86_64__)
#error "the emulator's context switch is x86-64 only"
#endif
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <utility>
#include <vector>
#include <algorithm>

#define RSG_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; } __attribute__((aligned(16)));
struct double4 { double x, y, z, w; } __attribute__((aligned(16)));
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline double4 make_double4(double x, double y, double z, double w) { double4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

// the built-in variables: plain globals, threadIdx is set by the scheduler at every fiber switch
inline uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline dim3 blockDim, gridDim;

// ---------------------------------------------------------------------------------------------
// fibers
namespace emu {

struct Fiber {
  void* sp = nullptr;     // saved stack pointer
  char* stack = nullptr;
  bool done = true;
  int wait = 0;           // 0 runnable, 1 at __syncthreads, 2 at a warp exchange
  uint3 tid{0, 0, 0};
  int lin = 0;            // linear thread index
  unsigned shfl_phase = 0;
  unsigned or_phase = 0;
};

struct Warp {
  double buf[2][32];
  unsigned long long ubuf[2][32];
  int arrived = 0, live = 0;
};

constexpr size_t kStack = 256 * 1024;

struct Block {
  std::vector<Fiber> f;
  std::vector<Warp> w;
  int nthreads = 0, live = 0, at_barrier = 0;
  int or_slot[3] = {0, 0, 0};
  std::vector<int> perm;
  void* sched_sp = nullptr;
  std::function<void()> body;
};

inline Block& blk() { static Block b; return b; }
inline Fiber*& cur() { static Fiber* c = nullptr; return c; }
inline uint3& bidx() { return blockIdx; }
inline dim3& bdim() { return blockDim; }
inline dim3& gdim() { return gridDim; }
inline std::vector<char>& dynsmem() { static std::vector<char> v; return v; }
inline void* dyn_smem() { return dynsmem().data(); }

// switch stacks: push callee-saved registers, store rsp to *from, load rsp from to, pop, ret
extern "C" void emu_switch(void** from, void* to);
#ifdef EMU_IMPLEMENT
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  subq $8, %rsp
  stmxcsr (%rsp)
  fnstcw 4(%rsp)
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  ldmxcsr (%rsp)
  fldcw 4(%rsp)
  addq $8, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
#endif

inline void yield_to_sched() {
  Fiber* me = cur();
  emu_switch(&me->sp, blk().sched_sp);
}

extern "C" inline void emu_fiber_entry() {
  Block& b = blk();
  b.body();
  Fiber* me = cur();
  me->done = true;
  --b.live;
  --b.w[me->lin >> 5].live;
  yield_to_sched();
  abort();  // never resumed
}

inline void prepare(Fiber& f) {
  if (!f.stack) f.stack = (char*)aligned_alloc(64, kStack);
  // initial frame: [mxcsr/fpcw][r15][r14][r13][r12][rbx][rbp][ret = entry][pad]
  uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
  uint64_t* s = (uint64_t*)top;
  *--s = 0;                                // alignment pad: after `ret` rsp % 16 == 8 as at a call
  *--s = (uint64_t)(void*)&emu_fiber_entry;  // return address
  for (int i = 0; i < 6; ++i) *--s = 0;    // rbp rbx r12..r15
  uint32_t cw[2];
  asm volatile("stmxcsr %0" : "=m"(cw[0]));
  uint16_t fcw;
  asm volatile("fnstcw %0" : "=m"(fcw));
  cw[1] = fcw;
  --s;
  memcpy(s, cw, 8);
  f.sp = s;
  f.done = false;
  f.wait = 0;
  f.shfl_phase = 0;
  f.or_phase = 0;
}

// run one block to completion
inline void run_block(const std::function<void()>& body, dim3 block) {
  Block& b = blk();
  const int n = (int)(block.x * block.y * block.z);
  if ((int)b.f.size() < n) b.f.resize(n);
  b.w.assign((n + 31) / 32, Warp());
  b.nthreads = n;
  b.live = n;
  b.at_barrier = 0;
  b.or_slot[0] = b.or_slot[1] = b.or_slot[2] = 0;
  b.body = body;
  for (int t = 0; t < n; ++t) {
    Fiber& f = b.f[t];
    prepare(f);
    f.lin = t;
    f.tid.x = t % block.x;
    f.tid.y = (t / block.x) % block.y;
    f.tid.z = t / (block.x * block.y);
    ++b.w[t >> 5].live;
  }
  // EMU_ORDER=reverse|random runs the runnable threads of a block in another order between
  // synchronisation points: results must not depend on it (a difference means a missing barrier)
  static const int order_mode = [] {
    const char* e = getenv("EMU_ORDER");
    return !e ? 0 : (!strcmp(e, "reverse") ? 1 : (!strcmp(e, "random") ? 2 : 0));
  }();
  static unsigned long long rng = 0x9E3779B97F4A7C15ull;
  std::vector<int>& perm = b.perm;
  perm.resize(n);
  for (int t = 0; t < n; ++t) perm[t] = order_mode == 1 ? n - 1 - t : t;
  while (b.live > 0) {
    bool progressed = false;
    if (order_mode == 2)
      for (int t = n - 1; t > 0; --t) {             // Fisher-Yates with xorshift64
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(perm[t], perm[(int)(rng % (unsigned)(t + 1))]);
      }
    for (int q = 0; q < n; ++q) {
      const int t = perm[q];
      Fiber& f = b.f[t];
      if (f.done || f.wait) continue;
      cur() = &f;
      threadIdx = f.tid;
      emu_switch(&b.sched_sp, f.sp);
      progressed = true;
    }
    // release a completed block barrier (exited threads count as arrived)
    if (b.live > 0 && b.at_barrier == b.live) {
      for (int t = 0; t < n; ++t)
        if (b.f[t].wait == 1) b.f[t].wait = 0;
      b.at_barrier = 0;
      progressed = true;
    }
    for (size_t wi = 0; wi < b.w.size(); ++wi) {
      Warp& w = b.w[wi];
      if (w.live > 0 && w.arrived == w.live) {
        for (int t = (int)wi * 32; t < std::min(n, (int)wi * 32 + 32); ++t)
          if (b.f[t].wait == 2) b.f[t].wait = 0;
        w.arrived = 0;
        progressed = true;
      }
    }
    if (!progressed && b.live > 0) {
      fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d live, %d at barrier\n", bidx().x, bidx().y, bidx().z, b.live,
              b.at_barrier);
      abort();
    }
  }
  cur() = nullptr;
}

inline void syncthreads() {
  Block& b = blk();
  cur()->wait = 1;
  ++b.at_barrier;
  yield_to_sched();
}
inline void warp_arrive() {
  Block& b = blk();
  Fiber* me = cur();
  me->wait = 2;
  ++b.w[me->lin >> 5].arrived;
  yield_to_sched();
}

template <class T>
inline T shfl(T v, int src_lane, int width = 32) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
  Fiber* me = cur();
  Warp& w = blk().w[me->lin >> 5];
  const int lane = me->lin & 31;
  const unsigned ph = me->shfl_phase++ & 1;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.ubuf[ph][lane] = bits;
  warp_arrive();
  T r = v;
  if (src_lane >= 0 && src_lane < 32) {
    const int base = lane & ~(width - 1);
    const int s = base + (src_lane & (width - 1));
    const int nlanes = std::min(32, blk().nthreads - (me->lin & ~31));
    if (s < nlanes) memcpy(&r, &w.ubuf[ph][s], sizeof(T));
  }
  return r;
}

// ---- streams, events, graphs -------------------------------------------------------------
struct Stream { int unused = 0; };
struct Graph { std::vector<std::function<void()>> ops; };
inline long long& launch_counter() { static long long n = 0; return n; }

// capture: while any stream is capturing, every submission is recorded (the library forks and joins
// its other streams off the captured one with events, so they belong to the same graph)
inline std::vector<std::function<void()>>*& active_rec() { static std::vector<std::function<void()>>* r = nullptr; return r; }
inline void submit(void*, std::function<void()> op) {
  if (active_rec()) active_rec()->push_back(std::move(op));
  else op();
}

inline void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  ++launch_counter();
  gdim() = grid;
  bdim() = block;
  if (dynsmem().size() < smem + 64) dynsmem().resize(smem + 64);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        bidx() = uint3{x, y, z};
        run_block(body, block);
      }
}

// `body` runs one thread of the kernel; its captures are the launch arguments, copied at the launch
inline void launch(dim3 grid, dim3 block, size_t smem, void* stream, std::function<void()> body) {
  submit(stream, [=]() { run_grid(grid, block, smem, body); });
}

}  // namespace emu

#define warpSize 32

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_arrive(); }
static inline int __syncthreads_or(int pred) {
  // three rotating slots: call n clears the slot of call n+1 before its barrier, so nobody can
  // still be reading it (its last use was call n-2) and nobody can already be writing it
  emu::Block& b = emu::blk();
  const unsigned n = emu::cur()->or_phase++;
  b.or_slot[(n + 1) % 3] = 0;
  if (pred) b.or_slot[n % 3] = 1;
  emu::syncthreads();
  return b.or_slot[n % 3];
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { return emu::shfl(v, src, width); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  return emu::shfl(v, (emu::cur()->lin & 31) ^ m, 32);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = emu::cur()->lin & 31;
  return emu::shfl(v, lane - (int)d >= 0 ? lane - (int)d : -1, 32);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = emu::cur()->lin & 31;
  return emu::shfl(v, lane + (int)d < 32 ? lane + (int)d : -1, 32);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  const int lane = emu::cur()->lin & 31;
  for (int s = 0; s < 32; ++s) {   // 32 exchanges: slow but exact
    int p = emu::shfl(pred ? 1 : 0, s);
    (void)lane;
    if (p) r |= 1u << s;
  }
  return r;
}

// atomics (one block runs at a time)
template <class T> static inline T atomicAdd(T* a, T v) { T o = *a; *a = o + v; return o; }
template <class T> static inline T atomicMin(T* a, T v) { T o = *a; if (v < o) *a = v; return o; }
template <class T> static inline T atomicMax(T* a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> static inline T atomicExch(T* a, T v) { T o = *a; *a = v; return o; }
template <class T> static inline T atomicOr(T* a, T v) { T o = *a; *a = o | v; return o; }

// bit casts and device intrinsics
static inline int __double2hiint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) {
  unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double x; memcpy(&x, &b, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; memcpy(&x, &b, 8); return x; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// CUDA's global min/max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline double max(double a, double b) { return std::fmax(a, b); }
using std::abs;
using std::exp;
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::log;
using std::log10;
using std::pow;
using std::sqrt;
using std::floor;
using std::isnan;
using std::isfinite;

// ---------------------------------------------------------------------------------------------
// runtime API (synchronous host emulation)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100 };
typedef void* cudaStream_t;
typedef struct EmuEvent { double t; }* cudaEvent_t;
typedef emu::Graph* cudaGraph_t;
typedef emu::Graph* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0, cudaHostAllocMapped = 2, cudaHostAllocDefault = 0 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrL2CacheSize = 38 };
struct cudaDeviceProp {
  char name[256];
  size_t totalGlobalMem, sharedMemPerBlockOptin;
  int multiProcessorCount, l2CacheSize, major, minor;
};

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = attr == cudaDevAttrMultiProcessorCount ? 148 : attr == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448 : 0;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  strcpy(p->name, "emulated sm_100a (host CPU)");
  p->totalGlobalMem = (size_t)180 << 30;
  p->sharedMemPerBlockOptin = 232448;
  p->multiProcessorCount = 148;
  p->l2CacheSize = 126 << 20;
  p->major = 10;
  return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n, unsigned = 0) { *p = (T*)aligned_alloc(256, (n + 255) / 256 * 256 + 256); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = (T*)h; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st = nullptr) {
  emu::submit(st, [=]() { memmove(d, s, n); });
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind,
                                            cudaStream_t st = nullptr) {
  emu::submit(st, [=]() { for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w); });
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr) {
  emu::submit(st, [=]() { memset(d, v, n); });
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emu::Stream(); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emu::Stream(); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete (emu::Stream*)s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new EmuEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
  if (emu::active_rec()) return cudaErrorInvalidValue;
  emu::active_rec() = new std::vector<std::function<void()>>();
  return cudaSuccess;
}
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
  if (!emu::active_rec()) return cudaErrorInvalidValue;
  *g = new emu::Graph();
  (*g)->ops = std::move(*emu::active_rec());
  delete emu::active_rec();
  emu::active_rec() = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long = 0) {
  *e = new emu::Graph(*g);
  return cudaSuccess;
}
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, void*, void*, size_t) {
  *e = new emu::Graph(*g);
  return cudaSuccess;
}
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s) {
  for (auto& op : e->ops) emu::submit(s, op);
  return cudaSuccess;
}
// cluster launches are not emulated (cooperative_groups.h stub): refuse them loudly
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttribute {
  cudaLaunchAttributeID id;
  struct { struct { unsigned x, y, z; } clusterDim; } val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes = 0;
  cudaStream_t stream = nullptr;
  cudaLaunchAttribute* attrs = nullptr;
  unsigned numAttrs = 0;
};
template <class K, class... A> static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t*, K, A&&...) {
  fprintf(stderr, "emu: cluster launch requested (set RSG_SCB_NO_CLUSTER=1)\n");
  return cudaErrorInvalidValue;
}
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
