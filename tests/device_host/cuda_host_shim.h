// cuda_host_shim.h — TEST INFRASTRUCTURE, not product.  Lets g++ compile the device headers of
// raytrace3_b200/csrc (device_math.cuh, scene.cuh, scene_prep.h) for the host, so the CPU test
// suite can hold the PRODUCT's own arithmetic (not a restatement of it) to the oracle without a
// GPU: tests/test_device_host_cpu.py.  Nothing under raytrace3_b200/ includes this file and the
// library never runs these functions on the host.
//
// A warp is emulated by 32 fibers (ucontext) on ONE host thread, switched round-robin at every warp
// collective (__ballot_sync, __shfl_sync, __shfl_xor_sync, __any_sync, __syncwarp, __syncthreads):
// lane l runs until its next collective and then hands over to lane (l+1) % 32, so when lane 31
// arrives every lane has, and lane 0 resumes behind the barrier.  That is exact for code whose
// collectives are called convergently with the full mask, which is how scene.cuh calls all of its
// collectives (and what the hardware requires of __syncwarp(FULL) anyway); a divergent collective
// would deadlock or mis-pair here instead of passing silently.  Values are exchanged through a
// 32-slot table between two such barriers.  A block of several warps is the same thing one level up:
// __syncthreads hands over to the next thread of the block, and a warp is only ever left at a block
// barrier or when it has finished.  One block at a time per host thread; blocks run on several host
// threads.
#pragma once
#define SPPM_HOST_EMULATION 1

#include <cuda_runtime.h>
#include <math.h>
#include <ucontext.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>

using std::max;
using std::min;

// nvcc's __noinline__ keyword; empty (not an attribute) so that libstdc++'s own
// __attribute__((__noinline__)) still parses after this point
#ifndef __noinline__
#define __noinline__
#endif

// ---- thread coordinates and the fiber warp ------------------------------------------------
struct DhDim {
  unsigned x = 0, y = 0, z = 0;
};
constexpr int DH_MAX_THREADS = 320;  // the largest block of the product (dep_scatter_kernel)
struct DhWarp {  // one emulated thread block: n_threads fibers, warps of 32 consecutive threads
  ucontext_t main_ctx;
  ucontext_t lane_ctx[DH_MAX_THREADS];
  bool done[DH_MAX_THREADS];
  int cur = 0;         // the thread that is running (threadIdx.x)
  int n_threads = 32;  // blockDim.x, a multiple of 32
  unsigned block = 0, grid = 1;  // blockIdx.x / gridDim.x
  uint64_t slot[DH_MAX_THREADS];  // exchange table, 32 entries per warp
  unsigned long long barriers = 0;
};
extern thread_local DhWarp* dh_warp;  // null: a lone lane (collectives degenerate to one participant)
struct DhThreadIdx {
  struct X {
    operator unsigned() const { return dh_warp ? (unsigned)dh_warp->cur : 0u; }
  } x;
};
struct DhBlockDim {
  struct X {
    operator unsigned() const { return dh_warp ? (unsigned)dh_warp->n_threads : 1u; }
  } x;
};
static const DhThreadIdx dh_thread_idx{};
static const DhBlockDim dh_block_dim{};
struct DhBlockIdx {
  struct X {
    operator unsigned() const { return dh_warp ? dh_warp->block : 0u; }
  } x;
};
struct DhGridDim {
  struct X {
    operator unsigned() const { return dh_warp ? dh_warp->grid : 1u; }
  } x;
};
static const DhBlockIdx dh_block_idx{};
static const DhGridDim dh_grid_dim{};
#define threadIdx dh_thread_idx
#define blockDim dh_block_dim
#define blockIdx dh_block_idx
#define gridDim dh_grid_dim
// `extern __shared__ unsigned char smem[];` of a kernel = the emulating host thread's own array
// (one block at a time per host thread), defined in device_host.cpp
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#undef __shared__
#define __shared__ thread_local

// warp collective: hand over to the next lane of the same warp; when lane 31 arrives every lane of the
// warp has, and lane 0 resumes behind the barrier
inline void dh_barrier() {
  DhWarp* w = dh_warp;
  if (!w) return;
  const int me = w->cur, next = (me & ~31) | ((me + 1) & 31);
  w->barriers += (me == 0);
  w->cur = next;
  swapcontext(&w->lane_ctx[me], &w->lane_ctx[next]);
}
// __syncthreads: hand over to the next thread of the block (a warp is only ever left at a block
// barrier or when it has finished, so the other warps are parked at the previous block barrier or
// have not started); when the last thread arrives thread 0 resumes
inline void dh_block_barrier() {
  DhWarp* w = dh_warp;
  if (!w) return;
  const int me = w->cur;
  int next = me;
  do next = (next + 1) % w->n_threads; while (w->done[next] && next != me);  // threads that exited count as arrived
  if (next == me) return;
  w->cur = next;
  swapcontext(&w->lane_ctx[me], &w->lane_ctx[next]);
}
template <class T>
inline T dh_exchange(T v, int src) {  // every lane publishes v, then reads lane src's value
  static_assert(sizeof(T) <= 8, "exchange slot is 64 bits");
  if (!dh_warp) return v;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  const int me = dh_warp->cur;
  dh_warp->slot[me] = bits;
  dh_barrier();
  const uint64_t got = dh_warp->slot[(me & ~31) | (src & 31)];
  dh_barrier();  // the table may be rewritten after everyone has read
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
inline void __syncwarp() { dh_barrier(); }
inline void __syncthreads() { dh_block_barrier(); }
inline unsigned __ballot_sync(unsigned, bool p) {
  if (!dh_warp) return p ? 1u : 0u;
  const int me = dh_warp->cur;
  dh_warp->slot[me] = p ? 1 : 0;
  dh_barrier();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)dh_warp->slot[(me & ~31) | l] << l;
  dh_barrier();
  return m;
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0; }
inline bool __all_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) == (dh_warp ? 0xffffffffu : 1u); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return dh_exchange(v, src); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta) {  // lanes below delta keep their own value
  const int me = dh_warp ? dh_warp->cur & 31 : 0;
  return dh_exchange(v, me >= delta ? me - delta : me);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int off) { return dh_exchange(v, (dh_warp ? dh_warp->cur & 31 : 0) ^ off); }

// ---- scalar intrinsics --------------------------------------------------------------------
template <class T>
inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline double __longlong_as_double(long long v) {
  double d;
  memcpy(&d, &v, 8);
  return d;
}
inline long long __double_as_longlong(double d) {
  long long v;
  memcpy(&v, &d, 8);
  return v;
}
inline float __double2float_rn(double d) { return (float)d; }
inline float __double2float_ru(double d) {  // round towards +inf
  float f = (float)d;
  return (double)f < d ? nextafterf(f, INFINITY) : f;
}
inline float __double2float_rd(double d) {  // round towards -inf
  float f = (float)d;
  return (double)f > d ? nextafterf(f, -INFINITY) : f;
}
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline void __trap() { abort(); }
// __fdividef has <= 2 ulp error on the device; the exact quotient lies inside that band
inline float __fdividef(float a, float b) { return a / b; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline double atomicAdd(double* p, double v) {
  uint64_t* q = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED);
  for (;;) {
    double o, nw;
    memcpy(&o, &old, 8);
    nw = o + v;
    uint64_t nb;
    memcpy(&nb, &nw, 8);
    if (__atomic_compare_exchange_n(q, &old, nb, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return o;
  }
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline int __double2int_rd(double d) { return (int)floor(d); }

// cooperative_groups::coalesced_threads() returns whichever lanes of the warp happen to be
// converged at the call; the lanes here run one at a time between collectives, so the group is the
// calling lane alone -- one of the groupings the hardware may form
namespace cooperative_groups {
struct coalesced_group {
  unsigned thread_rank() const { return 0; }
  unsigned size() const { return 1; }
  template <class T>
  T shfl(T v, int) const { return v; }
};
inline coalesced_group coalesced_threads() { return coalesced_group(); }
}  // namespace cooperative_groups
// stand-in for rcp.approx.ftz.f64 (scene.cuh rcp_newton): a reciprocal good to ~20 bits whose
// low mantissa word is zero, as PTX specifies; the two Newton-Raphson steps that follow it are
// the product's own
inline double dh_rcp_approx(double x) {
  long long b = __double_as_longlong(1.0 / x);
  b &= (long long)0xFFFFFFFF00000000ull;
  return __longlong_as_double(b);
}
