// Stand-in for <cuda_runtime.h> when the device headers of pyshocks_b200/csrc are compiled for
// the HOST by g++ (tests/host/fast_kernels_host.cpp): the CUDA keywords become nothing, one warp
// is 32 OS threads and a warp shuffle is an exchange through a 32-slot array between two
// barrier waits.  Test infrastructure only -- the product never includes this file.
#pragma once

#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#define PSK_HOST_EMU 1
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static  // one process-wide array; every emulated thread touches its own slots

typedef int cudaError_t;
constexpr int cudaSuccess = 0;
typedef void *cudaStream_t;

struct emu_uint3 {
  unsigned x, y, z;
};
struct double2 {
  double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {
struct Warp {
  pthread_barrier_t bar;
  unsigned long long slot[32];
};
extern thread_local emu_uint3 tid, bid, bdim, gdim;
extern thread_local Warp *warp;
extern thread_local int lane;

// every lane of the warp calls this at the same program point (full-mask shuffles only)
template <class T>
inline T exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  warp->slot[lane] = bits;
  pthread_barrier_wait(&warp->bar);
  bits = warp->slot[(src < 0 || src > 31) ? lane : src];
  pthread_barrier_wait(&warp->bar);
  T r;
  std::memcpy(&r, &bits, sizeof(T));
  return r;
}
}  // namespace emu

#define threadIdx emu::tid
#define blockIdx emu::bid
#define blockDim emu::bdim
#define gridDim emu::gdim

template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) { return emu::exchange(v, emu::lane - d); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, int d) { return emu::exchange(v, emu::lane + d); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::exchange(v, emu::lane ^ m); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, src & 31); }

inline unsigned long long atomicMax(unsigned long long *addr, unsigned long long v) {
  unsigned long long old = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(addr, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}

inline unsigned atomicAdd(unsigned *addr, unsigned v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned *addr, unsigned v) {
  unsigned old = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(addr, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}
inline double atomicAdd(double *addr, double v) {  // only one lane per warp calls it in the kernels under test
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED), want;
  double cur;
  do {
    std::memcpy(&cur, &old, 8);
    const double sum = cur + v;
    std::memcpy(&want, &sum, 8);
  } while (!__atomic_compare_exchange_n(a, &old, want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return cur;
}

inline long long __double_as_longlong(double x) {
  long long r;
  std::memcpy(&r, &x, 8);
  return r;
}
inline double __longlong_as_double(long long x) {
  double r;
  std::memcpy(&r, &x, 8);
  return r;
}
inline int __double2hiint(double x) { return static_cast<int>(__double_as_longlong(x) >> 32); }
inline double __hiloint2double(int hi, int lo) {
  const unsigned long long b = (static_cast<unsigned long long>(static_cast<unsigned>(hi)) << 32) | static_cast<unsigned>(lo);
  return __longlong_as_double(static_cast<long long>(b));
}
// compiled with -ffp-contract=off: the plain operators round once, like the _rn intrinsics
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }

// the seed MUFU.RCP64H delivers to fast_rcp: a reciprocal good to about 2^-23 (the low 29 bits of
// the exact quotient are dropped)
inline double psk_emu_rcp_seed(double x) {
  long long b = __double_as_longlong(1.0 / x);
  b &= ~((1ll << 29) - 1);
  return __longlong_as_double(b);
}

using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
