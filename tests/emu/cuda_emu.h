// TEST INFRASTRUCTURE -- not part of the product, never shipped, never a fallback.
// Minimal host stand-ins for the CUDA constructs the GENERATED BK1 kernel text uses, so that the emitter's output
// (reaction schedule, slot allocation, live-range splitting, constant folding ...) can be executed thread by
// thread on the CPU and compared with the oracle in the `-m "not gpu"` tests (tests/emu/emulate.py).  Threads of
// the generated kernels never communicate (one thread = one state), so running them one after another to
// completion is exact; __syncthreads() is a no-op.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __constant__ static const
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define CUDART_INF (__builtin_inf())

struct EmuDim3 {
  unsigned x = 0, y = 0, z = 0;
};
static EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct double2 {
  double x, y;
};
struct float2 {
  float x, y;
};

using std::max;
using std::min;

static inline int __double2hiint(double d)
{
  uint64_t b;
  memcpy(&b, &d, 8);
  return (int)(uint32_t)(b >> 32);
}
static inline int __double2loint(double d)
{
  uint64_t b;
  memcpy(&b, &d, 8);
  return (int)(uint32_t)(b & 0xffffffffu);
}
static inline double __hiloint2double(int hi, int lo)
{
  const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  memcpy(&d, &b, 8);
  return d;
}
template <class T>
static inline T __ldg(const T* p)
{
  return *p;
}
static inline void __syncthreads() {}

// stand-in for MUFU.RCP64H: a reciprocal good to ~2^-20 only (the hardware seed's measured error is 1e-6), so the
// Newton steps of kx_rcp are really exercised
static inline double kx_emu_rcp_seed(double a)
{
  const double r = 1.0 / a;
  return __hiloint2double(__double2hiint(r), 0);
}
