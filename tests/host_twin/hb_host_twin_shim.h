// TEST INFRASTRUCTURE. What hb_device.cuh / hb_tables.h need from the CUDA toolkit, for g++: the trace arithmetic of the
// product is compiled for the CPU from the very source nvcc compiles (-DHB_HOST_TWIN, tests/host_twin/hb_host_twin.cpp)
// and checked against the reference's golden vectors without a GPU. Build with -ffp-contract=off -frounding-math.
#ifndef HB_HOST_TWIN_SHIM_H_
#define HB_HOST_TWIN_SHIM_H_
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__

struct float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }

using std::max;
using std::min;

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmaf_rd(float a, float b, float c) {  // fma.rm: one rounding, toward -infinity
  const int mode = fegetround();
  fesetround(FE_DOWNWARD);
  volatile float va = a, vb = b, vc = c;
  const float r = fmaf(va, vb, vc);
  fesetround(mode);
  return r;
}
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }

#endif  // HB_HOST_TWIN_SHIM_H_
