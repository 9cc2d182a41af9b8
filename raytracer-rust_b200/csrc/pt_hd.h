// pt_hd.h — one set of device functions, two compilers.
//
// Every intersection / BSDF / RNG function of the core is written once as PT_HD inline code.  nvcc compiles
// it for sm_100a (the product, libptcore.so).  g++ compiles the very same headers into tests/hostsim (a
// TEST-ONLY harness, never loaded by the product) so the traversal and shading logic can be checked against
// the oracle in a container that has no GPU.
//
// Floating-point contract: the reference is rustc without target-cpu flags, i.e. every fp32 operation rounds
// once and a*b+c is never fused.  The CUDA translation unit is therefore built with --fmad=false and the host
// one with -ffp-contract=off; wherever fusing is harmless (conservative box tests) the code says fmaf()
// explicitly, which both compilers honour.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__
#else
#define PT_HD inline
#define PT_D inline
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
#endif

namespace pt {

PT_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
PT_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
PT_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
PT_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
// read-only 16-byte load (LDG.E.128.CONSTANT on the device)
PT_HD float4 ldg4(const float4 *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
// plain 16-byte load: the object table may have been staged into shared memory, where ld.global.nc must not be used
PT_HD float4 ld4(const float4 *p) { return *p; }
PT_HD bool isnan_f(float v) { return v != v; }
PT_HD bool isinf_f(float v) { return fabsf(v) == INFINITY; }

}  // namespace pt
