// pt_philox.h — counter-based RNG that replaces the reference's per-row StdRng (src/renderer.rs:91).
//
// Philox4x32-10 (Salmon et al., SC'11).  key = 64-bit render seed; counter = (pixel, sample, bounce, block).
// One block = four 24-bit uniforms, which covers every Material::scatter (the greediest, Plastic, draws 1 + 2).
// bounce = 0xffffffff is the camera jitter (renderer.rs:96-97: u first, then v).
#pragma once
#include "pt_hd.h"

namespace pt {

struct U4 {
  uint32_t x, y, z, w;
};

PT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

PT_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// rand's StandardUniform for f32 (24 high bits * 2^-24), the distribution of rng.random::<f32>()
PT_HD float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

struct Uniforms4 {
  float u[4];
};
PT_HD Uniforms4 philox_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block) {
  U4 c{pixel, sample, bounce, block};
  U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  Uniforms4 o;
  o.u[0] = u32_to_unit(r.x);
  o.u[1] = u32_to_unit(r.y);
  o.u[2] = u32_to_unit(r.z);
  o.u[3] = u32_to_unit(r.w);
  return o;
}

}  // namespace pt
