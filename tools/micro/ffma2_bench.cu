// Microbenchmark: issue rate of scalar FFMA / FMUL+FADD against the packed fma.rn.f32x2 / mul / add of sm_100a, alone
// and mixed with integer ALU work (the extend kernels are issue-bound with FMA and ALU pipes each ~35-55 % busy).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ffma2_bench tools/micro/ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float &rx, float &ry, float ax, float ay, float bx, float by, float cx, float cy) {
  asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
               : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
}
__device__ __forceinline__ void mul2(float &rx, float &ry, float ax, float ay, float bx, float by) {
  asm volatile("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
               : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ void add2(float &rx, float &ry, float ax, float ay, float bx, float by) {
  asm volatile("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd; }"
               : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}

// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 even under --fmad=false (seen in the SASS of the first
// version of this file), which changes the rounding.  Unfused packed arithmetic therefore has to be spelled as two
// FMAs the assembler cannot merge: x*a + (-0.0) is RN(x*a) exactly and y*1.0 + b is RN(y+b) exactly; the constants
// arrive as kernel arguments so that nothing folds them.
template <int MODE>
__global__ void k(float *out, int iters, float a, float b, float negzero, float one) {
  float x[8], y[8];
  unsigned u[4];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i, y[i] = x[i] + 0.5f;
  for (int i = 0; i < 4; i++) u[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) {  // scalar: mul then add, 16 flop-pairs = 32 instructions
        x[i] = x[i] * a;
        x[i] = x[i] + b;
        y[i] = y[i] * a;
        y[i] = y[i] + b;
      } else if (MODE == 1) {  // packed unfused mul, packed unfused add: 16 FFMA2
        fma2(x[i], y[i], x[i], y[i], a, a, negzero, negzero);
        fma2(x[i], y[i], x[i], y[i], one, one, b, b);
      } else if (MODE == 2) {  // scalar FFMA
        x[i] = fmaf(x[i], a, b);
        y[i] = fmaf(y[i], a, b);
      } else if (MODE == 3) {  // packed FFMA2
        fma2(x[i], y[i], x[i], y[i], a, a, b, b);
      } else if (MODE == 4) {  // scalar mul/add + as many integer ops
        x[i] = x[i] * a;
        x[i] = x[i] + b;
        y[i] = y[i] * a;
        y[i] = y[i] + b;
        u[i & 3] = (u[i & 3] ^ (u[(i + 1) & 3] >> 3)) + 0x9E3779B9u;
        u[(i + 2) & 3] = (u[(i + 2) & 3] & 0xff00ffu) | (u[i & 3] << 5);
      } else {  // packed unfused mul/add + the same integer ops
        fma2(x[i], y[i], x[i], y[i], a, a, negzero, negzero);
        fma2(x[i], y[i], x[i], y[i], one, one, b, b);
        u[i & 3] = (u[i & 3] ^ (u[(i + 1) & 3] >> 3)) + 0x9E3779B9u;
        u[(i + 2) & 3] = (u[(i + 2) & 3] & 0xff00ffu) | (u[i & 3] << 5);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
}

template <int MODE>
float run(float *d, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.0001f, -0.0f, 1.0f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.0001f, -0.0f, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float *d;
  cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const char *names[] = {"scalar FMUL+FADD (32 instr / iter)", "packed unfused mul + add as 2 FFMA2 (16 instr)", "scalar FFMA (16 instr)",
                         "packed fma.f32x2 (8 instr)", "scalar FMUL+FADD + 32 integer ops", "packed mul/add + 32 integer ops"};
  float t[6] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters)};
  for (int m = 0; m < 6; m++) {
    const double flops = 148.0 * 8 * 256 * iters * 8 * 4;  // mul+add on x and y = 4 flop per i
    printf("%-44s %8.3f ms  %7.2f Tflop/s (mul and add counted)\n", names[m], t[m], flops / t[m] / 1e9);
  }
  return 0;
}
