// Microbenchmark: read bandwidth of a RESIDENT set (L2) and of a streaming set (HBM) on one B200, SURVEY.md 8(d): "L2
// bandwidth is not in MEASURED_PEAKS.json — measure it with a resident-set read kernel on the first GPU run".
//
// Every thread reads 16-byte vectors (LDG.E.128, L1 bypassed) with a grid-stride over a buffer of S bytes, R passes in one launch; for
// S well below the L2 capacity all passes after the first are served by L2.  The access pattern is the one of the BVH
// fetches of k_traverse (16-byte loads, read-only path), but fully coalesced: this is the ceiling, not what a gather gets.
// A second variant reads 16-byte vectors at hashed (incoherent) addresses of the same buffer, one 32-byte sector per
// thread: the gather ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_bw tools/micro/l2_bw.cu && ./l2_bw > profiles/r2_l2_bw.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                  \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__global__ void __launch_bounds__(512) k_read(const uint4 *__restrict__ buf, size_t n_vec, int passes, uint4 *sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; p++) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);  // ld.global.cg: cached in L2 only, so a small set is not served by L1
      acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
    }
  }
  if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;  // never true for the fill pattern; keeps the loads alive
}

// incoherent 16-byte gathers: thread t reads vector hash(t, k) for k = 0..per_thread
__global__ void __launch_bounds__(512) k_gather(const uint4 *__restrict__ buf, size_t n_vec_pow2, int per_thread, uint4 *sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  unsigned long long x = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1ull;
  const size_t mask = n_vec_pow2 - 1;
#pragma unroll 4
  for (int k = 0; k < per_thread; k++) {
    x ^= x >> 12, x ^= x << 25, x ^= x >> 27;  // xorshift64*
    const size_t i = (size_t)((x * 0x2545F4914F6CDD1Dull) >> 20) & mask;
    const uint4 v = __ldcg(buf + i);
    acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int l2 = 0;
  CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
  const size_t max_bytes = (size_t)2 << 30;
  uint4 *buf, *sink;
  CK(cudaMalloc(&buf, max_bytes));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(buf, 0x5a, max_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 4;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"grid\": \"%d blocks x 512 threads, 16-byte loads (ld.global.cg, L1 bypassed)\",\n \"coalesced_read\": [", prop.name,
         prop.multiProcessorCount, l2, blocks);
  const size_t sizes_mb[] = {4, 8, 16, 32, 48, 64, 96, 128, 256, 2048};
  bool first = true;
  for (size_t mb : sizes_mb) {
    const size_t bytes = mb << 20, n_vec = bytes / 16;
    const int passes = (int)((((size_t)32 << 30) / bytes) < 4 ? 4 : (((size_t)32 << 30) / bytes > 400 ? 400 : ((size_t)32 << 30) / bytes));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
      k_read<<<blocks, 512>>>(buf, n_vec, 1, sink);  // warm the set into L2 (untimed)
      CK(cudaEventRecord(e0));
      k_read<<<blocks, 512>>>(buf, n_vec, passes, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
      if (gbs > best) best = gbs;
    }
    printf("%s\n  {\"set_mib\": %zu, \"passes\": %d, \"gb_per_s\": %.1f}", first ? "" : ",", mb, passes, best);
    first = false;
  }
  printf("],\n \"gather16\": [");
  first = true;
  const size_t gsizes_mb[] = {1, 16, 64, 128, 2048};
  for (size_t mb : gsizes_mb) {
    const size_t bytes = mb << 20, n_vec = bytes / 16;
    const int per_thread = 2048;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
      k_read<<<blocks, 512>>>(buf, n_vec, 1, sink);
      CK(cudaEventRecord(e0));
      k_gather<<<blocks, 512>>>(buf, n_vec, per_thread, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double loads = (double)blocks * 512 * per_thread;
      const double g = loads * 16.0 / (ms * 1e-3) / 1e9;
      if (g > best) best = g;
    }
    printf("%s\n  {\"set_mib\": %zu, \"gb_per_s_useful\": %.1f, \"gb_per_s_sectors\": %.1f}", first ? "" : ",", mb, best, best * 2.0);
    first = false;
  }
  printf("]}\n");
  CK(cudaGetLastError());
  return 0;
}
