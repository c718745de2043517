// shared-memory read bandwidth by access width (conflict-free patterns), B200
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int V>
__global__ void lds(float* out) {
  __shared__ __align__(16) float s[12288];
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) s[i] = i * 0.001f;
  __syncthreads();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int base = threadIdx.x * V;   // conflict-free: consecutive threads, consecutive vectors
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    const int off = (i & 7) * V * 4;  // small moving offset, keeps addresses in range and aligned
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* p = &s[(base + off + u * 1024 * V / 4 * 1) % (12288 - 4 * V)];
      if (V == 4) { float4 v = *reinterpret_cast<const float4*>(&s[((base + off) + u * 1024) & 8191]); a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
      else if (V == 2) { float2 v = *reinterpret_cast<const float2*>(&s[((base + off) + u * 1024) & 8191]); a0 += v.x; a1 += v.y; }
      else { a0 += s[((base + off) + u * 1024) & 8191]; }
      (void)p;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int bs : {256, 512, 1024}) {
    int grid = 148 * (2048 / bs);
    float t1 = timeit([&] { lds<1><<<grid, bs>>>(out); });
    float t2 = timeit([&] { lds<2><<<grid, bs>>>(out); });
    float t4 = timeit([&] { lds<4><<<grid, bs>>>(out); });
    double n = (double)grid * bs * ITERS * 8;
    printf("bs=%4d  LDS.32 %.1f  LDS.64 %.1f  LDS.128 %.1f  B/clk/SM (nominal %d MHz)\n", bs, n * 4 / t1 / 1e3 / 148 / (clk / 1e3),
           n * 8 / t2 / 1e3 / 148 / (clk / 1e3), n * 16 / t4 / 1e3 / 148 / (clk / 1e3), clk / 1000);
  }
  return 0;
}
