// Micro-benchmarks of the issue-rate facts the MED kernel design depends on (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b) {
  float2 x0 = make_float2(threadIdx.x, 1.f), x1 = make_float2(2.f, 3.f), x2 = make_float2(4.f, 5.f), x3 = make_float2(6.f, 7.f);
  float2 x4 = x0, x5 = x1, x6 = x2, x7 = x3;
  const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {  // scalar FFMA, 16 per iter (8 x float2 as 16 scalars)
#pragma unroll
      for (int u = 0; u < 1; ++u) {
        x0.x = fmaf(x0.x, a, b); x0.y = fmaf(x0.y, a, b); x1.x = fmaf(x1.x, a, b); x1.y = fmaf(x1.y, a, b);
        x2.x = fmaf(x2.x, a, b); x2.y = fmaf(x2.y, a, b); x3.x = fmaf(x3.x, a, b); x3.y = fmaf(x3.y, a, b);
        x4.x = fmaf(x4.x, a, b); x4.y = fmaf(x4.y, a, b); x5.x = fmaf(x5.x, a, b); x5.y = fmaf(x5.y, a, b);
        x6.x = fmaf(x6.x, a, b); x6.y = fmaf(x6.y, a, b); x7.x = fmaf(x7.x, a, b); x7.y = fmaf(x7.y, a, b);
      }
    } else if (MODE == 1) {  // packed FFMA2, 8 per iter = 16 flop-lanes
      x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B); x2 = __ffma2_rn(x2, A, B); x3 = __ffma2_rn(x3, A, B);
      x4 = __ffma2_rn(x4, A, B); x5 = __ffma2_rn(x5, A, B); x6 = __ffma2_rn(x6, A, B); x7 = __ffma2_rn(x7, A, B);
    } else if (MODE == 2) {  // MUFU.EX2 x16
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x4.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x4.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x5.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x5.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x6.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x6.y));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x7.x)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x7.y));
    } else if (MODE == 3) {  // scalar FADD x16
      x0.x += a; x0.y += a; x1.x += a; x1.y += a; x2.x += a; x2.y += a; x3.x += a; x3.y += a;
      x4.x += a; x4.y += a; x5.x += a; x5.y += a; x6.x += a; x6.y += a; x7.x += a; x7.y += a;
    } else if (MODE == 4) {  // packed FADD2 x8
      x0 = __fadd2_rn(x0, A); x1 = __fadd2_rn(x1, A); x2 = __fadd2_rn(x2, A); x3 = __fadd2_rn(x3, A);
      x4 = __fadd2_rn(x4, A); x5 = __fadd2_rn(x5, A); x6 = __fadd2_rn(x6, A); x7 = __fadd2_rn(x7, A);
    } else if (MODE == 5) {  // mixed: 8 FFMA + 8 FMNMX (fma pipe + alu pipe)
      x0.x = fmaf(x0.x, a, b); x0.y = fmaxf(x0.y, x0.x); x1.x = fmaf(x1.x, a, b); x1.y = fmaxf(x1.y, x1.x);
      x2.x = fmaf(x2.x, a, b); x2.y = fmaxf(x2.y, x2.x); x3.x = fmaf(x3.x, a, b); x3.y = fmaxf(x3.y, x3.x);
      x4.x = fmaf(x4.x, a, b); x4.y = fmaxf(x4.y, x4.x); x5.x = fmaf(x5.x, a, b); x5.y = fmaxf(x5.y, x5.x);
      x6.x = fmaf(x6.x, a, b); x6.y = fmaxf(x6.y, x6.x); x7.x = fmaf(x7.x, a, b); x7.y = fmaxf(x7.y, x7.x);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + x4.x + x4.y + x5.x +
                                               x5.y + x6.x + x6.y + x7.x + x7.y;
}
template <int V>
__global__ void lds(float* out) {  // shared-memory read bandwidth: V = 1 (LDS.32) or 4 (LDS.128), conflict-free
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  __syncthreads();
  float acc = 0.f;
  int idx = threadIdx.x * V;
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (V == 4) {
        float4 v = *reinterpret_cast<const float4*>(&s[(idx + u * 1024) & 8191]);
        acc += v.x + v.w;
      } else {
        acc += s[(idx + u * 256) & 8191];
      }
    }
    idx = (idx + 4) & 8191 & ~3;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int sm = 148, bs = 1024, grid = sm * 2;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char* names[] = {"FFMA x16", "FFMA2 x8 (16 lanes)", "MUFU.EX2 x16", "FADD x16", "FADD2 x8", "FFMA x8 + FMNMX x8"};
  float ms[6];
  ms[0] = timeit([&] { k<0><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  ms[1] = timeit([&] { k<1><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  ms[2] = timeit([&] { k<2><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  ms[3] = timeit([&] { k<3><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  ms[4] = timeit([&] { k<4><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  ms[5] = timeit([&] { k<5><<<grid, bs>>>(out, 1.0001f, 0.5f); });
  for (int m = 0; m < 6; ++m) {
    double lane_ops = (double)grid * bs * ITERS * 16;
    printf("%-24s %.3f ms  -> %.1f G lane-ops/s  = %.1f lane-ops/clk/SM @%d MHz nominal\n", names[m], ms[m],
           lane_ops / ms[m] / 1e6, lane_ops / ms[m] / 1e3 / sm / (clk / 1e3) , clk / 1000);
  }
  float t1 = timeit([&] { lds<1><<<grid, bs>>>(out); });
  float t4 = timeit([&] { lds<4><<<grid, bs>>>(out); });
  double b1 = (double)grid * bs * ITERS * 8 * 4, b4 = b1 * 4;
  printf("LDS.32  %.3f ms -> %.1f B/clk/SM\nLDS.128 %.3f ms -> %.1f B/clk/SM\n", t1, b1 / t1 / 1e3 / sm / (clk / 1e3), t4,
         b4 / t4 / 1e3 / sm / (clk / 1e3));
  return 0;
}
