// losses.cu -- fused loss kernels (forward value + analytic gradient), all HBM-bound streaming passes.
//
// Replaces the ~60 ATen/cuDNN launches per step behind /root/reference/loss_functions.py:52-109 and
// the inline mask / mirror arithmetic of /root/reference/Train_Stage2_K.py:295-324.
// Reductions are deterministic: per-block partials, folded in fixed order by the last block to finish.
#include "common.cuh"

namespace faln {
namespace {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 2048;
constexpr int kRedHeader = 4;  // floats reserved in front of the partials (ticket counter)

__device__ __forceinline__ void block_finish(float v, float* partials, float* out, float scale) {
  __shared__ float wsum[kRedThreads / 32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kRedThreads / 32 ? wsum[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) {
      partials[kRedHeader + blockIdx.x] = v;
      __threadfence();
      unsigned t = atomicAdd(reinterpret_cast<unsigned*>(partials), 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kRedThreads)
      s += reinterpret_cast<volatile float*>(partials)[kRedHeader + i];
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) wsum[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kRedThreads / 32; ++w) tot += wsum[w];
      out[0] = tot * scale;
      *reinterpret_cast<unsigned*>(partials) = 0u;
    }
  }
}

inline int red_grid(long long n) {
  long long g = (n + kRedThreads - 1) / kRedThreads;
  int cap = sm_count() * 8;
  if (cap > kRedMaxBlocks) cap = kRedMaxBlocks;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------ rec L1
__global__ void __launch_bounds__(kRedThreads) rec_l1_kernel(const float* __restrict__ synth,
                                                             const float* __restrict__ label,
                                                             const float* __restrict__ mask, float* __restrict__ blend,
                                                             float* out, float* partials, int B, int H, int W,
                                                             int flip_x) {
  const long long n = (long long)B * 3 * H * W;
  const long long hw = (long long)H * W;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
    const int x = (int)(i % W);
    const long long is = flip_x ? i - x + (W - 1 - x) : i;
    const float s = __ldg(synth + is), l = __ldg(label + i);
    float m = 1.f;
    if (mask) {
      const long long b = i / (3 * hw);
      m = __ldg(mask + b * hw + (i % hw));
    }
    acc += m * fabsf(s - l);
    if (blend) blend[i] = m * s + (1.f - m) * l;
  }
  block_finish(acc, partials, out, 1.0f / (float)n);
}

__global__ void __launch_bounds__(256) rec_l1_bwd_kernel(const float* __restrict__ synth, const float* __restrict__ label,
                                                         const float* __restrict__ mask, const float* __restrict__ g_blend,
                                                         float g_scale, const float* __restrict__ g_dev,
                                                         float* __restrict__ g_synth, int B, int H, int W, int flip_x) {
  const long long n = (long long)B * 3 * H * W;
  const long long hw = (long long)H * W;
  const float gs = (g_dev ? g_scale * __ldg(g_dev) : g_scale) / (float)n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const long long is = flip_x ? i - x + (W - 1 - x) : i;
    const float d = __ldg(synth + is) - __ldg(label + i);
    float m = 1.f;
    if (mask) {
      const long long b = i / (3 * hw);
      m = __ldg(mask + b * hw + (i % hw));
    }
    float g = gs * m * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    if (g_blend) g = fmaf(__ldg(g_blend + i), m, g);
    g_synth[is] = g;
  }
}

// ------------------------------------------------------------------------------------------ smoothness
struct SmoothArgs {
  const float* img;
  const float* disp;
  int B, H, W, x_lo, x_hi, flip_x;
  float gamma;
};

__device__ __forceinline__ float gray_at(const SmoothArgs& a, int b, int y, int x) {
  if (y < 0 || y >= a.H || x < a.x_lo || x >= a.x_hi) return 0.f;
  const size_t hw = (size_t)a.H * a.W;
  const float* p = a.img + (size_t)b * 3 * hw + (size_t)y * a.W + x;
  // getGrayscale(img + mean), /root/reference/loss_functions.py:73-77,104-109
  return 0.299f * (__ldg(p) + 0.411f) + 0.587f * (__ldg(p + hw) + 0.432f) + 0.114f * (__ldg(p + 2 * hw) + 0.45f);
}
__device__ __forceinline__ float disp_at(const SmoothArgs& a, int b, int y, int x) {
  if (y < 0 || y >= a.H || x < a.x_lo || x >= a.x_hi) return 0.f;
  const int xs = a.flip_x ? a.W - 1 - x : x;
  return __ldg(a.disp + ((size_t)b * a.H + y) * a.W + xs);
}
__device__ __forceinline__ bool in_win(const SmoothArgs& a, int y, int x) {
  return y >= 0 && y < a.H && x >= a.x_lo && x < a.x_hi;
}
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// edge weights exp(-gamma |2nd difference of gray|) at window pixel (y, x)
__device__ __forceinline__ void edge_w(const SmoothArgs& a, int b, int y, int x, float& wx, float& wy) {
  const float g = gray_at(a, b, y, x);
  const float ddx = 2.f * g - gray_at(a, b, y, x - 1) - gray_at(a, b, y, x + 1);
  const float ddy = 2.f * g - gray_at(a, b, y - 1, x) - gray_at(a, b, y + 1, x);
  wx = __expf(-a.gamma * fabsf(ddx));
  wy = __expf(-a.gamma * fabsf(ddy));
}

__global__ void __launch_bounds__(kRedThreads) smooth_kernel(const SmoothArgs a, float* out, float* partials) {
  const int ww = a.x_hi - a.x_lo;
  const long long n = (long long)a.B * a.H * ww;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
    const int x = a.x_lo + (int)(i % ww);
    const int y = (int)((i / ww) % a.H);
    const int b = (int)(i / ((long long)ww * a.H));
    float wx, wy;
    edge_w(a, b, y, x, wx, wy);
    const float d = disp_at(a, b, y, x);
    const float dx = d - disp_at(a, b, y, x + 1), dx1 = d - disp_at(a, b, y, x - 1);
    const float dy = d - disp_at(a, b, y - 1, x), dy1 = d - disp_at(a, b, y + 1, x);
    acc += (fabsf(dx) + fabsf(dx1)) * wx + (fabsf(dy) + fabsf(dy1)) * wy;
  }
  block_finish(acc, partials, out, 1.0f / (float)n);
}

__global__ void __launch_bounds__(256) smooth_bwd_kernel(const SmoothArgs a, float g_scale, const float* __restrict__ g_dev,
                                                         float* __restrict__ g_disp, int accumulate) {
  const int ww = a.x_hi - a.x_lo;
  const long long n = (long long)a.B * a.H * ww;
  const float gs = (g_dev ? g_scale * __ldg(g_dev) : g_scale) / (float)n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = a.x_lo + (int)(i % ww);
    const int y = (int)((i / ww) % a.H);
    const int b = (int)(i / ((long long)ww * a.H));
    const float d = disp_at(a, b, y, x);
    float wx, wy, g;
    edge_w(a, b, y, x, wx, wy);
    g = wx * (sgn(d - disp_at(a, b, y, x + 1)) + sgn(d - disp_at(a, b, y, x - 1))) +
        wy * (sgn(d - disp_at(a, b, y - 1, x)) + sgn(d - disp_at(a, b, y + 1, x)));
    float wx2, wy2;
    if (in_win(a, y, x - 1)) {  // q = p - x: dx_d(q) = d(q) - d(p)
      edge_w(a, b, y, x - 1, wx2, wy2);
      g -= wx2 * sgn(disp_at(a, b, y, x - 1) - d);
    }
    if (in_win(a, y, x + 1)) {  // q = p + x: dx1_d(q) = d(q) - d(p)
      edge_w(a, b, y, x + 1, wx2, wy2);
      g -= wx2 * sgn(disp_at(a, b, y, x + 1) - d);
    }
    if (in_win(a, y + 1, x)) {  // q = p + y: dy_d(q) = d(q) - d(q - y) = d(q) - d(p)
      edge_w(a, b, y + 1, x, wx2, wy2);
      g -= wy2 * sgn(disp_at(a, b, y + 1, x) - d);
    }
    if (in_win(a, y - 1, x)) {  // q = p - y: dy1_d(q) = d(q) - d(p)
      edge_w(a, b, y - 1, x, wx2, wy2);
      g -= wy2 * sgn(disp_at(a, b, y - 1, x) - d);
    }
    const int xs = a.flip_x ? a.W - 1 - x : x;
    float* dst = g_disp + ((size_t)b * a.H + y) * a.W + xs;
    *dst = accumulate ? *dst + gs * g : gs * g;
  }
}

// ------------------------------------------------------------------------------------------ mirror
__global__ void __launch_bounds__(kRedThreads) mirror_kernel(const float* __restrict__ disp, const float* __restrict__ mdisp,
                                                             const float* __restrict__ occ, const float* __restrict__ inv_max,
                                                             float* out, float* partials, int B, int H, int W, int x_lo,
                                                             int x_hi, int flip_x) {
  const int ww = x_hi - x_lo;
  const long long n = (long long)B * H * ww;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)kRedThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kRedThreads) {
    const int x = x_lo + (int)(i % ww);
    const long long r = i / ww;  // b*H + y
    const int b = (int)(r / H);
    const int xs = flip_x ? W - 1 - x : x;
    const float d = __ldg(disp + r * W + xs), md = __ldg(mdisp + r * W + x), o = __ldg(occ + r * W + x);
    acc += __ldg(inv_max + b) * (1.f - o) * fabsf(d - md);
  }
  block_finish(acc, partials, out, 1.0f / (float)n);
}

__global__ void __launch_bounds__(256) mirror_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ mdisp,
                                                         const float* __restrict__ occ, const float* __restrict__ inv_max,
                                                         float g_scale, const float* __restrict__ g_dev,
                                                         float* __restrict__ g_disp, int accumulate, int B, int H, int W,
                                                         int x_lo, int x_hi, int flip_x) {
  const int ww = x_hi - x_lo;
  const long long n = (long long)B * H * ww;
  const float gs = (g_dev ? g_scale * __ldg(g_dev) : g_scale) / (float)n;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = x_lo + (int)(i % ww);
    const long long r = i / ww;
    const int b = (int)(r / H);
    const int xs = flip_x ? W - 1 - x : x;
    const float d = __ldg(disp + r * W + xs), md = __ldg(mdisp + r * W + x), o = __ldg(occ + r * W + x);
    const float g = gs * __ldg(inv_max + b) * (1.f - o) * sgn(d - md);
    float* dst = g_disp + r * W + xs;
    *dst = accumulate ? *dst + g : g;
  }
}

// ------------------------------------------------------------------------------------------ MSE (bf16 features)
__global__ void __launch_bounds__(kRedThreads) mse_bf16_kernel(const __nv_bfloat162* __restrict__ a,
                                                               const __nv_bfloat162* __restrict__ b, long long n2,
                                                               float scale, float* out, float* partials) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)kRedThreads + threadIdx.x; i < n2; i += (long long)gridDim.x * kRedThreads) {
    const float2 x = __bfloat1622float2(a[i]), y = __bfloat1622float2(b[i]);
    const float d0 = x.x - y.x, d1 = x.y - y.y;
    acc = fmaf(d0, d0, fmaf(d1, d1, acc));
  }
  block_finish(acc, partials, out, scale);
}
__global__ void __launch_bounds__(256) mse_bf16_bwd_kernel(const __nv_bfloat162* __restrict__ a,
                                                           const __nv_bfloat162* __restrict__ b, long long n2, float gs,
                                                           const float* __restrict__ g_dev, __nv_bfloat162* __restrict__ g) {
  const float s = g_dev ? gs * __ldg(g_dev) : gs;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n2; i += (long long)gridDim.x * 256) {
    const float2 x = __bfloat1622float2(a[i]), y = __bfloat1622float2(b[i]);
    g[i] = __floats2bfloat162_rn(s * (x.x - y.x), s * (x.y - y.y));
  }
}

// ------------------------------------------------------------------------------------------ small helpers
__global__ void __launch_bounds__(256) inv_rowmax_kernel(const float* __restrict__ x, float* __restrict__ inv_max,
                                                         long long n_per) {
  __shared__ float wmax[8];
  const float* p = x + (size_t)blockIdx.x * n_per;
  float m = -INFINITY;
  for (long long i = threadIdx.x; i < n_per; i += 256) m = fmaxf(m, __ldg(p + i));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wmax[w]);
    inv_max[blockIdx.x] = 1.0f / m;
  }
}

__global__ void __launch_bounds__(256) occ_mask_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ out, long long n, int W, int flip_a, int flip_b,
                                                       int one_lo, int one_hi) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const long long ifl = i - x + (W - 1 - x);
    const float v = __ldg(a + (flip_a ? ifl : i)) * __ldg(b + (flip_b ? ifl : i));
    out[i] = (x >= one_lo && x < one_hi) ? 1.0f : v;
  }
}

inline int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_loss_partials_len(void) { return kRedHeader + kRedMaxBlocks; }

extern "C" int faln_loss_rec_l1(const float* synth, const float* label, const float* mask, float* blend, float* out,
                                float* partials, int B, int H, int W, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(synth && label && out && partials && B > 0 && H > 0 && W > 0, "faln_loss_rec_l1: bad argument");
  const long long n = (long long)B * 3 * H * W;
  rec_l1_kernel<<<red_grid(n), kRedThreads, 0, as_stream(stream)>>>(synth, label, mask, blend, out, partials, B, H, W,
                                                                     flip_x);
  return after_launch("rec_l1_kernel");
}

extern "C" int faln_loss_rec_l1_bwd(const float* synth, const float* label, const float* mask, const float* g_blend,
                                    float g_scale, const float* g_dev, float* g_synth, int B, int H, int W, int flip_x,
                                    faln_stream_t stream) {
  FALN_REQUIRE(synth && label && g_synth && B > 0 && H > 0 && W > 0, "faln_loss_rec_l1_bwd: bad argument");
  const long long n = (long long)B * 3 * H * W;
  rec_l1_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(synth, label, mask, g_blend, g_scale, g_dev, g_synth, B,
                                                                H, W, flip_x);
  return after_launch("rec_l1_bwd_kernel");
}

extern "C" int faln_loss_smooth(const float* img, const float* disp, float gamma, float* out, float* partials, int B,
                                int H, int W, int x_lo, int x_hi, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(img && disp && out && partials && B > 0 && H > 0 && 0 <= x_lo && x_lo < x_hi && x_hi <= W,
               "faln_loss_smooth: bad argument");
  SmoothArgs a{img, disp, B, H, W, x_lo, x_hi, flip_x, gamma};
  smooth_kernel<<<red_grid((long long)B * H * (x_hi - x_lo)), kRedThreads, 0, as_stream(stream)>>>(a, out, partials);
  return after_launch("smooth_kernel");
}

extern "C" int faln_loss_smooth_bwd(const float* img, const float* disp, float gamma, float g_scale, const float* g_dev,
                                    float* g_disp, int accumulate, int B, int H, int W, int x_lo, int x_hi, int flip_x,
                                    faln_stream_t stream) {
  FALN_REQUIRE(img && disp && g_disp && B > 0 && H > 0 && 0 <= x_lo && x_lo < x_hi && x_hi <= W,
               "faln_loss_smooth_bwd: bad argument");
  SmoothArgs a{img, disp, B, H, W, x_lo, x_hi, flip_x, gamma};
  smooth_bwd_kernel<<<ew_grid((long long)B * H * (x_hi - x_lo)), 256, 0, as_stream(stream)>>>(a, g_scale, g_dev, g_disp,
                                                                                              accumulate);
  return after_launch("smooth_bwd_kernel");
}

extern "C" int faln_loss_mirror(const float* disp, const float* mdisp, const float* occ, const float* inv_max, float* out,
                                float* partials, int B, int H, int W, int x_lo, int x_hi, int flip_x,
                                faln_stream_t stream) {
  FALN_REQUIRE(disp && mdisp && occ && inv_max && out && partials && B > 0 && 0 <= x_lo && x_lo < x_hi && x_hi <= W,
               "faln_loss_mirror: bad argument");
  mirror_kernel<<<red_grid((long long)B * H * (x_hi - x_lo)), kRedThreads, 0, as_stream(stream)>>>(
      disp, mdisp, occ, inv_max, out, partials, B, H, W, x_lo, x_hi, flip_x);
  return after_launch("mirror_kernel");
}

extern "C" int faln_loss_mirror_bwd(const float* disp, const float* mdisp, const float* occ, const float* inv_max,
                                    float g_scale, const float* g_dev, float* g_disp, int accumulate, int B, int H, int W,
                                    int x_lo, int x_hi, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(disp && mdisp && occ && inv_max && g_disp && B > 0 && 0 <= x_lo && x_lo < x_hi && x_hi <= W,
               "faln_loss_mirror_bwd: bad argument");
  mirror_bwd_kernel<<<ew_grid((long long)B * H * (x_hi - x_lo)), 256, 0, as_stream(stream)>>>(
      disp, mdisp, occ, inv_max, g_scale, g_dev, g_disp, accumulate, B, H, W, x_lo, x_hi, flip_x);
  return after_launch("mirror_bwd_kernel");
}

extern "C" int faln_mse_bf16(const void* a, const void* b, long long n, float* out, float* partials,
                             faln_stream_t stream) {
  FALN_REQUIRE(a && b && out && partials && n > 0 && (n & 1) == 0, "faln_mse_bf16: n must be even and > 0");
  mse_bf16_kernel<<<red_grid(n / 2), kRedThreads, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat162*>(a), static_cast<const __nv_bfloat162*>(b), n / 2, 1.0f / (float)n, out, partials);
  return after_launch("mse_bf16_kernel");
}

extern "C" int faln_mse_bf16_bwd(const void* a, const void* b, long long n, float g_scale, const float* g_dev, void* g_a,
                                 faln_stream_t stream) {
  FALN_REQUIRE(a && b && g_a && n > 0 && (n & 1) == 0, "faln_mse_bf16_bwd: n must be even and > 0");
  mse_bf16_bwd_kernel<<<ew_grid(n / 2), 256, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat162*>(a), static_cast<const __nv_bfloat162*>(b), n / 2, 2.0f * g_scale / (float)n, g_dev,
      static_cast<__nv_bfloat162*>(g_a));
  return after_launch("mse_bf16_bwd_kernel");
}

extern "C" int faln_inv_rowmax(const float* x, float* inv_max, int B, long long n_per_image, faln_stream_t stream) {
  FALN_REQUIRE(x && inv_max && B > 0 && n_per_image > 0, "faln_inv_rowmax: bad argument");
  inv_rowmax_kernel<<<B, 256, 0, as_stream(stream)>>>(x, inv_max, n_per_image);
  return after_launch("inv_rowmax_kernel");
}

extern "C" int faln_occ_mask(const float* a, const float* b, float* out, int B, int H, int W, int flip_a, int flip_b,
                             int one_lo, int one_hi, faln_stream_t stream) {
  FALN_REQUIRE(a && b && out && B > 0 && H > 0 && W > 0, "faln_occ_mask: bad argument");
  const long long n = (long long)B * H * W;
  occ_mask_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a, b, out, n, W, flip_a, flip_b, one_lo, one_hi);
  return after_launch("occ_mask_kernel");
}
