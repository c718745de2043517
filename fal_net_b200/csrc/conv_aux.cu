// conv_aux.cu -- the small CUDA-core kernels around the tcgen05 convolution: the 3-channel stem convolution
// (K = 27 is too thin for a tensor-core tile and its input is the fp32 NCHW image itself), nearest upsampling,
// 2x2 max-pooling, all on bf16 NHWC activations.  All are HBM-bound streaming kernels.
#include <stdlib.h>

#include "common.cuh"

namespace faln {
namespace {

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : (__expf(v) - 1.0f); }

// Stem: y[b,h,w,:] = act(bias + sum_{c<3,kh,kw} w[co,c,kh,kw] * x[b,c,h+kh-1,w+kw-1]), x fp32 NCHW (optionally read
// x-reversed), y bf16 NHWC.  Replaces conv0.0 (/root/reference/models/FAL_netB.py:99) and VGG conv1_1.
// A thread owns TWO horizontally adjacent pixels x 32 output channels: the 3 x 4 input window is loaded once for both, and
// every broadcast 128-bit weight load feeds four packed FFMA2 (two channel pairs x two pixels), so the kernel issues
// ~540 instructions per pixel and 32 channels instead of ~1,080 (one pixel x COUT channels per thread, scalar FFMA; ncu
// launch lists profiles/r1e_*: 88 us per Stage-1 launch against a 24 us FMA-pipe floor).  fp32 accumulation in the same
// (c, kh, kw) order as before.
template <int COUT, int ACT>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                        int B, int H, int W, int flip_x) {
  constexpr int G = COUT / 32;                        // channel groups of 32
  __shared__ __align__(16) float ws[27 * COUT];       // [tap][co], tap = c*9 + kh*3 + kw
  __shared__ __align__(16) float bs[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int co = i % COUT, t = i / COUT;
    ws[i] = __ldg(w + co * 27 + t);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int pw = (W + 1) / 2;                          // pixel pairs per row
  const int units = B * H * pw * G;                    // < 2^31 (checked on the host)
  const int hw = H * W;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int cg = u % G;
    int r = u / G;
    const int xw = (r % pw) * 2;
    r /= pw;
    const int yh = r % H;
    const int b = r / H;
    // column offsets / validity of the four window columns and row validity of the three window rows: once per unit
    int xo[4];
    bool xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = xw + k - 1;
      xv[k] = xx >= 0 && xx < W;
      xo[k] = flip_x ? W - 1 - xx : xx;
    }
    float2 acc0[16], acc1[16];
    {
      const float2* bp = reinterpret_cast<const float2*>(bs + cg * 32);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc0[j] = acc1[j] = bp[j];
    }
    const float* xb = x + (size_t)b * 3 * hw + (size_t)yh * W;
    const float* wp0 = ws + cg * 32;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {      // rolled: keeps the loop body (288 FFMA2) inside the instruction cache
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = yh + kh - 1;
        const bool yv = yy >= 0 && yy < H;
        const float* xr = xb + (kh - 1) * W;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (yv && xv[k]) ? __ldg(xr + xo[k]) : 0.f;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* wp = reinterpret_cast<const float4*>(wp0 + (kh * 3 + kw) * COUT);
          const float2 a0 = make_float2(v[kw], v[kw]), a1 = make_float2(v[kw + 1], v[kw + 1]);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w4 = wp[q];
            const float2 wl = make_float2(w4.x, w4.y), wh = make_float2(w4.z, w4.w);
            acc0[2 * q] = __ffma2_rn(a0, wl, acc0[2 * q]);
            acc0[2 * q + 1] = __ffma2_rn(a0, wh, acc0[2 * q + 1]);
            acc1[2 * q] = __ffma2_rn(a1, wl, acc1[2 * q]);
            acc1[2 * q + 1] = __ffma2_rn(a1, wh, acc1[2 * q + 1]);
          }
        }
      }
      xb += hw;
      wp0 += 9 * COUT;
    }
    const size_t pix = ((size_t)b * H + yh) * W + xw;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (xw + px >= W) break;
      uint4* dst = reinterpret_cast<uint4*>(y + (pix + px) * COUT + cg * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = px ? acc1[q * 4 + e] : acc0[q * 4 + e];
          float a0 = a.x, a1 = a.y;
          if (ACT == 1) { a0 = elu1(a0); a1 = elu1(a1); }
          else if (ACT == 2) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
          h2[e] = __floats2bfloat162_rn(a0, a1);
        }
        dst[q] = o;
      }
    }
  }
}

// Nearest upsampling with PyTorch's legacy 'nearest' index rule src = min(floor(dst * in/out), in-1)
// (F.interpolate(mode='nearest'), /root/reference/models/FAL_netB.py:58), bf16 NHWC, 16 bytes per thread.
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B,
                                                               int Hi, int Wi, int Ho, int Wo, int C8, float sh, float sw) {
  const long long n = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int hi = min((int)floorf(ho * sh), Hi - 1), wi = min((int)floorf(wo * sw), Wi - 1);
    dst[i] = __ldg(src + (((long long)b * Hi + hi) * Wi + wi) * C8 + c);
  }
}

// The same, one block row per OUTPUT row: the source row and the 64-bit bases are computed once per block, the column index
// with 32-bit arithmetic, and every thread keeps four 16-byte loads in flight.  (The flat kernel above spends ~150 instructions
// per 16 bytes on 64-bit div / mod: 563 us per Test pass of 8 x 375 x 1242 against ~175 us of DRAM time.)
__global__ void __launch_bounds__(256) upsample_nearest_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                                    int Hi, int Wi, int Ho, int Wo, int C8, float sh, float sw) {
  const int row = blockIdx.y;                       // b * Ho + ho
  const int b = row / Ho, ho = row - b * Ho;
  const int hi = min((int)floorf(ho * sh), Hi - 1);
  const uint4* srow = src + ((long long)b * Hi + hi) * Wi * C8;
  uint4* drow = dst + (long long)row * Wo * C8;
  const int rowlen = Wo * C8;
  for (int j0 = blockIdx.x * 1024 + threadIdx.x; j0 < rowlen; j0 += gridDim.x * 1024) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + 256 * k;
      if (j < rowlen) {
        const int wo = j / C8, c = j - wo * C8;
        const int wi = min((int)floorf(wo * sw), Wi - 1);
        v[k] = __ldg(srow + wi * C8 + c);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + 256 * k;
      if (j < rowlen) drow[j] = v[k];
    }
  }
}

// 2x2 / stride-2 max pooling (floor mode), bf16 NHWC.
__global__ void __launch_bounds__(256) maxpool2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int Hi,
                                                       int Wi, int C8) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long n = (long long)B * Ho * Wo * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const uint4* p = src + (((long long)b * Hi + 2 * ho) * Wi + 2 * wo) * C8 + c;
    uint4 a = __ldg(p), bq = __ldg(p + C8), cq = __ldg(p + (long long)Wi * C8), d = __ldg(p + (long long)Wi * C8 + C8);
    uint4 o;
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bq);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&cq);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&d);
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) o2[e] = __hmax2(__hmax2(a2[e], b2[e]), __hmax2(c2[e], d2[e]));
    dst[i] = o;
  }
}

// Backward of the nearest upsampling, fused with the producer's activation derivative:
//   g_lo[b,hl,wl,:] = act'(y_lo[b,hl,wl,:]) * ( [accum: g_lo_old] + sum over the hi-res pixels (h,w) whose nearest source is
//   (hl,wl) of g_hi[b,h,w,:] ).   The pre-image of a source row is a contiguous run of destination rows; it is found by
// testing the forward index rule itself on the few candidates, so forward and backward cannot disagree.
__global__ void __launch_bounds__(256) upsample_nearest_bwd_kernel(const uint4* __restrict__ g_hi, const uint4* __restrict__ y_lo,
                                                                   uint4* __restrict__ g_lo, int B, int Hl, int Wl, int Hh,
                                                                   int Wh, int C8, float sh, float sw, int dact, int accum) {
  const long long n = (long long)B * Hl * Wl * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wl = (int)(r % Wl);
    r /= Wl;
    const int hl = (int)(r % Hl);
    const int b = (int)(r / Hl);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (accum) {
      const uint4 u = g_lo[i];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        acc[2 * e] = f.x;
        acc[2 * e + 1] = f.y;
      }
    }
    const int h_lo = max(0, (int)floorf((float)hl / sh) - 2), h_hi = min(Hh - 1, (int)ceilf((float)(hl + 1) / sh) + 2);
    const int w_lo = max(0, (int)floorf((float)wl / sw) - 2), w_hi = min(Wh - 1, (int)ceilf((float)(wl + 1) / sw) + 2);
    for (int h = h_lo; h <= h_hi; ++h) {
      if (min((int)floorf(h * sh), Hl - 1) != hl) continue;
      for (int w = w_lo; w <= w_hi; ++w) {
        if (min((int)floorf(w * sw), Wl - 1) != wl) continue;
        const uint4 u = __ldg(g_hi + (((long long)b * Hh + h) * Wh + w) * C8 + c);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
    }
    if (dact) {
      const uint4 u = __ldg(y_lo + i);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        acc[2 * e] *= f.x > 0.f ? 1.f : (dact == 1 ? f.x + 1.f : 0.f);
        acc[2 * e + 1] *= f.y > 0.f ? 1.f : (dact == 1 ? f.y + 1.f : 0.f);
      }
    }
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
    g_lo[i] = o;
  }
}

// Backward of the 2x2 max-pool fused with the ReLU derivative of the pooled-from activation x (VGG slices):
//   g_x[b,h,w,:] = [x is the (first) maximum of its window] * [x > 0 if dact] * g_y[b,h/2,w/2,:];  odd trailing rows/cols get 0.
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g_y,
                                                           uint4* __restrict__ g_x, int B, int Hi, int Wi, int C8, int dact) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long n = (long long)B * Hi * Wi * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int w = (int)(r % Wi);
    r /= Wi;
    const int h = (int)(r % Hi);
    const int b = (int)(r / Hi);
    uint4 o = make_uint4(0, 0, 0, 0);
    const int ho = h >> 1, wo = w >> 1;
    if (ho < Ho && wo < Wo) {
      const uint4* p = x + (((long long)b * Hi + 2 * ho) * Wi + 2 * wo) * C8 + c;
      const uint4 q[4] = {__ldg(p), __ldg(p + C8), __ldg(p + (long long)Wi * C8), __ldg(p + (long long)Wi * C8 + C8)};
      const uint4 gy = __ldg(g_y + (((long long)b * Ho + ho) * Wo + wo) * C8 + c);
      const int me = (h & 1) * 2 + (w & 1);
      const __nv_bfloat16* gv = reinterpret_cast<const __nv_bfloat16*>(&gy);
      __nv_bfloat16* ov = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&q[k])[e]);
        int arg = 0;  // first maximum in window scan order (matches ATen's max_pool2d_with_indices tie-breaking)
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k] > v[arg]) arg = k;
        const bool on = arg == me && (!dact || v[me] > 0.f);
        ov[e] = on ? gv[e] : __float2bfloat16(0.f);
      }
    }
    g_x[i] = o;
  }
}

// The same, one thread per 2x2 WINDOW and 8-channel chunk: four x loads + one g_y load feed four stores (the per-pixel kernel
// above loads the whole window again for each of its four pixels and pays 64-bit div / mod per 16 bytes).  One block row per
// window row; odd trailing rows / columns are zero-filled by the threads of the last window row / column.
__global__ void __launch_bounds__(256) maxpool2_bwd_win_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g_y,
                                                               uint4* __restrict__ g_x, int Hi, int Wi, int C8, int dact) {
  const int Ho = Hi / 2, Wo = Wi / 2, Hw = (Hi + 1) / 2, Ww = (Wi + 1) / 2;   // windows incl. the partial trailing ones
  const int row = blockIdx.y;                       // b * Hw + window row
  const int b = row / Hw, ho = row - b * Hw;
  const int rowlen = Ww * C8;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int j = blockIdx.x * 256 + threadIdx.x; j < rowlen; j += gridDim.x * 256) {
    const int wo = j / C8, c = j - wo * C8;
    const int h0 = 2 * ho, w0 = 2 * wo;
    const long long base = (((long long)b * Hi + h0) * Wi + w0) * C8 + c;
    const long long dn = (long long)Wi * C8;
    if (ho < Ho && wo < Wo) {
      const uint4 q[4] = {__ldg(x + base), __ldg(x + base + C8), __ldg(x + base + dn), __ldg(x + base + dn + C8)};
      const uint4 gy = __ldg(g_y + (((long long)b * Ho + ho) * Wo + wo) * C8 + c);
      uint4 o[4];
      const unsigned short* gv = reinterpret_cast<const unsigned short*>(&gy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&q[k])[e]);
        // first maximum in window scan order (ATen's max_pool2d_with_indices tie-breaking)
        const bool g1 = v[1] > v[0];
        const float m01 = g1 ? v[1] : v[0];
        const bool g2 = v[2] > m01;
        const float m012 = g2 ? v[2] : m01;
        const bool g3 = v[3] > m012;
        const int arg = g3 ? 3 : (g2 ? 2 : (g1 ? 1 : 0));
        const float vmax = g3 ? v[3] : m012;
        const bool live = !dact || vmax > 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          reinterpret_cast<unsigned short*>(&o[k])[e] = (live && arg == k) ? gv[e] : (unsigned short)0;
      }
      g_x[base] = o[0];
      g_x[base + C8] = o[1];
      g_x[base + dn] = o[2];
      g_x[base + dn + C8] = o[3];
    } else {
      // partial trailing window: whatever pixels of it exist get a zero gradient (floor-mode pooling never reads them)
      if (h0 < Hi && w0 < Wi) g_x[base] = zero;
      if (h0 < Hi && w0 + 1 < Wi) g_x[base + C8] = zero;
      if (h0 + 1 < Hi && w0 < Wi) g_x[base + dn] = zero;
      if (h0 + 1 < Hi && w0 + 1 < Wi) g_x[base + dn + C8] = zero;
    }
  }
}

// Per-channel sums of a bf16 NHWC gradient (bias gradients): out[c] += sum over pixels of g[pix, c], fp32.
// Block = 256 threads = (256 / C8) pixel lanes x C8 channel groups of 8; shared-memory tree over the pixel lanes, then one
// atomicAdd per channel per block.
template <bool kUnroll>
__global__ void __launch_bounds__(256) channel_sum_kernel(const uint4* __restrict__ g, float* __restrict__ out, long long npix,
                                                          int C8, int Cstride8, int Cout) {
  __shared__ float sm[256 * 8];
  const int lanes = 256 / C8;                 // C8 in {4, 8, 16, 32, 64}: divides 256
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // four independent 16-byte loads in flight per thread (the single-load loop left the HBM pipe at ~1/3: 31 us per launch
  // in profiles/r1d_launches_stage1.txt)
  const long long step = (long long)gridDim.x * lanes;
  long long px = (long long)blockIdx.x * lanes + pl;
  auto add8 = [&](const uint4& u) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h2[e]);
      acc[2 * e] += f.x;
      acc[2 * e + 1] += f.y;
    }
  };
  for (; kUnroll && px + 3 * step < npix; px += 4 * step) {
    const uint4 u0 = __ldg(g + px * Cstride8 + cg);
    const uint4 u1 = __ldg(g + (px + step) * Cstride8 + cg);
    const uint4 u2 = __ldg(g + (px + 2 * step) * Cstride8 + cg);
    const uint4 u3 = __ldg(g + (px + 3 * step) * Cstride8 + cg);
    add8(u0);
    add8(u1);
    add8(u2);
    add8(u3);
  }
  for (; px < npix; px += step) add8(__ldg(g + px * Cstride8 + cg));
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  for (int s = lanes / 2; s > 0; s >>= 1) {
    if (pl < s) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sm[threadIdx.x * 8 + e] += sm[(threadIdx.x + s * C8) * 8 + e];
    }
    __syncthreads();
  }
  if (pl == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (cg * 8 + e < Cout) atomicAdd(out + cg * 8 + e, sm[threadIdx.x * 8 + e]);
  }
}

inline int ew_grid(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  long long cap = (long long)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_stem_conv(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int Cout,
                              int act, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "faln_stem_conv: bad argument");
  FALN_REQUIRE(Cout == 32 || Cout == 64, "faln_stem_conv: Cout must be 32 or 64 (got %d)", Cout);
  FALN_REQUIRE(act >= 0 && act <= 2, "faln_stem_conv: act must be 0 (none), 1 (ELU) or 2 (ReLU)");
  const long long units = (long long)B * H * ((W + 1) / 2) * (Cout / 32);
  FALN_REQUIRE(units < (1LL << 31) && (long long)B * 3 * H * W < (1LL << 31), "faln_stem_conv: tensor too large");
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(y);
  const int grid = ew_grid(units, 128);
  cudaStream_t st = as_stream(stream);
#define FALN_STEM(C, A) stem_conv_kernel<C, A><<<grid, 128, 0, st>>>(x, w, bias, out, B, H, W, flip_x)
  if (Cout == 32) {
    if (act == 0) FALN_STEM(32, 0); else if (act == 1) FALN_STEM(32, 1); else FALN_STEM(32, 2);
  } else {
    if (act == 0) FALN_STEM(64, 0); else if (act == 1) FALN_STEM(64, 1); else FALN_STEM(64, 2);
  }
#undef FALN_STEM
  return after_launch("stem_conv_kernel");
}

extern "C" int faln_upsample_nearest_nhwc(const void* src, void* dst, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                          faln_stream_t stream) {
  FALN_REQUIRE(src && dst && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "faln_upsample_nearest_nhwc: C %% 8");
  const long long n = (long long)B * Ho * Wo * (C / 8);
  static const bool flat = getenv("FALN_EW_FLAT") != nullptr;      // A/B switch: the first-generation flat-index kernels
  if (!flat && (long long)B * Ho <= 65535 && (long long)Wo * (C / 8) < (1LL << 30) && (long long)Wi * (C / 8) < (1LL << 30)) {
    const int rowlen = Wo * (C / 8);
    int gx = (rowlen + 1023) / 1024;
    if (gx > 8) gx = 8;
    upsample_nearest_rows_kernel<<<dim3(gx, B * Ho), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(src), static_cast<uint4*>(dst), Hi, Wi, Ho, Wo, C / 8, (float)Hi / (float)Ho,
        (float)Wi / (float)Wo);
    return after_launch("upsample_nearest_rows_kernel");
  }
  upsample_nearest_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), B, Hi, Wi, Ho, Wo, C / 8, (float)Hi / (float)Ho,
      (float)Wi / (float)Wo);
  return after_launch("upsample_nearest_kernel");
}

extern "C" int faln_maxpool2_nhwc(const void* src, void* dst, int B, int Hi, int Wi, int C, faln_stream_t stream) {
  FALN_REQUIRE(src && dst && B > 0 && Hi >= 2 && Wi >= 2 && C % 8 == 0, "faln_maxpool2_nhwc: bad argument");
  const long long n = (long long)B * (Hi / 2) * (Wi / 2) * (C / 8);
  maxpool2_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), B,
                                                                  Hi, Wi, C / 8);
  return after_launch("maxpool2_kernel");
}

extern "C" int faln_upsample_nearest_bwd_nhwc(const void* g_hi, const void* y_lo, void* g_lo, int B, int Hl, int Wl, int Hh,
                                              int Wh, int C, int dact, int accum, faln_stream_t stream) {
  FALN_REQUIRE(g_hi && g_lo && B > 0 && Hl > 0 && Wl > 0 && Hh > 0 && Wh > 0 && C % 8 == 0,
               "faln_upsample_nearest_bwd_nhwc: bad argument");
  FALN_REQUIRE((dact == 0) || y_lo, "faln_upsample_nearest_bwd_nhwc: dact needs the saved activation");
  const long long n = (long long)B * Hl * Wl * (C / 8);
  upsample_nearest_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(g_hi), static_cast<const uint4*>(y_lo), static_cast<uint4*>(g_lo), B, Hl, Wl, Hh, Wh, C / 8,
      (float)Hl / (float)Hh, (float)Wl / (float)Wh, dact, accum);
  return after_launch("upsample_nearest_bwd_kernel");
}

extern "C" int faln_maxpool2_bwd_nhwc(const void* x, const void* g_y, void* g_x, int B, int Hi, int Wi, int C, int dact,
                                      faln_stream_t stream) {
  FALN_REQUIRE(x && g_y && g_x && B > 0 && Hi >= 2 && Wi >= 2 && C % 8 == 0, "faln_maxpool2_bwd_nhwc: bad argument");
  const long long n = (long long)B * Hi * Wi * (C / 8);
  static const bool flat = getenv("FALN_EW_FLAT") != nullptr;      // A/B switch: the first-generation flat-index kernel
  const int Hw = (Hi + 1) / 2, Ww = (Wi + 1) / 2;
  if (!flat && (long long)B * Hw <= 65535 && (long long)Wi * (C / 8) < (1LL << 30)) {
    const int rowlen = Ww * (C / 8);
    int gx = (rowlen + 255) / 256;
    if (gx > 16) gx = 16;
    maxpool2_bwd_win_kernel<<<dim3(gx, B * Hw), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(x), static_cast<const uint4*>(g_y), static_cast<uint4*>(g_x), Hi, Wi, C / 8, dact);
    return after_launch("maxpool2_bwd_win_kernel");
  }
  maxpool2_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(g_y),
                                                                      static_cast<uint4*>(g_x), B, Hi, Wi, C / 8, dact);
  return after_launch("maxpool2_bwd_kernel");
}

extern "C" int faln_channel_sum_nhwc(const void* g, float* out, long long npix, int C, int Cstride, faln_stream_t stream) {
  FALN_REQUIRE(g && out && npix > 0 && C > 0 && Cstride % 8 == 0 && Cstride >= C, "faln_channel_sum_nhwc: bad argument");
  const int C8 = Cstride / 8;   // all channel groups of the tensor are walked; channels >= C are dropped at the end
  FALN_REQUIRE(C8 <= 64 && 256 % C8 == 0, "faln_channel_sum_nhwc: Cstride must be 8/16/32/64/128/256/512 (got %d)", Cstride);
  const int lanes = 256 / C8;
  long long grid = (npix + lanes - 1) / lanes;
  // Tuning aids (the kernel runs on the gradient side stream, concurrently with the data-gradient chain, so its
  // footprint matters as much as its own speed): FALN_CHSUM_CAP = blocks per SM, FALN_CHSUM_UNROLL = 0/1.
  // Measured on B200 (Stage-1 step, gpurun_out/s5_*): 8 blocks/SM 5.36 ms, 4 blocks/SM 5.16 ms, 2 blocks/SM with four
  // loads in flight 5.05 ms -- a small footprint leaves the SMs to the data-gradient chain.
  static const int cap_per_sm = getenv("FALN_CHSUM_CAP") ? atoi(getenv("FALN_CHSUM_CAP")) : 1;  // round 2: 1 (own stream, FALN_BIAS_STREAM): 4.37 -> 4.34 ms
  static const int unroll = getenv("FALN_CHSUM_UNROLL") ? atoi(getenv("FALN_CHSUM_UNROLL")) : 1;
  const long long cap = (long long)sm_count() * (cap_per_sm > 0 ? cap_per_sm : 1);
  if (grid > cap) grid = cap;
  if (unroll)
    channel_sum_kernel<true><<<(int)grid, 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(g), out, npix, C8, Cstride / 8, C);
  else
    channel_sum_kernel<false><<<(int)grid, 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(g), out, npix, C8, Cstride / 8, C);
  return after_launch("channel_sum_kernel");
}
