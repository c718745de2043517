// postproc.cu -- inference post-processing and validation metrics on the device.
//
//   ms_pp (/root/reference/Test_KITTI.py:287-300): flip + bilinear 2/3 down-scale (align_corners=True) of the input view,
//   a per-image 95th percentile of the disparity WITHOUT the reference's `.cpu().numpy()` host sync, and the blend with the
//   nearest-upsampled, un-flipped low-resolution disparity.
//   validate() (/root/reference/Train_Stage1_K.py:279-347, myUtils.py:138-150,196-277, loss_functions.py:124-173): RMSE of
//   the synthesised view, realEPE, and the seven KITTI depth errors, as per-image partial sums in fp64.
//
// Everything here is HBM / latency bound and tiny next to the network passes; the point is that nothing forces a
// device-to-host synchronisation inside the evaluation loop.
#include <cooperative_groups.h>

#include "common.cuh"

namespace faln {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// flip_x + bilinear resize, align_corners=True (ATen upsample_bilinear2d: src = dst * (in-1)/(out-1), fp32)
// in [BC,H,W] fp32 -> out [BC,Ho,Wo] fp32
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flip_resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                   int BC, int H, int W, int Ho, int Wo, float sy, float sx,
                                                                   int flip_x) {
  const long long n = (long long)BC * Ho * Wo;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    const int yo = (int)((i / Wo) % Ho);
    const long long bc = i / ((long long)Wo * Ho);
    const float h1r = sy * yo;
    const int h1 = (int)h1r;
    const int h1p = h1 < H - 1 ? 1 : 0;
    const float h1l = h1r - h1, h0l = 1.f - h1l;
    const float w1r = sx * xo;
    const int w1 = (int)w1r;
    const int w1p = w1 < W - 1 ? 1 : 0;
    const float w1l = w1r - w1, w0l = 1.f - w1l;
    const float* p = in + bc * (long long)H * W;
    const int xa = flip_x ? W - 1 - w1 : w1;
    const int xb = flip_x ? W - 1 - (w1 + w1p) : w1 + w1p;
    const float v00 = __ldg(p + (long long)h1 * W + xa), v01 = __ldg(p + (long long)h1 * W + xb);
    const float v10 = __ldg(p + (long long)(h1 + h1p) * W + xa), v11 = __ldg(p + (long long)(h1 + h1p) * W + xb);
    out[i] = h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// q-th percentile of every row of x [B, n] (numpy.percentile, method 'linear'), exact: three radix-select passes over the
// order-preserving integer image of the floats (11 + 11 + 10 bits) give the k-th order statistic, a fourth pass the
// next larger value when the (k+1)-th differs.
// A row is one thread-block CLUSTER of kPctCluster CTAs (round 2: one CTA per row took 226 us for 8 x 466 k values -- a
// single SM's load bandwidth plus 32-way same-address shared atomics, disparities fall into a few dozen of the 2048 first-pass
// bins): every CTA histograms its interleaved slice into its own shared memory with warp-aggregated atomics, the
// histograms are summed over distributed shared memory by every CTA (identical result everywhere, so no broadcast), and
// the bin search runs redundantly per CTA.
// out[b] = (float)(p + add), p computed in fp64 like numpy's lerp on its float64 virtual index.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// finds the bin holding rank `r` (0-based) in hist[0..nb), returns bin and rank inside it through shared scalars
__device__ void pick_bin(const unsigned* hist, int nb, unsigned long long r, unsigned* s_bin, unsigned long long* s_r,
                         unsigned* s_cnt) {
  // warp 0: lane l sums bins [l*nb/32, (l+1)*nb/32)
  if (threadIdx.x < 32) {
    const int per = nb / 32, lane = threadIdx.x;
    unsigned long long mine = 0;
    for (int j = 0; j < per; ++j) mine += hist[lane * per + j];
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned long long excl = incl - mine;
    if (r >= excl && r < incl) {  // exactly one lane
      unsigned long long c = excl;
      for (int j = 0; j < per; ++j) {
        const unsigned h = hist[lane * per + j];
        if (r < c + h) {
          *s_bin = lane * per + j;
          *s_r = r - c;
          *s_cnt = h;
          break;
        }
        c += h;
      }
    }
  }
}

constexpr int kPctCluster = 8;
constexpr int kPctUnroll = 8;

__global__ void __cluster_dims__(kPctCluster, 1, 1) __launch_bounds__(1024)
percentile_rows_kernel(const float* __restrict__ x, long long n, long long stride, double q, double add,
                       float* __restrict__ out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned hist[2048];    // this CTA's slice (read by the other CTAs of the cluster)
  __shared__ unsigned merged[2048];  // the row's histogram
  __shared__ unsigned s_bin, s_cnt, s_min;
  __shared__ unsigned long long s_r;
  const int rank = (int)cluster.block_rank();
  const int row_id = blockIdx.x / kPctCluster;
  const float* row = x + (long long)row_id * stride;
  const long long first = (long long)rank * 1024 + threadIdx.x, step = 1024LL * kPctCluster;
  const double vi = q * (double)(n - 1);
  unsigned long long k = (unsigned long long)floor(vi);
  const double t = vi - (double)k;
  unsigned prefix = 0;
  unsigned long long r = k;
  unsigned cnt_last = 0;
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const int nb = pass == 2 ? 1024 : 2048;
    const unsigned himask = pass == 0 ? 0u : (pass == 1 ? 0xffe00000u : 0xfffffc00u);
    for (int i = threadIdx.x; i < nb; i += 1024) hist[i] = 0;
    __syncthreads();
    for (long long i0 = first; i0 < n; i0 += step * kPctUnroll) {          // kPctUnroll independent loads in flight per thread
      float v[kPctUnroll];
#pragma unroll
      for (int u = 0; u < kPctUnroll; ++u) v[u] = i0 + u * step < n ? __ldg(row + i0 + u * step) : 0.f;
#pragma unroll
      for (int u = 0; u < kPctUnroll; ++u) {
        const unsigned o = f2ord(v[u]);
        if (i0 + u * step < n && (o & himask) == prefix) {
          const unsigned bin = (o >> shift) & (nb - 1);
          const unsigned peers = __match_any_sync(__activemask(), bin);   // one atomic per distinct bin of the warp
          if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
      }
    }
    cluster.sync();                                  // every slice histogram is complete and visible
    for (int i = threadIdx.x; i < nb; i += 1024) {
      unsigned sum = 0;
#pragma unroll
      for (int c = 0; c < kPctCluster; ++c) sum += cluster.map_shared_rank(hist, c)[i];
      merged[i] = sum;
    }
    cluster.sync();                                  // all remote reads of hist done before anyone zeroes it again
    pick_bin(merged, nb, r, &s_bin, &s_r, &s_cnt);
    __syncthreads();
    prefix |= s_bin << shift;
    r = s_r;
    cnt_last = s_cnt;
    __syncthreads();
  }
  const unsigned uk = prefix;  // exact k-th order statistic; r = its rank among the cnt_last copies of that value
  unsigned uk1 = uk;
  const bool need_next = t > 0.0 && r + 1 >= cnt_last;  // the (k+1)-th is the smallest value greater than uk (uniform over the cluster)
  if (need_next) {
    if (threadIdx.x == 0) s_min = 0xffffffffu;
    __syncthreads();
    unsigned m = 0xffffffffu;
    for (long long i0 = first; i0 < n; i0 += step * kPctUnroll) {
      float v[kPctUnroll];
#pragma unroll
      for (int u = 0; u < kPctUnroll; ++u) v[u] = i0 + u * step < n ? __ldg(row + i0 + u * step) : 0.f;
#pragma unroll
      for (int u = 0; u < kPctUnroll; ++u) {
        const unsigned o = f2ord(v[u]);
        if (i0 + u * step < n && o > uk && o < m) m = o;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, m);
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) {
      unsigned mm = 0xffffffffu;
      for (int c = 0; c < kPctCluster; ++c) mm = min(mm, *cluster.map_shared_rank(&s_min, c));
      uk1 = mm == 0xffffffffu ? uk : mm;
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    const double a = (double)ord2f(uk), b = (double)ord2f(uk1);
    // numpy _lerp: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5
    const double d = b - a;
    const double p = t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
    out[row_id] = (float)(p + add);
  }
  cluster.sync();   // no CTA leaves while its shared memory may still be read by rank 0
}

// ---------------------------------------------------------------------------------------------------------------
// ms_pp blend: out = (1 - norm) * d + norm * (up_mul * small[ny(y), nx(W-1-x)]),  norm = min(d / p[b], 1)
// nearest index = min((int)floorf(dst * scale), in - 1), scale = (float)in / out (ATen legacy 'nearest')
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mspp_blend_kernel(const float* __restrict__ disp, const float* __restrict__ small,
                                                         const float* __restrict__ p, float* __restrict__ out, int B, int H,
                                                         int W, int Hs, int Ws, float up_mul) {
  const long long n = (long long)B * H * W;
  const float sy = (float)Hs / H, sx = (float)Ws / W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((long long)W * H));
    const int ys = min((int)floorf(y * sy), Hs - 1);
    const int xs = min((int)floorf((W - 1 - x) * sx), Ws - 1);
    const float d = disp[i];
    const float d2 = up_mul * __ldg(small + ((long long)b * Hs + ys) * Ws + xs);
    float nrm = d / __ldg(p + b);
    nrm = nrm > 1.f ? 1.f : nrm;
    out[i] = (1.f - nrm) * d + nrm * d2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// KITTI depth errors (myUtils.py:196-232 after :234-277): per-image sums in fp64.
//   mode 0 (disps_to_depths_kitti2015): gt is a disparity map; depth = fb / (disp + (disp > 0 ? 0 : 1)) on both sides
//   mode 1 (disps_to_depths_kitti, Eigen): gt is a depth map; pred depth = fb / (disp + (disp > 0 ? 0 : 1))
// window [y0,y1) x [x0,x1) (the Eigen crop [H-219,H-4) x [44,1180); whole image for mode 0); valid where gt > 0.
// sums[b*8 + ..] = {count, abs_rel, sq_rel, sq, log_sq, n(a1), n(a2), n(a3)}
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kitti_errors_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                           double* __restrict__ sums, int H, int W, int y0, int y1, int x0,
                                                           int x1, int mode, double fb_gt, double fb_pred, double min_d,
                                                           double max_d) {
  const int b = blockIdx.y;
  const int ww = x1 - x0, hh = y1 - y0;
  const long long n = (long long)ww * hh;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = x0 + (int)(i % ww), y = y0 + (int)(i / ww);
    const long long o = ((long long)b * H + y) * W + x;
    const float g_raw = __ldg(gt + o), p_raw = __ldg(pred + o);
    if (!(g_raw > 0.f)) continue;                              // gt_mask
    double g = mode == 0 ? fb_gt / (double)g_raw : (double)g_raw;
    double p = fb_pred / ((double)p_raw + (p_raw > 0.f ? 0.0 : 1.0));
    p = p > max_d ? max_d : (p < min_d ? min_d : p);
    g = g > max_d ? max_d : (g < min_d ? min_d : g);
    const double th = fmax(g / p, p / g);
    const double diff = g - p, lg = log(g) - log(p);
    acc[0] += 1.0;
    acc[1] += fabs(diff) / g;
    acc[2] += diff * diff / g;
    acc[3] += diff * diff;
    acc[4] += lg * lg;
    acc[5] += th < 1.25 ? 1.0 : 0.0;
    acc[6] += th < 1.25 * 1.25 ? 1.0 : 0.0;
    acc[7] += th < 1.25 * 1.25 * 1.25 ? 1.0 : 0.0;
  }
  __shared__ double red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    if (v != 0.0) atomicAdd(sums + b * 8 + threadIdx.x, v);
  }
}

// realEPE (loss_functions.py:124-141,170-173): bilinear (align_corners=True) up-sampling of the 1-channel output to the
// target's size, |target - up|, masked mean over target != 0 when sparse.  sums = {sum_abs, count}
__global__ void __launch_bounds__(256) real_epe_kernel(const float* __restrict__ outp, const float* __restrict__ tgt,
                                                       double* __restrict__ sums, int B, int h, int w, int H, int W, float sy,
                                                       float sx, int sparse) {
  const long long n = (long long)B * H * W;
  double s = 0, c = 0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    const float t = __ldg(tgt + i);
    if (sparse && t == 0.f) continue;
    float v;
    if (h == H && w == W) {
      v = __ldg(outp + i);
    } else {
      const float h1r = sy * y, w1r = sx * x;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = h1 < h - 1 ? 1 : 0, w1p = w1 < w - 1 ? 1 : 0;
      const float h1l = h1r - h1, h0l = 1.f - h1l, w1l = w1r - w1, w0l = 1.f - w1l;
      const float* p = outp + b * (long long)h * w;
      v = h0l * (w0l * __ldg(p + (long long)h1 * w + w1) + w1l * __ldg(p + (long long)h1 * w + w1 + w1p)) +
          h1l * (w0l * __ldg(p + (long long)(h1 + h1p) * w + w1) + w1l * __ldg(p + (long long)(h1 + h1p) * w + w1 + w1p));
    }
    s += (double)fabsf(t - v);
    c += 1.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0 && c != 0.0) {
    atomicAdd(sums, s);
    atomicAdd(sums + 1, c);
  }
}

// get_rmse (myUtils.py:138-150): sum over all elements of (clamp((o + mean_c) * 255, 0, 255) - (l + mean_c) * 255)^2
__global__ void __launch_bounds__(256) rmse255_kernel(const float* __restrict__ o, const float* __restrict__ l,
                                                      double* __restrict__ sum, long long n, long long hw, float m0, float m1,
                                                      float m2) {
  double s = 0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)((i / hw) % 3);
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2);
    float a = (__ldg(o + i) + m) * 255.f;
    a = a > 255.f ? 255.f : (a < 0.f ? 0.f : a);
    const float bb = (__ldg(l + i) + m) * 255.f;
    const float d = a - bb;
    s += (double)(d * d);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}

// ---------------------------------------------------------------------------------------------------------------
// FAL_netA's maskR (/root/reference/models/FAL_netA.py:264): sum_n grid_sample(softmax(dlog0)_n, grid + x_of_n) with the
// grid built for align_corners=True (:231-234) but sampled with grid_sample's DEFAULT align_corners=False -- a true 2-D
// bilinear resampling (ix = ((gx + 1) * W - 1) / 2, iy = ((gy + 1) * H - 1) / 2, zero padding), clamped to <= 1.
// ATen's fp32 op order is replayed (grid_sampler_unnormalize, floor, corner weights).  Only FAL_netA with ret_subocc
// reaches this kernel; it gathers from global memory (not a roofline kernel).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maskr_noalign_kernel(const float* __restrict__ logits, const float* __restrict__ lse0,
                                                            const float* __restrict__ g0x, const float* __restrict__ g0y,
                                                            const float* __restrict__ x_of, float* __restrict__ out, int B,
                                                            int N, int H, int W, long long pitch) {
  const long long n_px = (long long)B * H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_px; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int b = (int)(i / ((long long)W * H));
    const float gy = __ldg(g0y + y);
    const float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 2.f);
    const float iy0f = floorf(iy);
    const int iy0 = (int)iy0f, iy1 = iy0 + 1;
    const float wy0 = __fadd_rn(__fadd_rn(iy0f, 1.f), -iy), wy1 = __fadd_rn(iy, -iy0f);
    const float gx0 = __ldg(g0x + x);
    float acc = 0.f;
    for (int n = 0; n < N; ++n) {
      const float gx = __fadd_rn(gx0, __ldg(x_of + b * N + n));
      const float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 2.f);
      const float ix0f = floorf(ix);
      const int ix0 = (int)ix0f, ix1 = ix0 + 1;
      const float wx0 = __fadd_rn(__fadd_rn(ix0f, 1.f), -ix), wx1 = __fadd_rn(ix, -ix0f);
      const float* L = logits + ((long long)b * N + n) * H * pitch;
      const float* S = lse0 + (long long)b * H * W;
      float v = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int yy = (t & 2) ? iy1 : iy0, xx = (t & 1) ? ix1 : ix0;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float p = __expf(__ldg(L + (long long)yy * pitch + xx) - __ldg(S + (long long)yy * W + xx));
          const float w = __fmul_rn((t & 1) ? wx1 : wx0, (t & 2) ? wy1 : wy0);
          v = __fadd_rn(v, __fmul_rn(p, w));
        }
      }
      acc += v;
    }
    out[i] = acc > 1.f ? 1.f : acc;
  }
}

int grid_for(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_flip_resize_bilinear(const float* in, float* out, int BC, int H, int W, int Ho, int Wo, int flip_x,
                                         faln_stream_t stream) {
  FALN_REQUIRE(in && out && BC > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "faln_flip_resize_bilinear: bad argument");
  const float sy = Ho > 1 ? (float)(H - 1) / (Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (Wo - 1) : 0.f;
  flip_resize_bilinear_kernel<<<grid_for((long long)BC * Ho * Wo), 256, 0, as_stream(stream)>>>(in, out, BC, H, W, Ho, Wo, sy,
                                                                                               sx, flip_x);
  return after_launch("flip_resize_bilinear_kernel");
}

extern "C" int faln_percentile_rows(const float* x, int B, long long n, long long stride, double q, double add, float* out,
                                    faln_stream_t stream) {
  FALN_REQUIRE(x && out && B > 0 && n > 0 && stride >= n && q >= 0.0 && q <= 1.0, "faln_percentile_rows: bad argument");
  percentile_rows_kernel<<<B * kPctCluster, 1024, 0, as_stream(stream)>>>(x, n, stride, q, add, out);
  return after_launch("percentile_rows_kernel");
}

extern "C" int faln_mspp_blend(const float* disp, const float* small, const float* p, float* out, int B, int H, int W, int Hs,
                               int Ws, float up_mul, faln_stream_t stream) {
  FALN_REQUIRE(disp && small && p && out && B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "faln_mspp_blend: bad argument");
  mspp_blend_kernel<<<grid_for((long long)B * H * W), 256, 0, as_stream(stream)>>>(disp, small, p, out, B, H, W, Hs, Ws, up_mul);
  return after_launch("mspp_blend_kernel");
}

extern "C" int faln_kitti_errors(const float* gt, const float* pred, double* sums, int B, int H, int W, int y0, int y1, int x0,
                                 int x1, int mode, double fb_gt, double fb_pred, double min_d, double max_d,
                                 faln_stream_t stream) {
  FALN_REQUIRE(gt && pred && sums && B > 0 && B <= 65535 && 0 <= y0 && y0 < y1 && y1 <= H && 0 <= x0 && x0 < x1 && x1 <= W &&
                   (mode == 0 || mode == 1), "faln_kitti_errors: bad window / mode");
  FALN_REQUIRE(cudaMemsetAsync(sums, 0, sizeof(double) * 8 * B, as_stream(stream)) == cudaSuccess, "faln_kitti_errors: memset");
  long long n = (long long)(y1 - y0) * (x1 - x0);
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  kitti_errors_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(gt, pred, sums, H, W, y0, y1, x0, x1, mode, fb_gt, fb_pred,
                                                                   min_d, max_d);
  return after_launch("kitti_errors_kernel");
}

extern "C" int faln_real_epe(const float* output, const float* target, double* sums, int B, int h, int w, int H, int W,
                             int sparse, faln_stream_t stream) {
  FALN_REQUIRE(output && target && sums && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "faln_real_epe: bad argument");
  FALN_REQUIRE(cudaMemsetAsync(sums, 0, sizeof(double) * 2, as_stream(stream)) == cudaSuccess, "faln_real_epe: memset");
  const float sy = H > 1 ? (float)(h - 1) / (H - 1) : 0.f;
  const float sx = W > 1 ? (float)(w - 1) / (W - 1) : 0.f;
  real_epe_kernel<<<grid_for((long long)B * H * W), 256, 0, as_stream(stream)>>>(output, target, sums, B, h, w, H, W, sy, sx,
                                                                                 sparse);
  return after_launch("real_epe_kernel");
}

extern "C" int faln_rmse255(const float* output, const float* label, double* sum, int B, int H, int W, float m0, float m1,
                            float m2, faln_stream_t stream) {
  FALN_REQUIRE(output && label && sum && B > 0 && H > 0 && W > 0, "faln_rmse255: bad argument");
  FALN_REQUIRE(cudaMemsetAsync(sum, 0, sizeof(double), as_stream(stream)) == cudaSuccess, "faln_rmse255: memset");
  const long long n = (long long)B * 3 * H * W;
  rmse255_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(output, label, sum, n, (long long)H * W, m0, m1, m2);
  return after_launch("rmse255_kernel");
}

extern "C" int faln_maskr_noalign(const float* logits, const float* lse0, const float* g0x, const float* g0y, const float* x_of,
                                  float* maskR, int B, int N, int H, int W, long long logit_pitch, faln_stream_t stream) {
  FALN_REQUIRE(logits && lse0 && g0x && g0y && x_of && maskR && B > 0 && N > 0 && H > 0 && W > 0 && logit_pitch >= W,
               "faln_maskr_noalign: bad argument");
  maskr_noalign_kernel<<<grid_for((long long)B * H * W), 256, 0, as_stream(stream)>>>(logits, lse0, g0x, g0y, x_of, maskR, B, N,
                                                                                     H, W, logit_pitch);
  return after_launch("maskr_noalign_kernel");
}
