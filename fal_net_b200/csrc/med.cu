// med.cu -- fused MED view synthesis for sm_100a: forward (pan + disparity + softmax stats
// [+ sub-occlusion masks]) and backward (g_pan, g_disp -> g_logits).
//
// Replaces /root/reference/models/FAL_netB.py:216-297 (2N grid_sample launches, an O(N^2) cat, two
// softmaxes over a materialised [B,N,H,W] volume, N warped images) by ONE kernel per direction that
// streams each logit plane exactly once per sweep:
//
//   * one CTA per image row (b, y), persistent over rows; 4 consecutive pixels per thread
//   * plane rows arrive in a shared-memory ring through 1-D bulk async copies (TMA unit, UBLKCP)
//     issued by a producer warp and tracked by full/empty mbarriers, so the HBM stream never waits
//     on the math warps
//   * softmax over the planes is an online softmax held in registers (per pixel: running max, sum,
//     disparity / colour accumulators); neither the probability volume nor a warped image exists
//   * the horizontal sub-pixel shift is a two-tap gather out of the staged row; sample coordinates
//     replay the reference's fp32 normalised-grid arithmetic bit for bit
//   * masks (Stage-2) need softmax normalisers of OTHER pixels of the row, so they take a second
//     sweep over the planes of the same row (re-read hits L2: the row was just streamed)
//   * backward is a single sweep: dot(x) = <g_pan(x), pan(x)> and the two log-sum-exps come from the
//     forward, and the adjoint of the zero-padded shift is a gather of the opposite shift.
#include <math.h>

#include "common.cuh"

namespace faln {
namespace {

constexpr int kPX = 4;         // pixels per thread
constexpr int kMaxN = 128;     // planes
constexpr int kMaxW = 2048;    // 512 threads x 4 px
constexpr int kPad = 8;        // front padding (floats) of every staged row
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLazy = 8.0f;  // lazy-rescale threshold of the online softmax, log2 units

struct MedParams {
  const float* logits;
  const float* image;
  const float* g0x;
  const float* x_of;
  const float* d_lvl;
  float* pan;
  float* disp;
  float* maskL;
  float* maskR;
  float* lse0;
  float* lsew;
  // backward only
  const float* pan_in;
  const float* disp_in;
  const float* lse0_in;
  const float* lsew_in;
  const float* g_pan;
  const float* g_disp;
  float* g_logits;
  long long g_pitch;
  int B, N, H, W;
  long long pitch;         // logits row pitch, elements
  long long logit_bytes;   // bytes addressable from `logits` (for the clamped tail of the last row)
  unsigned flags;
  int S;                   // ring slots
  int slotf;               // floats per ring slot
  int wpad;                // floats per staged per-row array
};

struct Layout {
  int off_full, off_empty, off_tab, off_img, off_ring, off_aux;
  int total;
};

__host__ __device__ inline Layout make_layout(int S, int slotf, int wpad, int n_img_rows, int n_aux_rows) {
  Layout l;
  int o = 0;
  l.off_full = o;
  o += S * 8;
  l.off_empty = o;
  o += S * 8;
  o = (o + 15) & ~15;
  l.off_tab = o;
  o += kMaxN * 4 * 4;  // xof, d, k0, special
  l.off_img = o;
  o += n_img_rows * wpad * 4;
  l.off_ring = o;
  o += S * slotf * 4;
  l.off_aux = o;
  o += n_aux_rows * wpad * 4;
  l.total = o;
  return l;
}

// v[j] = base[idx + j], j = 0..4, for idx with (idx & 3) == r (r warp-uniform); base 16B aligned.
__device__ __forceinline__ void load_win5(const float* base, int idx, int r, float v[5]) {
  const float4* p = reinterpret_cast<const float4*>(base + (idx - r));
  float4 w0 = p[0], w1 = p[1];
  switch (r) {
    case 0: v[0] = w0.x; v[1] = w0.y; v[2] = w0.z; v[3] = w0.w; v[4] = w1.x; break;
    case 1: v[0] = w0.y; v[1] = w0.z; v[2] = w0.w; v[3] = w1.x; v[4] = w1.y; break;
    case 2: v[0] = w0.z; v[1] = w0.w; v[2] = w1.x; v[3] = w1.y; v[4] = w1.z; break;
    default: v[0] = w0.w; v[1] = w1.x; v[2] = w1.y; v[3] = w1.z; v[4] = w1.w; break;
  }
}
__device__ __forceinline__ void load4(const float* base, int idx, int r, float v[4]) {
  if (r == 0) {
    float4 w = *reinterpret_cast<const float4*>(base + idx);
    v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
  } else {
    float t[5];
    load_win5(base, idx, r, t);
    v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; v[3] = t[3];
  }
}

// Sample coordinate of pixel x on plane with normalised offset xof -- the reference's fp32 pipeline
// (affine_grid value + offset, then ATen's ((g+1)/2)*(W-1)); the *0.5 is folded into cW = (W-1)/2,
// which is exact.  _rn intrinsics forbid FMA contraction.
__device__ __forceinline__ float coord_plus(float g0, float xof, float cW) {
  return __fmul_rn(__fadd_rn(__fadd_rn(g0, xof), 1.0f), cW);
}
__device__ __forceinline__ float coord_minus(float g0, float xof, float cW) {
  return __fmul_rn(__fadd_rn(__fsub_rn(g0, xof), 1.0f), cW);
}

__device__ __forceinline__ float tap(const float* row, int j, int W) {
  return (j >= 0 && j <= W - 1) ? row[j] : 0.0f;
}

__device__ __forceinline__ void store_row4(float* rowp, int xb, const float v[4], int W) {
  float* p = rowp + xb;
  if (xb + 3 < W) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 15) == 0) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if ((a & 7) == 0) {
      *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(p + 2) = make_float2(v[2], v[3]);
    } else {
      p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (xb + i < W) p[i] = v[i];
  }
}

// Cooperative load of a contiguous global row of W floats into a staged row (data at +kPad).
__device__ __forceinline__ void stage_row(float* dst, const float* src, int W, int tid, int nthr) {
  for (int x = tid; x < W; x += nthr) dst[kPad + x] = __ldg(src + x);
}

struct RowLoad {
  const float* src;   // 16B-aligned-down source
  uint32_t bytes;     // multiple of 16 (possibly clamped)
  int head;           // floats between src and the first element of the row
  int tail_from;      // first row element NOT covered by the bulk copy (== W when fully covered)
};

__device__ __forceinline__ RowLoad plan_row(const float* base, long long elem_off, int W, long long total_bytes) {
  RowLoad r;
  const char* b = reinterpret_cast<const char*>(base);
  long long start = elem_off * 4;
  long long al = start & ~15LL;
  long long end = start + (long long)W * 4;
  long long end_up = (end + 15) & ~15LL;
  r.head = (int)((start - al) >> 2);
  r.tail_from = W;
  if (end_up > total_bytes) {  // never read past the tensor: bulk-copy the aligned part, patch the tail
    end_up = end & ~15LL;
    r.tail_from = (int)((end_up - start) >> 2);
    if (r.tail_from < 0) r.tail_from = 0;
  }
  r.src = reinterpret_cast<const float*>(b + al);
  r.bytes = (uint32_t)(end_up > al ? end_up - al : 0);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Shared pieces of the consumer side
// ---------------------------------------------------------------------------------------------
struct PlaneTab {
  float* xof;
  float* d;
  int* k0;
  int* special;
};

__device__ __forceinline__ PlaneTab tab_ptrs(unsigned char* smem, const Layout& L) {
  PlaneTab t;
  float* f = reinterpret_cast<float*>(smem + L.off_tab);
  t.xof = f;
  t.d = f + kMaxN;
  t.k0 = reinterpret_cast<int*>(f + 2 * kMaxN);
  t.special = reinterpret_cast<int*>(f + 3 * kMaxN);
  return t;
}

// Level table of sample b: integer shift k0 = floor(s), and whether the shift is so close to an
// integer that fp32 rounding of the coordinate could move floor() across it for some pixel
// ("special": handled by the per-pixel generic path).
__device__ __forceinline__ void fill_tab(const MedParams& p, const PlaneTab& t, int b, int tid, int nthr) {
  const float cW = 0.5f * (float)(p.W - 1);
  const float delta = 4e-7f * (float)p.W + 2e-4f;
  for (int n = tid; n < p.N; n += nthr) {
    float xo = __ldg(p.x_of + (size_t)b * p.N + n);
    t.xof[n] = xo;
    t.d[n] = __ldg(p.d_lvl + (size_t)b * p.N + n);
    float s = xo * cW;
    float fl = floorf(s);
    float fr = s - fl;
    bool sp = !(fr > delta && fr < 1.0f - delta) || !(s >= 0.0f) || !(s < 1.0e6f) ||
              (p.flags & FALN_MED_FORCE_GENERIC);
    t.k0[n] = sp ? 0 : (int)fl;
    t.special[n] = sp ? 1 : 0;
  }
}

// Producer: stream `sweeps` x N plane rows of image row (b, y) through the ring.
__device__ __forceinline__ void produce_row(const MedParams& p, unsigned char* smem, const Layout& L, int b, int y,
                                            int sweeps, uint32_t& it) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_full);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + L.off_empty);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  for (int sw = 0; sw < sweeps; ++sw) {
    for (int n = 0; n < p.N; ++n, ++it) {
      int slot = it % p.S;
      uint32_t par = (it / p.S) & 1;
      mbar_wait(&empty[slot], par ^ 1);
      long long off = (((long long)b * p.N + n) * p.H + y) * p.pitch;
      RowLoad r = plan_row(p.logits, off, p.W, p.logit_bytes);
      float* dst = ring + (size_t)slot * p.slotf + kPad;  // element i of the row lands at dst[head + i]
      for (int i = r.tail_from; i < p.W; ++i) dst[r.head + i] = __ldg(p.logits + off + i);
      if (r.bytes) {
        mbar_arrive_expect_tx(&full[slot], r.bytes);
        bulk_g2s(dst, r.src, r.bytes, &full[slot]);
      } else {
        mbar_arrive(&full[slot]);
      }
    }
  }
}

__device__ __forceinline__ int row_head(const float* base, long long elem_off) {
  return (int)(((reinterpret_cast<uintptr_t>(base) + (unsigned long long)elem_off * 4ULL) & 15ULL) >> 2);
}

// Two-tap interpolation windows for 4 consecutive pixels at integer shift k (taps xb+k .. xb+k+4),
// zero beyond the right end of the row.  `row` points at element 0 of the row inside a staged buffer
// whose address is (16B-aligned base + off0) with off0 & 3 == r0.
__device__ __forceinline__ void win_right(const float* abase, int off0, int xb, int k, int W, float v[5]) {
  int j0 = xb + k;
  if (j0 > W - 1) {
#pragma unroll
    for (int j = 0; j < 5; ++j) v[j] = 0.0f;
    return;
  }
  int idx = off0 + j0;
  load_win5(abase, idx, idx & 3, v);
  if (j0 + 4 > W - 1) {
#pragma unroll
    for (int j = 1; j < 5; ++j)
      if (j0 + j > W - 1) v[j] = 0.0f;
  }
}
// Window at a (possibly negative) start j0 = xb + k, zero outside [0, W-1] on both sides.
__device__ __forceinline__ void win_any(const float* abase, int off0, int j0, int W, float v[5]) {
  if (j0 > W - 1 || j0 + 4 < 0) {
#pragma unroll
    for (int j = 0; j < 5; ++j) v[j] = 0.0f;
    return;
  }
  int idx = off0 + j0;  // off0 >= kPad keeps the aligned-down address inside the buffer for j0 >= -4
  load_win5(abase, idx, idx & 3, v);
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if (j0 + j < 0 || j0 + j > W - 1) v[j] = 0.0f;
}

struct Softmax4 {
  float m[kPX], z[kPX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < kPX; ++i) { m[i] = -INFINITY; z[i] = 0.0f; }
  }
};

// =============================================================================================
// Forward
// =============================================================================================
template <bool kMasks>
__global__ void __launch_bounds__(544, 1) med_fwd_kernel(const MedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Layout L = make_layout(p.S, p.slotf, p.wpad, 3, kMasks ? 4 : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_full);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + L.off_empty);
  float* img = reinterpret_cast<float*>(smem + L.off_img);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  float* aux = reinterpret_cast<float*>(smem + L.off_aux);
  const PlaneTab T = tab_ptrs(smem, L);

  const int ncons = blockDim.x - 32;  // consumer threads (producer = last warp)
  const int tid = threadIdx.x;
  const int rows = p.B * p.H;
  const int W = p.W, N = p.N;

  if (tid == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncons / 32);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  if (tid >= ncons) {
    // ------------------------------------------------------------------ producer warp
    if (tid == ncons) {
      uint32_t it = 0;
      for (int row = blockIdx.x; row < rows; row += gridDim.x)
        produce_row(p, smem, L, row / p.H, row % p.H, kMasks ? 2 : 1, it);
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int xb = tid * kPX;
  const bool active = xb < W;
  const float cW = 0.5f * (float)(W - 1);
  const int lane = tid & 31;
  float g0[kPX], xf[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) {
    int x = min(xb + i, W - 1);
    g0[i] = __ldg(p.g0x + x);
    xf[i] = (float)(xb + i);
  }
  uint32_t it = 0;
  int cur_b = -1;

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / p.H, y = row % p.H;
    named_bar_sync(1, ncons);  // previous row fully consumed: tables / image rows may be overwritten
    if (b != cur_b) {
      fill_tab(p, T, b, tid, ncons);
      cur_b = b;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      stage_row(img + c * p.wpad, p.image + (((size_t)b * 3 + c) * p.H + y) * W, W, tid, ncons);
    named_bar_sync(1, ncons);

    Softmax4 s0, sw;
    s0.init();
    sw.init();
    float dacc[kPX], pacc[3][kPX];
#pragma unroll
    for (int i = 0; i < kPX; ++i) { dacc[i] = 0.f; pacc[0][i] = pacc[1][i] = pacc[2][i] = 0.f; }

    // ---------------------------------------------------------------- sweep A: stats, disp, pan
    for (int n = 0; n < N; ++n, ++it) {
      const int slot = it % p.S;
      mbar_wait(&full[slot], (it / p.S) & 1);
      const float* sl = ring + (size_t)slot * p.slotf;  // 16B aligned
      const int head = row_head(p.logits, (((long long)b * N + n) * p.H + y) * p.pitch);
      const int off0 = kPad + head;
      if (active) {
        const float xof = T.xof[n];
        const float dn = T.d[n];
        float l[kPX], wl[kPX], sc[3][kPX];
        load4(sl, off0 + xb, (off0 + xb) & 3, l);
        if (!T.special[n]) {
          const int k0 = T.k0[n];
          const float k0f = (float)k0;
          float a[kPX], v[5];
          win_right(sl, off0, xb, k0, W, v);
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            float t = coord_plus(g0[i], xof, cW);
            a[i] = __fsub_rn(__fsub_rn(t, k0f), xf[i]);
            wl[i] = fmaf(a[i], v[i + 1] - v[i], v[i]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            win_right(img + c * p.wpad, kPad, xb, k0, W, v);
#pragma unroll
            for (int i = 0; i < kPX; ++i) sc[c][i] = fmaf(a[i], v[i + 1] - v[i], v[i]);
          }
        } else {
          const float* lrow = sl + off0;
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            float t = coord_plus(g0[i], xof, cW);
            float x0f = floorf(t);
            float a = t - x0f;
            int x0 = (int)x0f;
            float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
            wl[i] = fmaf(a, f1 - f0, f0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float* ir = img + c * p.wpad + kPad;
              float i0 = tap(ir, x0, W), i1 = tap(ir, x0 + 1, W);
              sc[c][i] = fmaf(a, i1 - i0, i0);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
          // un-warped softmax + disparity expectation (reference :216-226)
          float ls = l[i] * kLog2e;
          if (ls > s0.m[i] + kLazy) {
            float f = ex2f(s0.m[i] - ls);
            s0.z[i] *= f;
            dacc[i] *= f;
            s0.m[i] = ls;
          }
          float e = ex2f(ls - s0.m[i]);
          s0.z[i] += e;
          dacc[i] = fmaf(dn, e, dacc[i]);
          // warped softmax + colour blend (reference :245-248, 279-282)
          float ws = wl[i] * kLog2e;
          if (ws > sw.m[i] + kLazy) {
            float f = ex2f(sw.m[i] - ws);
            sw.z[i] *= f;
            pacc[0][i] *= f; pacc[1][i] *= f; pacc[2][i] *= f;
            sw.m[i] = ws;
          }
          float ew = ex2f(ws - sw.m[i]);
          sw.z[i] += ew;
          pacc[0][i] = fmaf(sc[0][i], ew, pacc[0][i]);
          pacc[1][i] = fmaf(sc[1][i], ew, pacc[1][i]);
          pacc[2][i] = fmaf(sc[2][i], ew, pacc[2][i]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }

    // ---------------------------------------------------------------- row results of sweep A
    float lse0s[kPX], lsews[kPX];  // log2-domain log-sum-exps, reused by sweep B
    if (active) {
      float o[kPX];
      const size_t r1 = ((size_t)b * p.H + y) * W;
#pragma unroll
      for (int i = 0; i < kPX; ++i) {
        lse0s[i] = s0.m[i] + lg2f(s0.z[i]);
        lsews[i] = sw.m[i] + lg2f(sw.z[i]);
      }
      if (p.disp) {
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = dacc[i] / s0.z[i];
        store_row4(p.disp + r1, xb, o, W);
      }
      if (p.pan) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int i = 0; i < kPX; ++i) o[i] = pacc[c][i] / sw.z[i];
          store_row4(p.pan + (((size_t)b * 3 + c) * p.H + y) * W, xb, o, W);
        }
      }
      if (p.lse0) {
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = lse0s[i] * kLn2;
        store_row4(p.lse0 + r1, xb, o, W);
      }
      if (p.lsew) {
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = lsews[i] * kLn2;
        store_row4(p.lsew + r1, xb, o, W);
      }
    }

    if (kMasks) {
      // -------------------------------------------------------------- sweep B: occlusion masks
      float mR[kPX], mL[kPX];
#pragma unroll
      for (int i = 0; i < kPX; ++i) mR[i] = mL[i] = 0.f;
      for (int n = 0; n < N; ++n, ++it) {
        const int slot = it % p.S;
        mbar_wait(&full[slot], (it / p.S) & 1);
        const float* sl = ring + (size_t)slot * p.slotf;
        const int head = row_head(p.logits, (((long long)b * N + n) * p.H + y) * p.pitch);
        const int off0 = kPad + head;
        float* EA = aux + (size_t)(n & 1) * 2 * p.wpad;  // softmax(L)_n        at every pixel of the row
        float* EB = EA + p.wpad;                          // softmax(warped L)_n at every pixel of the row
        const float xof = T.xof[n];
        const bool special = T.special[n] != 0;
        const int k0 = T.k0[n];
        float ap[kPX];  // +shift fractional weights (fast path)
        if (active) {
          float l[kPX], wl[kPX], e0[kPX], ew[kPX];
          load4(sl, off0 + xb, (off0 + xb) & 3, l);
          if (!special) {
            const float k0f = (float)k0;
            float v[5];
            win_right(sl, off0, xb, k0, W, v);
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_plus(g0[i], xof, cW);
              ap[i] = __fsub_rn(__fsub_rn(t, k0f), xf[i]);
              wl[i] = fmaf(ap[i], v[i + 1] - v[i], v[i]);
            }
          } else {
            const float* lrow = sl + off0;
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_plus(g0[i], xof, cW);
              float x0f = floorf(t);
              float a = t - x0f;
              int x0 = (int)x0f;
              float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
              wl[i] = fmaf(a, f1 - f0, f0);
            }
          }
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            bool ok = xb + i < W;
            e0[i] = ok ? ex2f(fmaf(l[i], kLog2e, -lse0s[i])) : 0.f;
            ew[i] = ok ? ex2f(fmaf(wl[i], kLog2e, -lsews[i])) : 0.f;
          }
          *reinterpret_cast<float4*>(EA + kPad + xb) = make_float4(e0[0], e0[1], e0[2], e0[3]);
          *reinterpret_cast<float4*>(EB + kPad + xb) = make_float4(ew[0], ew[1], ew[2], ew[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);  // the plane row itself is no longer needed
        named_bar_sync(2, ncons);
        if (active) {
          if (!special) {
            float v[5];
            win_right(EA, kPad, xb, k0, W, v);          // maskR: softmax(L)_n shifted by +s_n (:266)
#pragma unroll
            for (int i = 0; i < kPX; ++i) mR[i] += fmaf(ap[i], v[i + 1] - v[i], v[i]);
            const float k1f = (float)(k0 + 1);
            win_any(EB, kPad, xb - k0 - 1, W, v);        // maskL: softmax(warped)_n shifted by -s_n (:270-273)
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_minus(g0[i], xof, cW);
              float a = __fadd_rn(__fsub_rn(t, xf[i]), k1f);
              mL[i] += fmaf(a, v[i + 1] - v[i], v[i]);
            }
          } else {
            const float* ea = EA + kPad;
            const float* eb = EB + kPad;
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_plus(g0[i], xof, cW);
              float x0f = floorf(t);
              float a = t - x0f;
              int x0 = (int)x0f;
              float f0 = tap(ea, x0, W), f1 = tap(ea, x0 + 1, W);
              mR[i] += fmaf(a, f1 - f0, f0);
              t = coord_minus(g0[i], xof, cW);
              x0f = floorf(t);
              a = t - x0f;
              x0 = (int)x0f;
              f0 = tap(eb, x0, W);
              f1 = tap(eb, x0 + 1, W);
              mL[i] += fmaf(a, f1 - f0, f0);
            }
          }
        }
      }
      if (active) {
        const size_t r1 = ((size_t)b * p.H + y) * W;
        float o[kPX];
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = fminf(mL[i], 1.0f);
        store_row4(p.maskL + r1, xb, o, W);
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = fminf(mR[i], 1.0f);
        store_row4(p.maskR + r1, xb, o, W);
      }
    }
  }
}

// =============================================================================================
// Backward
// =============================================================================================
__global__ void __launch_bounds__(544, 1) med_bwd_kernel(const MedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Layout L = make_layout(p.S, p.slotf, p.wpad, 3, 6);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_full);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + L.off_empty);
  float* img = reinterpret_cast<float*>(smem + L.off_img);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  float* aux = reinterpret_cast<float*>(smem + L.off_aux);
  const PlaneTab T = tab_ptrs(smem, L);

  const int ncons = blockDim.x - 32;
  const int tid = threadIdx.x;
  const int rows = p.B * p.H;
  const int W = p.W, N = p.N;

  if (tid == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncons / 32);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  if (tid >= ncons) {
    if (tid == ncons) {
      uint32_t it = 0;
      for (int row = blockIdx.x; row < rows; row += gridDim.x) produce_row(p, smem, L, row / p.H, row % p.H, 1, it);
    }
    return;
  }

  const int xb = tid * kPX;
  const bool active = xb < W;
  const float cW = 0.5f * (float)(W - 1);
  const int lane = tid & 31;
  float g0[kPX], xf[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) {
    int x = min(xb + i, W - 1);
    g0[i] = __ldg(p.g0x + x);
    xf[i] = (float)(xb + i);
  }
  uint32_t it = 0;
  int cur_b = -1;

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / p.H, y = row % p.H;
    named_bar_sync(1, ncons);
    if (b != cur_b) {
      fill_tab(p, T, b, tid, ncons);
      cur_b = b;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      stage_row(img + c * p.wpad, p.image + (((size_t)b * 3 + c) * p.H + y) * W, W, tid, ncons);
    named_bar_sync(1, ncons);

    // per-pixel row constants
    float gp[3][kPX], dot[kPX], lsews[kPX], lse0s[kPX], gd[kPX], dsp[kPX];
    const size_t r1 = ((size_t)b * p.H + y) * W;
#pragma unroll
    for (int i = 0; i < kPX; ++i) {
      const bool ok = xb + i < W;
      const size_t x = r1 + min(xb + i, W - 1);
      dot[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t xc = (((size_t)b * 3 + c) * p.H + y) * W + min(xb + i, W - 1);
        gp[c][i] = (ok && p.g_pan) ? __ldg(p.g_pan + xc) : 0.f;
        dot[i] = fmaf(gp[c][i], p.g_pan ? __ldg(p.pan_in + xc) : 0.f, dot[i]);
      }
      lsews[i] = __ldg(p.lsew_in + x) * kLog2e;
      lse0s[i] = __ldg(p.lse0_in + x) * kLog2e;
      gd[i] = (ok && p.g_disp) ? __ldg(p.g_disp + x) : 0.f;
      dsp[i] = __ldg(p.disp_in + x);
    }

    for (int n = 0; n < N; ++n, ++it) {
      const int slot = it % p.S;
      mbar_wait(&full[slot], (it / p.S) & 1);
      const float* sl = ring + (size_t)slot * p.slotf;
      const long long roff = (((long long)b * N + n) * p.H + y);
      const int head = row_head(p.logits, roff * p.pitch);
      const int off0 = kPad + head;
      float* RA = aux + (size_t)(n & 1) * 3 * p.wpad;  // (1-a) * dwl  (or dwl on special planes)
      float* RB = RA + p.wpad;                          // a * dwl      (or a)
      float* RC = RB + p.wpad;                          //              (or x0 as float)
      const float xof = T.xof[n];
      const float dn = T.d[n];
      const bool special = T.special[n] != 0;
      const int k0 = T.k0[n];
      float l[kPX];
      if (active) {
        float wl[kPX], sc[3][kPX], a[kPX], x0s[kPX];
        load4(sl, off0 + xb, (off0 + xb) & 3, l);
        if (!special) {
          const float k0f = (float)k0;
          float v[5];
          win_right(sl, off0, xb, k0, W, v);
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            float t = coord_plus(g0[i], xof, cW);
            a[i] = __fsub_rn(__fsub_rn(t, k0f), xf[i]);
            wl[i] = fmaf(a[i], v[i + 1] - v[i], v[i]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            win_right(img + c * p.wpad, kPad, xb, k0, W, v);
#pragma unroll
            for (int i = 0; i < kPX; ++i) sc[c][i] = fmaf(a[i], v[i + 1] - v[i], v[i]);
          }
        } else {
          const float* lrow = sl + off0;
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            float t = coord_plus(g0[i], xof, cW);
            float x0f = floorf(t);
            a[i] = t - x0f;
            x0s[i] = x0f;
            int x0 = (int)x0f;
            float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
            wl[i] = fmaf(a[i], f1 - f0, f0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float* ir = img + c * p.wpad + kPad;
              float i0 = tap(ir, x0, W), i1 = tap(ir, x0 + 1, W);
              sc[c][i] = fmaf(a[i], i1 - i0, i0);
            }
          }
        }
        float ra[kPX], rb[kPX];
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
          const bool ok = xb + i < W;
          float P = ex2f(fmaf(wl[i], kLog2e, -lsews[i]));
          float dP = fmaf(gp[0][i], sc[0][i], fmaf(gp[1][i], sc[1][i], gp[2][i] * sc[2][i]));
          float dwl = ok ? P * (dP - dot[i]) : 0.f;
          if (!special) {
            // taps beyond the row end carry no gradient (zero padding)
            ra[i] = (xb + i + k0 <= W - 1) ? (1.0f - a[i]) * dwl : 0.f;
            rb[i] = (xb + i + k0 + 1 <= W - 1) ? a[i] * dwl : 0.f;
          } else {
            ra[i] = dwl;
            rb[i] = a[i];
          }
        }
        *reinterpret_cast<float4*>(RA + kPad + xb) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        *reinterpret_cast<float4*>(RB + kPad + xb) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        if (special) *reinterpret_cast<float4*>(RC + kPad + xb) = make_float4(x0s[0], x0s[1], x0s[2], x0s[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      named_bar_sync(2, ncons);
      if (active) {
        float g[kPX];
        if (!special) {
          float va[5], vb[5];
          win_any(RA, kPad, xb - k0, W, va);      // pixel x = j - k0 sampled tap x0 = j with weight (1-a)
          win_any(RB, kPad, xb - k0 - 1, W, vb);  // pixel x = j - k0 - 1 sampled tap x0+1 = j with weight a
#pragma unroll
          for (int i = 0; i < kPX; ++i) g[i] = va[i] + vb[i];
        } else {
          const float* rd = RA + kPad;
          const float* rw = RB + kPad;
          const float* rx = RC + kPad;
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            const int j = xb + i;
            float acc = 0.f;
            // x0(x) - x is within +-1 of the nominal floor; scan the row window that can reach j
            const int kn = (int)floorf(xof * cW);
            for (int x = j - kn - 2; x <= j - kn + 1; ++x) {
              if (x < 0 || x > W - 1) continue;
              const int x0 = (int)rx[x];
              const float aa = rw[x], dw = rd[x];
              if (x0 == j) acc += (1.0f - aa) * dw;
              if (x0 + 1 == j) acc += aa * dw;
            }
            g[i] = acc;
          }
        }
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
          float p0 = ex2f(fmaf(l[i], kLog2e, -lse0s[i]));
          g[i] = fmaf(p0 * gd[i], dn - dsp[i], g[i]);
        }
        store_row4(p.g_logits + roff * p.g_pitch, xb, g, W);
      }
    }
  }
}

// =============================================================================================
// Disparity-only epilogue (inference): pure streaming, no staging needed.
// =============================================================================================
__global__ void __launch_bounds__(256) med_disp_kernel(const float* __restrict__ logits, const float* __restrict__ d_lvl,
                                                       float* __restrict__ disp, int B, int N, int H, int W,
                                                       long long pitch) {
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const float* lp = logits + (((long long)b * N) * H + y) * pitch + x;
    const long long ps = (long long)H * pitch;
    float m = -INFINITY, z = 0.f, acc = 0.f;
#pragma unroll 7
    for (int n = 0; n < N; ++n) {
      float ls = __ldg(lp + n * ps) * kLog2e;
      if (ls > m + kLazy) {
        float f = ex2f(m - ls);
        z *= f;
        acc *= f;
        m = ls;
      }
      float e = ex2f(ls - m);
      z += e;
      acc = fmaf(__ldg(d_lvl + b * N + n), e, acc);
    }
    disp[i] = acc / z;
  }
}

int pick_config(MedParams& p, int n_aux_rows, int* threads, int* smem_bytes) {
  const int W = p.W;
  int groups = (W + kPX - 1) / kPX;
  int ncw = (groups + 31) / 32;
  *threads = (ncw + 1) * 32;
  p.wpad = ((W + 3) & ~3) + 2 * kPad;
  p.slotf = ((W + 3) & ~3) + 2 * kPad + 8;
  // ring depth: enough bytes in flight per SM to cover HBM latency (~64 KB/SM), within ~200 KB
  int per_slot = p.slotf * 4;
  int fixed = make_layout(0, p.slotf, p.wpad, 3, n_aux_rows).total;
  int S = 12;
  while (S > 3 && fixed + S * (per_slot + 16) > 100 * 1024) --S;
  p.S = S;
  *smem_bytes = make_layout(S, p.slotf, p.wpad, 3, n_aux_rows).total;
  return 0;
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_med_fwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                            const float* d_lvl, float* pan, float* disp, float* maskL, float* maskR, float* lse0,
                            float* lsew, int B, int N, int H, int W, long long logit_pitch, unsigned flags,
                            faln_stream_t stream) {
  FALN_REQUIRE(logits && image && g0x && x_of && d_lvl, "faln_med_fwd: null input pointer");
  FALN_REQUIRE(B > 0 && H > 0 && N >= 2 && N <= kMaxN, "faln_med_fwd: need 2 <= N <= %d (got %d)", kMaxN, N);
  FALN_REQUIRE(W >= 4 && W <= kMaxW, "faln_med_fwd: need 4 <= W <= %d (got %d)", kMaxW, W);
  FALN_REQUIRE(logit_pitch >= W, "faln_med_fwd: logit_pitch < W");
  FALN_REQUIRE((maskL == nullptr) == (maskR == nullptr), "faln_med_fwd: maskL and maskR go together");
  FALN_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "faln_med_fwd: logits must be 16-byte aligned");
  MedParams p{};
  p.logits = logits; p.image = image; p.g0x = g0x; p.x_of = x_of; p.d_lvl = d_lvl;
  p.pan = pan; p.disp = disp; p.maskL = maskL; p.maskR = maskR; p.lse0 = lse0; p.lsew = lsew;
  p.B = B; p.N = N; p.H = H; p.W = W; p.pitch = logit_pitch; p.flags = flags;
  p.logit_bytes = (((long long)B * N * H - 1) * logit_pitch + W) * 4;
  const bool masks = maskL != nullptr;
  int threads, smem;
  pick_config(p, masks ? 4 : 0, &threads, &smem);
  auto kern = masks ? med_fwd_kernel<true> : med_fwd_kernel<false>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > B * H) grid = B * H;
  kern<<<grid, threads, smem, as_stream(stream)>>>(p);
  return after_launch("med_fwd_kernel");
}

extern "C" int faln_med_bwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                            const float* d_lvl, const float* pan, const float* disp, const float* lse0,
                            const float* lsew, const float* g_pan, const float* g_disp, float* g_logits, int B,
                            int N, int H, int W, long long logit_pitch, long long g_pitch, unsigned flags,
                            faln_stream_t stream) {
  FALN_REQUIRE(logits && image && g0x && x_of && d_lvl && pan && disp && lse0 && lsew && g_logits,
               "faln_med_bwd: null pointer");
  FALN_REQUIRE(B > 0 && H > 0 && N >= 2 && N <= kMaxN, "faln_med_bwd: need 2 <= N <= %d (got %d)", kMaxN, N);
  FALN_REQUIRE(W >= 4 && W <= kMaxW, "faln_med_bwd: need 4 <= W <= %d (got %d)", kMaxW, W);
  FALN_REQUIRE(logit_pitch >= W && g_pitch >= W, "faln_med_bwd: pitch < W");
  FALN_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "faln_med_bwd: logits must be 16-byte aligned");
  MedParams p{};
  p.logits = logits; p.image = image; p.g0x = g0x; p.x_of = x_of; p.d_lvl = d_lvl;
  p.pan_in = pan; p.disp_in = disp; p.lse0_in = lse0; p.lsew_in = lsew;
  p.g_pan = g_pan; p.g_disp = g_disp; p.g_logits = g_logits; p.g_pitch = g_pitch;
  p.B = B; p.N = N; p.H = H; p.W = W; p.pitch = logit_pitch; p.flags = flags;
  p.logit_bytes = (((long long)B * N * H - 1) * logit_pitch + W) * 4;
  int threads, smem;
  pick_config(p, 6, &threads, &smem);
  cudaFuncSetAttribute(med_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, med_bwd_kernel, threads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > B * H) grid = B * H;
  med_bwd_kernel<<<grid, threads, smem, as_stream(stream)>>>(p);
  return after_launch("med_bwd_kernel");
}

extern "C" int faln_med_disp(const float* logits, const float* d_lvl, float* disp, int B, int N, int H, int W,
                             long long logit_pitch, faln_stream_t stream) {
  FALN_REQUIRE(logits && d_lvl && disp, "faln_med_disp: null pointer");
  FALN_REQUIRE(B > 0 && H > 0 && W > 0 && N >= 2, "faln_med_disp: bad shape");
  long long total = (long long)B * H * W;
  int grid = (int)((total + 255) / 256);
  int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  med_disp_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, d_lvl, disp, B, N, H, W, logit_pitch);
  return after_launch("med_disp_kernel");
}
