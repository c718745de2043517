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
//     disparity / colour accumulators) with lazy rescaling; neither the probability volume nor a
//     warped image exists
//   * the horizontal sub-pixel shift is a two-tap gather out of the staged row; sample coordinates
//     replay the reference's fp32 normalised-grid arithmetic bit for bit.  The image
//     row is staged as four phase-shifted, zero-padded copies so that every tap window is an aligned
//     128-bit shared load with no bounds logic, laid out as natural register pairs for the packed
//     FFMA2/FADD2/FMUL2 pipeline (two pixels per issue slot)
//   * masks (Stage-2) need softmax normalisers of OTHER pixels of the row, so they take a second
//     sweep over the planes of the same row (re-read hits L2: the row was just streamed)
//   * backward is a single sweep: dot(x) = <g_pan(x), pan(x)> and the two log-sum-exps come from the
//     forward, and the adjoint of the zero-padded shift is a gather of the opposite shift.
#include <math.h>

#include "common.cuh"
#include "med3.h"

namespace faln {
namespace {

constexpr int kPX = 4;         // pixels per thread
constexpr int kMaxN = 128;     // planes
constexpr int kMaxW = 2048;    // 512 threads x 4 px
constexpr int kPad = 8;        // front padding (floats) of every staged row
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLazy = 64.0f;  // lazy-rescale threshold of the online softmax, log2 units: the running reference max
                                // only moves when a logit exceeds it by 2^64 (fp32 has 2^127 of head-room), i.e. on the
                                // first plane and on pathological inputs; precision is unaffected (floating point)

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
  int S;                   // ring groups (one full/empty mbarrier pair per group)
  int G;                   // plane rows per group: one barrier wait / release per G planes
  int slotf;               // floats per ring slot
  int wpad;                // floats per staged per-row array (aux rows)
  int wcopy;               // floats per image phase copy
};

struct Layout {
  int off_full, off_empty, off_tab, off_img, off_ring, off_aux;
  int total;
};

__host__ __device__ inline Layout make_layout(int S, int G, int slotf, int wpad, int wcopy, int n_aux_rows) {
  Layout l;
  int o = 0;
  l.off_full = o;
  o += S * 8;
  l.off_empty = o;
  o += S * 8;
  o = (o + 15) & ~15;
  l.off_tab = o;
  o += kMaxN * 32;  // PlaneInfo[kMaxN]
  l.off_img = o;
  o += 12 * wcopy * 4;  // 4 phases x 3 channels
  l.off_ring = o;
  o += S * G * slotf * 4;
  l.off_aux = o;
  o += n_aux_rows * wpad * 4;
  l.total = o;
  return l;
}

// ------------------------------------------------------------------------------------------ packed fp32
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 ex2_2(float2 a) { return make_float2(ex2f(a.x), ex2f(a.y)); }
__device__ __forceinline__ float max4(float2 a, float2 b) { return fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)); }

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

// Sample coordinate of pixel x on a plane with normalised offset xof -- the reference's fp32 pipeline
// (affine_grid value + offset, then ATen's ((g+1)/2)*(W-1)); the *0.5 is folded into cW = (W-1)/2,
// which is exact.  _rn intrinsics (scalar and packed) forbid FMA contraction.
__device__ __forceinline__ float coord_plus(float g0, float xof, float cW) {
  return __fmul_rn(__fadd_rn(__fadd_rn(g0, xof), 1.0f), cW);
}
__device__ __forceinline__ float coord_minus(float g0, float xof, float cW) {
  return __fmul_rn(__fadd_rn(__fsub_rn(g0, xof), 1.0f), cW);
}
// fractional tap weight of a pixel pair on a plane whose integer shift is the constant k:
//   a = ((g0 + sxof) + 1) * cW + nk + nxf       (sxof = +-xof, nk = -k resp. +(k+1), nxf = -x)
// NOTE: the product is formed with two SCALAR __fmul_rn: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2, which skips the rounding of the un-normalised coordinate the reference performs (measured: 2e-4
// relative error on pan at W = 1242).  Scalar mul.rn / add.rn are never contracted.
__device__ __forceinline__ float2 frac2(float2 g0, float2 nxf, float sxof, float cW, float nk) {
  const float2 u = add2(add2(g0, splat(sxof)), splat(1.0f));
  const float2 t = make_float2(__fmul_rn(u.x, cW), __fmul_rn(u.y, cW));
  return add2(add2(t, splat(nk)), nxf);
}
// same for the opposite shift (-xof): here t ~ x - s, so subtract x FIRST (exact, small result) and add k+1 last;
// adding k+1 to t first would round at the larger magnitude (measured: 3e-5 on maskL).
__device__ __forceinline__ float2 frac2_minus(float2 g0, float2 nxf, float xof, float cW, float k1f) {
  const float2 u = add2(add2(g0, splat(-xof)), splat(1.0f));
  const float2 t = make_float2(__fmul_rn(u.x, cW), __fmul_rn(u.y, cW));
  return add2(add2(t, nxf), splat(k1f));
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

// Stage the three image rows of (b, y) as 4 phase-shifted, zero-padded copies:
//   copy[r][c][j] = image[c][j + r] (0 beyond the row), j in [0, wcopy).
// A tap window starting at pixel x + k is then the aligned float4 at copy[k & 3][c][x + (k & ~3)].
// All global loads of a thread are issued before the first shared store (one memory latency per row, not one per
// element): each thread owns float4 #tid (+ nthr, ...) of every channel row.
__device__ __forceinline__ float4 ld_img4(const float* src, int j, int W, bool vec) {
  if (vec && j + 3 < W) return __ldg(reinterpret_cast<const float4*>(src + j));
  float4 v;
  v.x = j < W ? __ldg(src + j) : 0.f;
  v.y = j + 1 < W ? __ldg(src + j + 1) : 0.f;
  v.z = j + 2 < W ? __ldg(src + j + 2) : 0.f;
  v.w = j + 3 < W ? __ldg(src + j + 3) : 0.f;
  return v;
}
__device__ __forceinline__ void st_phases(float* copies, int c, int wcopy, int j, float4 v) {
  if (j >= wcopy + 3) return;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* dst = copies + (r * 3 + c) * wcopy;
    if (j - r >= 0 && j - r < wcopy) dst[j - r] = v.x;
    if (j + 1 - r >= 0 && j + 1 - r < wcopy) dst[j + 1 - r] = v.y;
    if (j + 2 - r >= 0 && j + 2 - r < wcopy) dst[j + 2 - r] = v.z;
    if (j + 3 - r >= 0 && j + 3 - r < wcopy) dst[j + 3 - r] = v.w;
  }
}
__device__ __forceinline__ void stage_image(float* copies, const float* img_b, int y, int H, int W, int wcopy, int tid,
                                            int nthr) {
  const int nq = (wcopy + 3 + 3) / 4;  // float4 groups covering j in [0, wcopy + 3)
  const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(img_b) & 15) == 0;
  const float* s0 = img_b + (size_t)y * W;
  const float* s1 = s0 + (size_t)H * W;
  const float* s2 = s1 + (size_t)H * W;
  for (int q0 = 0; q0 < nq; q0 += 2 * nthr) {
    const int ja = (q0 + tid) * 4, jb = (q0 + nthr + tid) * 4;
    const float4 a0 = ld_img4(s0, ja, W, vec), a1 = ld_img4(s1, ja, W, vec), a2 = ld_img4(s2, ja, W, vec);
    const float4 b0 = ld_img4(s0, jb, W, vec), b1 = ld_img4(s1, jb, W, vec), b2 = ld_img4(s2, jb, W, vec);
    st_phases(copies, 0, wcopy, ja, a0);
    st_phases(copies, 1, wcopy, ja, a1);
    st_phases(copies, 2, wcopy, ja, a2);
    st_phases(copies, 0, wcopy, jb, b0);
    st_phases(copies, 1, wcopy, jb, b1);
    st_phases(copies, 2, wcopy, jb, b2);
  }
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
// Per-plane constants of sample b, precomputed once per (CTA, b) so the hot loop reads them with two broadcast
// 128-bit shared loads instead of recomputing ~20 integer instructions per plane.
struct __align__(16) PlaneInfo {
  float xof;    // normalised-grid offset of the level
  float d;      // disparity of the level (pixels)
  float nk0f;   // -(float)k0
  float k1f;    // (float)(k0 + 1)
  int k0;       // integer part of the pixel shift (fast path)
  int special;  // 1: shift within rounding distance of an integer -> generic per-pixel path
  int pa_off;   // float offset of image phase copy (k0 & 3), channel 0
  int pb_off;   // float offset of the copy holding taps 1..4
};
using PlaneTab = PlaneInfo*;

__device__ __forceinline__ PlaneTab tab_ptrs(unsigned char* smem, const Layout& L) {
  return reinterpret_cast<PlaneInfo*>(smem + L.off_tab);
}

// Level table of sample b: integer shift k0 = floor(s), and whether the shift is so close to an
// integer that fp32 rounding of the coordinate could move floor() across it for some pixel
// ("special": handled by the per-pixel generic path).
__device__ __forceinline__ void fill_tab(const MedParams& p, PlaneTab t, int b, int tid, int nthr) {
  const float cW = 0.5f * (float)(p.W - 1);
  const float delta = 4e-7f * (float)p.W + 2e-4f;
  for (int n = tid; n < p.N; n += nthr) {
    PlaneInfo pi;
    pi.xof = __ldg(p.x_of + (size_t)b * p.N + n);
    pi.d = __ldg(p.d_lvl + (size_t)b * p.N + n);
    float s = pi.xof * cW;
    float fl = floorf(s);
    float fr = s - fl;
    bool sp = !(fr > delta && fr < 1.0f - delta) || !(s >= 0.0f) || !(s < 1.0e6f) ||
              (p.flags & FALN_MED_FORCE_GENERIC);
    pi.k0 = sp ? 0 : (int)fl;
    pi.special = sp ? 1 : 0;
    pi.nk0f = -(float)pi.k0;
    pi.k1f = (float)(pi.k0 + 1);
    const int r = pi.k0 & 3;
    pi.pa_off = (r * 3) * p.wcopy;
    pi.pb_off = (r < 3) ? ((r + 1) * 3) * p.wcopy : 4;
    t[n] = pi;
  }
}


// ---------------------------------------------------------------------------------------------
// Clean-up mode (FALN_MED_CLEANUP_, internal): the third-generation forward (med3.cu) accumulates the softmax sums
// without a running maximum and leaves the rows whose sums left the safe range to the forward kernel of this file,
// marked by a NaN in lse0[b, 0, y, 128 w] (one mark per warp of 128 pixels).  In clean-up mode a CTA first builds the list of its rows that need work and exits
// at once when there is none (the normal case: one global-load latency).
// ---------------------------------------------------------------------------------------------
constexpr unsigned FALN_MED_CLEANUP_ = 0x100u;
constexpr int kTodoWords = 32;                       // up to 1024 rows per CTA

// All threads of the CTA.  Returns false when the CTA has nothing to do.  todo: bit i = i-th row of this CTA
// (row = blockIdx.x + i * gridDim.x) needs work.
__device__ __forceinline__ bool build_todo(const MedParams& p, const float* lse0, unsigned* todo) {
  const int rows = p.B * p.H;
  if (threadIdx.x < kTodoWords) todo[threadIdx.x] = 0u;
  __syncthreads();
  bool mine = false;
  for (int i = threadIdx.x; blockIdx.x + (long long)i * gridDim.x < rows; i += blockDim.x) {
    const int row = blockIdx.x + i * gridDim.x;
    // the third-generation forward marks a row per warp: NaN over lse0 of the warp's first pixel (every 128th)
    bool need = false;
    for (int x = 0; x < p.W; x += 128) {
      const float v = __ldcg(lse0 + (size_t)row * p.W + x);
      need = need || (v != v);
    }
    if (need) {
      atomicOr(&todo[i >> 5], 1u << (i & 31));
      mine = true;
    }
  }
  return __syncthreads_or(mine) != 0;
}
__device__ __forceinline__ bool todo_bit(const unsigned* todo, int i) { return (todo[i >> 5] >> (i & 31)) & 1u; }

// Producer: stream `sweeps` x N plane rows of image row (b, y) through the ring, G rows per mbarrier group.
__device__ __forceinline__ void produce_row(const MedParams& p, unsigned char* smem, const Layout& L, int b, int y,
                                            int sweeps, int& slot, uint32_t& par) {
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_full);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + L.off_empty);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const long long plane = (long long)p.H * p.pitch;
  for (int sw = 0; sw < sweeps; ++sw) {
    long long off = ((long long)b * p.N * p.H + y) * p.pitch;
    for (int n0 = 0; n0 < p.N; n0 += p.G) {
      const int cnt = min(p.G, p.N - n0);
      while (!mbar_try_wait(&empty[slot], par ^ 1)) __nanosleep(32);
      RowLoad r[4];
      uint32_t total = 0;
      for (int q = 0; q < cnt; ++q) {
        r[q] = plan_row(p.logits, off + q * plane, p.W, p.logit_bytes);
        total += r[q].bytes;
        float* dst = ring + ((size_t)slot * p.G + q) * p.slotf + kPad;  // element i of the row lands at dst[head + i]
        for (int i = r[q].tail_from; i < p.W; ++i) dst[r[q].head + i] = __ldg(p.logits + off + q * plane + i);
      }
      if (total) mbar_arrive_expect_tx(&full[slot], total);
      else mbar_arrive(&full[slot]);
      for (int q = 0; q < cnt; ++q)
        if (r[q].bytes) bulk_g2s(ring + ((size_t)slot * p.G + q) * p.slotf + kPad, r[q].src, r[q].bytes, &full[slot]);
      off += (long long)cnt * plane;
      if (++slot == p.S) { slot = 0; par ^= 1; }
    }
  }
}

// Two-tap interpolation window for 4 consecutive pixels at integer shift k (taps xb+k .. xb+k+4), zero beyond
// the right end of the row.  abase: 16B-aligned buffer, element j of the row at abase[off0 + j].
// `interior`: warp-uniform promise that no tap of this warp leaves the row.
__device__ __forceinline__ void win_right(const float* abase, int off0, int xb, int k, int W, bool interior, float v[5]) {
  const int j0 = xb + k;
  if (interior) {
    const int idx = off0 + j0;
    load_win5(abase, idx, idx & 3, v);
    return;
  }
  if (j0 > W - 1) {
#pragma unroll
    for (int j = 0; j < 5; ++j) v[j] = 0.0f;
    return;
  }
  const int idx = off0 + j0;
  load_win5(abase, idx, idx & 3, v);
#pragma unroll
  for (int j = 1; j < 5; ++j)
    if (j0 + j > W - 1) v[j] = 0.0f;
}
// Window at a (possibly negative) start j0, zero outside [0, W-1] on both sides.
__device__ __forceinline__ void win_any(const float* abase, int off0, int j0, int W, bool interior, float v[5]) {
  if (interior) {
    const int idx = off0 + j0;
    load_win5(abase, idx, idx & 3, v);
    return;
  }
  if (j0 > W - 1 || j0 + 4 < 0) {
#pragma unroll
    for (int j = 0; j < 5; ++j) v[j] = 0.0f;
    return;
  }
  const int idx = off0 + j0;  // off0 >= kPad keeps the aligned-down address inside the buffer for j0 >= -4
  load_win5(abase, idx, idx & 3, v);
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if (j0 + j < 0 || j0 + j > W - 1) v[j] = 0.0f;
}

// image tap windows out of the phase copies: A = taps 0..3 at pa[c*wcopy], Bq = taps 1..4 at pb[c*wcopy]
struct ImgWin {
  const float* pa;
  const float* pb;
};
__device__ __forceinline__ ImgWin img_win(const float* copies, const PlaneInfo& pi, int xb, int wz) {
  ImgWin w;
  const int i0 = min(xb + (pi.k0 & ~3), wz);
  w.pa = copies + pi.pa_off + i0;
  w.pb = copies + pi.pb_off + i0;
  return w;
}

// online-softmax update of one pixel when the lazy threshold is exceeded: nm = -(running max), log2 domain
__device__ __forceinline__ float rescale_factor(float& nm, float ls) {
  const float f = ex2f(-nm - ls);  // ex2(old_max - new_max); old_max = -inf -> 0
  nm = -ls;
  return f;
}

// =============================================================================================
// Forward
// =============================================================================================
template <bool kMasks, int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med_fwd_kernel(const MedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Layout L = make_layout(p.S, p.G, p.slotf, p.wpad, p.wcopy, kMasks ? 4 : 0);
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
  __shared__ unsigned todo[kTodoWords];
  const bool cleanup = (p.flags & FALN_MED_CLEANUP_) != 0;
  if (cleanup && !build_todo(p, p.lse0, todo)) return;

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
      int slot = 0;
      uint32_t par = 0;
      for (int row = blockIdx.x, it = 0; row < rows; row += gridDim.x, ++it) {
        if (cleanup && !todo_bit(todo, it)) continue;
        produce_row(p, smem, L, row / p.H, row % p.H, kMasks ? 2 : 1, slot, par);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int xb = tid * kPX;
  const bool active = xb < W;
  const bool ragged = xb + 3 >= W;           // this thread owns pixels beyond the row end
  const float cW = 0.5f * (float)(W - 1);
  const int lane = tid & 31;
  const int warp_last_xb = ((tid | 31)) * kPX;  // xb of lane 31 of this warp
  const int wz = ((W + 3) & ~3) + 4;         // start of a guaranteed all-zero, 16B-aligned stretch of every image copy
  const int wcopy = p.wcopy;
  float g0[kPX], xf[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) {
    int x = min(xb + i, W - 1);
    g0[i] = __ldg(p.g0x + x);
    xf[i] = (float)(xb + i);
  }
  const float2 g0p[2] = {make_float2(g0[0], g0[1]), make_float2(g0[2], g0[3])};
  const float2 nxf[2] = {make_float2(-xf[0], -xf[1]), make_float2(-xf[2], -xf[3])};
  int slot = 0;
  uint32_t par = 0;
  int cur_b = -1;
  const int plane_head_step = (int)(((long long)p.H * p.pitch) & 3);

  for (int row = blockIdx.x, it = 0; row < rows; row += gridDim.x, ++it) {
    if (cleanup && !todo_bit(todo, it)) continue;
    const int b = row / p.H, y = row % p.H;
    named_bar_sync(1, ncons);  // previous row fully consumed: tables / image rows may be overwritten
    if (b != cur_b) {
      fill_tab(p, T, b, tid, ncons);
      cur_b = b;
    }
    stage_image(img, p.image + (size_t)b * 3 * p.H * W, y, p.H, W, p.wcopy, tid, ncons);
    named_bar_sync(1, ncons);
    const int head0 = (int)(((reinterpret_cast<uintptr_t>(p.logits) >> 2) + ((long long)b * N * p.H + y) * p.pitch) & 3);

    float2 nm0[2], z0[2], dacc[2], nmw[2], zw[2], pacc[3][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      nm0[h] = nmw[h] = splat(INFINITY);
      z0[h] = dacc[h] = zw[h] = splat(0.f);
      pacc[0][h] = pacc[1][h] = pacc[2][h] = splat(0.f);
    }

    // ---------------------------------------------------------------- sweep A: stats, disp, pan
    int head = head0;
    for (int n = 0, q = 0; n < N; ++n) {
      if (q == 0) mbar_wait(&full[slot], par);
      const float* sl = ring + ((size_t)slot * p.G + q) * p.slotf;  // 16B aligned
      const int off0 = kPad + head;
      if (active) {
        const PlaneInfo pi = T[n];
        const float xof = pi.xof;
        const float dn = pi.d;
        float l[kPX];
        float2 wl[2], a[2];
        float2 sc0[3][2], sc1[3][2];  // image taps x0 / x0+1 per channel, pixel pairs
        bool blend_direct = false;
        load4(sl, off0 + xb, (off0 + xb) & 3, l);
        if (ragged) {
#pragma unroll
          for (int i = 0; i < kPX; ++i)
            if (xb + i >= W) l[i] = 0.f;
        }
        if (!pi.special) {
          const int k0 = pi.k0;
          const float nk0f = pi.nk0f;
          const bool interior = warp_last_xb + k0 + 4 <= W - 1;
          float v[5];
          win_right(sl, off0, xb, k0, W, interior, v);
          a[0] = frac2(g0p[0], nxf[0], xof, cW, nk0f);
          a[1] = frac2(g0p[1], nxf[1], xof, cW, nk0f);
          wl[0] = make_float2(fmaf(a[0].x, v[1] - v[0], v[0]), fmaf(a[0].y, v[2] - v[1], v[1]));
          wl[1] = make_float2(fmaf(a[1].x, v[3] - v[2], v[2]), fmaf(a[1].y, v[4] - v[3], v[3]));
          const ImgWin iw = img_win(img, pi, xb, wz);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float4 A = *reinterpret_cast<const float4*>(iw.pa + c * wcopy);
            const float4 Bq = *reinterpret_cast<const float4*>(iw.pb + c * wcopy);
            sc0[c][0] = make_float2(A.x, A.y);
            sc0[c][1] = make_float2(A.z, A.w);
            sc1[c][0] = make_float2(Bq.x, Bq.y);
            sc1[c][1] = make_float2(Bq.z, Bq.w);
          }
        } else {
          // generic per-pixel path (shift within rounding distance of an integer): resolve the two taps here
          blend_direct = true;
          const float* lrow = sl + off0;
          float wlv[kPX], av[kPX];
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            float t = coord_plus(g0[i], xof, cW);
            float x0f = floorf(t);
            av[i] = t - x0f;
            int x0 = (int)x0f;
            float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
            wlv[i] = fmaf(av[i], f1 - f0, f0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float* ir = img + c * p.wcopy;  // phase-0 copy = the row itself
              float i0 = tap(ir, x0, W), i1 = tap(ir, x0 + 1, W);
              if (i & 1) { (&sc0[c][i >> 1].x)[1] = i0; (&sc1[c][i >> 1].x)[1] = i1; }
              else { sc0[c][i >> 1].x = i0; sc1[c][i >> 1].x = i1; }
            }
          }
          a[0] = make_float2(av[0], av[1]);
          a[1] = make_float2(av[2], av[3]);
          wl[0] = make_float2(wlv[0], wlv[1]);
          wl[1] = make_float2(wlv[2], wlv[3]);
        }
        (void)blend_direct;
        const float2 l2[2] = {make_float2(l[0], l[1]), make_float2(l[2], l[3])};
        float2 arg0[2], argw[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          arg0[h] = fma2(l2[h], splat(kLog2e), nm0[h]);
          argw[h] = fma2(wl[h], splat(kLog2e), nmw[h]);
        }
        if (fmaxf(max4(arg0[0], arg0[1]), max4(argw[0], argw[1])) > kLazy) {
          // rare: some running max must move.  Per pixel, rescale the sums that depend on it.
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float& a0 = e ? arg0[h].y : arg0[h].x;
              if (a0 > kLazy) {
                float& nm = e ? nm0[h].y : nm0[h].x;
                const float f = rescale_factor(nm, (e ? l2[h].y : l2[h].x) * kLog2e);
                (e ? z0[h].y : z0[h].x) *= f;
                (e ? dacc[h].y : dacc[h].x) *= f;
                a0 = 0.f;
              }
              float& aw = e ? argw[h].y : argw[h].x;
              if (aw > kLazy) {
                float& nm = e ? nmw[h].y : nmw[h].x;
                const float f = rescale_factor(nm, (e ? wl[h].y : wl[h].x) * kLog2e);
                (e ? zw[h].y : zw[h].x) *= f;
#pragma unroll
                for (int c = 0; c < 3; ++c) (e ? pacc[c][h].y : pacc[c][h].x) *= f;
                aw = 0.f;
              }
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // un-warped softmax + disparity expectation (reference :216-226)
          const float2 e0 = ex2_2(arg0[h]);
          z0[h] = add2(z0[h], e0);
          dacc[h] = fma2(splat(dn), e0, dacc[h]);
          // warped softmax + colour blend (reference :245-248, 279-282): pan += ew*(1-a)*I[x0] + ew*a*I[x0+1]
          const float2 ew = ex2_2(argw[h]);
          zw[h] = add2(zw[h], ew);
          const float2 w1 = mul2(ew, a[h]);
          const float2 w0 = fma2(w1, splat(-1.0f), ew);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            pacc[c][h] = fma2(w0, sc0[c][h], pacc[c][h]);
            pacc[c][h] = fma2(w1, sc1[c][h], pacc[c][h]);
          }
        }
      }
      if (++q == p.G || n == N - 1) {  // last plane of the group: hand the G slots back to the producer
        q = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == p.S) { slot = 0; par ^= 1; }
      }
      head = (head + plane_head_step) & 3;
    }

    // ---------------------------------------------------------------- row results of sweep A
    float nlse0[kPX], nlsew[kPX];  // NEGATED log2-domain log-sum-exps, reused by sweep B
    if (active) {
      float o[kPX];
      const size_t r1 = ((size_t)b * p.H + y) * W;
      const float m0v[4] = {-nm0[0].x, -nm0[0].y, -nm0[1].x, -nm0[1].y};
      const float mwv[4] = {-nmw[0].x, -nmw[0].y, -nmw[1].x, -nmw[1].y};
      const float z0v[4] = {z0[0].x, z0[0].y, z0[1].x, z0[1].y};
      const float zwv[4] = {zw[0].x, zw[0].y, zw[1].x, zw[1].y};
#pragma unroll
      for (int i = 0; i < kPX; ++i) {
        nlse0[i] = -(m0v[i] + lg2f(z0v[i]));
        nlsew[i] = -(mwv[i] + lg2f(zwv[i]));
      }
      if (p.disp) {
        const float dv[4] = {dacc[0].x, dacc[0].y, dacc[1].x, dacc[1].y};
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = dv[i] / z0v[i];
        store_row4(p.disp + r1, xb, o, W);
      }
      if (p.pan) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pv[4] = {pacc[c][0].x, pacc[c][0].y, pacc[c][1].x, pacc[c][1].y};
#pragma unroll
          for (int i = 0; i < kPX; ++i) o[i] = pv[i] / zwv[i];
          store_row4(p.pan + (((size_t)b * 3 + c) * p.H + y) * W, xb, o, W);
        }
      }
      if (p.lse0) {
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = -nlse0[i] * kLn2;
        store_row4(p.lse0 + r1, xb, o, W);
      }
      if (p.lsew) {
#pragma unroll
        for (int i = 0; i < kPX; ++i) o[i] = -nlsew[i] * kLn2;
        store_row4(p.lsew + r1, xb, o, W);
      }
    }

    if (kMasks) {
      // -------------------------------------------------------------- sweep B: occlusion masks
      const float2 nl0[2] = {make_float2(nlse0[0], nlse0[1]), make_float2(nlse0[2], nlse0[3])};
      const float2 nlw[2] = {make_float2(nlsew[0], nlsew[1]), make_float2(nlsew[2], nlsew[3])};
      float mR[kPX], mL[kPX];
#pragma unroll
      for (int i = 0; i < kPX; ++i) mR[i] = mL[i] = 0.f;
      head = head0;
      for (int n = 0, q = 0; n < N; ++n) {
        if (q == 0) mbar_wait(&full[slot], par);
        const float* sl = ring + ((size_t)slot * p.G + q) * p.slotf;
        const int off0 = kPad + head;
        float* EA = aux + (size_t)(n & 1) * 2 * p.wpad;  // softmax(L)_n        at every pixel of the row
        float* EB = EA + p.wpad;                          // softmax(warped L)_n at every pixel of the row
        const PlaneInfo pi = T[n];
        const float xof = pi.xof;
        const bool special = pi.special != 0;
        const int k0 = pi.k0;
        const bool interior = warp_last_xb + k0 + 4 <= W - 1;
        float2 ap[2];  // +shift fractional weights (fast path)
        if (active) {
          float l[kPX];
          float2 wl[2];
          load4(sl, off0 + xb, (off0 + xb) & 3, l);
          if (!special) {
            float v[5];
            win_right(sl, off0, xb, k0, W, interior, v);
            ap[0] = frac2(g0p[0], nxf[0], xof, cW, pi.nk0f);
            ap[1] = frac2(g0p[1], nxf[1], xof, cW, pi.nk0f);
            wl[0] = make_float2(fmaf(ap[0].x, v[1] - v[0], v[0]), fmaf(ap[0].y, v[2] - v[1], v[1]));
            wl[1] = make_float2(fmaf(ap[1].x, v[3] - v[2], v[2]), fmaf(ap[1].y, v[4] - v[3], v[3]));
          } else {
            const float* lrow = sl + off0;
            float wlv[kPX];
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_plus(g0[i], xof, cW);
              float x0f = floorf(t);
              float aa = t - x0f;
              int x0 = (int)x0f;
              float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
              wlv[i] = fmaf(aa, f1 - f0, f0);
            }
            wl[0] = make_float2(wlv[0], wlv[1]);
            wl[1] = make_float2(wlv[2], wlv[3]);
          }
          float2 e0[2], ew[2];
          e0[0] = ex2_2(fma2(make_float2(l[0], l[1]), splat(kLog2e), nl0[0]));
          e0[1] = ex2_2(fma2(make_float2(l[2], l[3]), splat(kLog2e), nl0[1]));
          ew[0] = ex2_2(fma2(wl[0], splat(kLog2e), nlw[0]));
          ew[1] = ex2_2(fma2(wl[1], splat(kLog2e), nlw[1]));
          if (ragged) {
            if (xb + 1 >= W) { e0[0].y = 0.f; ew[0].y = 0.f; }
            if (xb + 2 >= W) { e0[1].x = 0.f; ew[1].x = 0.f; }
            if (xb + 3 >= W) { e0[1].y = 0.f; ew[1].y = 0.f; }
          }
          *reinterpret_cast<float4*>(EA + kPad + xb) = make_float4(e0[0].x, e0[0].y, e0[1].x, e0[1].y);
          *reinterpret_cast<float4*>(EB + kPad + xb) = make_float4(ew[0].x, ew[0].y, ew[1].x, ew[1].y);
        }
        if (++q == p.G || n == N - 1) {  // the plane rows of this group are no longer needed
          q = 0;
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
          if (++slot == p.S) { slot = 0; par ^= 1; }
        }
        head = (head + plane_head_step) & 3;
        named_bar_sync(2, ncons);
        if (active) {
          if (!special) {
            float v[5];
            win_right(EA, kPad, xb, k0, W, interior, v);          // maskR: softmax(L)_n shifted by +s_n (:266)
            mR[0] += fmaf(ap[0].x, v[1] - v[0], v[0]);
            mR[1] += fmaf(ap[0].y, v[2] - v[1], v[1]);
            mR[2] += fmaf(ap[1].x, v[3] - v[2], v[2]);
            mR[3] += fmaf(ap[1].y, v[4] - v[3], v[3]);
            const bool interior_l = (tid & ~31) * kPX - k0 - 1 >= 0 && warp_last_xb - k0 + 3 <= W - 1;
            win_any(EB, kPad, xb - k0 - 1, W, interior_l, v);      // maskL: softmax(warped)_n shifted by -s_n (:270-273)
            const float2 am0 = frac2_minus(g0p[0], nxf[0], xof, cW, pi.k1f);
            const float2 am1 = frac2_minus(g0p[1], nxf[1], xof, cW, pi.k1f);
            mL[0] += fmaf(am0.x, v[1] - v[0], v[0]);
            mL[1] += fmaf(am0.y, v[2] - v[1], v[1]);
            mL[2] += fmaf(am1.x, v[3] - v[2], v[2]);
            mL[3] += fmaf(am1.y, v[4] - v[3], v[3]);
          } else {
            const float* ea = EA + kPad;
            const float* eb = EB + kPad;
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              float t = coord_plus(g0[i], xof, cW);
              float x0f = floorf(t);
              float aa = t - x0f;
              int x0 = (int)x0f;
              float f0 = tap(ea, x0, W), f1 = tap(ea, x0 + 1, W);
              mR[i] += fmaf(aa, f1 - f0, f0);
              t = coord_minus(g0[i], xof, cW);
              x0f = floorf(t);
              aa = t - x0f;
              x0 = (int)x0f;
              f0 = tap(eb, x0, W);
              f1 = tap(eb, x0 + 1, W);
              mL[i] += fmaf(aa, f1 - f0, f0);
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
template <int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med_bwd_kernel(const MedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Layout L = make_layout(p.S, p.G, p.slotf, p.wpad, p.wcopy, 6);
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
      int slot = 0;
      uint32_t par = 0;
      for (int row = blockIdx.x; row < rows; row += gridDim.x) produce_row(p, smem, L, row / p.H, row % p.H, 1, slot, par);
    }
    return;
  }

  const int xb = tid * kPX;
  const bool active = xb < W;
  const float cW = 0.5f * (float)(W - 1);
  const int lane = tid & 31;
  const int warp_first_xb = (tid & ~31) * kPX, warp_last_xb = (tid | 31) * kPX;
  const int wz = ((W + 3) & ~3) + 4;
  const int wcopy = p.wcopy;
  float g0[kPX], xf[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) {
    int x = min(xb + i, W - 1);
    g0[i] = __ldg(p.g0x + x);
    xf[i] = (float)(xb + i);
  }
  const float2 g0p[2] = {make_float2(g0[0], g0[1]), make_float2(g0[2], g0[3])};
  const float2 nxf[2] = {make_float2(-xf[0], -xf[1]), make_float2(-xf[2], -xf[3])};
  int slot = 0;
  uint32_t par = 0;
  int cur_b = -1;
  const int plane_head_step = (int)(((long long)p.H * p.pitch) & 3);

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / p.H, y = row % p.H;
    named_bar_sync(1, ncons);
    if (b != cur_b) {
      fill_tab(p, T, b, tid, ncons);
      cur_b = b;
    }
    stage_image(img, p.image + (size_t)b * 3 * p.H * W, y, p.H, W, p.wcopy, tid, ncons);
    named_bar_sync(1, ncons);
    const int head0 = (int)(((reinterpret_cast<uintptr_t>(p.logits) >> 2) + ((long long)b * N * p.H + y) * p.pitch) & 3);

    // per-pixel row constants
    float gpv[3][kPX], dotv[kPX], nlsewv[kPX], nlse0v[kPX], gdv[kPX], dspv[kPX];
    const size_t r1 = ((size_t)b * p.H + y) * W;
#pragma unroll
    for (int i = 0; i < kPX; ++i) {
      const bool ok = xb + i < W;
      const size_t x = r1 + min(xb + i, W - 1);
      dotv[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t xc = (((size_t)b * 3 + c) * p.H + y) * W + min(xb + i, W - 1);
        gpv[c][i] = (ok && p.g_pan) ? __ldg(p.g_pan + xc) : 0.f;
        dotv[i] = fmaf(gpv[c][i], p.g_pan ? __ldg(p.pan_in + xc) : 0.f, dotv[i]);
      }
      nlsewv[i] = -__ldg(p.lsew_in + x) * kLog2e;
      nlse0v[i] = -__ldg(p.lse0_in + x) * kLog2e;
      gdv[i] = (ok && p.g_disp) ? __ldg(p.g_disp + x) : 0.f;
      dspv[i] = __ldg(p.disp_in + x);
    }
    float2 gp[3][2], ndot[2], nlw[2], nl0[2], gd[2], ndsp[2], okm[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gp[c][h] = make_float2(gpv[c][2 * h], gpv[c][2 * h + 1]);
      ndot[h] = make_float2(-dotv[2 * h], -dotv[2 * h + 1]);
      nlw[h] = make_float2(nlsewv[2 * h], nlsewv[2 * h + 1]);
      nl0[h] = make_float2(nlse0v[2 * h], nlse0v[2 * h + 1]);
      gd[h] = make_float2(gdv[2 * h], gdv[2 * h + 1]);
      ndsp[h] = make_float2(-dspv[2 * h], -dspv[2 * h + 1]);
      okm[h] = make_float2(xb + 2 * h < W ? 1.f : 0.f, xb + 2 * h + 1 < W ? 1.f : 0.f);
    }

    int head = head0;
    for (int n = 0, q = 0; n < N; ++n) {
      if (q == 0) mbar_wait(&full[slot], par);
      const float* sl = ring + ((size_t)slot * p.G + q) * p.slotf;
      const long long roff = (((long long)b * N + n) * p.H + y);
      const int off0 = kPad + head;
      float* RA = aux + (size_t)(n & 1) * 3 * p.wpad;  // (1-a) * dwl  (or dwl on special planes)
      float* RB = RA + p.wpad;                          // a * dwl      (or a)
      float* RC = RB + p.wpad;                          //              (or x0 as float)
      const PlaneInfo pi = T[n];
      const float xof = pi.xof;
      const float dn = pi.d;
      const bool special = pi.special != 0;
      const int k0 = pi.k0;
      float l[kPX];
      if (active) {
        load4(sl, off0 + xb, (off0 + xb) & 3, l);
        if (!special) {
          const bool interior = warp_last_xb + k0 + 4 <= W - 1;
          float v[5];
          float2 a[2], wl[2], dP[2];
          win_right(sl, off0, xb, k0, W, interior, v);
          a[0] = frac2(g0p[0], nxf[0], xof, cW, pi.nk0f);
          a[1] = frac2(g0p[1], nxf[1], xof, cW, pi.nk0f);
          wl[0] = make_float2(fmaf(a[0].x, v[1] - v[0], v[0]), fmaf(a[0].y, v[2] - v[1], v[1]));
          wl[1] = make_float2(fmaf(a[1].x, v[3] - v[2], v[2]), fmaf(a[1].y, v[4] - v[3], v[3]));
          const ImgWin iw = img_win(img, pi, xb, wz);
          float2 oma[2] = {fma2(a[0], splat(-1.f), splat(1.f)), fma2(a[1], splat(-1.f), splat(1.f))};
          dP[0] = dP[1] = splat(0.f);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float4 A = *reinterpret_cast<const float4*>(iw.pa + c * wcopy);
            const float4 Bq = *reinterpret_cast<const float4*>(iw.pb + c * wcopy);
            // sampled colour = (1-a) I[x0] + a I[x0+1];  dP += g_pan_c * colour
            float2 s0 = fma2(a[0], make_float2(Bq.x, Bq.y), mul2(oma[0], make_float2(A.x, A.y)));
            float2 s1 = fma2(a[1], make_float2(Bq.z, Bq.w), mul2(oma[1], make_float2(A.z, A.w)));
            dP[0] = fma2(gp[c][0], s0, dP[0]);
            dP[1] = fma2(gp[c][1], s1, dP[1]);
          }
          float2 ra[2], rb[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 P = ex2_2(fma2(wl[h], splat(kLog2e), nlw[h]));
            const float2 dwl = mul2(mul2(P, add2(dP[h], ndot[h])), okm[h]);
            rb[h] = mul2(a[h], dwl);
            ra[h] = fma2(rb[h], splat(-1.f), dwl);
          }
          if (!interior) {
            // taps beyond the row end carry no gradient (zero padding)
            float* rav = &ra[0].x;  // ra[0..1] are contiguous float2s
            float* rbv = &rb[0].x;
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
              if (xb + i + k0 > W - 1) (i < 2 ? (&ra[0].x)[i] : (&ra[1].x)[i - 2]) = 0.f;
              if (xb + i + k0 + 1 > W - 1) (i < 2 ? (&rb[0].x)[i] : (&rb[1].x)[i - 2]) = 0.f;
            }
            (void)rav; (void)rbv;
          }
          *reinterpret_cast<float4*>(RA + kPad + xb) = make_float4(ra[0].x, ra[0].y, ra[1].x, ra[1].y);
          *reinterpret_cast<float4*>(RB + kPad + xb) = make_float4(rb[0].x, rb[0].y, rb[1].x, rb[1].y);
        } else {
          const float* lrow = sl + off0;
          float ra[kPX], rb[kPX], x0s[kPX];
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            const bool ok = xb + i < W;
            float t = coord_plus(g0[i], xof, cW);
            float x0f = floorf(t);
            float aa = t - x0f;
            x0s[i] = x0f;
            int x0 = (int)x0f;
            float f0 = tap(lrow, x0, W), f1 = tap(lrow, x0 + 1, W);
            float wlv = fmaf(aa, f1 - f0, f0);
            float dPv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float* ir = img + c * p.wcopy;
              float i0 = tap(ir, x0, W), i1 = tap(ir, x0 + 1, W);
              dPv = fmaf(gpv[c][i], fmaf(aa, i1 - i0, i0), dPv);
            }
            float P = ex2f(fmaf(wlv, kLog2e, nlsewv[i]));
            ra[i] = ok ? P * (dPv - dotv[i]) : 0.f;
            rb[i] = aa;
          }
          *reinterpret_cast<float4*>(RA + kPad + xb) = make_float4(ra[0], ra[1], ra[2], ra[3]);
          *reinterpret_cast<float4*>(RB + kPad + xb) = make_float4(rb[0], rb[1], rb[2], rb[3]);
          *reinterpret_cast<float4*>(RC + kPad + xb) = make_float4(x0s[0], x0s[1], x0s[2], x0s[3]);
        }
      }
      if (++q == p.G || n == N - 1) {
        q = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == p.S) { slot = 0; par ^= 1; }
      }
      head = (head + plane_head_step) & 3;
      named_bar_sync(2, ncons);
      if (active) {
        float g[kPX];
        if (!special) {
          const bool interior_g = warp_first_xb - k0 - 1 >= 0 && warp_last_xb - k0 + 3 <= W - 1;
          float va[5], vb[5];
          win_any(RA, kPad, xb - k0, W, interior_g, va);      // pixel x = j - k0 sampled tap x0 = j with weight (1-a)
          win_any(RB, kPad, xb - k0 - 1, W, interior_g, vb);  // pixel x = j - k0 - 1 sampled tap x0+1 = j with weight a
#pragma unroll
          for (int i = 0; i < kPX; ++i) g[i] = va[i] + vb[i];
        } else {
          const float* rd = RA + kPad;
          const float* rw = RB + kPad;
          const float* rx = RC + kPad;
          const int kn = (int)floorf(xof * cW);
#pragma unroll
          for (int i = 0; i < kPX; ++i) {
            const int j = xb + i;
            float acc = 0.f;
            // x0(x) - x is within +-1 of the nominal floor; scan the row window that can reach j
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
        // disparity branch: p0_n * g_disp * (d_n - disp)
        const float2 p0a = ex2_2(fma2(make_float2(l[0], l[1]), splat(kLog2e), nl0[0]));
        const float2 p0b = ex2_2(fma2(make_float2(l[2], l[3]), splat(kLog2e), nl0[1]));
        const float2 ga = fma2(mul2(p0a, gd[0]), add2(splat(dn), ndsp[0]), make_float2(g[0], g[1]));
        const float2 gb = fma2(mul2(p0b, gd[1]), add2(splat(dn), ndsp[1]), make_float2(g[2], g[3]));
        const float go[4] = {ga.x, ga.y, gb.x, gb.y};
        store_row4(p.g_logits + roff * p.g_pitch, xb, go, W);
      }
    }
  }
}

// =============================================================================================
// Disparity-only epilogue (inference): pure streaming, no staging needed.  4 pixels per thread,
// all planes of a pixel quad in flight (MLP), float4 loads when the pitch allows.
// =============================================================================================
// Streaming form (round 2): the round-1 kernel visited one plane at a time with a data-dependent rescale branch, which kept
// only ~4 loads in flight per thread (0.47-0.57 of the HBM peak).  Here a thread owns a pixel quad and walks the planes in
// chunks of eight: the eight 128-bit loads are issued back to back, the running maximum moves at most once per chunk
// (one rescale per eight planes) and every plane costs one exp2 -- the kernel is a pure HBM stream of 4 (N + 1) B/px.
__global__ void __launch_bounds__(256) med_disp_kernel(const float* __restrict__ logits, const float* __restrict__ d_lvl,
                                                       float* __restrict__ disp, int B, int N, int H, int W,
                                                       long long pitch, int vec4) {
  constexpr int kChunk = 8;
  const int wq = (W + 3) / 4;
  const long long total = (long long)B * H * wq;
  const long long ps = (long long)H * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % wq);
    const long long r = i / wq;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    const int x = xq * 4;
    const float* lp = logits + (((long long)b * N) * H + y) * pitch + x;
    const float* dl = d_lvl + (size_t)b * N;
    float m[4], z[4], acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; z[k] = 0.f; acc[k] = 0.f; }
    const bool full = vec4 && x + 3 < W;
    for (int n0 = 0; n0 < N; n0 += kChunk) {
      float v[kChunk][4];
      float dn[kChunk];
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        const int n = n0 + j;
        if (n < N) {
          if (full) {
            const float4 q = __ldcs(reinterpret_cast<const float4*>(lp + n * ps));   // streamed once: evict first
            v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[j][k] = x + k < W ? __ldg(lp + n * ps + k) : 0.f;
          }
          dn[j] = __ldg(dl + n);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[j][k] = -INFINITY;
          dn[j] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float cm = v[0][k];
#pragma unroll
        for (int j = 1; j < kChunk; ++j) cm = fmaxf(cm, v[j][k]);
        cm *= kLog2e;
        if (cm > m[k]) {                       // at most once per chunk
          const float f = ex2f(m[k] - cm);     // exp2(-inf) = 0 on the first chunk
          z[k] *= f;
          acc[k] *= f;
          m[k] = cm;
        }
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
          const float e = ex2f(fmaf(v[j][k], kLog2e, -m[k]));
          z[k] += e;
          acc[k] = fmaf(dn[j], e, acc[k]);
        }
      }
    }
    float* o = disp + ((long long)b * H + y) * W + x;
    if (x + 3 < W && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0] / z[0], acc[1] / z[1], acc[2] / z[2], acc[3] / z[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x + k < W) o[k] = acc[k] / z[k];
    }
  }
}

// =============================================================================================
// Forward, fast path (no masks): rows 16-byte aligned, pitch % 4 == 0, pad columns [W, ceil4(W)) zero.
//
// ncu on the kernel above (profiles/r1d_med_full_*): 209 warp instructions per plane-iteration of which ~65 are the
// arithmetic / loads the algorithm needs; the rest is control flow of the per-plane alignment switches, bounds logic,
// address arithmetic and parameter re-loads.  This variant removes them:
//   * the online softmax is order-independent, so each row's planes are visited grouped by the alignment class
//     R = k0 & 3 of their integer shift; the loop body is instantiated per R and every tap window is two aligned
//     128-bit shared loads whose lanes are named at compile time (no switch, no moves)
//   * every ring slot keeps an all-zero tail and the rows' pad columns are zero, so a window that leaves the row reads
//     zeros: one index clamp replaces the interior / edge branches
//   * per-plane constants, ring pointers and loop state live in registers; the producer walks the same permuted order.
// Planes whose shift is within rounding distance of an integer ("special") are visited last on the generic path.
// =============================================================================================
struct FastAcc {
  float2 nm0[2], z0[2], dacc[2], nmw[2], zw[2], pacc[3][2];
};

__device__ __forceinline__ void lazy_rescale(FastAcc& A, float2 (&arg0)[2], float2 (&argw)[2], const float2 (&l2)[2],
                                             const float2 (&wl)[2]) {
  if (fmaxf(max4(arg0[0], arg0[1]), max4(argw[0], argw[1])) > kLazy) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float& a0 = e ? arg0[h].y : arg0[h].x;
        if (a0 > kLazy) {
          float& nm = e ? A.nm0[h].y : A.nm0[h].x;
          const float f = rescale_factor(nm, (e ? l2[h].y : l2[h].x) * kLog2e);
          (e ? A.z0[h].y : A.z0[h].x) *= f;
          (e ? A.dacc[h].y : A.dacc[h].x) *= f;
          a0 = 0.f;
        }
        float& aw = e ? argw[h].y : argw[h].x;
        if (aw > kLazy) {
          float& nm = e ? A.nmw[h].y : A.nmw[h].x;
          const float f = rescale_factor(nm, (e ? wl[h].y : wl[h].x) * kLog2e);
          (e ? A.zw[h].y : A.zw[h].x) *= f;
#pragma unroll
          for (int c = 0; c < 3; ++c) (e ? A.pacc[c][h].y : A.pacc[c][h].x) *= f;
          aw = 0.f;
        }
      }
    }
  }
}

__device__ __forceinline__ void accumulate_plane(FastAcc& A, float dn, const float2 (&arg0)[2], const float2 (&argw)[2],
                                                 const float2 (&a)[2], const float2 (&sc0)[3][2], const float2 (&sc1)[3][2]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float2 e0 = ex2_2(arg0[h]);
    A.z0[h] = add2(A.z0[h], e0);
    A.dacc[h] = fma2(splat(dn), e0, A.dacc[h]);
    const float2 ew = ex2_2(argw[h]);
    A.zw[h] = add2(A.zw[h], ew);
    const float2 w1 = mul2(ew, a[h]);
    const float2 w0 = fma2(w1, splat(-1.0f), ew);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A.pacc[c][h] = fma2(w0, sc0[c][h], A.pacc[c][h]);
      A.pacc[c][h] = fma2(w1, sc1[c][h], A.pacc[c][h]);
    }
  }
}

// One plane whose integer shift k0 has (k0 & 3) == R.  `row` points at element 0 of the staged plane row (16B aligned).
template <int R>
__device__ __forceinline__ void fast_plane(FastAcc& A, const float* row, const float* img, const PlaneInfo& pi, int xb,
                                           int wr, int wz, int wcopy, float cW, const float2 (&g0p)[2],
                                           const float2 (&nxf)[2]) {
  const float4 L = *reinterpret_cast<const float4*>(row + xb);
  // tap window: elements xb + k0 .. xb + k0 + 4 = lanes R .. R + 4 of the two aligned quads at xb + k0 - R; a window that
  // starts beyond the row is redirected to the zero tail
  const int b0 = min(xb + pi.k0 - R, wr);
  const float4 w0 = *reinterpret_cast<const float4*>(row + b0);
  const float4 w1 = *reinterpret_cast<const float4*>(row + b0 + 4);
  const float q[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float2 a[2], wl[2];
  a[0] = frac2(g0p[0], nxf[0], pi.xof, cW, pi.nk0f);
  a[1] = frac2(g0p[1], nxf[1], pi.xof, cW, pi.nk0f);
  wl[0] = make_float2(fmaf(a[0].x, q[R + 1] - q[R], q[R]), fmaf(a[0].y, q[R + 2] - q[R + 1], q[R + 1]));
  wl[1] = make_float2(fmaf(a[1].x, q[R + 3] - q[R + 2], q[R + 2]), fmaf(a[1].y, q[R + 4] - q[R + 3], q[R + 3]));
  // image taps out of the phase copies: A = taps 0..3 (copy R), B = taps 1..4 (copy R + 1, or copy 0 shifted by one quad)
  const int i0 = min(xb + (pi.k0 - R), wz);
  const float* pa = img + (R * 3) * wcopy + i0;
  const float* pb = img + (R < 3 ? (R + 1) * 3 * wcopy : 4) + i0;
  float2 sc0[3][2], sc1[3][2];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 Aq = *reinterpret_cast<const float4*>(pa + c * wcopy);
    const float4 Bq = *reinterpret_cast<const float4*>(pb + c * wcopy);
    sc0[c][0] = make_float2(Aq.x, Aq.y);
    sc0[c][1] = make_float2(Aq.z, Aq.w);
    sc1[c][0] = make_float2(Bq.x, Bq.y);
    sc1[c][1] = make_float2(Bq.z, Bq.w);
  }
  const float2 l2[2] = {make_float2(L.x, L.y), make_float2(L.z, L.w)};
  float2 arg0[2], argw[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    arg0[h] = fma2(l2[h], splat(kLog2e), A.nm0[h]);
    argw[h] = fma2(wl[h], splat(kLog2e), A.nmw[h]);
  }
  lazy_rescale(A, arg0, argw, l2, wl);
  accumulate_plane(A, pi.d, arg0, argw, a, sc0, sc1);
}

// generic per-pixel path (shift within rounding distance of an integer)
__device__ __forceinline__ void special_plane(FastAcc& A, const float* row, const float* img, const PlaneInfo& pi, int xb,
                                              int W, int wcopy, float cW, const float (&g0)[kPX]) {
  const float4 L = *reinterpret_cast<const float4*>(row + xb);
  float2 a[2], wl[2], sc0[3][2], sc1[3][2];
  float wlv[kPX], av[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) {
    const float t = coord_plus(g0[i], pi.xof, cW);
    const float x0f = floorf(t);
    av[i] = t - x0f;
    const int x0 = (int)x0f;
    const float f0 = tap(row, x0, W), f1 = tap(row, x0 + 1, W);
    wlv[i] = fmaf(av[i], f1 - f0, f0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* ir = img + c * wcopy;  // phase-0 copy = the row itself
      const float i0 = tap(ir, x0, W), i1 = tap(ir, x0 + 1, W);
      if (i & 1) { sc0[c][i >> 1].y = i0; sc1[c][i >> 1].y = i1; }
      else { sc0[c][i >> 1].x = i0; sc1[c][i >> 1].x = i1; }
    }
  }
  a[0] = make_float2(av[0], av[1]);
  a[1] = make_float2(av[2], av[3]);
  wl[0] = make_float2(wlv[0], wlv[1]);
  wl[1] = make_float2(wlv[2], wlv[3]);
  const float2 l2[2] = {make_float2(L.x, L.y), make_float2(L.z, L.w)};
  float2 arg0[2], argw[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    arg0[h] = fma2(l2[h], splat(kLog2e), A.nm0[h]);
    argw[h] = fma2(wl[h], splat(kLog2e), A.nmw[h]);
  }
  lazy_rescale(A, arg0, argw, l2, wl);
  accumulate_plane(A, pi.d, arg0, argw, a, sc0, sc1);
}

// Plane visiting order of sample b: classes R = 0..3 of the ordinary planes, then the special ones.  Both the producer
// and the consumers derive it from the same table, so the ring carries the planes in exactly the order they are consumed.
__device__ __forceinline__ void build_order(const PlaneInfo* T, int N, unsigned char* ord, int* cnt) {
  int o = 0;
  for (int cls = 0; cls < 5; ++cls) {
    int c = 0;
    for (int n = 0; n < N; ++n) {
      const int k = T[n].special ? 4 : (T[n].k0 & 3);
      if (k == cls) { ord[o++] = (unsigned char)n; ++c; }
    }
    cnt[cls] = c;
  }
}

template <int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med_fwd_fast_kernel(const MedParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Layout L = make_layout(p.S, p.G, p.slotf, p.wpad, p.wcopy, 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_full);
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem + L.off_empty);
  float* img = reinterpret_cast<float*>(smem + L.off_img);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  unsigned char* ord = smem + L.off_aux;                    // [kMaxN] consumer copy, [kMaxN] producer copy
  int* cnt = reinterpret_cast<int*>(smem + L.off_aux + 2 * kMaxN);
  const PlaneTab T = tab_ptrs(smem, L);
  PlaneInfo* Tp = reinterpret_cast<PlaneInfo*>(smem + L.off_aux + 2 * kMaxN + 64);   // producer's own table

  const int ncons = blockDim.x - 32;
  const int tid = threadIdx.x;
  const int rows = p.B * p.H;
  const int W = p.W, N = p.N;
  const int wr = (W + 3) & ~3;
  const int G = p.G, S = p.S, slotf = p.slotf;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncons / 32);
    }
    fence_barrier_init();
  }
  // zero the ring once: the front pad and the tail of every slot are never written again (bulk copies cover exactly
  // [kPad, kPad + wr) of a slot), so windows that leave the row read zeros
  for (int i = tid; i < S * G * slotf; i += blockDim.x) ring[i] = 0.f;
  fence_proxy_async();
  __syncthreads();

  if (tid >= ncons) {
    // ------------------------------------------------------------------ producer warp
    if (tid == ncons) {
      unsigned char* ordp = ord + kMaxN;
      int cntp[5];
      int slot = 0, cur_b = -1;
      uint32_t par = 0;
      const long long plane = (long long)p.H * p.pitch;
      for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int b = row / p.H, y = row % p.H;
        if (b != cur_b) {
          fill_tab(p, Tp, b, 0, 1);
          build_order(Tp, N, ordp, cntp);
          cur_b = b;
        }
        const float* src0 = p.logits + ((long long)b * N * p.H + y) * p.pitch;
        for (int i0 = 0; i0 < N; i0 += G) {
          const int c = min(G, N - i0);
          while (!mbar_try_wait(&empty[slot], par ^ 1)) __nanosleep(32);
          mbar_arrive_expect_tx(&full[slot], (uint32_t)(c * wr * 4));
          for (int q = 0; q < c; ++q)
            bulk_g2s(ring + ((size_t)slot * G + q) * slotf + kPad, src0 + ordp[i0 + q] * plane, (uint32_t)(wr * 4), &full[slot]);
          if (++slot == S) { slot = 0; par ^= 1; }
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int xb = tid * kPX;
  const bool active = xb < W;
  const float cW = 0.5f * (float)(W - 1);
  const int lane = tid & 31;
  const int wz = ((W + 3) & ~3) + 4;
  const int wcopy = p.wcopy;
  float g0[kPX];
#pragma unroll
  for (int i = 0; i < kPX; ++i) g0[i] = __ldg(p.g0x + min(xb + i, W - 1));
  const float2 g0p[2] = {make_float2(g0[0], g0[1]), make_float2(g0[2], g0[3])};
  const float2 nxf[2] = {make_float2(-(float)xb, -(float)(xb + 1)), make_float2(-(float)(xb + 2), -(float)(xb + 3))};
  int slot = 0, q = 0, cur_b = -1;
  uint32_t par = 0;

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / p.H, y = row % p.H;
    named_bar_sync(1, ncons);  // previous row fully consumed: tables / image rows may be overwritten
    if (b != cur_b) {
      fill_tab(p, T, b, tid, ncons);
      named_bar_sync(1, ncons);
      if (tid == 0) build_order(T, N, ord, cnt);
      cur_b = b;
    }
    stage_image(img, p.image + (size_t)b * 3 * p.H * W, y, p.H, W, p.wcopy, tid, ncons);
    named_bar_sync(1, ncons);

    FastAcc A;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      A.nm0[h] = A.nmw[h] = splat(INFINITY);
      A.z0[h] = A.dacc[h] = A.zw[h] = splat(0.f);
      A.pacc[0][h] = A.pacc[1][h] = A.pacc[2][h] = splat(0.f);
    }

    int i = 0;
    // one step of the ring: returns the staged row of the i-th plane in visiting order (element 0), after the barrier
    // wait of its group; `release` hands a finished group back to the producer
    auto acquire = [&]() -> const float* {
      if (q == 0) mbar_wait(&full[slot], par);
      return ring + ((size_t)slot * G + q) * slotf + kPad;
    };
    auto release = [&]() {
      ++i;
      if (++q == G || i == N) {
        q = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == S) { slot = 0; par ^= 1; }
      }
    };
    const int c0 = cnt[0], c1 = cnt[1], c2 = cnt[2], c3 = cnt[3], c4 = cnt[4];
    for (int c = 0; c < c0; ++c) {
      const float* rowp = acquire();
      if (active) fast_plane<0>(A, rowp, img, T[ord[i]], xb, wr, wz, wcopy, cW, g0p, nxf);
      release();
    }
    for (int c = 0; c < c1; ++c) {
      const float* rowp = acquire();
      if (active) fast_plane<1>(A, rowp, img, T[ord[i]], xb, wr, wz, wcopy, cW, g0p, nxf);
      release();
    }
    for (int c = 0; c < c2; ++c) {
      const float* rowp = acquire();
      if (active) fast_plane<2>(A, rowp, img, T[ord[i]], xb, wr, wz, wcopy, cW, g0p, nxf);
      release();
    }
    for (int c = 0; c < c3; ++c) {
      const float* rowp = acquire();
      if (active) fast_plane<3>(A, rowp, img, T[ord[i]], xb, wr, wz, wcopy, cW, g0p, nxf);
      release();
    }
    for (int c = 0; c < c4; ++c) {
      const float* rowp = acquire();
      if (active) special_plane(A, rowp, img, T[ord[i]], xb, W, wcopy, cW, g0);
      release();
    }

    if (active) {
      float o[kPX];
      const size_t r1 = ((size_t)b * p.H + y) * W;
      const float m0v[4] = {-A.nm0[0].x, -A.nm0[0].y, -A.nm0[1].x, -A.nm0[1].y};
      const float mwv[4] = {-A.nmw[0].x, -A.nmw[0].y, -A.nmw[1].x, -A.nmw[1].y};
      const float z0v[4] = {A.z0[0].x, A.z0[0].y, A.z0[1].x, A.z0[1].y};
      const float zwv[4] = {A.zw[0].x, A.zw[0].y, A.zw[1].x, A.zw[1].y};
      if (p.disp) {
        const float dv[4] = {A.dacc[0].x, A.dacc[0].y, A.dacc[1].x, A.dacc[1].y};
#pragma unroll
        for (int k = 0; k < kPX; ++k) o[k] = dv[k] / z0v[k];
        store_row4(p.disp + r1, xb, o, W);
      }
      if (p.pan) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pv[4] = {A.pacc[c][0].x, A.pacc[c][0].y, A.pacc[c][1].x, A.pacc[c][1].y};
#pragma unroll
          for (int k = 0; k < kPX; ++k) o[k] = pv[k] / zwv[k];
          store_row4(p.pan + (((size_t)b * 3 + c) * p.H + y) * W, xb, o, W);
        }
      }
      if (p.lse0) {
#pragma unroll
        for (int k = 0; k < kPX; ++k) o[k] = (m0v[k] + lg2f(z0v[k])) * kLn2;
        store_row4(p.lse0 + r1, xb, o, W);
      }
      if (p.lsew) {
#pragma unroll
        for (int k = 0; k < kPX; ++k) o[k] = (mwv[k] + lg2f(zwv[k])) * kLn2;
        store_row4(p.lsew + r1, xb, o, W);
      }
    }
  }
}

// Measured on B200 (640-px rows): 3 CTAs/SM at 96 registers beats 2 CTAs/SM at 112 registers (0.28 vs 0.33 ms): the
// kernel is latency- rather than issue-bound, so occupancy wins.  FALN_MED_TUNE_2CTA selects the 112-register variant.
int pick_config(MedParams& p, int n_aux_rows, int* threads, int* smem_bytes, bool* narrow) {
  const int W = p.W;
  int groups = (W + kPX - 1) / kPX;
  int ncw = (groups + 31) / 32;
  *threads = (ncw + 1) * 32;
  const int wr = (W + 3) & ~3;
  p.wpad = wr + 2 * kPad;
  p.slotf = wr + 2 * kPad + 8;
  p.wcopy = wr + 16;
  // ring: S groups of G plane rows (one barrier hand-shake per group).  Prefer G = 4 with 3 groups in flight; shrink the
  // group before giving up CTAs per SM.
  *narrow = *threads <= 288 && (p.flags & 2u);
  const int per_slot = p.slotf * 4;
  const int fixed = make_layout(0, 1, p.slotf, p.wpad, p.wcopy, n_aux_rows).total;
  const int budgets[3] = {(*narrow ? 110 : (W <= 700 ? 72 : (W <= 1400 ? 110 : 220))) * 1024, 110 * 1024, 220 * 1024};
  p.S = 0;
  for (int bi = 0; bi < 3 && p.S == 0; ++bi) {
    for (int G = 4; G >= 1 && p.S == 0; G >>= 1) {
      for (int S = 3; S >= 2; --S) {
        if (fixed + S * (G * per_slot + 16) <= budgets[bi]) {
          p.S = S;
          p.G = G;
          break;
        }
      }
    }
  }
  if (p.S == 0) { p.S = 2; p.G = 1; }
  *smem_bytes = make_layout(p.S, p.G, p.slotf, p.wpad, p.wcopy, n_aux_rows).total;
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
  bool narrow;
  const bool aligned_rows = logit_pitch % 4 == 0 && (W % 4 == 0 || (flags & FALN_MED_ZERO_PAD));
  // third generation (med3.cu): barrier-free register gathers, max-free softmax sums; needs the lse outputs (they carry the
  // "recompute this row" mark) and is followed by a clean-up launch of the robust kernel of this file
  if (!(flags & (FALN_MED_FORCE_GENERIC | FALN_MED_NO_FAST | FALN_MED_NO_V3)) && aligned_rows && lse0 && lsew &&
      (long long)B * H <= (long long)sm_count() * 32 * kTodoWords) {
    m3::M3Params q{};
    q.force_generic = (flags & FALN_MED_V3_GENERIC) ? 1 : 0;
    q.logits = logits; q.image = image; q.g0x = g0x; q.x_of = x_of; q.d_lvl = d_lvl;
    q.pan = pan; q.disp = disp; q.maskL = maskL; q.maskR = maskR; q.lse0 = lse0; q.lsew = lsew;
    q.B = B; q.N = N; q.H = H; q.W = W; q.pitch = logit_pitch;
    const int rc = m3::med3_launch_fwd(q, masks, as_stream(stream));
    if (rc < 0) return rc;
    if (rc == 1) {
      p.flags |= FALN_MED_CLEANUP_;
      pick_config(p, masks ? 4 : 0, &threads, &smem, &narrow);
      auto kern = narrow ? (masks ? med_fwd_kernel<true, 288, 112> : med_fwd_kernel<false, 288, 112>)
                         : (masks ? med_fwd_kernel<true, 544, 96> : med_fwd_kernel<false, 544, 96>);
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      int grid = sm_count();
      if (grid > B * H) grid = B * H;
      kern<<<grid, threads, smem, as_stream(stream)>>>(p);
      return after_launch("med_fwd_kernel<cleanup>");
    }
  }
  // fast path: no masks, 16-byte aligned rows whose pad columns [W, ceil4(W)) are zero (FALN_MED_ZERO_PAD promise, or
  // W % 4 == 0), N <= 128
  if (!masks && !(flags & (FALN_MED_FORCE_GENERIC | FALN_MED_NO_FAST)) && aligned_rows) {
    // aux region: two order lists + counts + the producer's own plane table (320 + kMaxN * 32 bytes)
    const int aux_need = 320 + kMaxN * 32, aux_row = (((W + 3) & ~3) + 2 * kPad) * 4;
    pick_config(p, (aux_need + aux_row - 1) / aux_row, &threads, &smem, &narrow);
    auto kern = threads <= 352 ? med_fwd_fast_kernel<352, 96> : med_fwd_fast_kernel<544, 96>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count() * per_sm;
    if (grid > B * H) grid = B * H;
    kern<<<grid, threads, smem, as_stream(stream)>>>(p);
    return after_launch("med_fwd_fast_kernel");
  }
  pick_config(p, masks ? 4 : 0, &threads, &smem, &narrow);
  // (Measured: capping 1242-px rows -- 352 threads -- at 88 registers so that two CTAs fit an SM is SLOWER, 0.70 vs 0.67 ms:
  // the spills cost more than the extra warps bring.)
  auto kern = narrow ? (masks ? med_fwd_kernel<true, 288, 112> : med_fwd_kernel<false, 288, 112>)
                     : (masks ? med_fwd_kernel<true, 544, 96> : med_fwd_kernel<false, 544, 96>);
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
  bool narrow;
  const bool aligned_rows = logit_pitch % 4 == 0 && (W % 4 == 0 || (flags & FALN_MED_ZERO_PAD));
  if (!(flags & (FALN_MED_FORCE_GENERIC | FALN_MED_NO_FAST | FALN_MED_NO_V3)) && aligned_rows) {
    m3::M3Params q{};
    q.force_generic = (flags & FALN_MED_V3_GENERIC) ? 1 : 0;
    q.logits = logits; q.image = image; q.g0x = g0x; q.x_of = x_of; q.d_lvl = d_lvl;
    q.pan_in = pan; q.disp_in = disp; q.lse0_in = lse0; q.lsew_in = lsew;
    q.g_pan = g_pan; q.g_disp = g_disp; q.g_logits = g_logits; q.g_pitch = g_pitch;
    q.B = B; q.N = N; q.H = H; q.W = W; q.pitch = logit_pitch;
    const int rc = m3::med3_launch_bwd(q, as_stream(stream));
    if (rc != 0) return rc < 0 ? rc : FALN_OK;
  }
  pick_config(p, 6, &threads, &smem, &narrow);
  auto kern = narrow ? med_bwd_kernel<288, 112> : med_bwd_kernel<544, 96>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > B * H) grid = B * H;
  kern<<<grid, threads, smem, as_stream(stream)>>>(p);
  return after_launch("med_bwd_kernel");
}

extern "C" int faln_med_disp(const float* logits, const float* d_lvl, float* disp, int B, int N, int H, int W,
                             long long logit_pitch, faln_stream_t stream) {
  FALN_REQUIRE(logits && d_lvl && disp, "faln_med_disp: null pointer");
  FALN_REQUIRE(B > 0 && H > 0 && W > 0 && N >= 2, "faln_med_disp: bad shape");
  long long total = (long long)B * H * ((W + 3) / 4);
  int grid = (int)((total + 255) / 256);
  int cap = sm_count() * 16;
  if (grid > cap) grid = cap;
  const int vec4 = (logit_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  med_disp_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, d_lvl, disp, B, N, H, W, logit_pitch, vec4);
  return after_launch("med_disp_kernel");
}
