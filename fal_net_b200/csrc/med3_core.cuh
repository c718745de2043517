// med3_core.cuh -- per-thread arithmetic of the third-generation MED kernels (med3.cu).
//
// Everything in this header is a pure function of "shared-memory" row pointers and registers, written so that it
// compiles both for sm_100a (nvcc) and for the host (g++, tests/host_emu/med3_emu.cpp).  The host build runs the
// same window / clamp / class logic thread by thread on the CPU and is compared with the oracle in the CPU test
// suite (tests/test_med3_emu.py), so the index arithmetic of the kernels is checked before it reaches a GPU.  The
// host build is TEST INFRASTRUCTURE: nothing in the product calls it.
//
// Math (reference /root/reference/models/FAL_netB.py:216-297, closed forms in SURVEY.md A.2/A.3); plane n of sample b
// has the pixel shift s_n = x_of_n * (W-1)/2 = k0 + frac, k0 = floor(s_n):
//   wl_n(x)   = lerp(L_n(x+k0), L_n(x+k0+1); a+(x))            warped logit,  a+(x) = fp32 replay of the sample coordinate
//   P_n(x)    = exp(wl_n(x)) / Zw(x),  Q_n(x) = exp(L_n(x)) / Z0(x)
//   disp(x)   = sum_n d_n Q_n(x)                                pan(x) = sum_n P_n(x) lerp(I(x+k0), I(x+k0+1); a+(x))
//   maskR(x)  = min(1, sum_n lerp(Q_n(x+k0), Q_n(x+k0+1); a+(x)))
//   maskL(x)  = min(1, sum_n lerp(P_n(x-k0-1), P_n(x-k0); a-(x)))
//   backward  g_n(j) = Q_n(j) g_disp(j) (d_n - disp(j)) + (1-a+(j-k0)) G_n(j-k0) + a+(j-k0-1) G_n(j-k0-1),
//             G_n(x) = P_n(x) (<g_pan(x), colour_n(x)> - <g_pan(x), pan(x)>)
// Taps outside [0, W-1] are zero (grid_sample zero padding).
//
// Design points (B200):
//   * planes are visited grouped by the alignment class R = k0 & 3 of their integer shift, so every five-tap window is
//     two aligned 128-bit shared loads whose lanes are named at compile time
//   * every staged row has zero (logits, g_pan, dot) or -inf (log-sum-exp rows) padding on both sides, and windows that
//     leave the row are redirected into the padding by ONE index clamp: no bounds branches in the plane loops
//   * masks and the backward are GATHERS computed in registers: a thread re-derives the probabilities of the (shifted)
//     pixels it needs from the un-shifted logit window it owns and the staged normaliser rows, so the plane loops contain
//     no CTA barrier and no shared-memory exchange (the second-generation kernels synchronised the CTA once per plane)
//   * the softmax sums are accumulated WITHOUT a running maximum (plain exp2 of the logit): a row whose sums leave
//     [2^-100, 2^120] is flagged (per warp of 128 pixels) and recomputed by the robust second-generation kernel
//     (med.cu, clean-up launch)
#pragma once

#if defined(__CUDACC__)
#include "common.cuh"
#define M3_FN __device__ __forceinline__
#define M3_HD __host__ __device__ inline
// 1 / x to 1 ulp (MUFU.RCP); the sums it divides are in [2^-100, 2^120]
__device__ __forceinline__ float m3_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#else
// ------------------------------------------------------------------------------------------ host shims
#include <math.h>
#include <stdint.h>
#include <string.h>
#define M3_FN static inline
#define M3_HD static inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
// compiled with -ffp-contract=off: plain operators are single IEEE operations
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
namespace faln {
static inline float ex2f(float x) { return exp2f(x); }
static inline float lg2f(float x) { return log2f(x); }
}  // namespace faln
static inline float m3_rcp(float x) { return 1.0f / x; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#endif

namespace faln {
namespace m3 {

constexpr int kPX = 4;        // pixels per thread
constexpr int kMaxN = 128;    // planes
constexpr int kMaxW = 2048;
constexpr int kPad = 8;       // floats of padding in front of every staged row
constexpr int kTail = 16;     // floats of padding behind the ceil4(W) payload of a ring slot
constexpr int kRowBack = 8;   // floats of padding behind the payload of a per-row array
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kZLo = 7.888609052210118e-31f;   // 2^-100
constexpr float kZHi = 1.329227995784916e+36f;   // 2^120

// One plane of the per-sample table, sorted by class (R = 0..3 ordinary, 4 = "special": shift within rounding distance
// of an integer, handled by the second-generation kernels).
struct __attribute__((aligned(16))) Ent {
  float xof;    // normalised-grid offset of the level
  float d;      // disparity of the level (pixels)
  float nk0f;   // -(float)k0
  int woff;     // k0 - R: quad-aligned part of the integer shift
  float k1f;    // (float)(k0 + 1)
  int src;      // original plane index n
  int cls;      // 0..3 = R, 4 = special
  int rsv;
};

M3_HD int ceil4(int w) { return (w + 3) & ~3; }
M3_HD int slot_floats(int W) { return ceil4(W) + kPad + kTail; }
M3_HD int row_floats(int W) { return ceil4(W) + kPad + kRowBack; }
// Image row staging: one record of kImgRec floats per pixel quad Q, holding the twelve float4
//   rec[(r*3 + c)*4 .. +3] = I_c[4Q + r .. 4Q + r + 3]          (r = phase 0..3, c = channel)
// so every tap window of every alignment class is an aligned 128-bit load at a COMPILE-TIME offset from one record
// pointer.  52 = 48 + 4 floats of padding: consecutive threads read consecutive records, and a 208-byte stride spreads the
// eight threads of a quarter-warp over all 32 banks (a 192-byte stride would be a 4-way conflict).
constexpr int kImgRec = 52;
M3_HD int img_floats(int W) { return (ceil4(W) / 4 + 4) * kImgRec; }

// ------------------------------------------------------------------------------------------ packed fp32
M3_FN float2 splat(float v) { return make_float2(v, v); }
M3_FN float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
M3_FN float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
M3_FN float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
M3_FN float2 ex2_2(float2 a) { return make_float2(ex2f(a.x), ex2f(a.y)); }
M3_FN float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
M3_FN void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Level table entry of plane n (same criterion as med.cu fill_tab: the shift is "special" when fp32 rounding of the
// sample coordinate could move floor() across an integer for some pixel of the row).
M3_FN Ent make_ent(float xof, float d, int n, int W, bool force_special) {
  const float cW = 0.5f * (float)(W - 1);
  const float delta = 4e-7f * (float)W + 2e-4f;
  Ent e;
  e.xof = xof;
  e.d = d;
  const float s = xof * cW;
  const float fl = floorf(s);
  const float fr = s - fl;
  const bool sp = !(fr > delta && fr < 1.0f - delta) || !(s >= 0.0f) || !(s < 1.0e6f) || force_special;
  const int k0 = sp ? 0 : (int)fl;
  e.cls = sp ? 4 : (k0 & 3);
  e.nk0f = -(float)k0;
  e.k1f = (float)(k0 + 1);
  e.woff = k0 - (k0 & 3);
  e.src = n;
  e.rsv = 0;
  return e;
}
// Position of plane n in the class-sorted order (stable within a class).
M3_FN int sorted_pos(const unsigned char* cls, int N, int n) {
  const int c = cls[n];
  int pos = 0;
  for (int m = 0; m < N; ++m) {
    const int cm = cls[m];
    pos += (cm < c || (cm == c && m < n)) ? 1 : 0;
  }
  return pos;
}

// Per-thread pixel context: the thread owns pixels xb .. xb+3 of the row.
struct PxCtx {
  int xb;
  float cW;         // (W-1)/2
  float2 g0p[2];    // affine-grid x of the own pixels
  float2 nxf[2];    // -(float)x of the own pixels
};

// Fractional tap weight of an own-pixel pair on a plane with integer shift k0 (reference fp32 op order: (g0 + xof) + 1,
// times (W-1)/2; minus floor).  The product uses scalar mul.rn: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2,
// which would skip a rounding the reference performs (med.cu, frac2).
M3_FN float2 frac_plus(float2 g0, float2 nxf, float xof, float cW, float nk0f) {
  const float2 u = add2(add2(g0, splat(xof)), splat(1.0f));
  const float2 t = make_float2(__fmul_rn(u.x, cW), __fmul_rn(u.y, cW));
  return add2(add2(t, splat(nk0f)), nxf);
}
// Same for the opposite shift (sample at x - s): subtract x first (exact), add k0 + 1 last.
M3_FN float2 frac_minus(float2 g0, float2 nxf, float xof, float cW, float k1f) {
  const float2 u = add2(add2(g0, splat(-xof)), splat(1.0f));
  const float2 t = make_float2(__fmul_rn(u.x, cW), __fmul_rn(u.y, cW));
  return add2(add2(t, nxf), splat(k1f));
}
// a+ of a FOREIGN pixel j whose grid value is g: t - (j + k0), with nm = -(float)(j + k0) (exact: j + k0 = floor(t)).
M3_FN float frac_at(float g, float xof, float cW, float nm) {
  const float t = __fmul_rn(__fadd_rn(__fadd_rn(g, xof), 1.0f), cW);
  return __fadd_rn(t, nm);
}

// v[i] = base[R + i], i = 0..4, base a 16-byte aligned quad pair
template <int R>
M3_FN void win5(const float* base, float v[5]) {
  const float4 a = ld4(base), b = ld4(base + 4);
  const float q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = q[R + i];
}
// un-shifted six-tap window of the own quad: row[xb-1 .. xb+4].  (Round 2 measured the two outer taps as warp shuffles of
// the neighbouring lanes' inner taps instead of the two scalar loads, whose 16-byte lane stride is a 4-way bank conflict:
// fwd+masks 0.366 -> 0.393 ms and bwd 0.278 -> 0.303 ms at 16x49x192x640 -- SLOWER; SHFL + activemask bookkeeping cost
// more issue slots than the conflicts cost wavefronts.  Loads kept.)
M3_FN void win6(const float* row, int xb, float v[6]) {
  const float4 a = ld4(row + xb);
  v[0] = row[xb - 1];
  v[1] = a.x; v[2] = a.y; v[3] = a.z; v[4] = a.w;
  v[5] = row[xb + 4];
}

// =============================================================================================
// Forward sweep A: softmax sums, disparity expectation, colour blend -- no running maximum.
// =============================================================================================
struct FwdAcc {
  float2 z0[2], dacc[2], zw[2], pacc[3][2];
};
M3_FN void fwd_acc_init(FwdAcc& A) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    A.z0[h] = A.dacc[h] = A.zw[h] = splat(0.f);
    A.pacc[0][h] = A.pacc[1][h] = A.pacc[2][h] = splat(0.f);
  }
}

// Write the own quad (pixels xb .. xb+3 of channel ch, zero beyond the row) into the image records.  Positions a row never
// writes (beyond ceil4(W) - r) keep the zeros of the kernel prologue.
M3_FN void stage_quad(float* img, int ch, int xb, float4 v) {
  const float e[4] = {v.x, v.y, v.z, v.w};
  float* rec = img + (xb >> 2) * kImgRec + ch * 4;
  st4(rec, v);
#pragma unroll
  for (int r = 1; r < 4; ++r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= r) rec[r * 12 + (k - r)] = e[k];                               // pixel xb+k-r lives in the own record
      else if (xb > 0) rec[r * 12 + (4 + k - r) - kImgRec] = e[k];            // ... or in the previous one
    }
  }
}

// One ordinary plane of class R.  `row`: element 0 of the staged plane row (16B aligned; zeros in front of it and
// behind ceil4(W)).  `img`: the image records of the row (stage_quad).
template <int R>
M3_FN void fwd_plane(FwdAcc& A, const float* row, const float* img, const Ent& e, const PxCtx& c, int wr) {
  const float4 L = ld4(row + c.xb);
  // taps xb+k0 .. xb+k0+4 are lanes R .. R+4 of the quad pair at xb + woff; a window that starts beyond the row is
  // redirected to the zero tail
  const int b0 = min(c.xb + e.woff, wr);
  float v[5];
  win5<R>(row + b0, v);
  float2 a[2], wl[2];
  a[0] = frac_plus(c.g0p[0], c.nxf[0], e.xof, c.cW, e.nk0f);
  a[1] = frac_plus(c.g0p[1], c.nxf[1], e.xof, c.cW, e.nk0f);
  wl[0] = make_float2(fmaf(a[0].x, v[1] - v[0], v[0]), fmaf(a[0].y, v[2] - v[1], v[1]));
  wl[1] = make_float2(fmaf(a[1].x, v[3] - v[2], v[2]), fmaf(a[1].y, v[4] - v[3], v[3]));
  const float2 e0a = ex2_2(mul2(make_float2(L.x, L.y), splat(kLog2e)));
  const float2 e0b = ex2_2(mul2(make_float2(L.z, L.w), splat(kLog2e)));
  const float2 ewa = ex2_2(mul2(wl[0], splat(kLog2e)));
  const float2 ewb = ex2_2(mul2(wl[1], splat(kLog2e)));
  A.z0[0] = add2(A.z0[0], e0a);
  A.z0[1] = add2(A.z0[1], e0b);
  A.dacc[0] = fma2(splat(e.d), e0a, A.dacc[0]);
  A.dacc[1] = fma2(splat(e.d), e0b, A.dacc[1]);
  A.zw[0] = add2(A.zw[0], ewa);
  A.zw[1] = add2(A.zw[1], ewb);
  // pan += ew (1-a) I[x0] + ew a I[x0+1]; image taps 0..3 = phase R of the record, taps 1..4 = phase R+1 (phase 0 of the next record)
  const float2 w1a = mul2(ewa, a[0]), w1b = mul2(ewb, a[1]);
  const float2 w0a = fma2(w1a, splat(-1.0f), ewa), w0b = fma2(w1b, splat(-1.0f), ewb);
  const float* rec = img + (b0 >> 2) * kImgRec;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float4 Aq = ld4(rec + (R * 3 + ch) * 4);
    const float4 Bq = ld4(rec + (R < 3 ? ((R + 1) * 3 + ch) * 4 : kImgRec + ch * 4));
    A.pacc[ch][0] = fma2(w0a, make_float2(Aq.x, Aq.y), A.pacc[ch][0]);
    A.pacc[ch][1] = fma2(w0b, make_float2(Aq.z, Aq.w), A.pacc[ch][1]);
    A.pacc[ch][0] = fma2(w1a, make_float2(Bq.x, Bq.y), A.pacc[ch][0]);
    A.pacc[ch][1] = fma2(w1b, make_float2(Bq.z, Bq.w), A.pacc[ch][1]);
  }
}

// Row results of sweep A.  Returns true when a softmax sum of an in-row pixel left the safe range (row must be
// recomputed by the robust kernel).  nl0/nlw: NEGATED log2-domain log-sum-exps (-inf for pixels beyond the row), the
// form sweep B consumes.
M3_FN bool fwd_finish(const FwdAcc& A, int xb, int W, float disp[4], float pan[3][4], float lse0[4], float lsew[4],
                      float nl0[4], float nlw[4]) {
  const float z0v[4] = {A.z0[0].x, A.z0[0].y, A.z0[1].x, A.z0[1].y};
  const float zwv[4] = {A.zw[0].x, A.zw[0].y, A.zw[1].x, A.zw[1].y};
  const float dv[4] = {A.dacc[0].x, A.dacc[0].y, A.dacc[1].x, A.dacc[1].y};
  bool bad = false;
  float rzw[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool in = xb + i < W;
    const bool ok = z0v[i] > kZLo && z0v[i] < kZHi && zwv[i] > kZLo && zwv[i] < kZHi;
    bad = bad || (in && !ok);
    const float l0 = lg2f(z0v[i]), lw = lg2f(zwv[i]);
    nl0[i] = in ? -l0 : -INFINITY;
    nlw[i] = in ? -lw : -INFINITY;
    lse0[i] = l0 * kLn2;
    lsew[i] = lw * kLn2;
    disp[i] = dv[i] * m3_rcp(z0v[i]);
    rzw[i] = m3_rcp(zwv[i]);
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float pv[4] = {A.pacc[ch][0].x, A.pacc[ch][0].y, A.pacc[ch][1].x, A.pacc[ch][1].y};
#pragma unroll
    for (int i = 0; i < 4; ++i) pan[ch][i] = pv[i] * rzw[i];
  }
  return bad;
}

// =============================================================================================
// Forward sweep B: the two sub-occlusion masks as register gathers.
//   nl0row / nlwrow: -log2(Z0), -log2(Zw) of every pixel of the row, -inf outside [0, W)
//   g0row:           affine-grid x of every pixel of the row
// =============================================================================================
template <int R>
M3_FN void mask_plane(float mR[4], float mL[4], const float* row, const float* nl0row, const float* nlwrow,
                      const float* g0row, const Ent& e, const PxCtx& c, float nxm1, int wr) {
  // ---- maskR(x) += lerp(Q(x+k0), Q(x+k0+1); a+(x)),  Q(j) = exp2(L(j) log2e - log2 Z0(j))
  {
    const int b0 = min(c.xb + e.woff, wr);
    float lw[5], nw[5], q[5];
    win5<R>(row + b0, lw);
    win5<R>(nl0row + b0, nw);
#pragma unroll
    for (int i = 0; i < 5; ++i) q[i] = ex2f(fmaf(lw[i], kLog2e, nw[i]));
    const float2 a0 = frac_plus(c.g0p[0], c.nxf[0], e.xof, c.cW, e.nk0f);
    const float2 a1 = frac_plus(c.g0p[1], c.nxf[1], e.xof, c.cW, e.nk0f);
    mR[0] += fmaf(a0.x, q[1] - q[0], q[0]);
    mR[1] += fmaf(a0.y, q[2] - q[1], q[1]);
    mR[2] += fmaf(a1.x, q[3] - q[2], q[2]);
    mR[3] += fmaf(a1.y, q[4] - q[3], q[3]);
  }
  // ---- maskL(x) += lerp(P(x-k0-1), P(x-k0); a-(x)),  P(j) = exp2(wl(j) log2e - log2 Zw(j)),
  //      wl(j) = lerp(L(j+k0), L(j+k0+1); a+(j)).  For j = xb-k0-1+i the logit taps are the UN-shifted L(xb-1+i), L(xb+i);
  //      the window of the foreign pixels starts at xb-k0-1 = (xb - woff - 4) + (3 - R).
  {
    const int bq = max(c.xb - e.woff - 4, -kPad);
    float gw[5], nw[5], lu[6], pw[5];
    win5<3 - R>(g0row + bq, gw);
    win5<3 - R>(nlwrow + bq, nw);
    win6(row, c.xb, lu);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float ap = frac_at(gw[i], e.xof, c.cW, nxm1 - (float)i);   // j + k0 = xb - 1 + i
      const float wl = fmaf(ap, lu[i + 1] - lu[i], lu[i]);
      pw[i] = ex2f(fmaf(wl, kLog2e, nw[i]));
    }
    const float2 m0 = frac_minus(c.g0p[0], c.nxf[0], e.xof, c.cW, e.k1f);
    const float2 m1 = frac_minus(c.g0p[1], c.nxf[1], e.xof, c.cW, e.k1f);
    mL[0] += fmaf(m0.x, pw[1] - pw[0], pw[0]);
    mL[1] += fmaf(m0.y, pw[2] - pw[1], pw[1]);
    mL[2] += fmaf(m1.x, pw[3] - pw[2], pw[2]);
    mL[3] += fmaf(m1.y, pw[4] - pw[3], pw[3]);
  }
}

// =============================================================================================
// Backward: g_logits of the own quad on one plane, as a register gather.
//   per-thread row constants: iw[c][0..5] = I_c(xb-1 .. xb+4) (zero outside the row), nl0 = -lse0 log2e,
//                             gd = g_disp, ndsp = -disp of the own pixels
//   staged rows: nlwrow (-lsew log2e, -inf outside the row), dotrow (<g_pan, pan>), gprow[c] (g_pan), g0row
// =============================================================================================
struct BwdCtx {
  float iw[3][6];
  float nl0[4], gd[4], ndsp[4];
};
template <int R>
M3_FN void bwd_plane(float g[4], const float* row, const float* nlwrow, const float* dotrow, const float* gprow,
                     int rowf, const float* g0row, const Ent& e, const PxCtx& c, const BwdCtx& t, float nxm1) {
  const int bq = max(c.xb - e.woff - 4, -kPad);
  float gw[5], nw[5], dw[5], lu[6], ap[5], G[5];
  win5<3 - R>(g0row + bq, gw);
  win5<3 - R>(nlwrow + bq, nw);
  win5<3 - R>(dotrow + bq, dw);
  win6(row, c.xb, lu);
#pragma unroll
  for (int i = 0; i < 5; ++i) {   // pixel x = xb-k0-1+i sampled the taps (xb-1+i, xb+i)
    ap[i] = frac_at(gw[i], e.xof, c.cW, nxm1 - (float)i);
    const float wl = fmaf(ap[i], lu[i + 1] - lu[i], lu[i]);
    G[i] = ex2f(fmaf(wl, kLog2e, nw[i]));   // P_n(x); 0 for x outside the row
  }
  float dP[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float gp[5];
    win5<3 - R>(gprow + ch * rowf + bq, gp);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float col = fmaf(ap[i], t.iw[ch][i + 1] - t.iw[ch][i], t.iw[ch][i]);
      dP[i] = fmaf(gp[i], col, dP[i]);
    }
  }
  float rb[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    G[i] = G[i] * (dP[i] - dw[i]);
    rb[i] = ap[i] * G[i];
  }
  // own pixel j = xb+i is tap x0 of pixel j-k0 (window index i+1, weight 1-a) and tap x0+1 of pixel j-k0-1 (index i, a)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float q0 = ex2f(fmaf(lu[i + 1], kLog2e, t.nl0[i]));   // Q_n(j)
    const float gdisp = (q0 * t.gd[i]) * (e.d + t.ndsp[i]);
    g[i] = gdisp + ((G[i + 1] - rb[i + 1]) + rb[i]);
  }
}

// =============================================================================================
// Generic per-pixel versions for "special" planes (class 4: shift within rounding distance of an integer, negative or
// absurdly large): floor() of the replayed fp32 coordinate is taken per pixel, taps are bounds-checked scalar reads.
// A few planes per sample at most (none for KITTI-shaped rows; plane 0 at 2048 px), so these favour clarity.
// =============================================================================================
#if defined(__CUDACC__)
#define M3_LDG(p) __ldg(p)
#else
#define M3_LDG(p) (*(p))
#endif
M3_FN float tapz(const float* row, int j, int W) { return (j >= 0 && j <= W - 1) ? row[j] : 0.0f; }
M3_FN float tapz_g(const float* row, int j, int W) { return (j >= 0 && j <= W - 1) ? M3_LDG(row + j) : 0.0f; }
M3_FN float coord_plus(float g0, float xof, float cW) { return __fmul_rn(__fadd_rn(__fadd_rn(g0, xof), 1.0f), cW); }
M3_FN float coord_minus(float g0, float xof, float cW) { return __fmul_rn(__fadd_rn(__fsub_rn(g0, xof), 1.0f), cW); }
// I_c[j] out of the image records (phase 0), zero outside the row
M3_FN float img_tap(const float* img, int ch, int j, int W) {
  return (j >= 0 && j <= W - 1) ? img[(j >> 2) * kImgRec + ch * 4 + (j & 3)] : 0.0f;
}

M3_FN void fwd_plane_generic(FwdAcc& A, const float* row, const float* img, const Ent& e, const PxCtx& c, int W) {
  const float4 L = ld4(row + c.xb);
  const float lv[4] = {L.x, L.y, L.z, L.w};
  const float g0[4] = {c.g0p[0].x, c.g0p[0].y, c.g0p[1].x, c.g0p[1].y};
  float e0[4], ew[4], w0[4], w1[4], i0[3][4], i1[3][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float t = coord_plus(g0[i], e.xof, c.cW);
    const float x0f = floorf(t);
    const float a = t - x0f;
    const int x0 = (int)x0f;
    const float f0 = tapz(row, x0, W), f1 = tapz(row, x0 + 1, W);
    const float wl = fmaf(a, f1 - f0, f0);
    e0[i] = ex2f(lv[i] * kLog2e);
    ew[i] = ex2f(wl * kLog2e);
    w1[i] = ew[i] * a;
    w0[i] = fmaf(w1[i], -1.0f, ew[i]);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      i0[ch][i] = img_tap(img, ch, x0, W);
      i1[ch][i] = img_tap(img, ch, x0 + 1, W);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float2 e0p = make_float2(e0[2 * h], e0[2 * h + 1]), ewp = make_float2(ew[2 * h], ew[2 * h + 1]);
    const float2 w0p = make_float2(w0[2 * h], w0[2 * h + 1]), w1p = make_float2(w1[2 * h], w1[2 * h + 1]);
    A.z0[h] = add2(A.z0[h], e0p);
    A.dacc[h] = fma2(splat(e.d), e0p, A.dacc[h]);
    A.zw[h] = add2(A.zw[h], ewp);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      A.pacc[ch][h] = fma2(w0p, make_float2(i0[ch][2 * h], i0[ch][2 * h + 1]), A.pacc[ch][h]);
      A.pacc[ch][h] = fma2(w1p, make_float2(i1[ch][2 * h], i1[ch][2 * h + 1]), A.pacc[ch][h]);
    }
  }
}

// P_n(j) = softmax of the warped logits at pixel j of a special plane (0 outside the row)
M3_FN float warped_prob_generic(const float* row, const float* nlwrow, const float* g0row, const Ent& e, float cW, int j,
                                int W, float* a_out, int* x0_out) {
  if (j < 0 || j > W - 1) {
    *a_out = 0.f;
    *x0_out = -4;
    return 0.0f;
  }
  const float t = coord_plus(g0row[j], e.xof, cW);
  const float x0f = floorf(t);
  const float a = t - x0f;
  const int x0 = (int)x0f;
  const float f0 = tapz(row, x0, W), f1 = tapz(row, x0 + 1, W);
  const float wl = fmaf(a, f1 - f0, f0);
  *a_out = a;
  *x0_out = x0;
  return ex2f(fmaf(wl, kLog2e, nlwrow[j]));
}

M3_FN void mask_plane_generic(float mR[4], float mL[4], const float* row, const float* nl0row, const float* nlwrow,
                              const float* g0row, const Ent& e, const PxCtx& c, int W) {
  const float g0[4] = {c.g0p[0].x, c.g0p[0].y, c.g0p[1].x, c.g0p[1].y};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // maskR: softmax(L)_n sampled at x + s
    float t = coord_plus(g0[i], e.xof, c.cW);
    float x0f = floorf(t);
    float a = t - x0f;
    int x0 = (int)x0f;
    const float q0 = (x0 >= 0 && x0 <= W - 1) ? ex2f(fmaf(row[x0], kLog2e, nl0row[x0])) : 0.f;
    const float q1 = (x0 + 1 >= 0 && x0 + 1 <= W - 1) ? ex2f(fmaf(row[x0 + 1], kLog2e, nl0row[x0 + 1])) : 0.f;
    mR[i] += fmaf(a, q1 - q0, q0);
    // maskL: softmax(warped L)_n sampled at x - s
    t = coord_minus(g0[i], e.xof, c.cW);
    x0f = floorf(t);
    a = t - x0f;
    x0 = (int)x0f;
    float au;
    int xu;
    const float p0 = warped_prob_generic(row, nlwrow, g0row, e, c.cW, x0, W, &au, &xu);
    const float p1 = warped_prob_generic(row, nlwrow, g0row, e, c.cW, x0 + 1, W, &au, &xu);
    mL[i] += fmaf(a, p1 - p0, p0);
  }
}

// img_rows: GLOBAL pointers to the three image rows of (b, y)
M3_FN void bwd_plane_generic(float g[4], const float* row, const float* nlwrow, const float* dotrow, const float* gprow,
                             int rowf, const float* g0row, const Ent& e, const PxCtx& c, const BwdCtx& t,
                             const float* img_r, const float* img_g, const float* img_b, int W) {
  const float4 L = ld4(row + c.xb);
  const float lv[4] = {L.x, L.y, L.z, L.w};
  const int kn = (int)floorf(e.xof * c.cW);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = c.xb + i;
    float acc = 0.f;
    // the pixels whose taps can be j: floor of their coordinate is within +-1 of the nominal shift
    for (int x = j - kn - 2; x <= j - kn + 1; ++x) {
      float a;
      int x0;
      const float P = warped_prob_generic(row, nlwrow, g0row, e, c.cW, x, W, &a, &x0);
      if (x0 != j && x0 + 1 != j) continue;
      float dP = 0.f;
      dP = fmaf(gprow[x], fmaf(a, tapz_g(img_r, x0 + 1, W) - tapz_g(img_r, x0, W), tapz_g(img_r, x0, W)), dP);
      dP = fmaf(gprow[rowf + x], fmaf(a, tapz_g(img_g, x0 + 1, W) - tapz_g(img_g, x0, W), tapz_g(img_g, x0, W)), dP);
      dP = fmaf(gprow[2 * rowf + x], fmaf(a, tapz_g(img_b, x0 + 1, W) - tapz_g(img_b, x0, W), tapz_g(img_b, x0, W)), dP);
      const float G = P * (dP - dotrow[x]);
      acc += (x0 == j) ? (1.0f - a) * G : a * G;
    }
    const float q0 = ex2f(fmaf(lv[i], kLog2e, t.nl0[i]));
    g[i] = (q0 * t.gd[i]) * (e.d + t.ndsp[i]) + acc;
  }
}

}  // namespace m3
}  // namespace faln
