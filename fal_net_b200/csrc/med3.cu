// med3.cu -- third-generation fused MED kernels for sm_100a (forward, forward + sub-occlusion masks, backward).
//
// Replaces /root/reference/models/FAL_netB.py:216-297 and its autograd.  Same data movement as med.cu (one persistent CTA
// per image row, plane rows streamed once through a shared-memory ring by 1-D bulk async copies from a producer warp),
// but the plane loops are barrier-free register gathers (see med3_core.cuh) and the per-plane control flow is gone:
//   * the producer warp builds the class-sorted plane table of a sample in parallel (one table per sample, handed over
//     through an mbarrier pair), rows are assigned to CTAs in contiguous chunks so a CTA meets at most two samples
//   * "special" planes (shift within rounding distance of an integer) are visited last, on per-pixel generic code
//   * no running maximum in the softmax sums; forward rows whose sums leave the safe range are left to the robust
//     second-generation kernel, launched right after in clean-up mode (it exits at once when there is nothing to do).
//     The signal is in the output itself: lse0[b, 0, y, 0] = NaN marks a row to recompute.
//   * register caps sized against the per-sub-partition register file so that 1242-px and 640-px rows run two CTAs per SM.
// Eligibility (checked on the host, med.cu): 16-byte aligned logit rows, pitch % 4 == 0, zero pad columns, lse outputs
// present.  Everything else stays on the second-generation kernels.
#include <math.h>
#include <stdlib.h>

#include "med3.h"
#include "med3_core.cuh"

namespace faln {
namespace m3 {
namespace {

struct Smem {
  int off_bar, off_cnt, off_cls, off_tab, off_img, off_rows, off_ring, total;
};
// n_rows: per-row arrays of row_floats(W) each; with_img: the image records of the row (med3_core.cuh, stage_quad)
__host__ __device__ inline Smem make_smem(int W, int S, int G, int n_rows, bool with_img) {
  Smem m;
  int o = 0;
  m.off_bar = o;   // full[S], empty[S], tab_full, tab_empty
  o += 256;
  m.off_cnt = o;   // int cnt[8]: planes per class 0..4
  o += 32;
  m.off_cls = o;   // class of plane n (bytes)
  o += kMaxN;
  m.off_tab = o;
  o += kMaxN * (int)sizeof(Ent);
  m.off_img = o;
  if (with_img) o += img_floats(W) * 4;
  m.off_rows = o;
  o += n_rows * row_floats(W) * 4;
  m.off_ring = o;
  o += S * G * slot_floats(W) * 4;
  m.total = o;
  return m;
}

__device__ __forceinline__ bool bar_red_or(int id, int nthreads, bool v) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, %2, %3, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(out)
      : "r"((uint32_t)v), "r"(id), "r"(nthreads)
      : "memory");
  return out != 0;
}

__device__ __forceinline__ void store_row4(float* rowp, int xb, const float v[4], int W) {
  float* p = rowp + xb;
  if (xb + 3 < W) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
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
// four consecutive floats of a global row, zero beyond W
__device__ __forceinline__ float4 load_row4(const float* rowp, int xb, int W) {
  const float* p = rowp + xb;
  if (xb + 3 < W && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v;
  v.x = xb < W ? __ldg(p) : 0.f;
  v.y = xb + 1 < W ? __ldg(p + 1) : 0.f;
  v.z = xb + 2 < W ? __ldg(p + 2) : 0.f;
  v.w = xb + 3 < W ? __ldg(p + 3) : 0.f;
  return v;
}

// ------------------------------------------------------------------------------------------ producer side
struct Pipe {
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tab_full;
  uint64_t* tab_empty;
};

// Whole producer warp: (re)build the sorted plane table of sample b.
__device__ __forceinline__ void build_table(const M3Params& p, int b, int lane, Ent* tab, unsigned char* cls, int* cnt) {
  Ent mine[kMaxN / 32];
#pragma unroll
  for (int j = 0; j < kMaxN / 32; ++j) {
    const int n = lane + 32 * j;
    if (n < p.N) {
      mine[j] = make_ent(__ldg(p.x_of + (size_t)b * p.N + n), __ldg(p.d_lvl + (size_t)b * p.N + n), n, p.W, p.force_generic != 0);
      cls[n] = (unsigned char)mine[j].cls;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < kMaxN / 32; ++j) {
    const int n = lane + 32 * j;
    if (n < p.N) tab[sorted_pos(cls, p.N, n)] = mine[j];
  }
  if (lane < 5) {
    int c = 0;
    for (int m = 0; m < p.N; ++m) c += cls[m] == lane ? 1 : 0;
    cnt[lane] = c;
  }
  __syncwarp();
}

// Producer warp main loop: rows [r0, r1) of the flattened (b, y) index, `sweeps` passes over the planes of every row.
__device__ __forceinline__ void producer(const M3Params& p, const Pipe& P, float* ring, Ent* tab, unsigned char* cls,
                                         int* cnt, int r0, int r1, int sweeps, int lane) {
  const int S = p.S, G = p.G, N = p.N;
  const int wr = ceil4(p.W), slotf = slot_floats(p.W);
  const long long plane = (long long)p.H * p.pitch;
  int slot = 0, cur_b = -1;
  uint32_t par = 0, te = 0;
  for (int row = r0; row < r1; ++row) {
    const int b = row / p.H, y = row % p.H;
    if (b != cur_b) {
      if (cur_b >= 0) {   // the consumers must be done with the previous sample's table
        mbar_wait(P.tab_empty, te);
        te ^= 1;
      }
      build_table(p, b, lane, tab, cls, cnt);
      if (lane == 0) mbar_arrive(P.tab_full);
      cur_b = b;
    }
    if (lane == 0) {
      // groups of up to G planes that never straddle a class boundary: the consumers' inner loops are class-pure
      const float* src0 = p.logits + ((long long)b * N * p.H + y) * p.pitch;
      for (int sw = 0; sw < sweeps; ++sw) {
        int i = 0;
        for (int k = 0; k < 5; ++k) {
          const int nk = cnt[k];
          for (int g0 = 0; g0 < nk; g0 += G) {
            const int c = min(G, nk - g0);
            while (!mbar_try_wait(&P.empty[slot], par ^ 1)) __nanosleep(20);
            mbar_arrive_expect_tx(&P.full[slot], (uint32_t)(c * wr * 4));
            for (int q = 0; q < c; ++q)
              bulk_g2s(ring + ((size_t)slot * G + q) * slotf + kPad, src0 + tab[i + q].src * plane, (uint32_t)(wr * 4),
                       &P.full[slot]);
            i += c;
            if (++slot == S) { slot = 0; par ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ consumer side
struct Ring {
  const float* ring;   // element 0 of the payload of slot 0, plane 0
  uint64_t* full;
  uint64_t* empty;
  const Ent* tab;
  int S, G, slotf, lane;
  int slot;
  uint32_t par;
  bool ready;          // the full barrier of the current slot was already seen complete (early try_wait)
  bool last_row;       // the CTA is on its last row
  int left;            // groups left in the current row (all sweeps)
};

// Groups per sweep of a row: every class is cut into groups of up to G planes (same rule as the producer).
__device__ __forceinline__ int groups_per_sweep(const int* cnt, int G) {
  int g = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) g += (cnt[k] + G - 1) / G;
  return g;
}

// All planes of one alignment class: groups of up to G planes share one mbarrier hand-shake; the inner loop is free of
// pipeline logic.  f(row, ent) processes one plane.  mbarrier.try_wait costs ~90 cycles even on a completed phase
// (B300_MICROARCH.md), so the test of the NEXT group's barrier is issued before this group's planes are processed and
// only its result is consumed afterwards.
template <int kUnroll, class F>
__device__ __forceinline__ void class_run(Ring& rg, const Ent*& ent, int count, bool active, F f) {
  for (int g0 = 0; g0 < count; g0 += rg.G) {
    const int c = min(rg.G, count - g0);
    if (!rg.ready) mbar_wait(&rg.full[rg.slot], rg.par);
    int ns = rg.slot + 1;
    uint32_t np = rg.par;
    if (ns == rg.S) { ns = 0; np ^= 1; }
    const bool has_next = !(rg.last_row && rg.left == 1);
    --rg.left;
    rg.ready = has_next ? mbar_try_wait(&rg.full[ns], np) : false;
    if (active) {
      const float* rowp = rg.ring + (size_t)rg.slot * rg.G * rg.slotf;
#pragma unroll kUnroll
      for (int q = 0; q < c; ++q) f(rowp + q * rg.slotf, ent[q]);
    }
    ent += c;
    __syncwarp();
    if (rg.lane == 0) mbar_arrive(&rg.empty[rg.slot]);
    rg.slot = ns;
    rg.par = np;
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Stage the three image rows of (b, y) into the image records; each thread handles its own quad.
__device__ __forceinline__ void stage_image(float* img, const float* img_b, int y, int H, int W, int xb) {
  float4 v[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) v[ch] = load_row4(img_b + ((size_t)ch * H + y) * W, xb, W);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) stage_quad(img, ch, xb, v[ch]);
}

// Common prologue: barriers, zeroed ring / image copies, -inf / zero padded row arrays, g0 row.
__device__ __forceinline__ void prologue(const M3Params& p, unsigned char* smem, const Smem& L, int n_inf_rows,
                                         int n_rows, bool with_img, int ncons) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&bars[s], 1);                 // full
      mbar_init(&bars[p.S + s], ncons / 32);  // empty
    }
    mbar_init(&bars[2 * p.S], 1);               // tab_full
    mbar_init(&bars[2 * p.S + 1], ncons / 32);  // tab_empty
    fence_barrier_init();
  }
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const int nring = p.S * p.G * slot_floats(p.W);
  for (int i = tid; i < nring; i += blockDim.x) ring[i] = 0.f;
  if (with_img) {
    float* img = reinterpret_cast<float*>(smem + L.off_img);
    for (int i = tid; i < img_floats(p.W); i += blockDim.x) img[i] = 0.f;
  }
  // row arrays: array 0 = g0 row (values), arrays 1 .. n_inf_rows = log-sum-exp rows (-inf), the rest zero
  float* rows = reinterpret_cast<float*>(smem + L.off_rows);
  const int rowf = row_floats(p.W);
  for (int i = tid; i < n_rows * rowf; i += blockDim.x) {
    const int a = i / rowf, j = i % rowf - kPad;
    float v = 0.f;
    if (a == 0) v = (j >= 0 && j < p.W) ? __ldg(p.g0x + j) : 0.f;
    else if (a <= n_inf_rows) v = -INFINITY;
    rows[i] = v;
  }
  fence_proxy_async();
  __syncthreads();
}

__device__ __forceinline__ void row_range(int rows, int& r0, int& r1) {
  r0 = (int)(((long long)rows * blockIdx.x) / gridDim.x);
  r1 = (int)(((long long)rows * (blockIdx.x + 1)) / gridDim.x);
}

// =============================================================================================
// Forward
// =============================================================================================
template <bool kMasks, int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med3_fwd_kernel(const M3Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kRows = kMasks ? 3 : 0;   // g0row, nl0row, nlwrow
  const Smem L = make_smem(p.W, p.S, p.G, kRows, true);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
  unsigned char* cls = smem + L.off_cls;
  Ent* tab = reinterpret_cast<Ent*>(smem + L.off_tab);
  float* img = reinterpret_cast<float*>(smem + L.off_img);
  float* rows_s = reinterpret_cast<float*>(smem + L.off_rows);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const Pipe P{bars, bars + p.S, bars + 2 * p.S, bars + 2 * p.S + 1};

  const int ncons = blockDim.x - 32;
  const int tid = threadIdx.x;
  const int W = p.W, H = p.H;
  const int wr = ceil4(W), rowf = row_floats(W);
  prologue(p, smem, L, kMasks ? 2 : 0, kRows, true, ncons);
  int r0, r1;
  row_range(p.B * H, r0, r1);

  if (tid >= ncons) {
    producer(p, P, ring, tab, cls, cnt, r0, r1, kMasks ? 2 : 1, tid - ncons);
    return;
  }

  PxCtx c;
  c.xb = tid * kPX;
  c.cW = 0.5f * (float)(W - 1);
  const bool active = c.xb < W;
  {
    float g0[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g0[i] = __ldg(p.g0x + min(c.xb + i, W - 1));
    c.g0p[0] = make_float2(g0[0], g0[1]);
    c.g0p[1] = make_float2(g0[2], g0[3]);
    c.nxf[0] = make_float2(-(float)c.xb, -(float)(c.xb + 1));
    c.nxf[1] = make_float2(-(float)(c.xb + 2), -(float)(c.xb + 3));
  }
  const float nxm1 = -(float)(c.xb - 1);
  const float* g0row = rows_s + kPad;
  float* nl0row = rows_s + rowf + kPad;
  float* nlwrow = rows_s + 2 * rowf + kPad;
  Ring rg{ring + kPad, P.full, P.empty, tab, p.S, p.G, slot_floats(W), tid & 31, 0, 0u, false, false, 0};
  int cur_b = -1;
  uint32_t tf = 0;

  for (int row = r0; row < r1; ++row) {
    const int b = row / H, y = row % H;
    if (b != cur_b) {
      if (cur_b >= 0) {
        __syncwarp();
        if (rg.lane == 0) mbar_arrive(P.tab_empty);
      }
      mbar_wait(P.tab_full, tf);
      tf ^= 1;
      cur_b = b;
    }
    rg.left = groups_per_sweep(cnt, p.G) * (kMasks ? 2 : 1);
    rg.last_row = row == r1 - 1;
    named_bar_sync(1, ncons);   // previous row fully consumed: image copies / row arrays may be overwritten
    if (active) stage_image(img, p.image + (size_t)b * 3 * H * W, y, H, W, c.xb);
    named_bar_sync(1, ncons);
    if (active && row + 1 < r1) {   // next row's image quads: in L2 by the time they are staged
      const int bn = (row + 1) / H, yn = (row + 1) % H;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) prefetch_l2(p.image + (((size_t)bn * 3 + ch) * H + yn) * W + min(c.xb, W - 1));
    }

    // ---------------------------------------------------------------- sweep A
    FwdAcc A;
    fwd_acc_init(A);
    {
      const Ent* ent = tab;
      class_run<2>(rg, ent, cnt[0], active, [&](const float* rowp, const Ent& e) { fwd_plane<0>(A, rowp, img, e, c, wr); });
      class_run<2>(rg, ent, cnt[1], active, [&](const float* rowp, const Ent& e) { fwd_plane<1>(A, rowp, img, e, c, wr); });
      class_run<2>(rg, ent, cnt[2], active, [&](const float* rowp, const Ent& e) { fwd_plane<2>(A, rowp, img, e, c, wr); });
      class_run<2>(rg, ent, cnt[3], active, [&](const float* rowp, const Ent& e) { fwd_plane<3>(A, rowp, img, e, c, wr); });
      class_run<1>(rg, ent, cnt[4], active, [&](const float* rowp, const Ent& e) { fwd_plane_generic(A, rowp, img, e, c, W); });
    }

    bool bad = false;
    const size_t r1o = ((size_t)b * H + y) * W;
    if (active) {
      float disp[4], pan[3][4], lse0[4], lsew[4], nl0[4], nlw[4];
      bad = fwd_finish(A, c.xb, W, disp, pan, lse0, lsew, nl0, nlw);
      if (p.disp) store_row4(p.disp + r1o, c.xb, disp, W);
      if (p.pan) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) store_row4(p.pan + (((size_t)b * 3 + ch) * H + y) * W, c.xb, pan[ch], W);
      }
      store_row4(p.lse0 + r1o, c.xb, lse0, W);
      store_row4(p.lsew + r1o, c.xb, lsew, W);
      if (kMasks) {
        st4(nl0row + c.xb, make_float4(nl0[0], nl0[1], nl0[2], nl0[3]));
        st4(nlwrow + c.xb, make_float4(nlw[0], nlw[1], nlw[2], nlw[3]));
      }
    }
    // one barrier per row: publishes the normaliser rows (masks) and ORs the overflow flags
    const bool any_bad = bar_red_or(2, ncons, bad);
    if (any_bad && tid == 0) p.lse0[r1o] = __int_as_float(0x7fc00000);   // row left to the clean-up launch

    if (kMasks) {
      // -------------------------------------------------------------- sweep B
      float mR[4] = {0.f, 0.f, 0.f, 0.f}, mL[4] = {0.f, 0.f, 0.f, 0.f};
      const Ent* ent = tab;
      class_run<1>(rg, ent, cnt[0], active,
                   [&](const float* rowp, const Ent& e) { mask_plane<0>(mR, mL, rowp, nl0row, nlwrow, g0row, e, c, nxm1, wr); });
      class_run<1>(rg, ent, cnt[1], active,
                   [&](const float* rowp, const Ent& e) { mask_plane<1>(mR, mL, rowp, nl0row, nlwrow, g0row, e, c, nxm1, wr); });
      class_run<1>(rg, ent, cnt[2], active,
                   [&](const float* rowp, const Ent& e) { mask_plane<2>(mR, mL, rowp, nl0row, nlwrow, g0row, e, c, nxm1, wr); });
      class_run<1>(rg, ent, cnt[3], active,
                   [&](const float* rowp, const Ent& e) { mask_plane<3>(mR, mL, rowp, nl0row, nlwrow, g0row, e, c, nxm1, wr); });
      class_run<1>(rg, ent, cnt[4], active,
                   [&](const float* rowp, const Ent& e) { mask_plane_generic(mR, mL, rowp, nl0row, nlwrow, g0row, e, c, W); });
      if (active) {
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fminf(mL[i], 1.0f);
        store_row4(p.maskL + r1o, c.xb, o, W);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fminf(mR[i], 1.0f);
        store_row4(p.maskR + r1o, c.xb, o, W);
      }
    }
  }
}

// =============================================================================================
// Backward
// =============================================================================================
template <int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med3_bwd_kernel(const M3Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kRows = 6;   // g0row, nlwrow, dotrow, gprow[3]
  const Smem L = make_smem(p.W, p.S, p.G, kRows, false);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
  unsigned char* cls = smem + L.off_cls;
  Ent* tab = reinterpret_cast<Ent*>(smem + L.off_tab);
  float* rows_s = reinterpret_cast<float*>(smem + L.off_rows);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const Pipe P{bars, bars + p.S, bars + 2 * p.S, bars + 2 * p.S + 1};

  const int ncons = blockDim.x - 32;
  const int tid = threadIdx.x;
  const int W = p.W, H = p.H, N = p.N;
  const int rowf = row_floats(W);
  prologue(p, smem, L, 1, kRows, false, ncons);
  int r0, r1;
  row_range(p.B * H, r0, r1);

  if (tid >= ncons) {
    producer(p, P, ring, tab, cls, cnt, r0, r1, 1, tid - ncons);
    return;
  }

  PxCtx c;
  c.xb = tid * kPX;
  c.cW = 0.5f * (float)(W - 1);
  c.g0p[0] = c.g0p[1] = c.nxf[0] = c.nxf[1] = make_float2(0.f, 0.f);   // own-pixel weights are not used by the backward
  const bool active = c.xb < W;
  const float nxm1 = -(float)(c.xb - 1);
  const float* g0row = rows_s + kPad;
  float* nlwrow = rows_s + rowf + kPad;
  float* dotrow = rows_s + 2 * rowf + kPad;
  float* gprow = rows_s + 3 * rowf + kPad;
  Ring rg{ring + kPad, P.full, P.empty, tab, p.S, p.G, slot_floats(W), tid & 31, 0, 0u, false, false, 0};
  int cur_b = -1;
  uint32_t tf = 0;

  for (int row = r0; row < r1; ++row) {
    const int b = row / H, y = row % H;
    if (b != cur_b) {
      if (cur_b >= 0) {
        __syncwarp();
        if (rg.lane == 0) mbar_arrive(P.tab_empty);
      }
      mbar_wait(P.tab_full, tf);
      tf ^= 1;
      cur_b = b;
    }

    // ---- per-row constants: own-pixel registers and the staged rows of the whole row
    BwdCtx t;
    const size_t r1o = ((size_t)b * H + y) * W;
    rg.left = groups_per_sweep(cnt, p.G);
    rg.last_row = row == r1 - 1;
    named_bar_sync(1, ncons);   // previous row fully consumed
    if (active) {
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const size_t rc = (((size_t)b * 3 + ch) * H + y) * W;
        const float4 gq = p.g_pan ? load_row4(p.g_pan + rc, c.xb, W) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 pq = load_row4(p.pan_in + rc, c.xb, W);
        dot[0] = fmaf(gq.x, pq.x, dot[0]);
        dot[1] = fmaf(gq.y, pq.y, dot[1]);
        dot[2] = fmaf(gq.z, pq.z, dot[2]);
        dot[3] = fmaf(gq.w, pq.w, dot[3]);
        st4(gprow + ch * rowf + c.xb, gq);
        const float4 iq = load_row4(p.image + rc, c.xb, W);
        t.iw[ch][0] = c.xb > 0 ? __ldg(p.image + rc + c.xb - 1) : 0.f;
        t.iw[ch][1] = iq.x; t.iw[ch][2] = iq.y; t.iw[ch][3] = iq.z; t.iw[ch][4] = iq.w;
        t.iw[ch][5] = c.xb + 4 < W ? __ldg(p.image + rc + c.xb + 4) : 0.f;
      }
      st4(dotrow + c.xb, make_float4(dot[0], dot[1], dot[2], dot[3]));
      const float4 lw = load_row4(p.lsew_in + r1o, c.xb, W);
      const float4 l0 = load_row4(p.lse0_in + r1o, c.xb, W);
      const float4 gd = p.g_disp ? load_row4(p.g_disp + r1o, c.xb, W) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 dp = load_row4(p.disp_in + r1o, c.xb, W);
      const float lwv[4] = {lw.x, lw.y, lw.z, lw.w}, l0v[4] = {l0.x, l0.y, l0.z, l0.w};
      const float gdv[4] = {gd.x, gd.y, gd.z, gd.w}, dpv[4] = {dp.x, dp.y, dp.z, dp.w};
      float nl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        nl[i] = c.xb + i < W ? -lwv[i] * kLog2e : -INFINITY;
        t.nl0[i] = -l0v[i] * kLog2e;
        t.gd[i] = gdv[i];
        t.ndsp[i] = -dpv[i];
      }
      st4(nlwrow + c.xb, make_float4(nl[0], nl[1], nl[2], nl[3]));
    }
    named_bar_sync(1, ncons);

    if (active && row + 1 < r1) {   // next row's per-pixel inputs: in L2 by the time they are loaded
      const int bn = (row + 1) / H, yn = (row + 1) % H;
      const int xq = min(c.xb, W - 1);
      const size_t rn = ((size_t)bn * H + yn) * W + xq;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const size_t rc = (((size_t)bn * 3 + ch) * H + yn) * W + xq;
        if (p.g_pan) prefetch_l2(p.g_pan + rc);
        prefetch_l2(p.pan_in + rc);
        prefetch_l2(p.image + rc);
      }
      prefetch_l2(p.lsew_in + rn);
      prefetch_l2(p.lse0_in + rn);
      prefetch_l2(p.disp_in + rn);
      if (p.g_disp) prefetch_l2(p.g_disp + rn);
    }
    const long long obase = ((long long)b * N * H + y) * p.g_pitch;
    const long long oplane = (long long)H * p.g_pitch;
    float* const obase_p = p.g_logits + obase + c.xb;
#define M3_BWD_CLASS(R)                                                                                          \
  class_run<1>(rg, ent, cnt[R], active, [&](const float* rowp, const Ent& e) {                                    \
    float g[4];                                                                                                   \
    bwd_plane<R>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, e, c, t, nxm1);                                     \
    store_row4(obase_p + e.src * oplane - c.xb, c.xb, g, W);                                                      \
  })
    const Ent* ent = tab;
    M3_BWD_CLASS(0);
    M3_BWD_CLASS(1);
    M3_BWD_CLASS(2);
    M3_BWD_CLASS(3);
#undef M3_BWD_CLASS
    {
      const float* ir = p.image + (((size_t)b * 3) * H + y) * W;
      class_run<1>(rg, ent, cnt[4], active, [&](const float* rowp, const Ent& e) {
        float g[4];
        bwd_plane_generic(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, e, c, t, ir, ir + (size_t)H * W, ir + 2 * (size_t)H * W, W);
        store_row4(obase_p + e.src * oplane - c.xb, c.xb, g, W);
      });
    }
  }
}

// ------------------------------------------------------------------------------------------ launch configuration
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// Threads, ring shape and CTAs/SM for a row width.  Returns the dynamic shared-memory size, or 0 when nothing fits.
// want_small: CTAs/SM to aim for on rows of <= 640 px (192 threads).  Wider rows run one CTA per SM at up to 128
// registers: measured on B200 at 1242 px, two CTAs at the 80 registers the sub-partition register file then allows spill
// into local memory and lose (fwd 0.400 vs 0.373 ms, masks 0.788 vs 0.678, bwd 0.702 vs 0.574; gpurun_out/s2_*).
// Ring shapes are tried in order of plane rows in flight, (S - 1) * G: the stream must cover the HBM latency
// (~35 KB per SM at 6.5 TB/s) while one group is being consumed.
int configure(M3Params& p, int n_rows, bool with_img, int want_small, int* threads, int* ctas) {
  const int groups = (p.W + kPX - 1) / kPX;
  const int ncw = (groups + 31) / 32;
  *threads = (ncw + 1) * 32;
  int want = *threads <= 192 ? want_small : 1;
  want = env_int("FALN_MED3_CTAS", want);
  if (want < 1) want = 1;
  if (*threads > 352) want = 1;
  else if (*threads > 192 && want > 2) want = 2;
  else if (want > 3) want = 3;
  static const int shapes[][2] = {{4, 4}, {4, 3}, {2, 5}, {2, 4}, {4, 2}, {2, 3}, {1, 5}, {2, 2}, {1, 3}, {1, 2}};   // {G, S}
  const int envS = env_int("FALN_MED3_S", 0), envG = env_int("FALN_MED3_G", 0);
  for (int ct = want; ct >= 1; --ct) {
    const int budget = (228 * 1024) / ct - 1024 - 64;
    for (const auto& sh : shapes) {
      const int G = envG > 0 ? envG : sh[0], S = envS > 1 ? (envS > 14 ? 14 : envS) : sh[1];
      const int need = make_smem(p.W, S, G, n_rows, with_img).total;
      if (need <= budget && need <= 227 * 1024) {
        p.S = S;
        p.G = G;
        *ctas = ct;
        return need;
      }
    }
  }
  return 0;
}

}  // namespace

// Returns 1 when the launch was made, 0 when the shape is not eligible (caller falls back), <0 on error.
int med3_launch_fwd(M3Params p, bool masks, cudaStream_t stream) {
  int threads = 0, ctas = 1;
  const int smem = configure(p, masks ? 3 : 0, true, 2, &threads, &ctas);
  if (!smem) return 0;
  int grid = sm_count() * ctas;
  if (grid > p.B * p.H) grid = p.B * p.H;
#define M3_LAUNCH_FWD(T, R)                                                                          \
  do {                                                                                               \
    auto kern = masks ? med3_fwd_kernel<true, T, R> : med3_fwd_kernel<false, T, R>;                  \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                   \
    kern<<<grid, threads, smem, stream>>>(p);                                                        \
  } while (0)
  // Register caps follow the PER-SUB-PARTITION register file (16,384 registers, warp w of a CTA on sub-partition w % 4):
  // k CTAs of nw warps put k * ceil(nw / 4) warps on sub-partition 0, so regs/thread <= 512 / (k * ceil(nw / 4)).
  // 6 warps (640 px): 3 CTAs -> 80, 2 -> 128; 11 warps (1242 px): 2 CTAs -> 80, 1 -> 128; 17 warps (2048 px): 96.
  if (threads <= 192 && ctas >= 3) M3_LAUNCH_FWD(192, 80);
  else if (threads <= 192) M3_LAUNCH_FWD(192, 128);
  else if (threads <= 352 && ctas >= 2) M3_LAUNCH_FWD(352, 80);
  else if (threads <= 352) M3_LAUNCH_FWD(352, 128);
  else M3_LAUNCH_FWD(544, 96);
#undef M3_LAUNCH_FWD
  const int rc = after_launch(masks ? "med3_fwd_kernel<masks>" : "med3_fwd_kernel");
  return rc == FALN_OK ? 1 : rc;
}

int med3_launch_bwd(M3Params p, cudaStream_t stream) {
  int threads = 0, ctas = 1;
  const int smem = configure(p, 6, false, 2, &threads, &ctas);
  if (!smem) return 0;
  int grid = sm_count() * ctas;
  if (grid > p.B * p.H) grid = p.B * p.H;
#define M3_LAUNCH_BWD(T, R)                                                                          \
  do {                                                                                               \
    auto kern = med3_bwd_kernel<T, R>;                                                               \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                   \
    kern<<<grid, threads, smem, stream>>>(p);                                                        \
  } while (0)
  if (threads <= 192 && ctas >= 3) M3_LAUNCH_BWD(192, 80);
  else if (threads <= 192) M3_LAUNCH_BWD(192, 128);
  else if (threads <= 352 && ctas >= 2) M3_LAUNCH_BWD(352, 80);
  else if (threads <= 352) M3_LAUNCH_BWD(352, 128);
  else M3_LAUNCH_BWD(544, 96);
#undef M3_LAUNCH_BWD
  const int rc = after_launch("med3_bwd_kernel");
  return rc == FALN_OK ? 1 : rc;
}

}  // namespace m3
}  // namespace faln
