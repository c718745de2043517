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
// n_rows: per-row arrays of row_floats(W) each (array 0 = the affine-grid row); n_img: image record buffers
__host__ __device__ inline Smem make_smem(int W, int S, int G, int n_rows, int n_img) {
  Smem m;
  int o = 0;
  m.off_bar = o;   // full[S], empty[S], tab_full, tab_empty, aux_full[2], aux_empty[2]
  o += 256;
  m.off_cnt = o;   // int cnt[8]: planes per class 0..4
  o += 32;
  m.off_cls = o;   // class of plane n (bytes)
  o += kMaxN;
  m.off_tab = o;
  o += kMaxN * (int)sizeof(Ent);
  m.off_img = o;
  o += n_img * img_floats(W) * 4;
  m.off_rows = o;
  o += n_rows * row_floats(W) * 4;
  m.off_ring = o;
  o += S * G * slot_floats(W) * 4;
  m.total = o;
  return m;
}

// mbarrier.try_wait with a suspend-time hint: the thread may sleep up to `ns` and is woken by the completing arrive,
// instead of re-issuing the test every few cycles (the hint-less spin of the producer lane was 5 % of all issued
// instructions in profiles/r1e_med3_full_640x192_N49.txt)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
  }
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

// ------------------------------------------------------------------------------------------ pipeline state
struct Pipe {
  uint64_t* full;        // [S]  ring group filled (bulk-copy transaction bytes)
  uint64_t* empty;       // [S]  ring group released by every consumer warp
  uint64_t* tab_full;    //      plane table of the sample published
  uint64_t* tab_empty;   //      every consumer warp is done with the previous sample's table
  uint64_t* aux_full;    // [2]  per-row staging buffer (image records / backward rows) published by the producer warp
  uint64_t* aux_empty;   // [2]  ... released by every consumer warp
};
__device__ __forceinline__ Pipe make_pipe(unsigned char* smem, const Smem& L, int S) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  return Pipe{bars, bars + S, bars + 2 * S, bars + 2 * S + 1, bars + 2 * S + 2, bars + 2 * S + 4};
}

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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Copy warp: rows [r0, r1) of the flattened (b, y) index, `sweeps` passes over the planes of every row.  Builds the plane
// table of every new sample (all lanes), then lane 0 issues the bulk copies of the plane rows, group by group.
__device__ __forceinline__ void copy_warp(const M3Params& p, const Pipe& P, float* ring, Ent* tab, unsigned char* cls,
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
        mbar_wait_sleep(P.tab_empty, te);
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
            mbar_wait_sleep(&P.empty[slot], par ^ 1);
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

// Staging warp (forward): the image records of rows [r0, r1) into the nbuf record buffers, one buffer per row, handed
// over through aux_full / aux_empty.  With two buffers row r+1 is staged while row r is consumed; with one, staging
// overlaps the consumers' epilogue / mask sweep.  Four quads per lane are loaded before the first is stored, so a row
// costs a few load latencies, not one per quad.
__device__ __forceinline__ void stage_warp(const M3Params& p, const Pipe& P, float* img0, int imgf, int r0, int r1, int lane) {
  const int nbuf = p.nbuf, W = p.W, H = p.H;
  const int nq = ceil4(W) / 4;
  uint32_t ae = 0;   // bit buf = parity of the next aux_empty phase of that buffer
  for (int row = r0, it = 0; row < r1; ++row, ++it) {
    const int b = row / H, y = row % H;
    const int buf = it % nbuf;
    const float* img_b = p.image + ((size_t)b * 3 * H + y) * W;
    // pull the row into L2 while the buffer is still in use
    for (int q = lane; q < nq; q += 32) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) prefetch_l2(img_b + (size_t)ch * H * W + min(4 * q, W - 1));
    }
    if (it >= nbuf) {   // the buffer holds row it - nbuf until every consumer warp released it
      mbar_wait_sleep(&P.aux_empty[buf], (ae >> buf) & 1u);
      ae ^= 1u << buf;
    }
    float* img = img0 + (size_t)buf * imgf;
    for (int q0 = lane; q0 < nq; q0 += 128) {
      float4 v[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + 32 * u;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          v[u][ch] = q < nq ? load_row4(img_b + (size_t)ch * H * W, 4 * q, W) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + 32 * u;
        if (q < nq) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) stage_quad(img, ch, 4 * q, v[u][ch]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&P.aux_full[buf]);
  }
}

// ------------------------------------------------------------------------------------------ consumer side
struct Ring {
  const float* ring;   // element 0 of the payload of slot 0, plane 0
  uint64_t* full;
  uint64_t* empty;
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

// Common prologue: barriers, zeroed ring / image records, padded row arrays (array 0 = affine-grid row, arrays whose bit is
// set in inf_mask = -inf, the rest zero).
__device__ __forceinline__ void prologue(const M3Params& p, unsigned char* smem, const Smem& L, unsigned inf_mask,
                                         int n_rows, int n_img, int ncons) {
  const Pipe P = make_pipe(smem, L, p.S);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(&P.full[s], 1);
      mbar_init(&P.empty[s], ncons / 32);
    }
    mbar_init(P.tab_full, 1);
    mbar_init(P.tab_empty, ncons / 32);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&P.aux_full[k], 1);
      mbar_init(&P.aux_empty[k], ncons / 32);
    }
    fence_barrier_init();
  }
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const int nring = p.S * p.G * slot_floats(p.W);
  for (int i = tid; i < nring; i += blockDim.x) ring[i] = 0.f;
  float* img = reinterpret_cast<float*>(smem + L.off_img);
  for (int i = tid; i < n_img * img_floats(p.W); i += blockDim.x) img[i] = 0.f;
  float* rows = reinterpret_cast<float*>(smem + L.off_rows);
  const int rowf = row_floats(p.W);
  for (int i = tid; i < n_rows * rowf; i += blockDim.x) {
    const int a = i / rowf, j = i % rowf - kPad;
    float v = 0.f;
    if (a == 0) v = (j >= 0 && j < p.W) ? __ldg(p.g0x + j) : 0.f;
    else if ((inf_mask >> a) & 1u) v = -INFINITY;
    rows[i] = v;
  }
  fence_proxy_async();
  __syncthreads();
}

__device__ __forceinline__ void row_range(int rows, int& r0, int& r1) {
  r0 = (int)(((long long)rows * blockIdx.x) / gridDim.x);
  r1 = (int)(((long long)rows * (blockIdx.x + 1)) / gridDim.x);
}

// Consumer side of the per-sample table hand-shake.
__device__ __forceinline__ void next_sample(const Pipe& P, int b, int& cur_b, uint32_t& tf, int lane) {
  if (b == cur_b) return;
  if (cur_b >= 0) {
    __syncwarp();
    if (lane == 0) mbar_arrive(P.tab_empty);
  }
  mbar_wait(P.tab_full, tf);
  tf ^= 1;
  cur_b = b;
}

// =============================================================================================
// Forward.  The consumer warps meet NO CTA-wide barrier in the plain forward (one per row with masks, to publish the
// normaliser rows): the image records are staged by the producer warp into one of two buffers, overflow marks are
// per warp, so a warp can be in the epilogue of row r while its neighbours still run planes.
// =============================================================================================
template <bool kMasks, int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med3_fwd_kernel(const M3Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kRows = kMasks ? 5 : 0;   // g0row, 2 x (nl0row, nlwrow)
  const Smem L = make_smem(p.W, p.S, p.G, kRows, p.nbuf);
  int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
  unsigned char* cls = smem + L.off_cls;
  Ent* tab = reinterpret_cast<Ent*>(smem + L.off_tab);
  float* img0 = reinterpret_cast<float*>(smem + L.off_img);
  float* rows_s = reinterpret_cast<float*>(smem + L.off_rows);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const Pipe P = make_pipe(smem, L, p.S);

  const int ncons = blockDim.x - 64;   // consumer threads; then the copy warp and the staging warp
  const int tid = threadIdx.x;
  const int W = p.W, H = p.H;
  const int wr = ceil4(W), rowf = row_floats(W), imgf = img_floats(W);
  prologue(p, smem, L, kMasks ? 0x1eu : 0u, kRows, p.nbuf, ncons);
  int r0, r1;
  row_range(p.B * H, r0, r1);

  if (tid >= ncons) {
    if (tid < ncons + 32) copy_warp(p, P, ring, tab, cls, cnt, r0, r1, kMasks ? 2 : 1, tid & 31);
    else stage_warp(p, P, img0, imgf, r0, r1, tid & 31);
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
  const int lane = tid & 31;
  Ring rg{ring + kPad, P.full, P.empty, p.S, p.G, slot_floats(W), lane, 0, 0u, false, false, 0};
  int cur_b = -1;
  uint32_t tf = 0, af = 0;   // af: bit buf = parity of the next aux_full phase of that buffer

  for (int row = r0, it = 0; row < r1; ++row, ++it) {
    const int b = row / H, y = row % H;
    next_sample(P, b, cur_b, tf, lane);
    rg.left = groups_per_sweep(cnt, p.G) * (kMasks ? 2 : 1);
    rg.last_row = row == r1 - 1;
    const int buf = it % p.nbuf;
    const float* img = img0 + (size_t)buf * imgf;
    mbar_wait(&P.aux_full[buf], (af >> buf) & 1u);
    af ^= 1u << buf;

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
    __syncwarp();
    if (lane == 0) mbar_arrive(&P.aux_empty[buf]);   // this warp is done with the image records of the row

    bool bad = false;
    const size_t r1o = ((size_t)b * H + y) * W;
    float* nl0row = rows_s + (1 + 2 * (it & 1)) * rowf + kPad;   // normaliser rows alternate between two buffers
    float* nlwrow = nl0row + rowf;
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
    // overflow mark of the warp: NaN over the lse0 of its first pixel (lane 0 owns it), read by the clean-up launch
    if (__any_sync(0xffffffffu, bad) && lane == 0 && active) p.lse0[r1o + c.xb] = __int_as_float(0x7fc00000);

    if (kMasks) {
      // -------------------------------------------------------------- sweep B
      // the only CTA-wide barrier of the row: every warp's normalisers are published.  The two row buffers alternate,
      // so a warp that runs ahead into the next row writes the other buffer; it cannot get two rows ahead of this barrier.
      named_bar_sync(2, ncons);
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
// Backward.  The per-pixel inputs of a row (g_pan, <g_pan, pan>, the normalisers, g_disp, disp, the image window) are
// loaded by the consumer threads themselves, one quad each -- 19 independent loads, one memory latency per row -- and
// published through four row arrays between two CTA barriers.  (Measured on B200: staging these thirteen input rows with
// one extra warp instead loses 25 % at 1242 px and 2x at 2048 px -- a single warp serialises ~10 load latencies per row.)
// =============================================================================================
template <int kMaxThreads, int kMaxRegs>
__global__ void __launch_bounds__(kMaxThreads) __maxnreg__(kMaxRegs) med3_bwd_kernel(const M3Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kRows = 6;   // g0row, nlwrow, dotrow, gprow[3]
  const Smem L = make_smem(p.W, p.S, p.G, kRows, 0);
  int* cnt = reinterpret_cast<int*>(smem + L.off_cnt);
  unsigned char* cls = smem + L.off_cls;
  Ent* tab = reinterpret_cast<Ent*>(smem + L.off_tab);
  float* rows_s = reinterpret_cast<float*>(smem + L.off_rows);
  float* ring = reinterpret_cast<float*>(smem + L.off_ring);
  const Pipe P = make_pipe(smem, L, p.S);

  const int ncons = blockDim.x - 32;   // consumer threads; then the copy warp
  const int tid = threadIdx.x;
  const int W = p.W, H = p.H, N = p.N;
  const int rowf = row_floats(W);
  prologue(p, smem, L, 1u << 1, kRows, 0, ncons);
  int r0, r1;
  row_range(p.B * H, r0, r1);

  if (tid >= ncons) {
    copy_warp(p, P, ring, tab, cls, cnt, r0, r1, 1, tid & 31);
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
  const int lane = tid & 31;
  Ring rg{ring + kPad, P.full, P.empty, p.S, p.G, slot_floats(W), lane, 0, 0u, false, false, 0};
  int cur_b = -1;
  uint32_t tf = 0;

  for (int row = r0; row < r1; ++row) {
    const int b = row / H, y = row % H;
    next_sample(P, b, cur_b, tf, lane);

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
    float* const obase_p = p.g_logits + obase;
#define M3_BWD_CLASS(R)                                                                                          \
  class_run<1>(rg, ent, cnt[R], active, [&](const float* rowp, const Ent& e) {                                    \
    float g[4];                                                                                                   \
    bwd_plane<R>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, e, c, t, nxm1);                                     \
    store_row4(obase_p + e.src * oplane, c.xb, g, W);                                                             \
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
        store_row4(obase_p + e.src * oplane, c.xb, g, W);
      });
    }
  }
}

// ------------------------------------------------------------------------------------------ launch configuration
int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// Threads, ring shape, staging buffers and CTAs/SM for a row width.  Returns the dynamic shared-memory size, or 0 when
// nothing fits.  want_small: CTAs/SM to aim for on rows of <= 640 px (192 threads).  Wider rows run one CTA per SM at up to
// 128 registers: measured on B200 at 1242 px, two CTAs at the 80 registers the sub-partition register file then allows
// spill into local memory and lose (fwd 0.400 vs 0.373 ms, masks 0.788 vs 0.678, bwd 0.702 vs 0.574; gpurun_out/s2_*).
// Shared memory per CTA = fixed + nbuf image-record buffers (forward) + rows_fixed row arrays + the ring.  Two staging buffers are preferred (the producer warp stages row r+1 while row r is
// consumed); ring shapes are tried in order of plane rows in flight, (S - 1) * G: the stream must cover the HBM latency
// (~35 KB per SM at 6.5 TB/s) while one group is being consumed.
int configure(M3Params& p, int rows_fixed, int img_per_buf, int max_buf, int n_prod, int want_small, int* threads, int* ctas) {
  const int groups = (p.W + kPX - 1) / kPX;
  const int ncw = (groups + 31) / 32;
  *threads = (ncw + n_prod) * 32;   // consumers + copy warp (+ staging warp)
  int want = *threads <= 224 ? want_small : 1;
  want = env_int("FALN_MED3_CTAS", want);
  if (want < 1) want = 1;
  if (*threads > 384) want = 1;
  else if (*threads > 224 && want > 2) want = 2;
  else if (want > 3) want = 3;
  static const int shapes[][2] = {{4, 4}, {4, 3}, {2, 5}, {2, 4}, {4, 2}, {2, 3}, {1, 5}, {2, 2}, {1, 3}, {1, 2}};   // {G, S}
  const int envS = env_int("FALN_MED3_S", 0), envG = env_int("FALN_MED3_G", 0), envB = env_int("FALN_MED3_NBUF", 0);
  for (int ct = want; ct >= 1; --ct) {
    const int budget = (228 * 1024) / ct - 1024 - 64;
    for (int nbuf = (envB == 1 ? 1 : max_buf); nbuf >= (envB == 2 ? max_buf : 1); --nbuf) {
      int best = -1, best_need = 0;
      for (int k = 0; k < (int)(sizeof(shapes) / sizeof(shapes[0])); ++k) {
        const int G = envG > 0 ? envG : shapes[k][0], S = envS > 1 ? (envS > 13 ? 13 : envS) : shapes[k][1];
        const int need = make_smem(p.W, S, G, rows_fixed, nbuf * img_per_buf).total;
        if (need <= budget && need <= 227 * 1024) {
          best = k;
          best_need = need;
          p.S = S;
          p.G = G;
          break;
        }
      }
      // two staging buffers only if they still leave a ring of at least 8 plane rows in flight, in groups of four
      if (best >= 0 && (nbuf == 1 || ((p.S - 1) * p.G >= 8 && p.G >= 4) || envB == 2)) {
        p.nbuf = nbuf;
        *ctas = ct;
        return best_need;
      }
    }
  }
  return 0;
}

}  // namespace

// Returns 1 when the launch was made, 0 when the shape is not eligible (caller falls back), <0 on error.
int med3_launch_fwd(M3Params p, bool masks, cudaStream_t stream) {
  int threads = 0, ctas = 1;
  // masks: one record buffer (the mask sweep leaves the staging warp a whole sweep to refill it; measured better)
  const int smem = configure(p, masks ? 5 : 0, 1, masks ? 1 : 2, 2, 2, &threads, &ctas);
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
  // 7 warps (640 px): 3 CTAs -> 80, 2 -> 128; 12 warps (1242 px): 2 CTAs -> 80, 1 -> 128; 18 warps (2048 px): 96.
  if (threads <= 224 && ctas >= 3) M3_LAUNCH_FWD(224, 80);
  else if (threads <= 224) M3_LAUNCH_FWD(224, 128);
  else if (threads <= 384 && ctas >= 2) M3_LAUNCH_FWD(384, 80);
  else if (threads <= 384) M3_LAUNCH_FWD(384, 128);
  else M3_LAUNCH_FWD(576, 96);
#undef M3_LAUNCH_FWD
  const int rc = after_launch(masks ? "med3_fwd_kernel<masks>" : "med3_fwd_kernel");
  return rc == FALN_OK ? 1 : rc;
}

int med3_launch_bwd(M3Params p, cudaStream_t stream) {
  int threads = 0, ctas = 1;
  const int smem = configure(p, 6, 0, 1, 1, 2, &threads, &ctas);
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
