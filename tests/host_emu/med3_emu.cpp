// med3_emu.cpp -- HOST emulation of the per-thread code of the third-generation MED kernels.
//
// TEST INFRASTRUCTURE ONLY (built and loaded by tests/test_med3_emu.py; nothing in the product links it).
// It compiles fal_net_b200/csrc/med3_core.cuh -- the very functions the CUDA kernels in med3.cu call -- with g++ and
// runs them "thread" by "thread" over host arrays laid out exactly like the kernels' shared memory (padded ring slots,
// image records, -inf / zero padded row arrays, class-sorted plane table).  What it checks is the part of the kernels
// that cannot be seen by reading them: window indices, clamps into the padding, alignment classes, edge pixels.  The
// pipeline around it (bulk copies, mbarriers, launch shapes) is only exercised on the GPU.
//
//   g++ -O2 -ffp-contract=off -shared -fPIC -o libmed3_emu.so med3_emu.cpp
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../fal_net_b200/csrc/med3_core.cuh"

using namespace faln::m3;

namespace {

struct Table {
  std::vector<Ent> tab;
  int cnt[5];
};

// same two phases as build_table() in med3.cu (there: one lane per plane, __syncwarp between the phases)
Table build_table(const float* x_of, const float* d_lvl, int b, int N, int W, bool force) {
  Table t;
  std::vector<Ent> mine(N);
  std::vector<unsigned char> cls(N);
  for (int n = 0; n < N; ++n) {
    mine[n] = make_ent(x_of[(size_t)b * N + n], d_lvl[(size_t)b * N + n], n, W, force);
    cls[n] = (unsigned char)mine[n].cls;
  }
  t.tab.resize(N);
  for (int n = 0; n < N; ++n) t.tab[sorted_pos(cls.data(), N, n)] = mine[n];
  for (int c = 0; c < 5; ++c) {
    t.cnt[c] = 0;
    for (int n = 0; n < N; ++n) t.cnt[c] += cls[n] == c;
  }
  return t;
}

// ring slot of one plane row: [kPad zeros][ceil4(W) payload, pad columns zero][kTail zeros]
std::vector<float> make_slot(const float* row, int W) {
  std::vector<float> s(slot_floats(W), 0.f);
  memcpy(s.data() + kPad, row, sizeof(float) * W);
  return s;
}

float4 load_row4(const float* rowp, int xb, int W) {
  float4 v;
  v.x = xb < W ? rowp[xb] : 0.f;
  v.y = xb + 1 < W ? rowp[xb + 1] : 0.f;
  v.z = xb + 2 < W ? rowp[xb + 2] : 0.f;
  v.w = xb + 3 < W ? rowp[xb + 3] : 0.f;
  return v;
}
void store_row4(float* rowp, int xb, const float v[4], int W) {
  for (int i = 0; i < 4; ++i)
    if (xb + i < W) rowp[xb + i] = v[i];
}

PxCtx make_ctx(int tid, int W, const float* g0x) {
  PxCtx c;
  c.xb = tid * kPX;
  c.cW = 0.5f * (float)(W - 1);
  float g0[4];
  for (int i = 0; i < 4; ++i) g0[i] = g0x[min(c.xb + i, W - 1)];
  c.g0p[0] = make_float2(g0[0], g0[1]);
  c.g0p[1] = make_float2(g0[2], g0[3]);
  c.nxf[0] = make_float2(-(float)c.xb, -(float)(c.xb + 1));
  c.nxf[1] = make_float2(-(float)(c.xb + 2), -(float)(c.xb + 3));
  return c;
}

// row arrays as the kernel prologue leaves them: array 0 = g0 row, 1..n_inf = -inf, rest zero
std::vector<float> make_rows(int n_rows, int n_inf, int W, const float* g0x) {
  const int rowf = row_floats(W);
  std::vector<float> r((size_t)n_rows * rowf);
  for (int i = 0; i < n_rows * rowf; ++i) {
    const int a = i / rowf, j = i % rowf - kPad;
    float v = 0.f;
    if (a == 0) v = (j >= 0 && j < W) ? g0x[j] : 0.f;
    else if (a <= n_inf) v = -INFINITY;
    r[i] = v;
  }
  return r;
}

}  // namespace

extern "C" int emu_med3_fwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                            const float* d_lvl, float* pan, float* disp, float* maskL, float* maskR, float* lse0,
                            float* lsew, int B, int N, int H, int W, int masks, int force_generic) {
  const int wr = ceil4(W), rowf = row_floats(W);
  const int ncons = ((W + 3) / 4 + 31) / 32 * 32;
  int flagged = 0;
  std::vector<float> img(img_floats(W), 0.f);
  std::vector<float> rows = make_rows(3, 2, W, g0x);
  const float* g0row = rows.data() + kPad;
  float* nl0row = rows.data() + rowf + kPad;
  float* nlwrow = rows.data() + 2 * rowf + kPad;
  for (int b = 0; b < B; ++b) {
    const Table T = build_table(x_of, d_lvl, b, N, W, force_generic != 0);
    for (int y = 0; y < H; ++y) {
      std::vector<std::vector<float>> slots(N);
      for (int i = 0; i < N; ++i) slots[i] = make_slot(logits + (((size_t)b * N + T.tab[i].src) * H + y) * W, W);
      for (int tid = 0; tid < ncons; ++tid) {
        const int xb = tid * kPX;
        if (xb >= W) continue;
        for (int ch = 0; ch < 3; ++ch)
          stage_quad(img.data(), ch, xb, load_row4(image + (((size_t)b * 3 + ch) * H + y) * W, xb, W));
      }
      const size_t r1o = ((size_t)b * H + y) * W;
      std::vector<char> warp_bad(ncons / 32, 0);   // the kernel votes per warp (128 pixels) and marks lse0 of its first pixel
      for (int tid = 0; tid < ncons; ++tid) {
        PxCtx c = make_ctx(tid, W, g0x);
        if (c.xb >= W) continue;
        FwdAcc A;
        fwd_acc_init(A);
        for (int i = 0; i < N; ++i) {
          const float* rowp = slots[i].data() + kPad;
          switch (T.tab[i].cls) {
            case 0: fwd_plane<0>(A, rowp, img.data(), T.tab[i], c, wr); break;
            case 1: fwd_plane<1>(A, rowp, img.data(), T.tab[i], c, wr); break;
            case 2: fwd_plane<2>(A, rowp, img.data(), T.tab[i], c, wr); break;
            case 3: fwd_plane<3>(A, rowp, img.data(), T.tab[i], c, wr); break;
            default: fwd_plane_generic(A, rowp, img.data(), T.tab[i], c, W); break;
          }
        }
        float dv[4], pv[3][4], l0[4], lw[4], nl0[4], nlw[4];
        if (fwd_finish(A, c.xb, W, dv, pv, l0, lw, nl0, nlw)) warp_bad[tid / 32] = 1;
        store_row4(disp + r1o, c.xb, dv, W);
        for (int ch = 0; ch < 3; ++ch) store_row4(pan + (((size_t)b * 3 + ch) * H + y) * W, c.xb, pv[ch], W);
        store_row4(lse0 + r1o, c.xb, l0, W);
        store_row4(lsew + r1o, c.xb, lw, W);
        st4(nl0row + c.xb, make_float4(nl0[0], nl0[1], nl0[2], nl0[3]));
        st4(nlwrow + c.xb, make_float4(nlw[0], nlw[1], nlw[2], nlw[3]));
      }
      bool any_bad = false;
      for (int w = 0; w < ncons / 32; ++w)
        if (warp_bad[w]) {
          lse0[r1o + 128 * w] = NAN;
          any_bad = true;
        }
      if (any_bad) ++flagged;
      if (!masks) continue;
      for (int tid = 0; tid < ncons; ++tid) {
        PxCtx c = make_ctx(tid, W, g0x);
        if (c.xb >= W) continue;
        const float nxm1 = -(float)(c.xb - 1);
        float mR[4] = {0, 0, 0, 0}, mL[4] = {0, 0, 0, 0};
        for (int i = 0; i < N; ++i) {
          const float* rowp = slots[i].data() + kPad;
          switch (T.tab[i].cls) {
            case 0: mask_plane<0>(mR, mL, rowp, nl0row, nlwrow, g0row, T.tab[i], c, nxm1, wr); break;
            case 1: mask_plane<1>(mR, mL, rowp, nl0row, nlwrow, g0row, T.tab[i], c, nxm1, wr); break;
            case 2: mask_plane<2>(mR, mL, rowp, nl0row, nlwrow, g0row, T.tab[i], c, nxm1, wr); break;
            case 3: mask_plane<3>(mR, mL, rowp, nl0row, nlwrow, g0row, T.tab[i], c, nxm1, wr); break;
            default: mask_plane_generic(mR, mL, rowp, nl0row, nlwrow, g0row, T.tab[i], c, W); break;
          }
        }
        float o[4];
        for (int i = 0; i < 4; ++i) o[i] = fminf(mL[i], 1.0f);
        store_row4(maskL + r1o, c.xb, o, W);
        for (int i = 0; i < 4; ++i) o[i] = fminf(mR[i], 1.0f);
        store_row4(maskR + r1o, c.xb, o, W);
      }
    }
  }
  return flagged;
}

extern "C" int emu_med3_bwd(const float* logits, const float* image, const float* g0x, const float* x_of,
                            const float* d_lvl, const float* pan, const float* disp, const float* lse0,
                            const float* lsew, const float* g_pan, const float* g_disp, float* g_logits, int B, int N,
                            int H, int W, int force_generic) {
  const int rowf = row_floats(W);
  const int ncons = ((W + 3) / 4 + 31) / 32 * 32;
  std::vector<float> rows = make_rows(6, 1, W, g0x);
  const float* g0row = rows.data() + kPad;
  float* nlwrow = rows.data() + rowf + kPad;
  float* dotrow = rows.data() + 2 * rowf + kPad;
  float* gprow = rows.data() + 3 * rowf + kPad;
  for (int b = 0; b < B; ++b) {
    const Table T = build_table(x_of, d_lvl, b, N, W, force_generic != 0);
    for (int y = 0; y < H; ++y) {
      std::vector<std::vector<float>> slots(N);
      for (int i = 0; i < N; ++i) slots[i] = make_slot(logits + (((size_t)b * N + T.tab[i].src) * H + y) * W, W);
      const size_t r1o = ((size_t)b * H + y) * W;
      std::vector<BwdCtx> ctx(ncons);
      for (int tid = 0; tid < ncons; ++tid) {
        const int xb = tid * kPX;
        if (xb >= W) continue;
        BwdCtx& t = ctx[tid];
        float dot[4] = {0, 0, 0, 0};
        for (int ch = 0; ch < 3; ++ch) {
          const size_t rc = (((size_t)b * 3 + ch) * H + y) * W;
          const float4 gq = load_row4(g_pan + rc, xb, W);
          const float4 pq = load_row4(pan + rc, xb, W);
          dot[0] = fmaf(gq.x, pq.x, dot[0]);
          dot[1] = fmaf(gq.y, pq.y, dot[1]);
          dot[2] = fmaf(gq.z, pq.z, dot[2]);
          dot[3] = fmaf(gq.w, pq.w, dot[3]);
          st4(gprow + ch * rowf + xb, gq);
          const float4 iq = load_row4(image + rc, xb, W);
          t.iw[ch][0] = xb > 0 ? image[rc + xb - 1] : 0.f;
          t.iw[ch][1] = iq.x; t.iw[ch][2] = iq.y; t.iw[ch][3] = iq.z; t.iw[ch][4] = iq.w;
          t.iw[ch][5] = xb + 4 < W ? image[rc + xb + 4] : 0.f;
        }
        st4(dotrow + xb, make_float4(dot[0], dot[1], dot[2], dot[3]));
        const float4 lw = load_row4(lsew + r1o, xb, W), l0 = load_row4(lse0 + r1o, xb, W);
        const float4 gd = load_row4(g_disp + r1o, xb, W), dp = load_row4(disp + r1o, xb, W);
        const float lwv[4] = {lw.x, lw.y, lw.z, lw.w}, l0v[4] = {l0.x, l0.y, l0.z, l0.w};
        const float gdv[4] = {gd.x, gd.y, gd.z, gd.w}, dpv[4] = {dp.x, dp.y, dp.z, dp.w};
        float nl[4];
        for (int i = 0; i < 4; ++i) {
          nl[i] = xb + i < W ? -lwv[i] * kLog2e : -INFINITY;
          t.nl0[i] = -l0v[i] * kLog2e;
          t.gd[i] = gdv[i];
          t.ndsp[i] = -dpv[i];
        }
        st4(nlwrow + xb, make_float4(nl[0], nl[1], nl[2], nl[3]));
      }
      for (int tid = 0; tid < ncons; ++tid) {
        PxCtx c = make_ctx(tid, W, g0x);
        if (c.xb >= W) continue;
        const float nxm1 = -(float)(c.xb - 1);
        for (int i = 0; i < N; ++i) {
          const float* rowp = slots[i].data() + kPad;
          float g[4];
          switch (T.tab[i].cls) {
            case 0: bwd_plane<0>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, T.tab[i], c, ctx[tid], nxm1); break;
            case 1: bwd_plane<1>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, T.tab[i], c, ctx[tid], nxm1); break;
            case 2: bwd_plane<2>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, T.tab[i], c, ctx[tid], nxm1); break;
            case 3: bwd_plane<3>(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, T.tab[i], c, ctx[tid], nxm1); break;
            default: {
              const float* ir = image + (((size_t)b * 3) * H + y) * W;
              bwd_plane_generic(g, rowp, nlwrow, dotrow, gprow, rowf, g0row, T.tab[i], c, ctx[tid], ir, ir + (size_t)H * W,
                                ir + 2 * (size_t)H * W, W);
            } break;
          }
          store_row4(g_logits + (((size_t)b * N + T.tab[i].src) * H + y) * W, c.xb, g, W);
        }
      }
    }
  }
  return 0;
}
