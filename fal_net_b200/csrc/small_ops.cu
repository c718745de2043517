// small_ops.cu -- the small device-side helpers that keep library (ATen / cuBLAS) launches out of a training step.
//
// Round 1's launch list (profiles/r1e_launches_stage1.txt) showed ~75 ATen launches per Stage-1 step (258 per Stage-2
// step): exp / log / linspace of the level tables recomputed every forward, the logit-fold einsums on cuBLAS / cutlass SIMT
// sgemm + magma, zero-fills, strided copies and a dozen one-element multiplies / adds of the loss arithmetic.  Each is a
// few microseconds of a serial launch chain.  The kernels here replace them one for one:
//
//   level_tables_kernel      d_n, x_of_n of /root/reference/models/FAL_netB.py:204-205,224-225,241 (same fp32 op order)
//   fold_logit_conv_kernel   W' = W0 . W_iconv1 (:174,190,215 are linear with nothing between them), written directly as the
//                            bf16 forward pack [Np,3,3,C] and the data-gradient pack [C,3,3,Np]
//   fold_grad_kernel         the adjoint: dW0 = <gW', W_iconv1>, dW_iconv1 = W0^T gW', accumulated into the gradient arena
//   const_table_kernel       border-class sums of the constant max_disp/100 channel's weights (:145,208-209)
//   const_wgrad_kernel       weight gradient of that channel from the nine border-class sums of the output gradient
//   scalar_combine / scale   loss = sum_i w_i * term_i and its adjoint (Train_Stage1_K.py:258, Train_Stage2_K.py:309-327)
#include "common.cuh"

namespace faln {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// Level tables.  The reference evaluates, per level n (c = n / (N - 1), a Python double; c - 1 reaches the fp32 kernel
// rounded to fp32):  d = max_disp * exp(log(max_disp / min_disp) * (c - 1)),
//                    x_of = x_pix_max * exp(log(x_pix_max / x_pix_min) * (c - 1)),  x_pix_* = 2 * *_disp / W
// as separate fp32 ATen kernels, i.e. every operation rounds to fp32: replayed here with explicit _rn intrinsics (no FMA
// contraction) and the same libm entry points ATen's CUDA kernels call (expf / logf, not the fast intrinsics).
// ---------------------------------------------------------------------------------------------------------------
__global__ void level_tables_kernel(const float* __restrict__ min_disp, const float* __restrict__ max_disp,
                                    float* __restrict__ d, float* __restrict__ xo, int B, int N, float Wf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i % N;
  const float cm1 = (float)((double)n / (double)(N - 1) - 1.0);
  const float mn = min_disp[b], mx = max_disp[b];
  // `tensor / python_scalar` on CUDA is a multiplication by the fp32 reciprocal of the scalar (ATen div_true_kernel_cuda's
  // CPU-scalar fast path), not a division: replayed as such
  const float inv_w = __frcp_rn(Wf);
  const float xmin = __fmul_rn(__fmul_rn(2.f, mn), inv_w), xmax = __fmul_rn(__fmul_rn(2.f, mx), inv_w);
  d[i] = __fmul_rn(mx, expf(__fmul_rn(logf(__fdiv_rn(mx, mn)), cm1)));
  xo[i] = __fmul_rn(xmax, expf(__fmul_rn(logf(__fdiv_rn(xmax, xmin)), cm1)));
}

// element (o, c, kh, kw) of a 4-D weight given its four strides (the optimiser arena stores 3x3 weights KRSC, plain
// parameters are NCHW-contiguous: both are strided views of the same logical tensor)
struct W4 {
  long long so, sc, sh, sw;
  __device__ __forceinline__ long long at(int o, int c, int kh, int kw) const { return o * so + c * sc + kh * sh + kw * sw; }
};

// W'[o, c, t] = sum_m W0[o, m] * Wi1[m, c, t];  fwd pack [Np, 9, C] bf16 (rows >= N zero), dgrad pack [Cp, 9, Np] bf16
// (rows >= C zero; columns >= N zero).  A 49 x 49 x 864 matmul: a block owns kFoldCols (c, t) columns, stages W0 and its slice
// of W_iconv1 in shared memory, and 256 threads produce the Np x kFoldCols outputs.
constexpr int kFoldCols = 32;
__global__ void __launch_bounds__(256) fold_logit_conv_kernel(const float* __restrict__ w0, const float* __restrict__ wi1, W4 s,
                                                              __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dg,
                                                              int N, int C, int Np, int Cp) {
  extern __shared__ float sm[];
  float* sw0 = sm;                 // [N][N]
  float* swi = sm + N * N;         // [N][kFoldCols]
  const int col0 = blockIdx.x * kFoldCols;
  for (int i = threadIdx.x; i < N * N; i += 256) sw0[i] = w0[i];
  for (int i = threadIdx.x; i < N * kFoldCols; i += 256) {
    const int m = i / kFoldCols, col = col0 + i % kFoldCols;
    const int c = col / 9, t = col % 9;
    swi[i] = (c < C) ? __ldg(wi1 + s.at(m, c, t / 3, t % 3)) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Np * kFoldCols; i += 256) {
    const int o = i / kFoldCols, j = i % kFoldCols, col = col0 + j;
    const int c = col / 9, t = col % 9;
    if (c >= Cp) continue;
    float acc = 0.f;
    if (o < N && c < C)
      for (int m = 0; m < N; ++m) acc = fmaf(sw0[o * N + m], swi[m * kFoldCols + j], acc);
    const __nv_bfloat16 h = __float2bfloat16(acc);
    if (c < C) fwd[((long long)o * 9 + t) * C + c] = h;
    dg[((long long)c * 9 + t) * Np + o] = h;
  }
}

// gwf [N, 9, C] fp32 (KRSC, what the weight-gradient kernel wrote for the folded layer).
//   blocks [0, nA):      dWi1[m, c, t] += sum_o W0[o, m] * gwf[o, t, c]         (a block owns kFoldCols columns)
//   blocks [nA, nA+N):   dW0[o, m]     += sum_{c,t} gwf[o, t, c] * Wi1[m, c, t]  (a block owns row o: gwf[o] staged, one warp per m)
__global__ void __launch_bounds__(256) fold_grad_kernel(const float* __restrict__ gwf, const float* __restrict__ w0,
                                                        const float* __restrict__ wi1, W4 s, float* __restrict__ g_wi1, W4 gs,
                                                        float* __restrict__ g_w0, int N, int C, int nA) {
  extern __shared__ float sm[];
  if ((int)blockIdx.x < nA) {
    float* sw0 = sm;               // [N][N]
    float* sg = sm + N * N;        // [N][kFoldCols]
    const int col0 = blockIdx.x * kFoldCols;
    for (int i = threadIdx.x; i < N * N; i += 256) sw0[i] = w0[i];
    for (int i = threadIdx.x; i < N * kFoldCols; i += 256) {
      const int o = i / kFoldCols, col = col0 + i % kFoldCols;
      const int c = col / 9, t = col % 9;
      sg[i] = (c < C) ? __ldg(gwf + ((long long)o * 9 + t) * C + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * kFoldCols; i += 256) {
      const int m = i / kFoldCols, j = i % kFoldCols, col = col0 + j;
      const int c = col / 9, t = col % 9;
      if (c >= C) continue;
      float acc = 0.f;
      for (int o = 0; o < N; ++o) acc = fmaf(sw0[o * N + m], sg[o * kFoldCols + j], acc);
      g_wi1[gs.at(m, c, t / 3, t % 3)] += acc;
    }
    return;
  }
  const int o = blockIdx.x - nA;
  float* sgo = sm;                 // [9 * C]: gwf[o, :, :]
  for (int i = threadIdx.x; i < 9 * C; i += 256) sgo[i] = __ldg(gwf + (long long)o * 9 * C + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < N; m += 8) {
    float acc = 0.f;
    for (int i = lane; i < 9 * C; i += 32) {
      const int t = i / C, c = i % C;
      acc = fmaf(sgo[i], __ldg(wi1 + s.at(m, c, t / 3, t % 3)), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) g_w0[o * N + m] += acc;
  }
}

// ctab[cls, o] = sum over the taps of border class cls = rc * 4 + cc (bit0: first tap outside, bit1: last tap outside) that
// stay inside the image of w[o, ch, kh, kw]
__global__ void const_table_kernel(const float* __restrict__ w, W4 s, int ch, float* __restrict__ ctab, int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * Cout) return;
  const int cls = i / Cout, o = i % Cout, rc = cls >> 2, cc = cls & 3;
  float acc = 0.f;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const bool out = (kh == 0 && (rc & 1)) || (kh == 2 && (rc & 2)) || (kw == 0 && (cc & 1)) || (kw == 2 && (cc & 2));
      if (!out) acc += __ldg(w + s.at(o, ch, kh, kw));
    }
  ctab[i] = acc;
}

// dW[o, ch, kh, kw] += sum_b value[b] * sum over the border classes (r, c) whose tap (kh, kw) lands inside the image of
// S[b, r, c, o]   (S = nine border-class sums of the output gradient, csrc/conv_aux.cu border_sum_kernel)
__global__ void const_wgrad_kernel(const float* __restrict__ S, const float* __restrict__ value, float* __restrict__ dW, W4 gs,
                                   int ch, int B, int Cout, int last_row_clipped, int last_col_clipped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 9) return;
  const int o = i / 9, kh = (i % 9) / 3, kw = i % 3;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    float sb = 0.f;
    for (int r = 0; r < 3; ++r) {
      const bool vr = !(kh == 0 && r == 0) && !(kh == 2 && r == 2 && last_row_clipped);
      if (!vr) continue;
      for (int c = 0; c < 3; ++c) {
        const bool vc = !(kw == 0 && c == 0) && !(kw == 2 && c == 2 && last_col_clipped);
        if (vc) sb += __ldg(S + (((long long)b * 3 + r) * 3 + c) * Cout + o);
      }
    }
    acc = fmaf(__ldg(value + b), sb, acc);
  }
  dW[gs.at(o, ch, kh, kw)] += acc;
}

// Folded weights of an up-sample + conv block (csrc/conv_tc.cu, faln_conv3x3_up2_fwd / _dgrad): sums of the original taps in
// fp32, rounded once to bf16.  fwd [Cout_pad][16][Cin]: tap (ph*2+pw)*4 + a*2+b = sum over G(ph,a) x G(pw,b);
// dgrad [Cin_pad][16][Cout_pad]: tap (r+1)*4 + (c+1) = sum over Gr(r) x Gr(c).  One thread per (co, ci).
__device__ __forceinline__ void up2_group(int idx, int& lo, int& hi) {   // G / Gr sets as [lo, hi] tap ranges
  // idx 0: {0}, 1: {1,2}, 2: {0,1}, 3: {2}
  lo = (idx == 1) ? 1 : (idx == 3 ? 2 : 0);
  hi = (idx == 0) ? 0 : (idx == 2 ? 1 : 2);
}
__device__ __forceinline__ void pack_up2_body(const long long i, const float* __restrict__ w, const W4 s,
                                              __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dg, int Cout, int Cin,
                                              int Cout_pad, int Cin_pad) {
  if (i >= (long long)Cout_pad * Cin_pad) return;
  const int ci = (int)(i % Cin_pad), co = (int)(i / Cin_pad);
  float k[3][3];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) k[kh][kw] = (co < Cout && ci < Cin) ? __ldg(w + s.at(co, ci, kh, kw)) : 0.f;
  // forward: classes (ph, pw), taps (a, b): row group index = ph*2 + a -> {0}, {1,2}, {0,1}, {2}
  if (ci < Cin) {
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            int r0, r1, c0, c1;
            up2_group(ph * 2 + a, r0, r1);
            up2_group(pw * 2 + b, c0, c1);
            float acc = 0.f;
            for (int kh = r0; kh <= r1; ++kh)
              for (int kw = c0; kw <= c1; ++kw) acc += k[kh][kw];
            fwd[((long long)co * 16 + (ph * 2 + pw) * 4 + a * 2 + b) * Cin + ci] = __float2bfloat16(acc);
          }
  }
  // data gradient: window offsets r, c in {-1, 0, 1, 2} -> Gr(-1) = {2}, Gr(0) = {1,2}, Gr(1) = {0,1}, Gr(2) = {0}
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gmap[4] = {3, 1, 2, 0};   // offset -1 -> {2}, 0 -> {1,2}, +1 -> {0,1}, +2 -> {0}
      int r0, r1, c0, c1;
      up2_group(gmap[r], r0, r1);
      up2_group(gmap[c], c0, c1);
      float acc = 0.f;
      for (int kh = r0; kh <= r1; ++kh)
        for (int kw = c0; kw <= c1; ++kw) acc += k[kh][kw];
      dg[((long long)ci * 16 + r * 4 + c) * Cout_pad + co] = __float2bfloat16(acc);
    }
}

struct Terms {
  const float* p[8];
  float w[8];
  int n;
};
__global__ void scalar_combine_kernel(Terms t, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < t.n; ++i) acc = fmaf(t.w[i], *t.p[i], acc);
    *out = acc;
  }
}
__global__ void scalar_scale_kernel(const float* __restrict__ g, Terms t, float* __restrict__ out) {
  if (threadIdx.x < t.n && blockIdx.x == 0) out[threadIdx.x] = t.w[threadIdx.x] * (*g);
}

__global__ void __launch_bounds__(256) pack_up2_kernel(const float* __restrict__ w, W4 s, __nv_bfloat16* __restrict__ fwd,
                                                       __nv_bfloat16* __restrict__ dg, int Cout, int Cin, int Cout_pad, int Cin_pad) {
  pack_up2_body(blockIdx.x * 256LL + threadIdx.x, w, s, fwd, dg, Cout, Cin, Cout_pad, Cin_pad);
}

// All deconv levels' folded packs in ONE launch (six launches of ~5 us each sat in the forward's main chain): every block finds
// its job by its block index.
struct Up2PackJob {
  const float* w;
  W4 s;
  __nv_bfloat16* fwd;
  __nv_bfloat16* dg;
  int Cout, Cin, Cout_pad, Cin_pad, block0;
};
struct Up2PackBatch {
  Up2PackJob job[8];
  int njobs;
};
__global__ void __launch_bounds__(256) pack_up2_multi_kernel(const __grid_constant__ Up2PackBatch b) {
  int j = 0;
  for (int k = 1; k < b.njobs; ++k)
    if ((int)blockIdx.x >= b.job[k].block0) j = k;
  const Up2PackJob& q = b.job[j];
  pack_up2_body(((long long)blockIdx.x - q.block0) * 256LL + threadIdx.x, q.w, q.s, q.fwd, q.dg, q.Cout, q.Cin, q.Cout_pad,
                q.Cin_pad);
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_level_tables(const float* min_disp, const float* max_disp, float* d_lvl, float* x_of, int B, int N, int W,
                                 faln_stream_t stream) {
  FALN_REQUIRE(min_disp && max_disp && d_lvl && x_of && B > 0 && N >= 2 && W > 0, "faln_level_tables: bad argument");
  level_tables_kernel<<<(B * N + 127) / 128, 128, 0, as_stream(stream)>>>(min_disp, max_disp, d_lvl, x_of, B, N, (float)W);
  return after_launch("level_tables_kernel");
}

extern "C" int faln_fold_logit_conv(const float* w0, const float* w_iconv1, long long so, long long sc, long long sh,
                                    long long sw, void* fwd_pack, void* dgrad_pack, int N, int C, int Np, int Cp,
                                    faln_stream_t stream) {
  FALN_REQUIRE(w0 && w_iconv1 && fwd_pack && dgrad_pack && N > 0 && N <= 96 && C > 0 && Np >= N && Cp >= C,
               "faln_fold_logit_conv: bad argument (N <= 96)");
  // zero rows [N, Np) of the forward pack
  if (Np > N)
    FALN_REQUIRE(cudaMemsetAsync(static_cast<__nv_bfloat16*>(fwd_pack) + (size_t)N * 9 * C, 0, (size_t)(Np - N) * 9 * C * 2,
                                 as_stream(stream)) == cudaSuccess, "faln_fold_logit_conv: memset failed");
  const W4 s{so, sc, sh, sw};
  fold_logit_conv_kernel<<<(Cp * 9 + kFoldCols - 1) / kFoldCols, 256, (size_t)(N * N + N * kFoldCols) * 4, as_stream(stream)>>>(
      w0, w_iconv1, s, static_cast<__nv_bfloat16*>(fwd_pack), static_cast<__nv_bfloat16*>(dgrad_pack), N, C, Np, Cp);
  return after_launch("fold_logit_conv_kernel");
}

extern "C" int faln_fold_logit_conv_bwd(const float* gwf, const float* w0, const float* w_iconv1, long long so, long long sc,
                                        long long sh, long long sw, float* g_w_iconv1, long long gso, long long gsc,
                                        long long gsh, long long gsw, float* g_w0, int N, int C, faln_stream_t stream) {
  FALN_REQUIRE(gwf && w0 && w_iconv1 && g_w_iconv1 && g_w0 && N > 0 && N <= 96 && C > 0, "faln_fold_logit_conv_bwd: bad argument");
  const W4 s{so, sc, sh, sw}, gs{gso, gsc, gsh, gsw};
  const int nA = (C * 9 + kFoldCols - 1) / kFoldCols;
  size_t smem = (size_t)(N * N + N * kFoldCols) * 4;
  if (smem < (size_t)9 * C * 4) smem = (size_t)9 * C * 4;
  FALN_REQUIRE(smem <= 48 * 1024, "faln_fold_logit_conv_bwd: layer too wide for the staging buffers");
  fold_grad_kernel<<<nA + N, 256, smem, as_stream(stream)>>>(gwf, w0, w_iconv1, s, g_w_iconv1, gs, g_w0, N, C, nA);
  return after_launch("fold_grad_kernel");
}

extern "C" int faln_const_channel_table(const float* w, long long so, long long sc, long long sh, long long sw, int channel,
                                        float* ctab, int Cout, faln_stream_t stream) {
  FALN_REQUIRE(w && ctab && Cout > 0 && channel >= 0, "faln_const_channel_table: bad argument");
  const W4 s{so, sc, sh, sw};
  const_table_kernel<<<(16 * Cout + 127) / 128, 128, 0, as_stream(stream)>>>(w, s, channel, ctab, Cout);
  return after_launch("const_table_kernel");
}

extern "C" int faln_const_channel_wgrad(const float* border_sums, const float* value, float* dW, long long so, long long sc,
                                        long long sh, long long sw, int channel, int B, int Cout, int last_row_clipped,
                                        int last_col_clipped, faln_stream_t stream) {
  FALN_REQUIRE(border_sums && value && dW && B > 0 && Cout > 0 && channel >= 0, "faln_const_channel_wgrad: bad argument");
  const W4 gs{so, sc, sh, sw};
  const_wgrad_kernel<<<(Cout * 9 + 127) / 128, 128, 0, as_stream(stream)>>>(border_sums, value, dW, gs, channel, B, Cout,
                                                                           last_row_clipped, last_col_clipped);
  return after_launch("const_wgrad_kernel");
}

extern "C" int faln_scalar_combine(const float* const* terms, const float* weights, int n, float* out, faln_stream_t stream) {
  FALN_REQUIRE(terms && weights && out && n > 0 && n <= 8, "faln_scalar_combine: 1..8 terms");
  Terms t{};
  t.n = n;
  for (int i = 0; i < n; ++i) {
    FALN_REQUIRE(terms[i] != nullptr, "faln_scalar_combine: null term");
    t.p[i] = terms[i];
    t.w[i] = weights[i];
  }
  scalar_combine_kernel<<<1, 32, 0, as_stream(stream)>>>(t, out);
  return after_launch("scalar_combine_kernel");
}

extern "C" int faln_scalar_scale(const float* g, const float* weights, int n, float* out, faln_stream_t stream) {
  FALN_REQUIRE(g && weights && out && n > 0 && n <= 8, "faln_scalar_scale: 1..8 terms");
  Terms t{};
  t.n = n;
  for (int i = 0; i < n; ++i) t.w[i] = weights[i];
  scalar_scale_kernel<<<1, 32, 0, as_stream(stream)>>>(g, t, out);
  return after_launch("scalar_scale_kernel");
}

extern "C" int faln_pack_up2_weights(const float* w, long long so, long long sc, long long sh, long long sw, void* fwd_pack,
                                     void* dgrad_pack, int Cout, int Cin, int Cout_pad, int Cin_pad, faln_stream_t stream) {
  FALN_REQUIRE(w && fwd_pack && dgrad_pack && Cout > 0 && Cin > 0 && Cout_pad >= Cout && Cin_pad >= Cin,
               "faln_pack_up2_weights: bad argument");
  const W4 s{so, sc, sh, sw};
  const long long n = (long long)Cout_pad * Cin_pad;
  pack_up2_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(w, s, static_cast<__nv_bfloat16*>(fwd_pack),
                                                                        static_cast<__nv_bfloat16*>(dgrad_pack), Cout, Cin,
                                                                        Cout_pad, Cin_pad);
  return after_launch("pack_up2_kernel");
}

// The folded packs of several deconv layers in one launch; each job is what one faln_pack_up2_weights call takes.
extern "C" int faln_pack_up2_weights_multi(const faln_up2_pack_job_t* jobs, int njobs, faln_stream_t stream) {
  FALN_REQUIRE(jobs && njobs > 0 && njobs <= 8, "faln_pack_up2_weights_multi: 1..8 jobs");
  static thread_local Up2PackBatch b;
  int blocks = 0;
  for (int i = 0; i < njobs; ++i) {
    const faln_up2_pack_job_t& j = jobs[i];
    FALN_REQUIRE(j.w && j.fwd_pack && j.dgrad_pack && j.Cout > 0 && j.Cin > 0 && j.Cout_pad >= j.Cout && j.Cin_pad >= j.Cin,
                 "faln_pack_up2_weights_multi: bad job");
    b.job[i].w = j.w;
    b.job[i].s = W4{j.so, j.sc, j.sh, j.sw};
    b.job[i].fwd = static_cast<__nv_bfloat16*>(j.fwd_pack);
    b.job[i].dg = static_cast<__nv_bfloat16*>(j.dgrad_pack);
    b.job[i].Cout = j.Cout; b.job[i].Cin = j.Cin; b.job[i].Cout_pad = j.Cout_pad; b.job[i].Cin_pad = j.Cin_pad;
    b.job[i].block0 = blocks;
    blocks += (int)(((long long)j.Cout_pad * j.Cin_pad + 255) / 256);
  }
  b.njobs = njobs;
  pack_up2_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(b);
  return after_launch("pack_up2_multi_kernel");
}
