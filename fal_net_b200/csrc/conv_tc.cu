// conv_tc.cu -- 3x3 convolution as a tcgen05 / TMEM implicit GEMM (bf16 in, fp32 accumulate), sm_100a.
//
// Replaces the cuDNN calls behind nn.Conv2d in /root/reference/models/FAL_netB.py:99-127,144-174 and the VGG
// slices of /root/reference/loss_functions.py:21-29, with bias / ELU / ReLU / residual fused in the epilogue
// (the reference runs them as separate ATen kernels, :47,59,79).
//
// GEMM view:  D[pixel, cout] = sum over (tap, cin) of  A[pixel shifted by tap, cin] * Wt[cout, tap, cin]
//   * activations are NHWC bf16; an output tile is 8 rows x 16 columns = 128 pixels = the UMMA M dimension
//   * per K step (one filter tap x one block of BK input channels) the TMA unit fetches
//       A: a 4-D box [1, 8*s, 16*s, BK] of the input at the tap's offset (element stride s = conv stride),
//          out-of-image taps zero-filled by the TMA bounds check = the conv's zero padding
//       B: a 2-D box [BN, BK] of the KRSC weight matrix
//     both land in shared memory in the 128B(64B)-swizzled K-major layout tcgen05.mma consumes directly
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator; tcgen05.commit
//     releases the smem stage / signals the epilogue through mbarriers
//   * a second (concatenated) source is just more K steps with its own tensor map: torch.cat never happens
//   * epilogue warps read TMEM with tcgen05.ld, add bias / residual, apply ELU / ReLU, and write either bf16
//     NHWC or fp32 planar (the logits layout the MED kernels stream).
#include "tc_common.cuh"

namespace faln {
namespace {

constexpr int kTH = 8;  // output tile: kTH x kTW = 8 x 16 pixels = 128 = UMMA_M

// One "tap class": the filter taps that contribute to the output pixels (r * out_mul + oh, c * out_mul + ow).
//   forward / stride-1 dgrad: one class with 9 taps; stride-2 dgrad: four parity classes with 1, 2, 2, 4 taps.
// Tap t reads the input tile at offset (dh[t], dw[t]) and the weight columns of filter tap wt[t].
struct TapClass {
  int n;
  signed char dh[9], dw[9], wt[9];
  signed char oh, ow;
};

struct ConvParams {
  int B, H, W, Ho, Wo;     // input dims (TMA coordinates), tile-grid dims (r, c)
  int C1, C2, Cin, Cout;   // Cin = weight-row length per tap (elements)
  int stride, act, planar; // stride = element stride of the input box
  int tiles_w, tiles_h;
  int kblocks1, kblocks2;  // BK-channel blocks of source 1 / source 2
  int ncls;
  TapClass cls[4];
  int out_mul, out_H, out_W;  // output pixel = (r * out_mul + oh, c * out_mul + ow) inside an out_H x out_W image
  int accum;                  // 1: add to what the output tensor already holds (gradient accumulation)
  int dact;                   // 0 none, 1: multiply by ELU'(ysave), 2: multiply by ReLU'(ysave)  (backward epilogues)
  const __nv_bfloat16* ysave; // saved activation output the derivative is taken at, NHWC with ysave_c channels
  int ysave_c, res_c;
  const float* bias;
  const float* ctab;    // [16][Cout] border-class sums of the constant-channel weights, or NULL
  const float* cscale;  // [B] value of the constant channel per sample
  const __nv_bfloat16* residual;
  void* out;
  long long out_pitch;
  int out_c;  // channel count (stride) of the NHWC output tensor
};

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major), [32,46) SBO >> 4 = 8 rows * row bytes,
//   [46,48) version = 1 (sm_100), [61,64) layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  constexpr uint64_t row_bytes = BK * 2;
  constexpr uint64_t sbo = 8 * row_bytes;
  constexpr uint64_t layout = (BK == 64) ? 2 : 4;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ULL << 16) | ((sbo >> 4) << 32) | (1ULL << 46) | (layout << 61);
}
// tcgen05 instruction descriptor, kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format BF16
// (bits 7, 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : (__expf(v) - 1.0f); }

template <int BK, int BN, int STAGES>
struct SmemLayout {
  static constexpr int kA = 128 * BK * 2;
  static constexpr int kB = BN * BK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kBars = 1024;  // barriers + tmem pointer
  static constexpr int kTotal = kBars + STAGES * kStage + 1024 /* alignment slack */;
};

template <int BK, int BN, int STAGES>
__global__ void __launch_bounds__(192)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using SL = SmemLayout<BK, BN, STAGES>;
  extern __shared__ unsigned char smem_raw[];
  // barriers first, then 1024B-aligned operand stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + SL::kBars + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
  const int n0 = blockIdx.y * BN;
  const int ho0 = th * kTH, wo0 = tw * kTW;
  const TapClass& tc = p.cls[blockIdx.z];
  const int ksteps = tc.n * (p.kblocks1 + p.kblocks2);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW);
    if (p.kblocks2) prefetch_tmap(&tmA2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < tc.n; ++t) {
        const int tap = tc.wt[t];
        const int hi = ho0 * p.stride + tc.dh[t], wi = wo0 * p.stride + tc.dw[t];
        for (int cb = 0; cb < p.kblocks1 + p.kblocks2; ++cb) {
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char* a_dst = stages + (size_t)s * SL::kStage;
          unsigned char* b_dst = a_dst + SL::kA;
          mbar_arrive_expect_tx(&full[s], SL::kStage);
          if (cb < p.kblocks1) {
            tma_load_4d(a_dst, &tmA1, cb * BK, wi, hi, b, &full[s]);
            tma_load_2d(b_dst, &tmW, tap * p.Cin + cb * BK, n0, &full[s]);
          } else {
            tma_load_4d(a_dst, &tmA2, (cb - p.kblocks1) * BK, wi, hi, b, &full[s]);
            tma_load_2d(b_dst, &tmW, tap * p.Cin + p.C1 + (cb - p.kblocks1) * BK, n0, &full[s]);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int s = 0;
      uint32_t ph = 0;
      for (int k = 0; k < ksteps; ++k) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(stages + (size_t)s * SL::kStage);
        const uint64_t adesc = make_desc<BK>(a_addr);
        const uint64_t bdesc = make_desc<BK>(a_addr + SL::kA);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
        }
        umma_commit(&empty[s]);  // frees the smem stage when these MMAs have read it
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_full);  // accumulator complete
    }
  } else {
    // ================================================================= epilogue (warps 2..5)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;       // row of the tile = pixel
    const int ho = (ho0 + m / kTW) * p.out_mul + tc.oh, wo = (wo0 + m % kTW) * p.out_mul + tc.ow;
    const bool valid = ho < p.out_H && wo < p.out_W;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const size_t pix = ((size_t)b * p.out_H + ho) * p.out_W + wo;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, r);
      const int cg = n0 + c0;  // first global output channel of this chunk
      if (!valid || cg >= p.Cout) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cg + j < p.Cout) v[j] += __ldg(p.bias + cg + j);
      }
      if (p.ctab) {
        // a spatially constant extra input channel (the max_disp/100 plane of reference :145,208-209) contributes
        // value * (sum of its weights over the taps that fall inside the image): a per-border-class bias
        const int rc = (ho == 0 ? 1 : 0) | (p.stride * ho + 1 > p.H - 1 ? 2 : 0);
        const int cc = (wo == 0 ? 1 : 0) | (p.stride * wo + 1 > p.W - 1 ? 2 : 0);
        const float* t = p.ctab + (size_t)(rc * 4 + cc) * p.Cout + cg;
        const float sc = __ldg(p.cscale + b);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cg + j < p.Cout) v[j] = fmaf(sc, __ldg(t + j), v[j]);
      }
      if (p.accum) {
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.out) + pix * p.out_c + cg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u = rp[q];
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h2[e]);
            v[q * 8 + 2 * e] += f.x;
            v[q * 8 + 2 * e + 1] += f.y;
          }
        }
      }
      if (p.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.res_c + cg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u = __ldg(rp + q);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h2[e]);
            v[q * 8 + 2 * e] += f.x;
            v[q * 8 + 2 * e + 1] += f.y;
          }
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = elu1(v[j]);
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.dact) {
        // backward epilogue: gradient w.r.t. the pre-activation = gradient * act'(pre), from the SAVED OUTPUT y:
        // ELU'(pre) = 1 (y > 0) else y + 1;  ReLU'(pre) = [y > 0]
        const uint4* yp = reinterpret_cast<const uint4*>(p.ysave + pix * p.ysave_c + cg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u = __ldg(yp + q);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h2[e]);
            const float d0 = f.x > 0.f ? 1.f : (p.dact == 1 ? f.x + 1.f : 0.f);
            const float d1 = f.y > 0.f ? 1.f : (p.dact == 1 ? f.y + 1.f : 0.f);
            v[q * 8 + 2 * e] *= d0;
            v[q * 8 + 2 * e + 1] *= d1;
          }
        }
      }
      if (p.planar) {
        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cg + j < p.Cout) o[(((size_t)b * p.Cout + cg + j) * p.out_H + ho) * p.out_pitch + wo] = v[j];
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + cg;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
          *reinterpret_cast<uint4*>(o + q * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// ------------------------------------------------------------------------------------------ host: tensor maps
bool make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int BK, int BN) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BK, int BN, int STAGES>
int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const ConvParams& p, cudaStream_t st) {
  using SL = SmemLayout<BK, BN, STAGES>;
  auto kern = conv3x3_tc_kernel<BK, BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal);
    attr_set = true;
  }
  dim3 grid(p.tiles_w * p.tiles_h * p.B, (p.Cout + BN - 1) / BN, p.ncls);
  kern<<<grid, 192, SL::kTotal, st>>>(a1, a2, w, p);
  return after_launch("conv3x3_tc_kernel");
}

// Stage counts: a CTA owns ONE 128-pixel tile and runs only 9 * Cin / BK k-steps, so what hides latency is several CTAs
// per SM (prologue, TMA latency and epilogue of different tiles overlapping), not a deep ring: 3-4 stages, sized so that
// 2-5 CTAs fit in shared memory (and their accumulators in the 512 TMEM columns).
int dispatch(int BK, int BN, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& wm, const ConvParams& p,
             cudaStream_t st) {
  if (BK == 64) {
    switch (BN) {
      case 256: return launch<64, 256, 4>(a1, a2, wm, p, st);
      case 128: return launch<64, 128, 3>(a1, a2, wm, p, st);
      case 64: return launch<64, 64, 3>(a1, a2, wm, p, st);
      default: return launch<64, 32, 3>(a1, a2, wm, p, st);
    }
  }
  switch (BN) {
    case 256: return launch<32, 256, 4>(a1, a2, wm, p, st);
    case 128: return launch<32, 128, 3>(a1, a2, wm, p, st);
    case 64: return launch<32, 64, 4>(a1, a2, wm, p, st);
    default: return launch<32, 32, 4>(a1, a2, wm, p, st);
  }
}

}  // namespace
}  // namespace faln

using namespace faln;

// x [B,H,W,C1] bf16 NHWC, x2 [B,H,W,C2] or NULL (channel-concatenated after x), w [Cout_pad, 3, 3, C1+C2] bf16 (KRSC, rows
// beyond Cout zero), bias [Cout] fp32 or NULL, residual [B,Ho,Wo,out_c] bf16 or NULL.
// y: bf16 NHWC [B,Ho,Wo,out_c] (planar = 0) or fp32 planar [B,Cout,Ho,out_pitch] (planar = 1).
extern "C" int faln_conv3x3_fwd(const void* x, const void* x2, const void* w, const float* bias, const float* ctab,
                                const float* cscale, const void* residual, void* y, int B, int H, int W, int C1, int C2, int Cout, int Cout_pad, int stride, int act,
                                int planar, long long out_pitch, int out_c, faln_stream_t stream) {
  FALN_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "faln_conv3x3_fwd: null pointer / bad shape");
  FALN_REQUIRE(stride == 1 || stride == 2, "faln_conv3x3_fwd: stride must be 1 or 2");
  FALN_REQUIRE((ctab == nullptr) == (cscale == nullptr), "faln_conv3x3_fwd: ctab and cscale go together");
  FALN_REQUIRE(C1 % 32 == 0 && C2 % 32 == 0 && C1 > 0 && (x2 != nullptr) == (C2 > 0),
               "faln_conv3x3_fwd: channel counts must be multiples of 32 (got %d + %d)", C1, C2);
  FALN_REQUIRE(Cout > 0 && Cout_pad >= Cout && Cout_pad % 32 == 0, "faln_conv3x3_fwd: Cout_pad must be a multiple of 32");
  FALN_REQUIRE(planar || (out_c % 8 == 0 && out_c >= Cout && Cout % 32 == 0),
               "faln_conv3x3_fwd: NHWC output needs Cout %% 32 == 0 and out_c %% 8 == 0, out_c >= Cout");
  FALN_REQUIRE(!planar || out_pitch >= (W - 1) / stride + 1, "faln_conv3x3_fwd: out_pitch too small");
  const int BK = (C1 % 64 == 0 && C2 % 64 == 0) ? 64 : 32;
  int BN = Cout_pad % 256 == 0 ? 256 : (Cout_pad % 128 == 0 ? 128 : (Cout_pad % 64 == 0 ? 64 : 32));
  ConvParams p{};
  p.B = B; p.H = H; p.W = W;
  p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  p.C1 = C1; p.C2 = C2; p.Cin = C1 + C2; p.Cout = Cout;
  p.stride = stride; p.act = act; p.planar = planar;
  p.tiles_w = (p.Wo + kTW - 1) / kTW; p.tiles_h = (p.Ho + kTH - 1) / kTH;
  p.kblocks1 = C1 / BK; p.kblocks2 = C2 / BK;
  p.bias = bias; p.ctab = ctab; p.cscale = cscale; p.residual = static_cast<const __nv_bfloat16*>(residual); p.out = y;
  p.out_pitch = out_pitch; p.out_c = out_c;
  p.ncls = 1;
  p.cls[0].n = 9;
  for (int t = 0; t < 9; ++t) {
    p.cls[0].dh[t] = (signed char)(t / 3 - 1);
    p.cls[0].dw[t] = (signed char)(t % 3 - 1);
    p.cls[0].wt[t] = (signed char)t;
  }
  p.cls[0].oh = p.cls[0].ow = 0;
  p.out_mul = 1; p.out_H = p.Ho; p.out_W = p.Wo;
  p.accum = 0; p.dact = 0; p.ysave = nullptr; p.ysave_c = 0; p.res_c = out_c;
  CUtensorMap a1, a2, wm;
  if (!make_act_map(&a1, x, B, H, W, C1, BK, stride) || !make_w_map(&wm, w, Cout_pad, 9 * (C1 + C2), BK, BN) ||
      (x2 && !make_act_map(&a2, x2, B, H, W, C2, BK, stride))) {
    set_error("faln_conv3x3_fwd: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  if (!x2) a2 = a1;
  return dispatch(BK, BN, a1, a2, wm, p, as_stream(stream));
}

// Data gradient of the 3x3 convolution (pad 1, stride 1 or 2) on the same tcgen05 kernel:
//   gx[b,h,w,ci] = sum_{kh,kw,co} g[b,(h+1-kh)/s,(w+1-kw)/s,co] * W[co,kh,kw,ci]      (terms with non-integer /s dropped)
// g  [B,Hg,Wg,Cg] bf16 NHWC: gradient w.r.t. the conv's pre-activation output (Cg = padded Cout, multiple of 32)
// wd [Cx_pad,3,3,Cg] bf16: the weights re-packed per INPUT channel (rows) -- for a concatenated input the caller passes
//    the row range of one source and calls once per source
// gx [B,H,W,gx_c] bf16 NHWC, channels [0,Cx) written (accum: added to); optional fused epilogue
//    gx = (gx_old + dgrad + residual) * act'(ysave)   -- the chain rule through the producer's ELU / ReLU.
// Stride 2 runs the four output-parity classes (1, 2, 2, 4 contributing taps) as blockIdx.z of one launch.
extern "C" int faln_conv3x3_dgrad(const void* g, const void* wd, void* gx, const void* residual, const void* ysave, int B,
                                  int H, int W, int Cg, int Cx, int Cx_pad, int stride, int accum, int dact, int gx_c,
                                  int res_c, int ysave_c, faln_stream_t stream) {
  FALN_REQUIRE(g && wd && gx && B > 0 && H > 0 && W > 0, "faln_conv3x3_dgrad: null pointer / bad shape");
  FALN_REQUIRE(stride == 1 || stride == 2, "faln_conv3x3_dgrad: stride must be 1 or 2");
  FALN_REQUIRE(Cg > 0 && Cg % 32 == 0, "faln_conv3x3_dgrad: Cg must be a multiple of 32 (got %d)", Cg);
  FALN_REQUIRE(Cx > 0 && Cx % 32 == 0 && Cx_pad >= Cx && Cx_pad % 32 == 0 && gx_c >= Cx && gx_c % 8 == 0,
               "faln_conv3x3_dgrad: Cx must be a multiple of 32 and fit the output tensor");
  FALN_REQUIRE((dact == 0) == (ysave == nullptr), "faln_conv3x3_dgrad: dact and ysave go together");
  const int Hg = (H - 1) / stride + 1, Wg = (W - 1) / stride + 1;
  const int BK = (Cg % 64 == 0) ? 64 : 32;
  const int BN = Cx_pad % 256 == 0 ? 256 : (Cx_pad % 128 == 0 ? 128 : (Cx_pad % 64 == 0 ? 64 : 32));
  ConvParams p{};
  p.B = B; p.H = Hg; p.W = Wg;
  p.Ho = (H + stride - 1) / stride; p.Wo = (W + stride - 1) / stride;   // tile grid over one parity class
  p.C1 = Cg; p.C2 = 0; p.Cin = Cg; p.Cout = Cx;
  p.stride = 1; p.act = 0; p.planar = 0;
  p.tiles_w = (p.Wo + kTW - 1) / kTW; p.tiles_h = (p.Ho + kTH - 1) / kTH;
  p.kblocks1 = Cg / BK; p.kblocks2 = 0;
  p.bias = nullptr; p.ctab = nullptr; p.cscale = nullptr;
  p.residual = static_cast<const __nv_bfloat16*>(residual); p.out = gx; p.out_pitch = 0; p.out_c = gx_c;
  p.out_mul = stride; p.out_H = H; p.out_W = W;
  p.accum = accum; p.dact = dact; p.ysave = static_cast<const __nv_bfloat16*>(ysave); p.ysave_c = ysave_c; p.res_c = res_c;
  if (stride == 1) {
    p.ncls = 1;
    p.cls[0].n = 9;
    for (int t = 0; t < 9; ++t) {
      p.cls[0].dh[t] = (signed char)(1 - t / 3);
      p.cls[0].dw[t] = (signed char)(1 - t % 3);
      p.cls[0].wt[t] = (signed char)t;
    }
    p.cls[0].oh = p.cls[0].ow = 0;
  } else {
    p.ncls = 4;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        TapClass& c = p.cls[ph * 2 + pw];
        c.n = 0; c.oh = (signed char)ph; c.ow = (signed char)pw;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            if (((ph + 1 - kh) & 1) || ((pw + 1 - kw) & 1)) continue;
            c.dh[c.n] = (signed char)((ph + 1 - kh) / 2);
            c.dw[c.n] = (signed char)((pw + 1 - kw) / 2);
            c.wt[c.n] = (signed char)(kh * 3 + kw);
            ++c.n;
          }
      }
  }
  CUtensorMap a1, wm;
  if (!make_act_map(&a1, g, B, Hg, Wg, Cg, BK, 1) || !make_w_map(&wm, wd, Cx_pad, 9 * Cg, BK, BN)) {
    set_error("faln_conv3x3_dgrad: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  return dispatch(BK, BN, a1, a1, wm, p, as_stream(stream));
}
