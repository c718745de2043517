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
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace faln {
namespace {

constexpr int kTH = 8;  // output tile: kTH x kTW = 8 x 16 pixels = 128 = UMMA_M

// One "tap class": the filter taps that contribute to the output pixels (r * out_mul + oh, c * out_mul + ow).
//   forward / stride-1 dgrad: one class with 9 taps; stride-2 dgrad: four parity classes with 1, 2, 2, 4 taps.
// Tap t reads the input tile at offset (dh[t], dw[t]) and the weight columns of filter tap wt[t].
struct TapClass {
  int n;
  signed char dh[16], dw[16], wt[16];   // up to 16 taps: the 4x4 gather of the upsample-folded data gradient
  signed char oh, ow;
};

struct ConvParams {
  int B, H, W, Ho, Wo;     // input dims (TMA coordinates), tile-grid dims (r, c)
  int C1, C2, Cin, Cout;   // Cin = weight-row length per tap (elements)
  int stride, act, planar; // stride = element stride of the input box
  int tiles_w, tiles_h;
  int kblocks1, kblocks2;  // BK-channel blocks of source 1 / source 2
  int ncls, nblk;         // tap classes, N blocks of BN output channels
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
  // inference epilogue of the logits layer (row kernel only): disp = sum_n d_lvl[b,n] * softmax_n(logits + bias)
  const float* disp_lvl;  // [B, Cout] disparity of each level
  float* disp_out;        // [B, 1, H, W] fp32; when set nothing else is written (the logits never reach HBM)
  int tma_out;            // row kernel: bf16 NHWC output staged in shared memory and written by TMA (tensor map tmY)
  int tile_h, tile_b;     // split-K kernel: the 128-pixel tile is tile_b images x tile_h rows x 16 columns (tile_b * tile_h = 8)
  int tiles_b;            // split-K kernel: ceil(B / tile_b)
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

// ELU with the flush-to-zero hardware exponential (MUFU.EX2 without the denormal range fix-up __expf compiles to: 5 instead
// of 10 instructions per value in the epilogues, which are issue-bound -- DESIGN.md 5)
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : (ex2f(v * 1.4426950408889634f) - 1.0f); }

template <int BK, int BN, int STAGES>
struct SmemLayout {
  static constexpr int kA = 128 * BK * 2;
  static constexpr int kB = BN * BK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kBars = 1024;  // barriers + tmem pointer
  static constexpr int kTotal = kBars + STAGES * kStage + 1024 /* alignment slack */;
};

// Epilogue of CW (32 or 16) consecutive output channels of one pixel: accumulator registers r -> bias / constant-channel
// table / accumulate / residual / activation / activation derivative -> bf16 NHWC or fp32 planar store.
// breg: the CW bias values of this thread's channel chunk held in registers (row kernel: the chunk never changes during the
// life of the CTA, so they are loaded once instead of CW scalar loads per tile -- measured: 121 -> see profiles/r2_*), or
// nullptr to read p.bias per call (tile kernel: the N block changes from work item to work item).
template <int CW>
__device__ __forceinline__ void epilogue_values(const ConvParams& p, const uint32_t (&r)[CW], const float* breg, int cg,
                                                size_t pix, int b, int ho, int wo, float (&v)[CW], bool bias_done = false) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
    if (breg) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] += breg[j];
    } else if (p.bias && !bias_done) {
      if (cg + CW <= p.Cout && (reinterpret_cast<uintptr_t>(p.bias + cg) & 15) == 0) {   // whole chunk: 128-bit loads
        const float4* bp = reinterpret_cast<const float4*>(p.bias + cg);
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          const float4 bq = __ldg(bp + q);
          v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (cg + j < p.Cout) v[j] += __ldg(p.bias + cg + j);
      }
    }
    if (p.ctab) {
      // a spatially constant extra input channel (the max_disp/100 plane of reference :145,208-209) contributes
      // value * (sum of its weights over the taps that fall inside the image): a per-border-class bias
      const int rc = (ho == 0 ? 1 : 0) | (p.stride * ho + 1 > p.H - 1 ? 2 : 0);
      const int cc = (wo == 0 ? 1 : 0) | (p.stride * wo + 1 > p.W - 1 ? 2 : 0);
      const float* t = p.ctab + (size_t)(rc * 4 + cc) * p.Cout + cg;
      const float sc = __ldg(p.cscale + b);
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (cg + j < p.Cout) v[j] = fmaf(sc, __ldg(t + j), v[j]);
    }
    if (p.accum) {
      const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.out) + pix * p.out_c + cg);
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) {
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
      for (int q = 0; q < CW / 8; ++q) {
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
      for (int j = 0; j < CW; ++j) v[j] = elu1(v[j]);
    } else if (p.act == 2) {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.dact) {
      // backward epilogue: gradient w.r.t. the pre-activation = gradient * act'(pre), from the SAVED OUTPUT y:
      // ELU'(pre) = 1 (y > 0) else y + 1;  ReLU'(pre) = [y > 0]
      const uint4* yp = reinterpret_cast<const uint4*>(p.ysave + pix * p.ysave_c + cg);
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) {
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
}

template <int CW>
__device__ __forceinline__ void store_direct(const ConvParams& p, const float (&v)[CW], int cg, size_t pix, int b, int ho,
                                             int wo) {
    if (p.planar) {
      float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (cg + j < p.Cout) o[(((size_t)b * p.Cout + cg + j) * p.out_H + ho) * p.out_pitch + wo] = v[j];
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + cg;
#pragma unroll
      for (int q = 0; q < CW / 8; ++q) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(o + q * 8) = u;
      }
    }
}

template <int CW>
__device__ __forceinline__ void epilogue_store(const ConvParams& p, const uint32_t (&r)[CW], int cg, size_t pix, int b, int ho,
                                               int wo) {
  float v[CW];
  epilogue_values<CW>(p, r, nullptr, cg, pix, b, ho, wo, v);
  store_direct<CW>(p, v, cg, pix, b, ho, wo);
}

// bf16 NHWC output staged through shared memory and written by the TMA unit (row kernel).  A thread owns pixel row m of the
// [128 px][BN ch] tile and CW consecutive channels starting at c0; the 16-byte chunks go where the tensor map's swizzle
// (128B for 128-byte rows, 64B for 64-byte rows; a function of the shared address, tile base 1024-byte aligned) expects them,
// which also makes the 32 lanes of a warp hit 32 different bank groups.
template <int BN, int CW>
__device__ __forceinline__ void stage_tile_row(unsigned char* tile, int m, int c0, const float (&v)[CW]) {
  constexpr int kRowBytes = BN * 2;
#pragma unroll
  for (int q = 0; q < CW / 8; ++q) {
    uint4 u;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
    const int chunk = c0 / 8 + q;
    const int phys = kRowBytes == 128 ? (chunk ^ (m & 7)) : (chunk ^ ((m >> 1) & 3));
    *reinterpret_cast<uint4*>(tile + (size_t)m * kRowBytes + phys * 16) = u;
  }
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Persistent, warp-specialised kernel: a CTA walks work items (tap class, N block, 128-pixel tile) with stride gridDim.x.
//   warp 0      TMA producer: runs ahead across tiles through the STAGES-deep smem ring
//   warp 1      MMA issuer: accumulates tile i into TMEM buffer i % 2 while ...
//   warps 2-9   ... the eight epilogue warps (two per TMEM lane quadrant, alternating 32-column chunks) drain buffer
//               (i - 1) % 2 (tcgen05.ld, bias / residual / activation, stores).  Round 2: eight instead of four -- on the
//               layers that give a CTA a single work item the epilogue is not overlapped with anything, and four warps
//               (one per sub-partition, latency-bound) took ~6 us for a 128 x 256 tile
// so neither the TMA latency nor the epilogue of a tile is exposed (measured on the one-tile-per-CTA version: ~6 us of
// serial latency per tile against 0.6 us of tensor work).
template <int BN>
constexpr int tc_epi_warps() { return BN >= 256 ? 16 : (BN >= 128 ? 8 : 4); }   // wide N tiles: 2 / 4 epilogue warps per TMEM lane quadrant

template <int BK, int BN, int STAGES>
__global__ void __launch_bounds__(64 + 32 * tc_epi_warps<BN>())
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using SL = SmemLayout<BK, BN, STAGES>;
  constexpr int kAcc = 2;                                   // TMEM accumulator buffers
  constexpr uint32_t kTmemCols = (kAcc * BN) < 32 ? 32 : (kAcc * BN);
  extern __shared__ unsigned char smem_raw[];
  // barriers first, then 1024B-aligned operand stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + kAcc;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + kAcc);
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + SL::kBars + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.tiles_w * p.tiles_h * p.B;
  const int total = tiles * p.nblk * p.ncls;                // work item w = (cls * tiles + tile) * nblk + nb
  const int kb = p.kblocks1 + p.kblocks2;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW);
    if (p.kblocks2) prefetch_tmap(&tmA2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], tc_epi_warps<BN>());         // one arrival per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                 // everything above overlapped the previous kernel's tail; its results are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nb = w % p.nblk, tile = (w / p.nblk) % tiles;
        const TapClass& tc = p.cls[w / (p.nblk * tiles)];
        const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
        const int n0 = nb * BN;
        for (int t = 0; t < tc.n; ++t) {
          const int tap = tc.wt[t];
          const int hi = th * kTH * p.stride + tc.dh[t], wi = tw * kTW * p.stride + tc.dw[t];
          for (int cb = 0; cb < kb; ++cb) {
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
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int ksteps = p.cls[w / (p.nblk * tiles)].n * kb;
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int k = 0; k < ksteps; ++k) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stages + (size_t)s * SL::kStage);
          const uint64_t adesc = make_desc<BK>(a_addr);
          const uint64_t bdesc = make_desc<BK>(a_addr + SL::kA);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in the (addr >> 4) field
            umma_bf16(tacc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0);
          }
          umma_commit(&empty[s]);  // frees the smem stage when these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ================================================================= epilogue (warps 2..9)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    constexpr int kChunkStep = 32 * (tc_epi_warps<BN>() / 4);
    const int half = (warp - 2) >> 2;     // which warp of the quadrant (0 when there is one): even / odd 32-column chunks
    const int m = quad * 32 + lane;       // row of the tile = pixel
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int nb = w % p.nblk, tile = (w / p.nblk) % tiles;
      const TapClass& tc = p.cls[w / (p.nblk * tiles)];
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
      const int n0 = nb * BN;
      const int ho = (th * kTH + m / kTW) * p.out_mul + tc.oh, wo = (tw * kTW + m % kTW) * p.out_mul + tc.ow;
      const bool valid = ho < p.out_H && wo < p.out_W;
      const size_t pix = ((size_t)b * p.out_H + ho) * p.out_W + wo;
      const int a = it & 1;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += kChunkStep) {
        uint32_t r[32];
        tmem_ld32(tacc + c0, r);
        const int cg = n0 + c0;  // first global output channel of this chunk
        if (!valid || cg >= p.Cout) continue;
        epilogue_store<32>(p, r, cg, pix, b, ho, wo);
      }
      // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ split-K cluster kernel
// Small feature maps (3x10 ... 12x40 at the KITTI crop) give the tile kernel at most a few dozen 128-pixel tiles with K loops
// of 36 ... 72 steps: every CTA pulls its 16 KB + BN * 128 B per step through ONE SM's L2 port (~64 B/clk measured), 14 - 24 us
// per layer where cuDNN needs 7 - 10 (per-layer table).  Here ONE work item (tap class, N block, tile) belongs to
// a thread-block CLUSTER of SPLIT CTAs:
//   * CTA `rank` runs K steps [rank * ksteps / SPLIT, (rank + 1) * ksteps / SPLIT) of the same (tap, channel block) loop
//     into its own TMEM accumulator -- SPLIT SMs' worth of L2 ports and tensor pipes per tile, wide N tiles stay affordable
//   * the partial accumulators are reduce-scattered over DISTRIBUTED SHARED MEMORY in 16-column chunks: chunk c belongs to
//     rank (c / kSub) % SPLIT; a CTA sends the chunks it does not own straight from TMEM registers into the owner's staging
//     buffer (st.shared::cluster), one cluster barrier, then every CTA finishes its own columns (bias, residual,
//     activation ... the same epilogue_values as everywhere) -- no atomics, no second kernel, fixed summation order
//   * the tile itself spans images when the map is shorter than 8 rows (tile_b x tile_h x 16 pixels, a 4-D TMA box with a
//     batch extent), so a 3x10 map fills 60 of the 128 MMA rows instead of 30.
// One work item per cluster, no persistence (these layers have fewer items than SMs); warp roles as in the tile kernel.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int BN, int SPLIT>
struct SplitCfg {
  static constexpr int kChunks = BN / 16;                              // 16-column chunks of the tile
  static constexpr int kOwn = kChunks / SPLIT;                         // chunks a CTA finishes
  static constexpr int kSub = kOwn >= 2 ? 2 : 1;                       // epilogue warps per TMEM lane quadrant
  static constexpr int kThreads = 64 + 128 * kSub;
  static constexpr int kStaging = (SPLIT - 1) * kOwn * 4 * 128 * 16;   // [sender slot][own chunk][float4 q][row] float4
  static constexpr int kStage = 128 * 64 * 2 + BN * 64 * 2;            // BK = 64
  static constexpr int kStagesFit = (218 * 1024 - kStaging) / kStage;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kTotal = 1024 + kStages * kStage + kStaging + 1024;
  static_assert(kChunks % SPLIT == 0 && kOwn % kSub == 0 && kStages >= 2, "unsupported split-K shape");
};

template <int BN, int SPLIT>
__global__ void __launch_bounds__(SplitCfg<BN, SPLIT>::kThreads)
conv3x3_splitk_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using CF = SplitCfg<BN, SPLIT>;
  constexpr int BK = 64, STAGES = CF::kStages, kSub = CF::kSub;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1024 + 1023) & ~uintptr_t(1023));
  unsigned char* staging = stages + (size_t)STAGES * CF::kStage;      // same offset in every CTA of the cluster

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int w = blockIdx.x / SPLIT;                                   // work item = (cls * tiles + tile) * nblk + nb
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int kb = p.kblocks1 + p.kblocks2;
  const int nb = w % p.nblk, tile = (w / p.nblk) % tiles;
  const TapClass& tc = p.cls[w / (p.nblk * tiles)];
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tb = tile / (p.tiles_w * p.tiles_h);
  const int n0 = nb * BN;
  const int ksteps = tc.n * kb;
  const int k_lo = rank * ksteps / SPLIT, k_hi = (rank + 1) * ksteps / SPLIT;   // the host guarantees ksteps >= SPLIT

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
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  cluster_arrive();   // #1: "this CTA runs, its shared memory exists" -- waited for before the first remote store

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int k = k_lo; k < k_hi; ++k) {
        const int t = k / kb, cb = k - t * kb;
        const int tap = tc.wt[t];
        const int hi = th * p.tile_h * p.stride + tc.dh[t], wi = tw * kTW * p.stride + tc.dw[t];
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* a_dst = stages + (size_t)s * CF::kStage;
        unsigned char* b_dst = a_dst + 128 * BK * 2;
        mbar_arrive_expect_tx(&full[s], CF::kStage);
        if (cb < p.kblocks1) {
          tma_load_4d(a_dst, &tmA1, cb * BK, wi, hi, tb * p.tile_b, &full[s]);
          tma_load_2d(b_dst, &tmW, tap * p.Cin + cb * BK, n0, &full[s]);
        } else {
          tma_load_4d(a_dst, &tmA2, (cb - p.kblocks1) * BK, wi, hi, tb * p.tile_b, &full[s]);
          tma_load_2d(b_dst, &tmW, tap * p.Cin + p.C1 + (cb - p.kblocks1) * BK, n0, &full[s]);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int s = 0;
      uint32_t ph = 0;
      for (int k = k_lo; k < k_hi; ++k) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(stages + (size_t)s * CF::kStage);
        const uint64_t adesc = make_desc<BK>(a_addr);
        const uint64_t bdesc = make_desc<BK>(a_addr + 128 * BK * 2);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k != k_lo) || kk != 0);
        umma_commit(&empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_full);
    }
  }
  __syncwarp();
  cluster_wait();     // #1

  // ---- epilogue warps 2 .. 2 + 4 * kSub: quadrant = warp & 3 (the TMEM lanes a warp may read), sub = which of the quadrant's warps
  const int quad = warp & 3, sub = (warp - 2) >> 2;
  const int m = quad * 32 + lane;                                     // tile row = pixel (bb, hh, ww)
  const int per_img = p.tile_h * kTW;
  const int bb = m / per_img, hh = (m / kTW) % p.tile_h, ww = m % kTW;
  const int b = tb * p.tile_b + bb;
  const int ho = (th * p.tile_h + hh) * p.out_mul + tc.oh, wo = (tw * kTW + ww) * p.out_mul + tc.ow;
  const bool valid = b < p.B && ho < p.out_H && wo < p.out_W;
  const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t stg = smem_u32(staging);
  if (warp >= 2) {
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = sub; c < CF::kChunks; c += kSub) {                   // send the chunks other ranks own
      const int owner = (c / kSub) % SPLIT;
      if (owner == rank) continue;
      uint32_t r[16];
      tmem_ld16(tacc + c * 16, r);
      if (!valid) continue;
      const int slot = (rank - owner + SPLIT) % SPLIT - 1;
      const int lc = (c / (kSub * SPLIT)) * kSub + sub;                // index among the owner's chunks
      const uint32_t dst = mapa_u32(stg + (uint32_t)(((slot * CF::kOwn + lc) * 4) * 128 + m) * 16, (uint32_t)owner);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) st_cluster_v4(dst + qd * 128 * 16, r[4 * qd], r[4 * qd + 1], r[4 * qd + 2], r[4 * qd + 3]);
    }
  }
  __syncwarp();
  cluster_arrive();   // #2: release the partial sums ...
  cluster_wait();     //     ... and acquire the ones sent to this CTA
  if (warp >= 2) {
    const size_t pix = ((size_t)b * p.out_H + ho) * p.out_W + wo;
#pragma unroll 1
    for (int c = sub; c < CF::kChunks; c += kSub) {
      if ((c / kSub) % SPLIT != rank) continue;
      uint32_t r[16];
      tmem_ld16(tacc + c * 16, r);
      const int cg = n0 + c * 16;
      if (!valid || cg >= p.Cout) continue;
      const int lc = (c / (kSub * SPLIT)) * kSub + sub;
#pragma unroll
      for (int slot = 0; slot < SPLIT - 1; ++slot) {
        const float4* src = reinterpret_cast<const float4*>(staging) + ((slot * CF::kOwn + lc) * 4) * 128 + m;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const float4 a = src[qd * 128];
          r[4 * qd] = __float_as_uint(__uint_as_float(r[4 * qd]) + a.x);
          r[4 * qd + 1] = __float_as_uint(__uint_as_float(r[4 * qd + 1]) + a.y);
          r[4 * qd + 2] = __float_as_uint(__uint_as_float(r[4 * qd + 2]) + a.z);
          r[4 * qd + 3] = __float_as_uint(__uint_as_float(r[4 * qd + 3]) + a.w);
        }
      }
      epilogue_store<16>(p, r, cg, pix, b, ho, wo);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ row-tile kernel
// Wide, few-channel layers (full and half resolution) are bound by L2 -> SM traffic in the tile kernel above: every 128-pixel
// tile re-fetches its input patch once per filter tap (9 x 16 KB) plus the weights (9 x 8 KB), ~8.9 TB/s measured against
// the ~12 TB/s L2 cap.  Here a tile is ONE output row x 128 pixels, so that
//   * the input arrives as ONE halo box per channel block: [BK, 130, 3] (rows h-1..h+1, columns w0-1..w0+128); the A
//     operand of tap (dh, dw) is the window starting at halo row (dh+1)*130 + (dw+1) -- 128 consecutive shared-memory rows,
//     a legal K-major operand with the usual SBO (the 128B/64B swizzle is a function of the shared address, so a window
//     may start on any row)
//   * the weights of all nine taps stay resident in shared memory for the life of the (persistent) CTA
//   * the output row segment is contiguous in NHWC memory.
// Per 128 pixels the SM now pulls 50 KB (3x the algorithmic input, the overlap served by L2) instead of 216 KB.
// Eight epilogue warps (two per TMEM lane quadrant, half of the columns each) drain the double-buffered accumulator.
constexpr int kRowW = 128, kHaloCols = kRowW + 2, kHaloRows = 3;

template <int BK, int BN, int STAGES>
struct RowSmem {
  static constexpr int kHalo = (kHaloCols * kHaloRows * BK * 2 + 1023) / 1024 * 1024;
  static constexpr int kW = BN * BK * 2;                       // one (tap, channel block) weight tile
  static constexpr int kBars = 4096;                           // barriers (first 1 KB) + disparity-epilogue exchange (3 KB)
  static constexpr int kOut = 128 * BN * 2;                    // one staged output tile [128 px][BN ch] bf16 (x2: double buffer)
  static int total(int kb) { return kBars + 1024 + 9 * kb * kW + STAGES * kHalo + 2 * kOut; }
};

template <int BK, int BN, int STAGES>
__global__ void __launch_bounds__(320)
conv3x3_row_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY, const ConvParams p) {
  using SL = RowSmem<BK, BN, STAGES>;
  constexpr int kAcc = 2;
  constexpr uint32_t kTmemCols = (kAcc * BN) < 32 ? 32 : (kAcc * BN);
  constexpr int CW = BN / 2;                                   // columns per epilogue warp
  extern __shared__ unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + kAcc;
  uint64_t* wfull = acc_empty + kAcc;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);
  float* exch = reinterpret_cast<float*>(smem_raw + 1024);    // [2][128][3]: partial softmax sums of the second column half
  unsigned char* wsm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + SL::kBars + 1023) & ~uintptr_t(1023));
  const int kb = p.kblocks1 + p.kblocks2;
  unsigned char* stages = wsm + (size_t)9 * kb * SL::kW;      // kW is a multiple of 1024 (BN * BK * 2 >= 2048)
  unsigned char* otile = stages + (size_t)STAGES * SL::kHalo; // 2 x kOut, 1024-byte aligned (kHalo is a multiple of 1024)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.tiles_w * p.H * p.B;                    // work item = (b, h, 128-pixel segment)
  const TapClass& tc = p.cls[0];

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW);
    if (p.kblocks2) prefetch_tmap(&tmA2);
    if (p.tma_out) prefetch_tmap(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 8);                            // one arrival per epilogue warp
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                 // everything above overlapped the previous kernel's tail; its results are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, 9 * kb * SL::kW);
      for (int t = 0; t < 9; ++t)
        for (int cb = 0; cb < kb; ++cb) {
          const int col = tc.wt[t] * p.Cin + (cb < p.kblocks1 ? cb * BK : p.C1 + (cb - p.kblocks1) * BK);
          tma_load_2d(wsm + (size_t)(t * kb + cb) * SL::kW, &tmW, col, 0, wfull);
        }
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int tw = w % p.tiles_w, h = (w / p.tiles_w) % p.H, b = w / (p.tiles_w * p.H);
        for (int cb = 0; cb < kb; ++cb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], kHaloCols * kHaloRows * BK * 2);
          unsigned char* dst = stages + (size_t)s * SL::kHalo;
          if (cb < p.kblocks1) tma_load_4d(dst, &tmA1, cb * BK, tw * kRowW - 1, h - 1, b, &full[s]);
          else tma_load_4d(dst, &tmA2, (cb - p.kblocks1) * BK, tw * kRowW - 1, h - 1, b, &full[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      uint32_t aoff[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) aoff[t] = (uint32_t)(((tc.dh[t] + 1) * kHaloCols + (tc.dw[t] + 1)) * BK * 2) >> 4;
      const uint64_t wdesc0 = make_desc<BK>(smem_u32(wsm));
      mbar_wait(wfull, 0);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int cb = 0; cb < kb; ++cb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t hdesc = make_desc<BK>(smem_u32(stages + (size_t)s * SL::kHalo));
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint64_t adesc = hdesc + aoff[t];
            const uint64_t bdesc = wdesc0 + (uint64_t)(((t * kb + cb) * SL::kW) >> 4);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(tacc, adesc + 2 * kk, bdesc + 2 * kk, idesc, (cb | t | kk) != 0);
          }
          umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ================================================================= epilogue (warps 2 .. EW + 1)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                           // column chunk 0 .. EW / 4 - 1 of the quadrant
    const int m = quad * 32 + lane;
    const bool issuer = warp == 2 && lane == 0;                 // issues the TMA stores of the staged output tiles
    // this thread's CW output channels never change: keep their bias in registers for the life of the CTA
    float breg[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) breg[j] = (p.bias && p.disp_out == nullptr && half * CW + j < p.Cout) ? __ldg(p.bias + half * CW + j) : 0.f;
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int tw = w % p.tiles_w, h = (w / p.tiles_w) % p.H, b = w / (p.tiles_w * p.H);
      const int wo = tw * kRowW + m;
      const bool valid = wo < p.W;
      const size_t pix = ((size_t)b * p.H + h) * p.W + wo;
      const int a = it & 1;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
      if (BN == 64 && p.disp_out != nullptr) {
        // fused softmax-expectation (reference :216-229).  The two warps of a lane quadrant each reduce 32 planes
        // (bias padded with -inf and d_lvl with 0 beyond N, so no per-plane predicates), exchange their partial
        // (max, sum, weighted sum) through shared memory and the first one writes the disparity.
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t)(a * BN + half * 32) + ((uint32_t)(quad * 32) << 16), r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[a]);
        const float4* bp = reinterpret_cast<const float4*>(p.bias) + half * 8;
        const float4* dp = reinterpret_cast<const float4*>(p.disp_lvl + (size_t)b * 64) + half * 8;
        float v[32];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = __ldg(bp + q);
          v[4 * q] = __uint_as_float(r[4 * q]) + bq.x;
          v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
          v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
          v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
          mx = fmaxf(fmaxf(fmaxf(mx, v[4 * q]), fmaxf(v[4 * q + 1], v[4 * q + 2])), v[4 * q + 3]);
        }
        const float nm = -mx * 1.4426950408889634f;
        float z = 0.f, acc = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 dq = __ldg(dp + q);
          const float e0 = ex2f(fmaf(v[4 * q], 1.4426950408889634f, nm)), e1 = ex2f(fmaf(v[4 * q + 1], 1.4426950408889634f, nm));
          const float e2 = ex2f(fmaf(v[4 * q + 2], 1.4426950408889634f, nm)), e3 = ex2f(fmaf(v[4 * q + 3], 1.4426950408889634f, nm));
          z += (e0 + e1) + (e2 + e3);
          acc = fmaf(dq.x, e0, fmaf(dq.y, e1, fmaf(dq.z, e2, fmaf(dq.w, e3, acc))));
        }
        float* ex = exch + ((it & 1) * 128 + m) * 3;
        if (half == 1) {
          ex[0] = mx; ex[1] = z; ex[2] = acc;
        }
        named_bar_sync(2 + quad, 64);
        if (half == 0 && valid) {
          const float m1 = ex[0], z1 = ex[1], a1 = ex[2];
          const float mm = fmaxf(mx, m1);                     // half 0 always holds real planes: mm is finite
          const float f0 = ex2f((mx - mm) * 1.4426950408889634f), f1 = ex2f((m1 - mm) * 1.4426950408889634f);
          p.disp_out[pix] = (acc * f0 + a1 * f1) / (z * f0 + z1 * f1);
        }
        continue;
      }
      uint32_t r[CW];
      const uint32_t taddr = tmem_base + (uint32_t)(a * BN + half * CW) + ((uint32_t)(quad * 32) << 16);
      if (CW == 32) tmem_ld32(taddr, reinterpret_cast<uint32_t(&)[32]>(r));
      else tmem_ld16(taddr, reinterpret_cast<uint32_t(&)[16]>(r));
      // registers hold the data: the accumulator buffer can be refilled while this warp does the arithmetic and stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
      const int cg = half * CW;
      if (p.tma_out) {
        // every thread owns a row of the staged tile (rows beyond the image are clipped by the tensor map on the way out)
        float v[CW];
        if (valid) epilogue_values<CW>(p, r, breg, cg, pix, b, h, wo, v);
        else {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = 0.f;
        }
        unsigned char* tile = otile + (size_t)(it & 1) * SL::kOut;
        stage_tile_row<BN, CW>(tile, m, cg, v);
        fence_proxy_async();                                  // generic-proxy writes -> visible to the TMA (async proxy)
        if (issuer) tma_store_wait_read0();                   // the store issued one tile ago has finished reading smem
        named_bar_sync(1, 256);                               // (so the OTHER buffer is free for the next tile's writes)
        if (issuer) tma_store_3d(&tmY, tile, 0, tw * kRowW, b * p.H + h);
      } else if (valid && cg < p.Cout) {
        float v[CW];
        epilogue_values<CW>(p, r, breg, cg, pix, b, h, wo, v);
        store_direct<CW>(p, v, cg, pix, b, h, wo);
      }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ column-walk kernel
// The row kernel above still reads the A operand once per tap: 9 x (4 KB A + CO * 32 B of B) per 128-pixel K16 step, 192 B per
// tensor clock at CO = 64 against the SM's 128 B/clk shared-memory port ("what bounds the convolutions").  Here a
// CTA walks DOWN a 128-pixel column strip and turns the three vertical taps into extra N columns instead of extra A reads:
// input row r contributes to the output rows r, r + 1 and r - 1 (vertical tap offsets 0, -1, +1), so ONE instruction with
// N = 3 * CO multiplies the row's window by the three taps' weights [B0 | B1 | B2] and accumulates into three TMEM
// accumulators that live in slots (output row mod 3) of a 3 * CO-column bank; the horizontal taps stay shifted windows of
// the row.  Per K16 step and horizontal tap the port now moves 4 KB of A + 3 * CO * 32 B of B for 3x the MACs: 104 B per
// tensor clock at CO = 64 (tensor-bound), and every input row is fetched ONCE (a [BK, 130, 1] box) instead of three times.
//   * slot order: the three slots hold (B0,B1,B2), (B2,B0,B1) or (B1,B2,B0) for r mod 3 = 0, 1, 2; blocks that are adjacent
//     both in the weight tile and in TMEM go out as one instruction (N = 3 CO, or 2 CO + CO)
//   * the block of the NEW output row r + 1 starts its accumulator (accumulate flag off at the row's first K step, which is
//     therefore issued block by block); output row q is complete after input row q + 1 and is drained by the epilogue warps
//     while the tensor pipe works on the CTA's SECOND strip: two banks (2 x 3 x CO TMEM columns), row-steps alternate
//   * work = the flattened (image, column strip, row) sequence cut into 2 x gridDim.x chains; a chain restarts (one extra,
//     one-block row-step above and below) where it crosses into another strip; rows -1 and H are zero boxes (TMA bounds).
// Epilogue arithmetic, weight residency and the TMA-staged output are the row kernel's (the store is issued by a dedicated
// warp here).  fwd and stride-1 dgrad differ only in TapClass.
// STATUS (round 2, measured on B200; opt-in with FALN_CONV_COL=1, off by default): parity-green on every row-kernel test
// shape plus six of its own (tests/test_conv_gpu.py), but not yet faster than the row kernel: 64 -> 64 at 8x192x640 94 us vs
// 81 us, 32 -> 32 47 vs 38 us, 96 -> 64 138 vs 126 us.  clock64() timelines of one CTA (recorded while developing it) show
// where the time goes: the tensor side needs ~1,600 clk per row-step (12 - 24 instructions; the row kernel needs ~2,180 clk
// for its 36) but the eight epilogue warps need ~2,350 clk per drained row (bias + ELU ~1,190 clk -- MUFU-bound at 512 --,
// staging ~380, TMEM read / zero / barriers ~450) and with ONE CTA per SM (384 TMEM columns) nothing else overlaps them, whereas
// the row kernel's epilogue (~1,800 clk per tile) hides under its longer MMA phase.  What would make it pay: sixteen epilogue
// warps (16 columns each) and a packed-half exponential, or two CTAs per SM with the weights shared through a cluster.
// Lessons already folded in: a single copy of each role's loop body (two unrolled copies overflowed the instruction cache: 30 % of
// the stall samples were no_inst), 32-bit descriptor arithmetic and no per-instruction predicates in the issuing thread (the first
// version spent ~1,100 instructions per row-step there), TMEM zeroing by the epilogue warps instead of accumulate-flag logic.
struct ColWalker {
  int t, t1, H, tiles_w;
  int b, tw, h0, h1, r;
  bool in_run;
  __device__ __forceinline__ void init(long long t0_, long long t1_, int H_, int tiles_w_) {
    t = (int)t0_; t1 = (int)t1_; H = H_; tiles_w = tiles_w_; in_run = false;
    b = tw = h0 = h1 = r = 0;
  }
  // next row-step (input row r of the run of output rows [h0, h1] in strip (b, tw)); false when the chain is exhausted
  __device__ __forceinline__ bool next() {
    if (in_run && r <= h1) { ++r; return true; }
    if (in_run) { t += h1 - h0 + 1; in_run = false; }
    if (t >= t1) return false;
    h0 = t % H;
    const int col = t / H;
    tw = col % tiles_w; b = col / tiles_w;
    h1 = h0 + (t1 - t) - 1;
    if (h1 > H - 1) h1 = H - 1;
    r = h0 - 1;
    in_run = true;
    return true;
  }
};

template <int BK, int CO>
struct ColSmem {
  static constexpr int kRow = (kHaloCols * BK * 2 + 1023) / 1024 * 1024;   // one input row [130 px][BK] (swizzled rows)
  static constexpr int kW = CO * BK * 2;                                  // one (tap, channel block) weight tile
  static constexpr int kBars = 1024;
  static constexpr int kOut = 128 * CO * 2;
  static int total(int kb, int stages) { return kBars + 1024 + 9 * kb * kW + stages * kRow + 2 * kOut; }
};

template <int BK, int CO, int EW>
__global__ void __launch_bounds__(96 + 32 * EW)
conv3x3_col_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY, const ConvParams p,
                   const int nstages) {
  using SL = ColSmem<BK, CO>;
  constexpr int kMaxStages = 12;
  constexpr uint32_t kTmemCols = 6 * CO <= 256 ? 256 : 512;
  constexpr int kBank = kTmemCols / 2;                          // TMEM columns between the two banks (3 * CO used)
  constexpr int CW = CO / (EW / 4);                             // columns per epilogue warp: EW / 4 warps share a lane quadrant
  static_assert(CW == 32 || CW == 16, "epilogue chunk");
  extern __shared__ unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* wfull = acc_empty + 2;
  uint64_t* out_full = wfull + 1;                              // staged output tile complete (EW epilogue warps)
  uint64_t* out_empty = out_full + 2;                          // ... and read out by the TMA store (store warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem_raw + 512);     // [CO] bias (zeros when the layer has none)
  unsigned char* wsm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + SL::kBars + 1023) & ~uintptr_t(1023));
  const int kb = p.kblocks1 + p.kblocks2;
  unsigned char* stages = wsm + (size_t)9 * kb * SL::kW;
  unsigned char* otile = stages + (size_t)nstages * SL::kRow;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = (long long)p.tiles_w * p.H * p.B;
  const TapClass& tc = p.cls[0];
  // the CTA's two chains; every role walks them alternately with ONE copy of its loop body (`cur` / `oth` swap each turn:
  // two unrolled copies of the epilogue overflowed the instruction cache -- 30 % of all stall samples were no_inst)
  ColWalker cur, oth;
  {
    const long long c = 2LL * blockIdx.x, nch = 2LL * gridDim.x;
    cur.init(c * T / nch, (c + 1) * T / nch, p.H, p.tiles_w);
    oth.init((c + 1) * T / nch, (c + 2) * T / nch, p.H, p.tiles_w);
  }
  bool alive_cur = cur.next(), alive_oth = oth.next();
  int be = 0;                                                   // bank of `cur`
  auto swap_chains = [&]() {
    const ColWalker tmp = cur; cur = oth; oth = tmp;
    const bool ta = alive_cur; alive_cur = alive_oth; alive_oth = ta;
    be ^= 1;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW);
    if (p.kblocks2) prefetch_tmap(&tmA2);
    if (p.tma_out) prefetch_tmap(&tmY);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], EW);
    }
    mbar_init(wfull, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&out_full[a], EW);
      mbar_init(&out_empty[a], 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + CO) sbias[threadIdx.x - 64] = (p.bias && (int)threadIdx.x - 64 < p.Cout) ? __ldg(p.bias + threadIdx.x - 64) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, 9 * kb * SL::kW);
      for (int t = 0; t < 9; ++t) {
        const int dwi = tc.dw[t] + 1, j = tc.dh[t] == 0 ? 0 : (tc.dh[t] < 0 ? 1 : 2);
        for (int cb = 0; cb < kb; ++cb) {
          const int col = tc.wt[t] * p.Cin + (cb < p.kblocks1 ? cb * BK : p.C1 + (cb - p.kblocks1) * BK);
          tma_load_2d(wsm + (size_t)((cb * 3 + dwi) * 3 + j) * SL::kW, &tmW, col, 0, wfull);
        }
      }
      int s = 0;
      uint32_t ph = 0;
      while (alive_cur || alive_oth) {
        if (alive_cur) {
          const ColWalker& k = cur;
          for (int cb = 0; cb < kb; ++cb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&full[s], kHaloCols * BK * 2);
            unsigned char* dst = stages + (size_t)s * SL::kRow;
            if (cb < p.kblocks1) tma_load_4d(dst, &tmA1, cb * BK, k.tw * kRowW - 1, k.r, k.b, &full[s]);
            else tma_load_4d(dst, &tmA2, (cb - p.kblocks1) * BK, k.tw * kRowW - 1, k.r, k.b, &full[s]);
            if (++s == nstages) { s = 0; ph ^= 1; }
          }
          alive_cur = cur.next();
        }
        swap_chains();
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      // Every row-step issues ALL three blocks with the accumulate flag on -- the accumulators are zeroed by the epilogue
      // warps (tcgen05.st after a slot is drained; all three slots after a run's last drain and at kernel start), so rows
      // outside the run only ever add into slots that are re-zeroed before they carry a real row.  That keeps this loop
      // free of per-instruction predicates: the issuing thread was the bottleneck of the first version of this kernel
      // (~1100 instructions per row-step for 36 MMAs; ncu: it never waited for data or for the epilogue).
      // Block j sits in slot (rho + j) % 3 for j = 0, 1 and (rho + 2) % 3 for j = 2, rho = r mod 3:
      //   rho 0: [B0 B1 B2] -> slots 0,1,2 (one instruction);  rho 1: [B0 B1] -> 1,2 and [B2] -> 0;  rho 2: [B1 B2] -> 0,1 and [B0] -> 2
      constexpr uint32_t idesc1 = make_idesc<CO>(), idesc2 = make_idesc<2 * CO>(), idesc3 = make_idesc<3 * CO>();
      constexpr uint32_t kDescHi = (uint32_t)((((uint64_t)(8 * BK * 2) >> 4) << 32 | (1ULL << 46) | ((uint64_t)(BK == 64 ? 2 : 4) << 61)) >> 32);
      const uint32_t w_lo = ((smem_u32(wsm) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t st_lo = ((smem_u32(stages) & 0x3FFFF) >> 4) | (1u << 16);
      mbar_wait(wfull, 0);
      tc_fence_after();
      int s = 0;
      uint32_t ph = 0;
      int nd_cur = 1, nd_oth = 1;                               // drains signalled per bank (+1: the initial zeroing)
      while (alive_cur || alive_oth) {
        if (alive_cur) {
          const ColWalker& k = cur;
          const int rho = (k.r + 3) % 3;                          // r >= -1
          // segment A (always) and segment B (rho != 0): TMEM column offset, weight-block offset (>> 4), instruction descriptor
          const uint32_t bank = tmem_base + (uint32_t)(be * kBank);
          const uint32_t dA = bank + (uint32_t)((rho == 2 ? 0 : rho) * CO), dB = bank + (uint32_t)((rho == 1 ? 0 : 2) * CO);
          const uint32_t bA = rho == 2 ? (uint32_t)(SL::kW >> 4) : 0u, bB = rho == 1 ? (uint32_t)((2 * SL::kW) >> 4) : 0u;
          const uint32_t iA = rho == 0 ? idesc3 : idesc2;
          const bool two = rho != 0;
          const bool edge = k.r <= k.h0;
          mbar_wait(&acc_empty[be], (uint32_t)((nd_cur & 1) ^ 1));   // the slot the new output row takes is drained and zeroed
          tc_fence_after();
          for (int cb = 0; cb < kb; ++cb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a_lo = st_lo + (uint32_t)(s * (SL::kRow >> 4));
            const uint32_t b_lo = w_lo + (uint32_t)(cb * ((9 * SL::kW) >> 4));
            if (!edge) {
#pragma unroll
              for (int dwi = 0; dwi < 3; ++dwi) {
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                  const uint32_t al = a_lo + (uint32_t)(((dwi * BK * 2) >> 4) + 2 * kk);
                  const uint32_t bl = b_lo + (uint32_t)(((dwi * 3 * SL::kW) >> 4) + 2 * kk);
                  umma_bf16_lo(dA, al, bl + bA, kDescHi, iA);
                  if (two) umma_bf16_lo(dB, al, bl + bB, kDescHi, idesc1);
                }
              }
            } else {
              // the first two row-steps of a run (r = h0 - 1, h0): only the blocks of rows inside the run, one by one -- the
              // slots of rows h0 + 1 and h0 + 2 must still be zero when their first contribution arrives
#pragma unroll
              for (int dwi = 0; dwi < 3; ++dwi) {
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                  const uint32_t al = a_lo + (uint32_t)(((dwi * BK * 2) >> 4) + 2 * kk);
                  const uint32_t bl = b_lo + (uint32_t)(((dwi * 3 * SL::kW) >> 4) + 2 * kk);
                  if (k.r == k.h0) umma_bf16_lo(bank + (uint32_t)(rho * CO), al, bl, kDescHi, idesc1);
                  umma_bf16_lo(bank + (uint32_t)(((rho + 1) % 3) * CO), al, bl + (uint32_t)(SL::kW >> 4), kDescHi, idesc1);
                }
              }
            }
            umma_commit(&empty[s]);
            if (++s == nstages) { s = 0; ph ^= 1; }
          }
          if (k.r >= k.h0 + 1) {                                  // output row r - 1 is complete
            umma_commit(&acc_full[be]);
            ++nd_cur;
          }
          alive_cur = cur.next();
        }
        swap_chains();
        { const int tn = nd_cur; nd_cur = nd_oth; nd_oth = tn; }
      }
    }
  } else if (warp == 2 + EW) {
    // ================================================================= TMA store warp (staged bf16 NHWC output)
    if (lane == 0 && p.tma_out) {
      int it = 0;
      while (alive_cur || alive_oth) {
        if (alive_cur) {
          const ColWalker& k = cur;
          if (k.r >= k.h0 + 1) {
            mbar_wait(&out_full[it & 1], (uint32_t)((it >> 1) & 1));
            tma_store_3d(&tmY, otile + (size_t)(it & 1) * SL::kOut, 0, k.tw * kRowW, k.b * p.H + k.r - 1);
            if (it > 0) {
              tma_store_wait_read1();
              mbar_arrive(&out_empty[(it - 1) & 1]);
            }
            ++it;
          }
          alive_cur = cur.next();
        }
        swap_chains();
      }
      tma_store_wait_all();
    }
  } else {
    // ================================================================= epilogue (warps 2 .. EW + 1)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                           // column chunk 0 .. EW / 4 - 1 of the quadrant
    const int m = quad * 32 + lane;
    // zero this warp's part (its 32 TMEM lanes x CW columns) of all six accumulator slots, then open both banks
    const uint32_t tpart = tmem_base + (uint32_t)(half * CW) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
    for (int sl = 0; sl < 6; ++sl) tmem_zero<CW>(tpart + (uint32_t)((sl / 3) * kBank + (sl % 3) * CO));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&acc_empty[0]);
      mbar_arrive(&acc_empty[1]);
    }
    int nd_cur = 0, nd_oth = 0;
    int it = 0;
    while (alive_cur || alive_oth) {
      if (alive_cur) {
        const ColWalker& k = cur;
        if (k.r >= k.h0 + 1) {
          const int h = k.r - 1, tw = k.tw, b = k.b;
          const int wo = tw * kRowW + m;
          const bool valid = wo < p.W;
          const size_t pix = ((size_t)b * p.H + h) * p.W + wo;
          mbar_wait(&acc_full[be], (uint32_t)(nd_cur & 1));
          ++nd_cur;
          tc_fence_after();
          uint32_t r[CW];
          const uint32_t taddr = tmem_base + (uint32_t)(be * kBank + (h % 3) * CO + half * CW) + ((uint32_t)(quad * 32) << 16);
          if (CW == 32) tmem_ld32(taddr, reinterpret_cast<uint32_t(&)[32]>(r));
          else tmem_ld16(taddr, reinterpret_cast<uint32_t(&)[16]>(r));
          // the drained slot carries output row r + 2 next: zero it; after the run's last row zero the whole bank (rows
          // outside the run have added into the other two slots)
          if (k.r == k.h1 + 1) {
#pragma unroll
            for (int sl = 0; sl < 3; ++sl) tmem_zero<CW>(tpart + (uint32_t)(be * kBank + sl * CO));
          } else {
            tmem_zero<CW>(taddr);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[be]);
          const int cg = half * CW;
          {
            const float4* sb = reinterpret_cast<const float4*>(sbias + cg);   // warp-uniform address: one broadcast per load
#pragma unroll
            for (int qd = 0; qd < CW / 4; ++qd) {
              const float4 bq = sb[qd];
              r[4 * qd] = __float_as_uint(__uint_as_float(r[4 * qd]) + bq.x);
              r[4 * qd + 1] = __float_as_uint(__uint_as_float(r[4 * qd + 1]) + bq.y);
              r[4 * qd + 2] = __float_as_uint(__uint_as_float(r[4 * qd + 2]) + bq.z);
              r[4 * qd + 3] = __float_as_uint(__uint_as_float(r[4 * qd + 3]) + bq.w);
            }
          }
          if (p.tma_out) {
            float v[CW];
            if (valid) epilogue_values<CW>(p, r, nullptr, cg, pix, b, h, wo, v, true);
            else {
#pragma unroll
              for (int j = 0; j < CW; ++j) v[j] = 0.f;
            }
            unsigned char* tile = otile + (size_t)(it & 1) * SL::kOut;
            mbar_wait(&out_empty[it & 1], (uint32_t)(((it >> 1) & 1) ^ 1));   // the store of two tiles ago has read this buffer
            stage_tile_row<CO, CW>(tile, m, cg, v);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&out_full[it & 1]);
          } else if (valid && cg < p.Cout) {
            float v[CW];
            epilogue_values<CW>(p, r, nullptr, cg, pix, b, h, wo, v, true);
            store_direct<CW>(p, v, cg, pix, b, h, wo);
          }
          ++it;
        }
        alive_cur = cur.next();
      }
      swap_chains();
      { const int tn = nd_cur; nd_cur = nd_oth; nd_oth = tn; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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
int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, ConvParams p, cudaStream_t st) {
  using SL = SmemLayout<BK, BN, STAGES>;
  auto kern = conv3x3_tc_kernel<BK, BN, STAGES>;
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::kTotal);
    // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that contains tcgen05.alloc (it cannot know
    // the column count), so the co-residency is computed here: shared memory (228 KB per SM, 1 KB reserved per CTA),
    // TMEM columns (512 per SM) and registers (64 K per SM, __launch_bounds__(320)).
    int n = (228 * 1024) / (SL::kTotal + 1024);
    const int tmem_cols = 2 * BN < 32 ? 32 : 2 * BN;
    if (n > 512 / tmem_cols) n = 512 / tmem_cols;          // co-resident CTAs must all get their TMEM columns
    if (n > 6) n = 6;
    if (tc_epi_warps<BN>() == 8 && n > 2) n = 2;            // 320 threads x ~72 registers: two CTAs per SM fit the register file
    if (tc_epi_warps<BN>() == 16 && n > 1) n = 1;           // 576 threads
    if (getenv("FALN_DEBUG")) fprintf(stderr, "conv3x3_tc_kernel<%d,%d,%d>: smem %d, %d CTAs/SM\n", BK, BN, STAGES, SL::kTotal, n);
    per_sm = n < 1 ? 1 : n;
  }
  p.nblk = (p.Cout + BN - 1) / BN;
  const long long total = (long long)p.tiles_w * p.tiles_h * p.B * p.nblk * p.ncls;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > total) grid = total;
  launch_pdl(kern, dim3((unsigned)grid), dim3(64 + 32 * tc_epi_warps<BN>()), (size_t)SL::kTotal, st, a1, a2, w, p);
  return after_launch("conv3x3_tc_kernel");
}

bool make_row_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int BK, int rows = kHaloRows) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)kHaloCols, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 NHWC output [B*H rows][W][out_c] with a [1][128][BN] store box (rows = 128 consecutive pixels of one image row)
bool make_out_map(CUtensorMap* m, const void* ptr, long long rows, int W, int out_c, int BN) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)out_c, (cuuint64_t)W, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)out_c * 2, (cuuint64_t)W * out_c * 2};
  cuuint32_t box[3] = {(cuuint32_t)BN, (cuuint32_t)kRowW, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, BN == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BK, int BN, int STAGES>
int launch_row(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const CUtensorMap& ym, ConvParams p,
               cudaStream_t st) {
  using SL = RowSmem<BK, BN, STAGES>;
  auto kern = conv3x3_row_kernel<BK, BN, STAGES>;
  const int smem = SL::total(p.kblocks1 + p.kblocks2);
  static int attr = 0;
  if (attr < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = smem;
  }
  int per_sm = (228 * 1024) / (smem + 1024);      // see launch(): the occupancy API answers 1 for tcgen05.alloc kernels
  const int tmem_cols = 2 * BN < 32 ? 32 : 2 * BN;
  if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  p.tiles_w = (p.W + kRowW - 1) / kRowW;
  p.nblk = 1;
  const long long total = (long long)p.tiles_w * p.H * p.B;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > total) grid = total;
  launch_pdl(kern, dim3((unsigned)grid), dim3(320), (size_t)smem, st, a1, a2, w, ym, p);
  return after_launch("conv3x3_row_kernel");
}

template <int BK, int CO, int EW>
int launch_col(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const CUtensorMap& ym, ConvParams p,
               cudaStream_t st) {
  using SL = ColSmem<BK, CO>;
  auto kern = conv3x3_col_kernel<BK, CO, EW>;
  const int kb = p.kblocks1 + p.kblocks2;
  // two CTAs per SM when the 6 * CO accumulator columns leave room (CO = 32) and two copies of the weights + rings fit
  int per_sm = 6 * CO <= 256 ? 2 : 1;
  int budget = (228 * 1024) / per_sm - 1024;
  int stages = (budget - SL::total(kb, 0)) / SL::kRow;
  if (per_sm == 2 && stages < 3) {
    per_sm = 1;
    budget = 227 * 1024;
    stages = (budget - SL::total(kb, 0)) / SL::kRow;
  }
  if (stages > 8) stages = 8;
  if (stages < 2) return 1;                                     // does not fit: caller falls back to the row kernel
  const int smem = SL::total(kb, stages);
  static int attr = 0;
  if (attr < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = smem;
  }
  p.tiles_w = (p.W + kRowW - 1) / kRowW;
  p.nblk = 1;
  const long long total = (long long)p.tiles_w * p.H * p.B;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > total / 12) grid = total / 12;                      // at least ~12 output rows per chain (restart overhead)
  if (grid < 1) grid = 1;
  if (getenv("FALN_DEBUG")) fprintf(stderr, "conv3x3_col_kernel<%d,%d,%d>: smem %d, %d stages, %d CTAs/SM, grid %lld\n", BK, CO, EW, smem, stages, per_sm, grid);
  launch_pdl(kern, dim3((unsigned)grid), dim3(96 + 32 * EW), (size_t)smem, st, a1, a2, w, ym, p, stages);
  const int rc = after_launch("conv3x3_col_kernel");
  return rc == 0 ? 0 : rc;
}

// Stride-1 layers with one N block (Cout_pad <= 64) on wide maps whose nine-tap weights fit in shared memory beside two
// halo stages.  Returns 1 if the layer was launched here, 0 if the caller should use the tile kernel, < 0 on error.
int try_row_kernel(const void* x, const void* x2, const void* wptr, int wrows, ConvParams& p, int BK, int C2, cudaStream_t st) {
  static const bool disabled = getenv("FALN_CONV_NO_ROW") != nullptr;
  if (disabled || p.stride != 1 || p.ncls != 1 || p.cls[0].n != 9 || p.out_mul != 1) return 0;
  if (wrows != 32 && wrows != 64) return 0;
  if (p.disp_out && wrows != 64) return 0;
  if (p.W < 192) return 0;
  const int kb = p.kblocks1 + p.kblocks2;
  const int halo = (kHaloCols * kHaloRows * BK * 2 + 1023) / 1024 * 1024;
  if (2048 + 4096 + 9 * kb * wrows * BK * 2 + (BK == 64 ? 2 : 3) * halo + 2 * 128 * wrows * 2 > 226 * 1024) return 0;
  CUtensorMap a1, a2, wm, ym;
  if (!make_row_map(&a1, x, p.B, p.H, p.W, p.C1, BK) || !make_w_map(&wm, wptr, wrows, 9 * p.Cin, BK, wrows) ||
      (x2 && !make_row_map(&a2, x2, p.B, p.H, p.W, C2, BK))) {
    set_error("conv3x3 row kernel: cuTensorMapEncodeTiled failed");
    return FALN_ERR_LAUNCH;
  }
  if (!x2) a2 = a1;
  // staged TMA store of the bf16 NHWC output: whole-chunk writes only (Cout == the N tile), 16-byte aligned rows
  static const bool no_tma_out = getenv("FALN_CONV_NO_TMA_OUT") != nullptr;
  p.tma_out = 0;
  ym = a1;
  if (!no_tma_out && !p.planar && !p.disp_out && p.Cout == wrows && p.out_c % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && make_out_map(&ym, p.out, (long long)p.B * p.H, p.W, p.out_c, wrows))
    p.tma_out = 1;
  // column-walk kernel (vertical taps as N columns, each input row fetched once): opt-in with FALN_CONV_COL=1 -- parity-green
  // but, as measured, still slower than the row kernel (see the kernel's header)
  static const int use_col = getenv("FALN_CONV_COL") ? atoi(getenv("FALN_CONV_COL")) : 0;
  if (use_col && !p.disp_out && p.H >= 8) {
    CUtensorMap c1, c2;
    if (!make_row_map(&c1, x, p.B, p.H, p.W, p.C1, BK, 1) || (x2 && !make_row_map(&c2, x2, p.B, p.H, p.W, C2, BK, 1))) {
      set_error("conv3x3 column-walk kernel: cuTensorMapEncodeTiled failed");
      return FALN_ERR_LAUNCH;
    }
    if (!x2) c2 = c1;
    int rc;
    // Sixteen epilogue warps (four per lane quadrant, 16 columns each; FALN_COL_EW=16) were measured for CO = 64, where one
    // CTA per SM leaves the eight warps alone with the drain: SLOWER (64 -> 64 at 8x192x640 forward 92.7 vs 88.2 us, data
    // gradient 89.0 vs 85.7 us; row kernel 80.4 / 79.6 us) -- the drain is not what holds this kernel back.  Eight is the default.
    static const int ew64 = getenv("FALN_COL_EW") ? atoi(getenv("FALN_COL_EW")) : 8;
    if (wrows == 64 && ew64 == 16)
      rc = BK == 64 ? launch_col<64, 64, 16>(c1, c2, wm, ym, p, st) : launch_col<32, 64, 16>(c1, c2, wm, ym, p, st);
    else if (BK == 64) rc = wrows == 64 ? launch_col<64, 64, 8>(c1, c2, wm, ym, p, st) : launch_col<64, 32, 8>(c1, c2, wm, ym, p, st);
    else rc = wrows == 64 ? launch_col<32, 64, 8>(c1, c2, wm, ym, p, st) : launch_col<32, 32, 8>(c1, c2, wm, ym, p, st);
    if (rc <= 0) return rc == 0 ? 1 : rc;                       // rc == 1: does not fit, fall through to the row kernel
  }
  int rc;
  if (BK == 64) rc = wrows == 64 ? launch_row<64, 64, 2>(a1, a2, wm, ym, p, st) : launch_row<64, 32, 2>(a1, a2, wm, ym, p, st);
  else rc = wrows == 64 ? launch_row<32, 64, 3>(a1, a2, wm, ym, p, st) : launch_row<32, 32, 2>(a1, a2, wm, ym, p, st);
  return rc == 0 ? 1 : rc;
}

// Small feature maps give few 128-pixel tiles: a wide N tile then leaves most SMs idle while a handful of CTAs walk a
// long K loop alone (measured: 16 CTAs, 22-45 us for the 3x10 bottleneck layers).  Narrow the N tile until the grid
// covers the SMs; the A tile is re-read by the extra CTAs out of L2.
int narrow_bn(int BN, int tiles, int cout_pad, int ncls) {
  // FALN_CONV_NARROW_PCT: work items (in % of the SM count) below which the N tile is halved.  Measured on the Stage-1 step
  // (gpurun_out/s16_*): 200 % -> 4.655 ms, 100 % -> 4.657 ms, 50 % -> 4.606 ms.
  static const int pct = getenv("FALN_CONV_NARROW_PCT") ? atoi(getenv("FALN_CONV_NARROW_PCT")) : 50;
  while (BN > 64 && tiles * (cout_pad / BN) * ncls < sm_count() * pct / 100) BN >>= 1;
  return BN;
}

// Ring depth: the producer runs ahead across tiles, so 3-4 stages with two CTAs per SM keep the HBM stream busy on the
// large maps; when the whole layer is at most two work items per SM (small maps, long K loops) a deeper ring with one
// CTA per SM hides the per-step TMA latency instead.
int dispatch(int BK, int BN, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& wm, const ConvParams& p,
             cudaStream_t st) {
  const long long total = (long long)p.tiles_w * p.tiles_h * p.B * ((p.Cout + BN - 1) / BN) * p.ncls;
  // FALN_CONV_DEEP_PCT: work items (in % of the SM count) up to which the deep-ring / one-CTA-per-SM variants are used
  static const int deep_pct = getenv("FALN_CONV_DEEP_PCT") ? atoi(getenv("FALN_CONV_DEEP_PCT")) : 200;
  const bool deep = total <= (long long)sm_count() * deep_pct / 100;
  if (BK == 64) {
    switch (BN) {
      case 256: return launch<64, 256, 4>(a1, a2, wm, p, st);
      case 128: return deep ? launch<64, 128, 6>(a1, a2, wm, p, st) : launch<64, 128, 3>(a1, a2, wm, p, st);
      case 64: return deep ? launch<64, 64, 8>(a1, a2, wm, p, st) : launch<64, 64, 2>(a1, a2, wm, p, st);
      default: return launch<64, 32, 2>(a1, a2, wm, p, st);
    }
  }
  switch (BN) {
    case 256: return launch<32, 256, 4>(a1, a2, wm, p, st);
    case 128: return launch<32, 128, 4>(a1, a2, wm, p, st);
    case 64: return launch<32, 64, 3>(a1, a2, wm, p, st);
    default: return launch<32, 32, 3>(a1, a2, wm, p, st);
  }
}

// ---- split-K cluster kernel: host side -------------------------------------------------------------------------
template <int BN, int SPLIT>
int launch_split(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const ConvParams& p, long long items,
                 cudaStream_t st, int* max_clusters) {
  using CF = SplitCfg<BN, SPLIT>;
  auto kern = conv3x3_splitk_kernel<BN, SPLIT>;
  static int resident = 0;   // clusters the device can hold at once (GPC topology x one CTA per SM at this shared-memory size)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(items > 0 ? items * SPLIT : SPLIT));
  cfg.blockDim = dim3(CF::kThreads);
  cfg.dynamicSmemBytes = CF::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = SPLIT;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (resident == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kTotal);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = (sm_count() * 7 / 8) / SPLIT;   // conservative: GPCs are not all the same size
    }
    resident = n;
    if (getenv("FALN_DEBUG")) fprintf(stderr, "conv3x3_splitk_kernel<%d,%d>: smem %d, %d stages, %d resident clusters\n", BN, SPLIT, CF::kTotal, CF::kStages, n);
  }
  if (max_clusters) {        // query only
    *max_clusters = resident;
    return 0;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a1, a2, w, p);
  if (e != cudaSuccess) {
    set_error("conv3x3_splitk_kernel launch failed: %s", cudaGetErrorString(e));
    return FALN_ERR_LAUNCH;
  }
  return after_launch("conv3x3_splitk_kernel");
}

int split_call(int BN, int SPLIT, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& w, const ConvParams& p,
               long long items, cudaStream_t st, int* max_clusters) {
  if (BN == 128) {
    if (SPLIT == 8) return launch_split<128, 8>(a1, a2, w, p, items, st, max_clusters);
    if (SPLIT == 4) return launch_split<128, 4>(a1, a2, w, p, items, st, max_clusters);
    if (SPLIT == 2) return launch_split<128, 2>(a1, a2, w, p, items, st, max_clusters);
  } else if (BN == 64) {
    if (SPLIT == 4) return launch_split<64, 4>(a1, a2, w, p, items, st, max_clusters);
    if (SPLIT == 2) return launch_split<64, 2>(a1, a2, w, p, items, st, max_clusters);
  }
  return FALN_ERR_ARG;
}

// Decides between the tile kernel (BN_cur as narrowed by the caller) and the split-K cluster kernel with a cost model in
// "KB through one SM's L2 port" (the measured bound of these layers): K steps per CTA x (A tile + B tile) + the partial sums
// exchanged over DSMEM.  Only layers whose work items fit one wave are considered.  Returns 1 if launched, 0 if the caller
// should go on with the tile kernel, < 0 on error.  FALN_CONV_SPLITK=0 disables; FALN_CONV_SPLITK_PCT = cost ratio (%) below
// which the split kernel is taken.
int try_splitk(const void* x, const void* x2, int C2, const void* wptr, int wrows, int wcols, ConvParams p, int BK, int BN_cur,
               cudaStream_t st) {
  static const int enabled = getenv("FALN_CONV_SPLITK") ? atoi(getenv("FALN_CONV_SPLITK")) : 1;
  static const int pct = getenv("FALN_CONV_SPLITK_PCT") ? atoi(getenv("FALN_CONV_SPLITK_PCT")) : 75;
  if (!enabled || BK != 64 || p.disp_out) return 0;
  const int kb = p.kblocks1 + p.kblocks2;
  int min_steps = 1 << 30, max_steps = 0;
  for (int c = 0; c < p.ncls; ++c) {
    min_steps = p.cls[c].n * kb < min_steps ? p.cls[c].n * kb : min_steps;
    max_steps = p.cls[c].n * kb > max_steps ? p.cls[c].n * kb : max_steps;
  }
  const long long cur_items = (long long)p.tiles_w * p.tiles_h * p.B * ((p.Cout + BN_cur - 1) / BN_cur) * p.ncls;
  if (cur_items > sm_count()) return 0;
  const double cost_cur = (double)max_steps * (16 + BN_cur / 8);
  const int th = p.Ho >= 5 ? 8 : (p.Ho >= 3 ? 4 : 2), tb = 8 / th;
  const int tiles_h = (p.Ho + th - 1) / th, tiles_b = (p.B + tb - 1) / tb;
  const long long tiles = (long long)p.tiles_w * tiles_h * tiles_b;
  static const int cand[5][2] = {{128, 8}, {128, 4}, {128, 2}, {64, 4}, {64, 2}};
  int best = -1;
  double best_cost = cost_cur * pct / 100.0;
  for (int i = 0; i < 5; ++i) {
    const int BN = cand[i][0], S = cand[i][1];
    if (wrows % BN != 0 || min_steps < S) continue;
    const long long items = tiles * ((p.Cout + BN - 1) / BN) * p.ncls;
    int resident = 0;
    CUtensorMap dummy{};
    if (split_call(BN, S, dummy, dummy, dummy, p, 0, st, &resident) != 0 || items > resident) continue;
    const double cost = (double)((max_steps + S - 1) / S) * (16 + BN / 8) + 2.0 * (BN / 2) * (S - 1) / S;
    if (cost < best_cost) {
      best_cost = cost;
      best = i;
    }
  }
  if (best < 0) return 0;
  const int BN = cand[best][0], S = cand[best][1];
  p.tile_h = th; p.tile_b = tb; p.tiles_h = tiles_h; p.tiles_b = tiles_b;
  p.nblk = (p.Cout + BN - 1) / BN;
  CUtensorMap a1, a2, wm;
  if (!make_act_map(&a1, x, p.B, p.H, p.W, p.C1, BK, p.stride, th, tb) || !make_w_map(&wm, wptr, wrows, wcols, BK, BN) ||
      (x2 && !make_act_map(&a2, x2, p.B, p.H, p.W, C2, BK, p.stride, th, tb))) {
    set_error("conv3x3 split-K kernel: cuTensorMapEncodeTiled failed");
    return FALN_ERR_LAUNCH;
  }
  if (!x2) a2 = a1;
  const int rc = split_call(BN, S, a1, a2, wm, p, tiles * p.nblk * p.ncls, st, nullptr);
  return rc == 0 ? 1 : rc;
}

}  // namespace
}  // namespace faln

using namespace faln;

// x [B,H,W,C1] bf16 NHWC, x2 [B,H,W,C2] or NULL (channel-concatenated after x), w [Cout_pad, 3, 3, C1+C2] bf16 (KRSC, rows
// beyond Cout zero), bias [Cout] fp32 or NULL, residual [B,Ho,Wo,out_c] bf16 or NULL.
// y: bf16 NHWC [B,Ho,Wo,out_c] (planar = 0) or fp32 planar [B,Cout,Ho,out_pitch] (planar = 1).
static int conv3x3_fwd_impl(const void* x, const void* x2, const void* w, const float* bias, const float* ctab,
                            const float* cscale, const void* residual, void* y, int B, int H, int W, int C1, int C2, int Cout,
                            int Cout_pad, int stride, int act, int planar, long long out_pitch, int out_c,
                            const float* disp_lvl, float* disp_out, faln_stream_t stream) {
  FALN_REQUIRE(x && w && (y || disp_out) && B > 0 && H > 0 && W > 0, "faln_conv3x3_fwd: null pointer / bad shape");
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
  BN = narrow_bn(BN, B * ((p.Wo + kTW - 1) / kTW) * ((p.Ho + kTH - 1) / kTH), Cout_pad, 1);
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
  p.disp_lvl = disp_lvl; p.disp_out = disp_out;
  {
    const int rr = try_row_kernel(x, x2, w, Cout_pad, p, BK, C2, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  if (disp_out) {
    set_error("faln_conv3x3_logits_disp: layer not eligible for the row-tile kernel (needs stride 1, W >= 192, Cout_pad == 64)");
    return FALN_ERR_ARG;
  }
  {
    const int rr = try_splitk(x, x2, C2, w, Cout_pad, 9 * (C1 + C2), p, BK, BN, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  CUtensorMap a1, a2, wm;
  if (!make_act_map(&a1, x, B, H, W, C1, BK, stride) || !make_w_map(&wm, w, Cout_pad, 9 * (C1 + C2), BK, BN) ||
      (x2 && !make_act_map(&a2, x2, B, H, W, C2, BK, stride))) {
    set_error("faln_conv3x3_fwd: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  if (!x2) a2 = a1;
  return dispatch(BK, BN, a1, a2, wm, p, as_stream(stream));
}

extern "C" int faln_conv3x3_fwd(const void* x, const void* x2, const void* w, const float* bias, const float* ctab,
                                const float* cscale, const void* residual, void* y, int B, int H, int W, int C1, int C2, int Cout, int Cout_pad, int stride, int act,
                                int planar, long long out_pitch, int out_c, faln_stream_t stream) {
  return conv3x3_fwd_impl(x, x2, w, bias, ctab, cscale, residual, y, B, H, W, C1, C2, Cout, Cout_pad, stride, act, planar,
                          out_pitch, out_c, nullptr, nullptr, stream);
}

// Inference form of the logits layer: disp[b,0,h,w] = sum_n d_lvl[b,n] * softmax_n(conv3x3(cat(x, x2)) + bias)_n, fused into
// the epilogue of the row-tile kernel so that the N logit planes are never written (reference: conv :174,215 + softmax and
// expectation :216-229).  bias is [64] padded with -inf beyond N, d_lvl is [B,64] padded with 0 (so the epilogue needs no
// per-plane predicates).  Needs W >= 192 and N <= 64; returns FALN_ERR_ARG otherwise (the caller then writes the logits
// and calls faln_med_disp).
extern "C" int faln_conv3x3_logits_disp(const void* x, const void* x2, const void* w, const float* bias, const float* d_lvl,
                                        float* disp, int B, int H, int W, int C1, int C2, int N, int Cout_pad,
                                        faln_stream_t stream) {
  FALN_REQUIRE(d_lvl && disp && bias && N > 0 && N <= 64 && Cout_pad == 64, "faln_conv3x3_logits_disp: need N <= 64 padded to 64 rows");
  FALN_REQUIRE(((reinterpret_cast<uintptr_t>(d_lvl) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
               "faln_conv3x3_logits_disp: bias [64] and d_lvl [B,64] must be 16-byte aligned");
  return conv3x3_fwd_impl(x, x2, w, bias, nullptr, nullptr, nullptr, nullptr, B, H, W, C1, C2, N, Cout_pad, 1, 0, 1, W, N,
                          d_lvl, disp, stream);
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
  int BN = Cx_pad % 256 == 0 ? 256 : (Cx_pad % 128 == 0 ? 128 : (Cx_pad % 64 == 0 ? 64 : 32));
  ConvParams p{};
  p.B = B; p.H = Hg; p.W = Wg;
  p.Ho = (H + stride - 1) / stride; p.Wo = (W + stride - 1) / stride;   // tile grid over one parity class
  BN = narrow_bn(BN, B * ((p.Wo + kTW - 1) / kTW) * ((p.Ho + kTH - 1) / kTH), Cx_pad, stride == 2 ? 4 : 1);
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
  {
    const int rr = try_row_kernel(g, nullptr, wd, Cx_pad, p, BK, 0, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  {
    const int rr = try_splitk(g, nullptr, 0, wd, Cx_pad, 9 * Cg, p, BK, BN, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  CUtensorMap a1, wm;
  if (!make_act_map(&a1, g, B, Hg, Wg, Cg, BK, 1) || !make_w_map(&wm, wd, Cx_pad, 9 * Cg, BK, BN)) {
    set_error("faln_conv3x3_dgrad: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  return dispatch(BK, BN, a1, a1, wm, p, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------------------------
// Nearest-neighbour 2x up-sampling folded INTO the 3x3 convolution that follows it (the reference's ``deconv`` block,
// /root/reference/models/FAL_netB.py:51-60: F.interpolate(nearest) then conv3x3).  For an exact 2x size the up-sampled
// image repeats every source pixel 2x2, so output pixel (2i+ph, 2j+pw) sees only a 2x2 neighbourhood of SOURCE pixels and the
// nine taps collapse into four per output parity class:
//     y[2i+ph, 2j+pw] = sum_{a,b in {0,1}} Wf[ph,pw][a][b] . x[i + a + ph - 1, j + b + pw - 1],
//     Wf[ph,pw][a][b] = sum_{kh in G(ph,a)} sum_{kw in G(pw,b)} W[kh][kw],  G(0,0)={0}, G(0,1)={1,2}, G(1,0)={0,1}, G(1,1)={2}
// (zero padding of the up-sampled image == zero fill of the source box).  2.25x fewer MMAs, the up-sampled tensor is never
// written or read, and the launch runs as four tap classes of the tile kernel (like the stride-2 data gradient).
// w: [Cout_pad][16][C1] bf16, virtual tap (ph*2+pw)*4 + a*2+b (faln_pack_up2_weights).  y: bf16 NHWC [B,2H,2W,out_c].
// ------------------------------------------------------------------------------------------------------------------
extern "C" int faln_conv3x3_up2_fwd(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C1,
                                    int Cout, int Cout_pad, int act, int out_c, faln_stream_t stream) {
  FALN_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "faln_conv3x3_up2_fwd: null pointer / bad shape");
  FALN_REQUIRE(C1 > 0 && C1 % 32 == 0 && Cout > 0 && Cout % 32 == 0 && Cout_pad >= Cout && Cout_pad % 32 == 0 && out_c >= Cout &&
                   out_c % 8 == 0, "faln_conv3x3_up2_fwd: channel counts must be multiples of 32");
  const int BK = (C1 % 64 == 0) ? 64 : 32;
  int BN = Cout_pad % 256 == 0 ? 256 : (Cout_pad % 128 == 0 ? 128 : (Cout_pad % 64 == 0 ? 64 : 32));
  ConvParams p{};
  p.B = B; p.H = H; p.W = W;
  p.Ho = H; p.Wo = W;                                     // tile grid over the SOURCE pixels (one parity class at a time)
  BN = narrow_bn(BN, B * ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH), Cout_pad, 4);
  p.C1 = C1; p.C2 = 0; p.Cin = C1; p.Cout = Cout;
  p.stride = 1; p.act = act; p.planar = 0;
  p.tiles_w = (W + kTW - 1) / kTW; p.tiles_h = (H + kTH - 1) / kTH;
  p.kblocks1 = C1 / BK; p.kblocks2 = 0;
  p.bias = bias; p.out = y; p.out_c = out_c; p.res_c = out_c;
  p.out_mul = 2; p.out_H = 2 * H; p.out_W = 2 * W;
  p.ncls = 4;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      TapClass& c = p.cls[ph * 2 + pw];
      c.n = 4; c.oh = (signed char)ph; c.ow = (signed char)pw;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          c.dh[a * 2 + b] = (signed char)(a + ph - 1);
          c.dw[a * 2 + b] = (signed char)(b + pw - 1);
          c.wt[a * 2 + b] = (signed char)((ph * 2 + pw) * 4 + a * 2 + b);
        }
    }
  {
    const int rr = try_splitk(x, nullptr, 0, w, Cout_pad, 16 * C1, p, BK, BN, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  CUtensorMap a1, wm;
  if (!make_act_map(&a1, x, B, H, W, C1, BK, 1) || !make_w_map(&wm, w, Cout_pad, 16 * C1, BK, BN)) {
    set_error("faln_conv3x3_up2_fwd: cuTensorMapEncodeTiled failed");
    return FALN_ERR_LAUNCH;
  }
  return dispatch(BK, BN, a1, a1, wm, p, as_stream(stream));
}

// Data gradient of the folded block: gradient w.r.t. the LOW-resolution input directly (the nearest-upsample backward --
// a 2x2 sum -- and the 3x3 data gradient fused):
//     gx[i, j] = sum_{r,c in {-1,0,1,2}} V[r][c]^T . g[2i + r, 2j + c],   V[r][c] = sum_{kh in Gr(r)} sum_{kw in Gr(c)} W[kh][kw],
//     Gr(-1) = {2}, Gr(0) = {1,2}, Gr(1) = {0,1}, Gr(2) = {0}
// i.e. a stride-2 gather over the high-resolution gradient with a 4x4 window (one tap class of 16 taps, box element
// stride 2).  wd: [Cx_pad][16][Cg] bf16, virtual tap (r+1)*4 + (c+1).  Epilogue as faln_conv3x3_dgrad.
extern "C" int faln_conv3x3_up2_dgrad(const void* g, const void* wd, void* gx, const void* ysave, int B, int H, int W, int Cg,
                                      int Cx, int dact, int gx_c, int ysave_c, faln_stream_t stream) {
  FALN_REQUIRE(g && wd && gx && B > 0 && H > 0 && W > 0, "faln_conv3x3_up2_dgrad: null pointer / bad shape");
  FALN_REQUIRE(Cg > 0 && Cg % 32 == 0 && Cx > 0 && Cx % 32 == 0 && gx_c >= Cx && gx_c % 8 == 0,
               "faln_conv3x3_up2_dgrad: channel counts must be multiples of 32");
  FALN_REQUIRE((dact == 0) == (ysave == nullptr), "faln_conv3x3_up2_dgrad: dact and ysave go together");
  const int BK = (Cg % 64 == 0) ? 64 : 32;
  int BN = Cx % 256 == 0 ? 256 : (Cx % 128 == 0 ? 128 : (Cx % 64 == 0 ? 64 : 32));
  ConvParams p{};
  p.B = B; p.H = 2 * H; p.W = 2 * W;                      // the tensor the TMA reads: the high-resolution gradient
  p.Ho = H; p.Wo = W;
  BN = narrow_bn(BN, B * ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH), Cx, 1);
  p.C1 = Cg; p.C2 = 0; p.Cin = Cg; p.Cout = Cx;
  p.stride = 2; p.act = 0; p.planar = 0;
  p.tiles_w = (W + kTW - 1) / kTW; p.tiles_h = (H + kTH - 1) / kTH;
  p.kblocks1 = Cg / BK; p.kblocks2 = 0;
  p.out = gx; p.out_c = gx_c; p.res_c = gx_c;
  p.out_mul = 1; p.out_H = H; p.out_W = W;
  p.dact = dact; p.ysave = static_cast<const __nv_bfloat16*>(ysave); p.ysave_c = ysave_c;
  p.ncls = 1;
  p.cls[0].n = 16; p.cls[0].oh = p.cls[0].ow = 0;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      p.cls[0].dh[r * 4 + c] = (signed char)(r - 1);
      p.cls[0].dw[r * 4 + c] = (signed char)(c - 1);
      p.cls[0].wt[r * 4 + c] = (signed char)(r * 4 + c);
    }
  {
    const int rr = try_splitk(g, nullptr, 0, wd, Cx, 16 * Cg, p, BK, BN, as_stream(stream));
    if (rr != 0) return rr > 0 ? 0 : rr;
  }
  CUtensorMap a1, wm;
  if (!make_act_map(&a1, g, B, 2 * H, 2 * W, Cg, BK, 2) || !make_w_map(&wm, wd, Cx, 16 * Cg, BK, BN)) {
    set_error("faln_conv3x3_up2_dgrad: cuTensorMapEncodeTiled failed");
    return FALN_ERR_LAUNCH;
  }
  return dispatch(BK, BN, a1, a1, wm, p, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core stem (round 2).  The first layer of the encoder (3 -> 32, ELU; /root/reference/models/FAL_netB.py:99) and of
// VGG19 (3 -> 64, ReLU; /root/reference/loss_functions.py:21) read the fp32 NCHW image.  The fp32-FMA stem kernel
// (conv_aux.cu) is FMA-bound: 27 x Cout FMAs per pixel, 139 us per 16 images at 64 channels against a ~45 us store floor.
// Here the 3x3x3 patch of every pixel becomes a 32-wide bf16 K vector (27 taps + 5 zeros; zero padding at the borders;
// optional horizontal flip) in ONE pass over the image, and the layer is a K = 32 GEMM on the tile kernel (one tap class
// with a single centre tap), bias / activation in its epilogue.  An extra block of the im2col launch packs the weights.
// ------------------------------------------------------------------------------------------------------------------
namespace faln {
namespace {
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ col, __nv_bfloat16* __restrict__ wpack,
                                                          int B, int H, int W, int Cout, int Cout_pad, int flip_x) {
  if (blockIdx.x == gridDim.x - 1) {               // weight pack: [Cout_pad][32], k = (kh*3 + kw)*3 + c
    for (int i = threadIdx.x; i < Cout_pad * 32; i += 256) {
      const int co = i / 32, k = i % 32;
      float v = 0.f;
      if (co < Cout && k < 27) {
        const int c = k % 3, t = k / 3;
        v = __ldg(w + (co * 3 + c) * 9 + t);
      }
      wpack[i] = __float2bfloat16(v);
    }
    return;
  }
  const long long npx = (long long)B * H * W, hw = (long long)H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < npx; i += (long long)(gridDim.x - 1) * 256) {
    const int xo = (int)(i % W), y = (int)((i / W) % H);
    const long long b = i / hw;
    const float* p = x + b * 3 * hw;
    __align__(16) __nv_bfloat16 v[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = xo + t % 3 - 1;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const int xs = flip_x ? W - 1 - xx : xx;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[t * 3 + c] = __float2bfloat16(in ? __ldg(p + c * hw + (long long)yy * W + xs) : 0.f);
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = __float2bfloat16(0.f);
    uint4* o = reinterpret_cast<uint4*>(col + i * 32);
    const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = src[q];
  }
}
}  // namespace
}  // namespace faln


namespace faln {
namespace {
// ------------------------------------------------------------------------------------------ fused tensor-core stem
// The 3 -> 32 / 3 -> 64 first layer (conv0.0 of the encoder, conv1_1 of the VGG slices) as ONE kernel on tcgen05: the fp32-FMA
// stem kernel (conv_aux.cu) spends 864 (1728) FMAs per pixel and runs 4-5x off both its HBM and its issue bound, and the
// im2col + GEMM pair above writes and re-reads a 64 B / px patch tensor.  Here a CTA of 128 threads owns 128 consecutive
// pixels of an image row per turn; every thread gathers the 27 taps of ITS pixel straight from the fp32 NCHW image, splits
// each into a bf16 high and low part (hi + lo carries 16 mantissa bits: the image is not rounded to bf16) and writes the row
//     A[pixel] = [ hi(27) 0(5) | lo(27) 0(5) ]   (K = 64 bf16 = one 128-byte swizzle row)
// of a K-major SWIZZLE_128B operand tile in shared memory -- the patch matrix never leaves the SM.  The weights sit beside it
// as B[co] = [ w(27) 0(5) | w(27) 0(5) ] (bf16 like every other layer's), four K = 16 instructions produce the 128 x CO fp32
// tile in TMEM, and the same threads read their pixel's row back (tcgen05.ld), add the bias, apply ELU / ReLU and store
// bf16 NHWC.  No pipeline inside the CTA: 25 KB of shared memory and CO TMEM columns let eight CTAs share an SM and hide each
// other's gather / MMA / store phases.
template <int CO, int ACT>
__global__ void __launch_bounds__(128, 8)
stem_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, int B, int H, int W, int tiles_w, long long total, int flip_x) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* At = base;                              // 128 rows x 128 B
  unsigned char* Bt = base + 128 * 128;                  // CO rows x 128 B
  float* sbias = reinterpret_cast<float*>(Bt + CO * 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbias + CO);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  // one swizzled operand row: 8 chunks of 8 bf16; chunk j of row r lives at physical chunk j ^ (r & 7)
  auto store_row = [](unsigned char* tile, int r, const uint32_t (&pk)[32]) {
    unsigned char* row = tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  };
  auto pack2 = [](float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  };

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, CO);
  if (tid < CO) {
    sbias[tid] = bias ? __ldg(bias + tid) : 0.f;
    float wv[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) wv[k] = k < 27 ? __ldg(w + tid * 27 + k) : 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) pk[k] = pk[16 + k] = pack2(wv[2 * k], wv[2 * k + 1]);
    store_row(Bt, tid, pk);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint64_t adesc = make_desc<64>(smem_u32(At)), bdesc = make_desc<64>(smem_u32(Bt));
  constexpr uint32_t idesc = make_idesc<CO>();
  const long long hw = (long long)H * W;
  uint32_t phase = 0;

  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int tw = (int)(t % tiles_w);
    const long long row = t / tiles_w;                   // b * H + y
    const int yy0 = (int)(row % H);
    const long long b = row / H;
    const int xo = tw * 128 + tid;
    // ---- gather the 27 taps of pixel (b, yy0, xo): fp32 NCHW, zero padding, optional horizontal flip of the source
    // (column / row offsets and their bounds are shared by the three channels: ~3 instructions per tap)
    float v[32];
    const float* pb = x + b * 3 * hw;
    int xoff[3], yoff[3];
    bool okx[3], oky[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int xx = xo + d - 1, yy = yy0 + d - 1;
      okx[d] = xx >= 0 && xx < W;
      oky[d] = yy >= 0 && yy < H;
      xoff[d] = flip_x ? W - 1 - xx : xx;
      yoff[d] = yy * W;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pc = pb + c * hw;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          v[c * 9 + dy * 3 + dx] = (oky[dy] && okx[dx]) ? __ldg(pc + (yoff[dy] + xoff[dx])) : 0.f;
      }
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      const float2 hf = __bfloat1622float2(h);
      pk[k] = *reinterpret_cast<const uint32_t*>(&h);
      pk[16 + k] = pack2(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
    }
    store_row(At, tid, pk);
    fence_proxy_async();                                 // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: this thread's pixel = TMEM lane 32 * warp + lane, CO columns, sixteen at a time -> bf16 row in the
    // (now free) A tile, swizzled so that a quarter-warp's 16-byte stores hit distinct banks; then the tile -- one CONTIGUOUS
    // run of valid_px * CO * 2 bytes in the NHWC output -- is copied out with fully coalesced 16-byte stores (a thread
    // writing its own 64 / 128-byte row straight to global memory touches 32 different sectors per instruction: twice the
    // L1 -> L2 store traffic, measured)
    constexpr int RB = CO * 2, CPR = RB / 16;            // row bytes, 16-byte chunks per row
    const int swz = CPR == 8 ? (tid & 7) : ((tid >> 1) & 3);
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
      float f[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(sbias + c0 + 4 * q);
        f[4 * q] = __uint_as_float(r[4 * q]) + bq.x;
        f[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
        f[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
        f[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (ACT == 1) f[j] = elu1(f[j]);
        if (ACT == 2) f[j] = fmaxf(f[j], 0.f);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
        *reinterpret_cast<uint4*>(At + tid * RB + (((c0 / 8 + q) ^ swz) << 4)) =
            make_uint4(pack2(f[8 * q], f[8 * q + 1]), pack2(f[8 * q + 2], f[8 * q + 3]), pack2(f[8 * q + 4], f[8 * q + 5]),
                       pack2(f[8 * q + 6], f[8 * q + 7]));
    }
    tc_fence_before();
    __syncthreads();                                     // the staged tile is complete (and everyone has read the accumulator)
    {
      const int valid_px = min(128, W - tw * 128);
      uint4* out = reinterpret_cast<uint4*>(y + ((row * W) + tw * 128) * CO);
      const int nchunks = valid_px * CPR;
#pragma unroll
      for (int k = 0; k < CPR; ++k) {
        const int i = tid + 128 * k;
        if (i < nchunks) {
          const int pp = i / CPR, q = i % CPR;
          const int sw = CPR == 8 ? (pp & 7) : ((pp >> 1) & 3);
          out[i] = *reinterpret_cast<const uint4*>(At + pp * RB + ((q ^ sw) << 4));
        }
      }
    }
    __syncthreads();                                     // the copy-out has read the tile: the next gather may overwrite A
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, CO);
  }
}
}  // namespace
}  // namespace faln

// The stem as one tcgen05 kernel (stem_mma_kernel): x [B,3,H,W] fp32 NCHW, w [Cout,3,3,3] fp32, bias [Cout] or NULL ->
// y [B,H,W,Cout] bf16 NHWC, Cout = 32 or 64.  Same contract as faln_stem_conv (conv_aux.cu), which it replaces by default.
extern "C" int faln_stem_conv_mma(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int Cout,
                                  int act, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "faln_stem_conv_mma: bad argument");
  FALN_REQUIRE(Cout == 32 || Cout == 64, "faln_stem_conv_mma: Cout must be 32 or 64 (got %d)", Cout);
  FALN_REQUIRE(act >= 0 && act <= 2, "faln_stem_conv_mma: act must be 0 (none), 1 (ELU) or 2 (ReLU)");
  FALN_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "faln_stem_conv_mma: y must be 16-byte aligned");
  const int tiles_w = (W + 127) / 128;
  const long long total = (long long)B * H * tiles_w;
  const int smem = 1024 + 128 * 128 + Cout * 128 + Cout * 4 + 64;
  static const int per_sm = getenv("FALN_STEM_CTAS") ? atoi(getenv("FALN_STEM_CTAS")) : 8;
  long long grid = (long long)sm_count() * (per_sm > 0 ? per_sm : 8);
  if (grid > total) grid = total;
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(y);
#define FALN_STEM_MMA(C, A) stem_mma_kernel<C, A><<<(int)grid, 128, smem, st>>>(x, w, bias, out, B, H, W, tiles_w, total, flip_x)
  if (Cout == 32) {
    if (act == 0) FALN_STEM_MMA(32, 0); else if (act == 1) FALN_STEM_MMA(32, 1); else FALN_STEM_MMA(32, 2);
  } else {
    if (act == 0) FALN_STEM_MMA(64, 0); else if (act == 1) FALN_STEM_MMA(64, 1); else FALN_STEM_MMA(64, 2);
  }
#undef FALN_STEM_MMA
  return after_launch("stem_mma_kernel");
}

// x [B,3,H,W] fp32 NCHW, w [Cout,3,3,3] fp32, bias [Cout] or NULL -> y [B,H,W,Cout] bf16 NHWC (Cout = 32 or 64).
// col: scratch [B,H,W,32] bf16; wpack: scratch [Cout,32] bf16 (both written by this call).
extern "C" int faln_stem_conv_tc(const float* x, const float* w, const float* bias, void* y, void* col, void* wpack, int B,
                                 int H, int W, int Cout, int act, int flip_x, faln_stream_t stream) {
  FALN_REQUIRE(x && w && y && col && wpack && B > 0 && H > 0 && W > 0, "faln_stem_conv_tc: bad argument");
  FALN_REQUIRE(Cout == 32 || Cout == 64, "faln_stem_conv_tc: Cout must be 32 or 64 (got %d)", Cout);
  FALN_REQUIRE(act >= 0 && act <= 2, "faln_stem_conv_tc: act must be 0 (none), 1 (ELU) or 2 (ReLU)");
  const long long npx = (long long)B * H * W;
  long long grid = (npx + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  stem_im2col_kernel<<<(int)grid + 1, 256, 0, as_stream(stream)>>>(x, w, static_cast<__nv_bfloat16*>(col),
                                                                 static_cast<__nv_bfloat16*>(wpack), B, H, W, Cout, Cout, flip_x);
  int rc = after_launch("stem_im2col_kernel");
  if (rc != FALN_OK) return rc;
  ConvParams p{};
  p.B = B; p.H = H; p.W = W; p.Ho = H; p.Wo = W;
  p.C1 = 32; p.C2 = 0; p.Cin = 32; p.Cout = Cout;
  p.stride = 1; p.act = act; p.planar = 0;
  p.tiles_w = (W + kTW - 1) / kTW; p.tiles_h = (H + kTH - 1) / kTH;
  p.kblocks1 = 1; p.kblocks2 = 0;
  p.bias = bias; p.out = y; p.out_c = Cout; p.res_c = Cout;
  p.out_mul = 1; p.out_H = H; p.out_W = W;
  p.ncls = 1;
  p.cls[0].n = 1; p.cls[0].dh[0] = 0; p.cls[0].dw[0] = 0; p.cls[0].wt[0] = 0; p.cls[0].oh = p.cls[0].ow = 0;
  CUtensorMap a1, wm;
  if (!make_act_map(&a1, col, B, H, W, 32, 32, 1) || !make_w_map(&wm, wpack, Cout, 32, 32, Cout)) {
    set_error("faln_stem_conv_tc: cuTensorMapEncodeTiled failed");
    return FALN_ERR_LAUNCH;
  }
  return dispatch(32, Cout, a1, a1, wm, p, as_stream(stream));
}
