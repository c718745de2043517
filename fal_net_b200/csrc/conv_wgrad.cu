// conv_wgrad.cu -- weight gradient of the 3x3 convolutions as a tcgen05 / TMEM GEMM over pixels, sm_100a.
//
// Replaces the cuDNN wgrad calls autograd issues for nn.Conv2d in /root/reference/models/FAL_netB.py:99-127 (backward of
// :144-174) and the ~3 ATen cast / add / cat launches around each of them.
//
//   dW[co, kh, kw, ci] = sum over (b, ho, wo) of  g[b, ho, wo, co] * x[b, ho*s + kh - 1, wo*s + kw - 1, ci]
//
// GEMM view (per filter tap): D[ci, co] = X_tap^T[ci, pixel] * G[pixel, co] -- the reduction dimension K is the PIXEL
// axis.  Both tensors are NHWC bf16, i.e. the channel (M resp. N) index is the contiguous one: both operands are
// "MN-major" for tcgen05.mma, which the instruction descriptor supports for 16-bit types (a_major = b_major = 1).
//
//   * a CTA owns ALL NINE taps of one block of XC input channels x NB output channels and a strided subset of the
//     64-pixel chunks (4 rows x 16 columns of the gradient map) -- split-K over pixels
//   * per chunk the TMA unit fetches
//       g : one box [NB channels, 16, 4]                       (64 channels -> 128B swizzle, 32 -> 64B swizzle)
//       x : stride 1, 64-channel blocks: ONE halo box [64 channels, 18, 6] -- every tap's window is a sub-rectangle of
//           it, addressed by the start address / LBO of the matrix descriptors.  Measured on B200: the 128B swizzle is a
//           function of the absolute shared-memory address, so a window may start on any 128-byte row with the
//           descriptor's base-offset field left 0 (setting it to (start >> 7) & 7 gives wrong results);
//           otherwise nine boxes [XC channels, 16*s, 4*s] at the taps' offsets with element stride s.
//           Out-of-image elements are zero-filled by the TMA bounds check = the conv's zero padding
//     a box lands as pixel rows of 128 (64) bytes = the canonical MN-major swizzled atoms: 8 pixel rows x one swizzle row
//     per atom, atoms 1024 (512) bytes apart along K (descriptor SBO)
//   * the M = 128 rows of one tcgen05.mma are TWO (four) taps' windows (descriptor LBO = distance between them), so
//     64- and 32-channel blocks fill the instruction; each instruction row owns NB TMEM columns (5 x 64 = 320 columns)
//   * the epilogue warps read the accumulators with tcgen05.ld and reduce them into the fp32 gradient tensor
//     [Cout, 3, 3, Cin_total] (KRSC, the layout of the flat gradient arena the fused Adam reads: lanes = consecutive ci,
//     so every red.global.add.f32 is a coalesced 128-byte request); gradients of a concatenated input are two calls with
//     different column offsets.  No workspace, no separate reduction kernel.
//   * the bias gradient rides along: nine taps leave one window slot of the last instruction row unused (three for 32-channel
//     blocks); that slot reads a block of bf16 ONES, so its accumulator rows are sum_pixels 1 * g[pixel, co] -- the bias
//     gradient of the layer, from operands the kernel streams anyway.  The CTAs of input-channel block 0 add one row of it to
//     dbias.  (Before: a separate channel-sum launch per biased layer re-read every gradient map, 13 launches = 2.7 % of the
//     Stage-1 step even on its own stream.)
#include <stdlib.h>

#include "tc_common.cuh"

namespace faln {
namespace {

constexpr int kCR = 4;            // gradient-map rows per chunk
constexpr int kP = kCR * kTW;     // 64 pixels per chunk = 4 UMMA K-steps
constexpr int kHaloW = kTW + 2, kHaloH = kCR + 2;

struct WgradParams {
  int B, Hg, Wg;            // gradient map
  int stride;
  int tiles_w, tiles_h, chunks;
  int n_cib, n_cob;         // channel blocks
  int NB;                   // output channels per block (= channels of one g box)
  int irows;                // instruction rows: ceil(9 taps / (128 / XC))
  int halo;                 // 1: x arrives as one halo box per chunk
  int x_bytes;              // bytes reserved per stage for x (1024-aligned), tx_x: bytes the TMA delivers
  int tx_x;
  int Cx, Cout;             // real channel counts (bounds of what is written)
  int ci_off, Cin_tot;      // column offset / row length of dW
  int stages;
  float* dW;
  float* dbias;             // optional [Cout] fp32: += sum over pixels of g (the conv's bias gradient), from the spare tap slot
  int ones_off;             // halo path: byte offset (from the 1024-aligned stage base) of the 2 KB block of bf16 ones
};

// MN-major swizzled shared-memory matrix descriptor: start >> 4, LBO >> 4 at [16,30) (distance between the 64- / 32-channel
// atoms along M/N), SBO >> 4 at [32,46) (distance between 8-pixel atoms along K), version 1 at [46,48), base offset at
// [49,52) (left 0, see above), layout type at [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
template <int ROWB>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t base_off = 0) {
  constexpr uint64_t sbo = 8 * ROWB;
  constexpr uint64_t layout = (ROWB == 128) ? 2 : 4;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((sbo >> 4) << 32) |
         (1ULL << 46) | ((uint64_t)(base_off & 7) << 49) | (layout << 61);
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, both MN-major (bits 15, 16), N >> 3 at [17,23), M = 128
__device__ __forceinline__ uint32_t make_idesc_mn(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// The body of one weight-gradient CTA: channel block `blk` (= cob * n_cib + cib), split `split` of `nsplit` over the chunks.
// Shared by the one-layer kernel and the batched one (several small layers in one grid).
template <int XROWB, int GROWB, bool HALO>
__device__ __forceinline__ void wgrad_cta(const CUtensorMap& tmX, const CUtensorMap& tmG, const WgradParams& p, const int blk,
                                          const int split, const int nsplit) {
  constexpr int XC = XROWB / 2;                        // channels per x box
  constexpr int NB = GROWB / 2;                        // output channels per block = channels of the g box
  constexpr int XBOX = kP * XROWB, GBOX = kP * GROWB;  // bytes per (non-halo) box
  constexpr int SPR = 128 / XC;                        // taps per instruction row
  constexpr int IROWS = (9 + SPR - 1) / SPR;           // instruction rows (5 for 64-channel blocks, 3 for 32-channel ones)
  extern __shared__ unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.x_bytes + GBOX;
  const int tx_bytes = p.tx_x + GBOX;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cib = blk % p.n_cib, cob = blk / p.n_cib;
  const int my_chunks = (p.chunks - split + nsplit - 1) / nsplit;  // chunks split, split + nsplit, ...
  constexpr uint32_t tmem_cols = IROWS * NB;
  constexpr uint32_t alloc_cols = tmem_cols <= 32 ? 32 : tmem_cols <= 64 ? 64 : tmem_cols <= 128 ? 128 : tmem_cols <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, alloc_cols);
  if (p.dbias) {
    // bf16 ones for the spare tap slot (generic-proxy stores, made visible to the tensor core's async proxy below):
    // halo path: one 2 KB block (two 8-pixel atoms) behind the ring; box path: slot 9 of every stage, which no TMA load writes
    constexpr uint32_t kOnes = 0x3F803F80u;
    if (HALO) {
      uint4* o = reinterpret_cast<uint4*>(stages + p.ones_off);
      for (int i = threadIdx.x; i < 2048 / 16; i += 192) o[i] = make_uint4(kOnes, kOnes, kOnes, kOnes);
    } else {
      for (int s = 0; s < p.stages; ++s) {
        uint4* o = reinterpret_cast<uint4*>(stages + (size_t)s * stage_bytes + 9 * XBOX);
        for (int i = threadIdx.x; i < XBOX / 16; i += 192) o[i] = make_uint4(kOnes, kOnes, kOnes, kOnes);
      }
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                 // prologue overlapped the previous kernel's tail (common.cuh: programmatic dependent launch)
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_chunks; ++it) {
        const int c = split + it * nsplit;
        const int tw = c % p.tiles_w, th = (c / p.tiles_w) % p.tiles_h, b = c / (p.tiles_w * p.tiles_h);
        const int ho0 = th * kCR, wo0 = tw * kTW;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* xs = stages + (size_t)s * stage_bytes;
        unsigned char* gs = xs + p.x_bytes;
        mbar_arrive_expect_tx(&full[s], tx_bytes);
        tma_load_4d(gs, &tmG, cob * NB, wo0, ho0, b, &full[s]);
        if (HALO) {
          tma_load_4d(xs, &tmX, cib * XC, wo0 - 1, ho0 - 1, b, &full[s]);
        } else {
          const int wi = wo0 * p.stride - 1, hi = ho0 * p.stride - 1;
#pragma unroll
          for (int t = 0; t < 9; ++t) tma_load_4d(xs + t * XBOX, &tmX, cib * XC, wi + t % 3, hi + t / 3, b, &full[s]);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(NB);
      // everything but the stage base address is a compile-time constant: descriptors are (template + base >> 4)
      const uint64_t b_tmpl = make_desc_mn<GROWB>(0, GBOX);
      const uint64_t a_tmpl_box = make_desc_mn<XROWB>(0, XBOX);
      const uint32_t ones_base = smem_u32(stages + p.ones_off);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_chunks; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xs = smem_u32(stages + (size_t)s * stage_bytes);
        const uint64_t xs16 = (uint64_t)(xs >> 4), gs16 = (uint64_t)((xs + p.x_bytes) >> 4);
#pragma unroll
        for (int k = 0; k < kP / 16; ++k) {
          // K-step k = the 16 pixels of chunk row k: two 8-pixel atoms, 8 swizzle rows apart (SBO)
          const uint64_t bdesc = b_tmpl + gs16 + (uint64_t)(k * 16 * GROWB / 16);
#pragma unroll
          for (int j = 0; j < IROWS; ++j) {
            uint64_t adesc;
            if (HALO && XC == 32) {
              // 32-channel blocks: instruction row j = the taps of COLUMN kw = j; its four 32-channel atoms are the windows
              // of kh = 0, 1, 2 (one halo row = 18 pixels = the uniform LBO apart) and a spare one (kh = 3, dropped)
              const uint32_t o0 = (uint32_t)((k * kHaloW + j) * XROWB);
              adesc = make_desc_mn<XROWB>(0, kHaloW * XROWB) + xs16 + (uint64_t)(o0 >> 4);
            } else if (HALO) {
              // instruction row j = taps 2j, 2j+1: windows of the halo tile starting at halo row (k + kh) * 18 + kw
              const int t0 = 2 * j, t1 = 2 * j + 1;
              const uint32_t o0 = (uint32_t)(((k + t0 / 3) * kHaloW + t0 % 3) * XROWB);
              uint32_t o1 = (uint32_t)(((k + t1 / 3) * kHaloW + t1 % 3) * XROWB);
              if (t1 == 9 && p.dbias) o1 = ones_base - xs;       // the spare slot: the block of ones (behind every stage)
              adesc = make_desc_mn<XROWB>(0, o1 - o0) + xs16 + (uint64_t)(o0 >> 4);
            } else {
              adesc = a_tmpl_box + xs16 + (uint64_t)((j * SPR * XBOX + k * 16 * XROWB) >> 4);
            }
            umma_bf16(tmem_base + j * NB, adesc, bdesc, idesc, (it | k) != 0);
          }
        }
        umma_commit(&empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (my_chunks > 0) {
    // ================================================================= epilogue (warps 2..5): red.add into dW (KRSC)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;   // accumulator row inside an instruction row
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // CTAs of the same channel block reduce into the same addresses: start each split at a different (row, column
    // group) so they do not sweep the output in lock-step
    constexpr int kSteps = IROWS * (NB / 32);
    for (int q = 0; q < kSteps; ++q) {
      const int step = (q + split) % kSteps;
      const int j = step / (NB / 32), c0 = (step % (NB / 32)) * 32;
      const int tap = (HALO && XC == 32) ? (m / XC < 3 ? (m / XC) * 3 + j : 9) : j * SPR + m / XC;
      const int ci = cib * XC + (m % XC);
      const bool row_ok = tap < 9 && ci < p.Cx;
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + j * NB + c0, r);
        if (tap == 9 && (m % XC) == 0 && cib == 0 && p.dbias) {   // first row of the ones slot: sum over this CTA's pixels of g
#pragma unroll
          for (int n = 0; n < 32; ++n)
            if (cob * NB + c0 + n < p.Cout) atomicAdd(p.dbias + cob * NB + c0 + n, __uint_as_float(r[n]));
        }
        if (!row_ok) continue;
        const int co0 = cob * NB + c0;
        float* dst = p.dW + ((size_t)co0 * 9 + tap) * p.Cin_tot + p.ci_off + ci;
#pragma unroll
        for (int n = 0; n < 32; ++n)
          if (co0 + n < p.Cout) atomicAdd(dst + (size_t)n * 9 * p.Cin_tot, __uint_as_float(r[n]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, alloc_cols);
  }
}


template <int XROWB, int GROWB, bool HALO>
__global__ void __launch_bounds__(192)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgradParams p) {
  wgrad_cta<XROWB, GROWB, HALO>(tmX, tmG, p, blockIdx.y, blockIdx.x, gridDim.x);
}

// Several (small) layers in ONE grid.  On the 3x10 ... 12x40 maps a weight-gradient launch is a chain of latencies -- barrier
// init, TMEM allocation, first TMA round trip, a handful of chunks, the reduction: ~13 us + 0.6 us per chunk whatever the layer
// -- during which its CTAs hold an SM's shared memory and all 512 TMEM columns and keep the data-gradient chain off it.  Ten such
// launches in a row cost ~180 us of that; as one grid (every CTA finds its job by its block index) they cost one.
constexpr int kMaxBatch = 12;
struct WgradJob {
  CUtensorMap mx, mg;
  WgradParams p;
  int splits, cta0;               // CTAs [cta0, cta0 + splits * n_cib * n_cob) belong to this job
};
struct WgradBatch {
  WgradJob job[kMaxBatch];
  int njobs;
};
template <int XROWB, int GROWB, bool HALO>
__global__ void __launch_bounds__(192)
conv3x3_wgrad_multi_kernel(const __grid_constant__ WgradBatch batch) {
  int j = 0;
  for (int k = 1; k < batch.njobs; ++k)
    if ((int)blockIdx.x >= batch.job[k].cta0) j = k;
  const WgradJob& job = batch.job[j];
  const int local = (int)blockIdx.x - job.cta0;
  wgrad_cta<XROWB, GROWB, HALO>(job.mx, job.mg, job.p, local / job.splits, local % job.splits, job.splits);
}

// ------------------------------------------------------------------------------------------ folded deconv weight gradient
// Weight gradient of "nearest 2x up-sampling, then conv3x3" (the reference's deconv block, models/FAL_netB.py:51-60) WITHOUT the
// up-sampled tensor -- the counterpart of faln_conv3x3_up2_fwd / _dgrad (conv_tc.cu).  With up(h)[y, x] = h[y >> 1, x >> 1],
// output pixel (2i + py, 2j + px) and tap (kh, kw) read h[i + a, j + b], a = floor((py + kh - 1) / 2), b likewise: per output
// parity class only a 2 x 2 neighbourhood of SOURCE pixels is touched (a in {-1, 0} for py = 0, {0, 1} for py = 1), so
//     dW[kh, kw] = sum over (py, px) of S[py, px][a(py, kh)][b(px, kw)],
//     S[py, px][a][b][ci, co] = sum_{i, j} h[i + a, j + b, ci] * g[2i + py, 2j + px, co]
// = sixteen quarter-resolution correlations instead of nine full-resolution ones (2.25 x fewer MMAs), and the low-resolution
// map is read instead of a 4 x larger one that first had to be written.  Per 4 x 16 chunk of the LOW-resolution map the TMA unit
// fetches the [64, 18, 6] halo box of h (the nine (a, b) windows are sub-rectangles of it, as in the stride-1 kernel above) and
// FOUR parity-sub-sampled boxes of g (element strides 2, 2).  Instruction row ir of class (py, px) pairs the windows
// (a_ir, b_lo), (a_ir, b_hi) -- 128 bytes apart in the halo tile -- into the M = 128 rows; 4 classes x 2 rows x 64 columns fill
// the 512 TMEM columns.  The epilogue folds the classes that feed the same tap in registers (the (py, ir) pairs of a kh live in
// different TMEM columns of the same lanes; the kw = 1 tap is fed by b_hi of px = 0 and b_lo of px = 1, i.e. by different lane
// halves, and takes two reductions): 12 tap blocks of red.add per CTA against 9 for a plain 3 x 3 kernel.
__device__ __forceinline__ void wgrad_up2_cta(const CUtensorMap& tmX, const CUtensorMap& tmG, const WgradParams& p, const int blk,
                                              const int split, const int nsplit) {
  constexpr int XROWB = 128, GROWB = 128, NB = 64, XC = 64;
  constexpr int GBOX = kP * GROWB;
  constexpr int kHaloBytes = kHaloW * kHaloH * XROWB;
  extern __shared__ unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  unsigned char* stages = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 256 + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.x_bytes + 4 * GBOX;
  const int tx_bytes = kHaloBytes + 4 * GBOX;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cib = blk % p.n_cib, cob = blk / p.n_cib;
  const int my_chunks = (p.chunks - split + nsplit - 1) / nsplit;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_chunks; ++it) {
        const int c = split + it * nsplit;
        const int tw = c % p.tiles_w, th = (c / p.tiles_w) % p.tiles_h, b = c / (p.tiles_w * p.tiles_h);
        const int i0 = th * kCR, j0 = tw * kTW;                  // low-resolution origin of the chunk
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* xs = stages + (size_t)s * stage_bytes;
        unsigned char* gs = xs + p.x_bytes;
        mbar_arrive_expect_tx(&full[s], tx_bytes);
        tma_load_4d(xs, &tmX, cib * XC, j0 - 1, i0 - 1, b, &full[s]);
#pragma unroll
        for (int cl = 0; cl < 4; ++cl)                           // class (py, px) = (cl >> 1, cl & 1): g[2i + py, 2j + px]
          tma_load_4d(gs + cl * GBOX, &tmG, cob * NB, 2 * j0 + (cl & 1), 2 * i0 + (cl >> 1), b, &full[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(NB);
      const uint64_t b_tmpl = make_desc_mn<GROWB>(0, GBOX);
      const uint64_t a_tmpl = make_desc_mn<XROWB>(0, XROWB);     // LBO: the b_hi window starts one pixel (128 bytes) after b_lo
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_chunks; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xs = smem_u32(stages + (size_t)s * stage_bytes);
        const uint64_t xs16 = (uint64_t)(xs >> 4), gs16 = (uint64_t)((xs + p.x_bytes) >> 4);
#pragma unroll
        for (int k = 0; k < kP / 16; ++k) {
#pragma unroll
          for (int cl = 0; cl < 4; ++cl) {
            const int py = cl >> 1, px = cl & 1;
            const uint64_t bdesc = b_tmpl + gs16 + (uint64_t)((cl * GBOX + k * 16 * GROWB) >> 4);
#pragma unroll
            for (int ir = 0; ir < 2; ++ir) {
              // halo row of chunk row k shifted by a = ir - 1 (py = 0) or ir (py = 1): k + a + 1; halo column of b_lo: px
              const uint32_t o0 = (uint32_t)(((k + ir + py) * kHaloW + px) * XROWB);
              umma_bf16(tmem_base + (uint32_t)((cl * 2 + ir) * NB), a_tmpl + xs16 + (uint64_t)(o0 >> 4), bdesc, idesc,
                        (it | k) != 0);
            }
          }
        }
        umma_commit(&empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (my_chunks > 0) {
    // ================================================================= epilogue (warps 2..5): fold the classes, red.add into dW
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int half = m >> 6;                                     // 0: the b_lo windows, 1: the b_hi windows (warp-uniform)
    const int ci = cib * XC + (m & 63);
    const bool row_ok = ci < p.Cx;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int q = 0; q < 6; ++q) {
      const int step = (q + split) % 6;                          // CTAs of one channel block start at different places
      const int kh = step >> 1, c0 = (step & 1) * 32;
      // the (py, ir) accumulators of this kh: py = 0 uses a = -1 (ir 0) for kh = 0 and a = 0 (ir 1) otherwise; py = 1 uses
      // a = 0 (ir 0) for kh = 0, 1 and a = 1 (ir 1) for kh = 2
      const int ir0 = kh == 0 ? 0 : 1, ir1 = kh == 2 ? 1 : 0;
      uint32_t r[32];
      float s0[32], s1[32];                                      // sums over py of the px = 0 / px = 1 classes
      tmem_ld32(tlane + (uint32_t)(((0 * 2 + 0) * 2 + ir0) * NB + c0), r);
#pragma unroll
      for (int n = 0; n < 32; ++n) s0[n] = __uint_as_float(r[n]);
      tmem_ld32(tlane + (uint32_t)(((1 * 2 + 0) * 2 + ir1) * NB + c0), r);
#pragma unroll
      for (int n = 0; n < 32; ++n) s0[n] += __uint_as_float(r[n]);
      tmem_ld32(tlane + (uint32_t)(((0 * 2 + 1) * 2 + ir0) * NB + c0), r);
#pragma unroll
      for (int n = 0; n < 32; ++n) s1[n] = __uint_as_float(r[n]);
      tmem_ld32(tlane + (uint32_t)(((1 * 2 + 1) * 2 + ir1) * NB + c0), r);
#pragma unroll
      for (int n = 0; n < 32; ++n) s1[n] += __uint_as_float(r[n]);
      if (!row_ok) continue;
      // b_lo is b = -1 <-> kw = 0 for px = 0 and b = 0 <-> kw in {0, 1} for px = 1; b_hi is b = 0 <-> kw in {1, 2} for px = 0 and
      // b = 1 <-> kw = 2 for px = 1
      const int kw_both = half == 0 ? 0 : 2;                     // the tap both px classes feed from this lane half
      const int co0 = cob * NB + c0;
      float* dst = p.dW + ((size_t)co0 * 9 + kh * 3) * p.Cin_tot + p.ci_off + ci;
#pragma unroll
      for (int n = 0; n < 32; ++n) {
        if (co0 + n < p.Cout) {
          float* d = dst + (size_t)n * 9 * p.Cin_tot;
          atomicAdd(d + (size_t)kw_both * p.Cin_tot, s0[n] + s1[n]);
          atomicAdd(d + (size_t)p.Cin_tot, half == 0 ? s1[n] : s0[n]);        // kw = 1
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


__global__ void __launch_bounds__(192)
conv3x3_wgrad_up2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgradParams p) {
  wgrad_up2_cta(tmX, tmG, p, blockIdx.y, blockIdx.x, gridDim.x);
}
// several folded deconv layers in one grid (see conv3x3_wgrad_multi_kernel)
__global__ void __launch_bounds__(192) conv3x3_wgrad_up2_multi_kernel(const __grid_constant__ WgradBatch batch) {
  int j = 0;
  for (int k = 1; k < batch.njobs; ++k)
    if ((int)blockIdx.x >= batch.job[k].cta0) j = k;
  const WgradJob& job = batch.job[j];
  const int local = (int)blockIdx.x - job.cta0;
  wgrad_up2_cta(job.mx, job.mg, job.p, local / job.splits, local % job.splits, job.splits);
}

// ------------------------------------------------------------------------------------------ stem weight gradient
// Weight (and bias) gradient of the 3 -> 32 first layer from the fp32 NCHW image itself.  The generic kernel above needs the
// image as a 32-channel bf16 NHWC tensor (a 63 MB transpose per step) and then spends a full 32 x 32-channel launch on 27 x 32
// numbers.  Here the patch rows of the fused stem forward (conv_tc.cu, stem_mma_kernel: [hi(27) 1 0(4) | lo(27) 0(5)] per
// pixel, K = taps there) are rebuilt in shared memory and read as the MN-major operand of dW^T[tap, co] = sum_pixels
// patch[pixel, tap] * g[pixel, co]: M = the 64 hi / lo tap slots (+ 64 rows that are never read), N = 32, K = the 128 pixels of
// the tile = eight instructions; the accumulator stays in TMEM for all tiles of the persistent CTA and is reduced once at the
// end (hi + lo rows add into the same element; slot 27 holds ones, so its row is the bias gradient).
__global__ void __launch_bounds__(128, 4)
stem_wgrad_mma_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ g, float* __restrict__ dW,
                      float* __restrict__ dbias, int B, int H, int W, int tiles_w, long long total) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* At = base;                              // 128 pixel rows x 128 B (64 tap slots)
  unsigned char* Bt = base + 2 * 128 * 128;              // 128 pixel rows x 64 B (32 output channels); [At + 16 KB, Bt): rows 64..127 of M
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bt + 128 * 64);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto pack2 = [](float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 32);
  {  // the never-read second M atom: finite values (zeros) so that nothing odd is ever multiplied
    uint4* z = reinterpret_cast<uint4*>(At + 128 * 128);
    for (int i = tid; i < 128 * 128 / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = make_idesc_mn(32);
  const uint64_t a_tmpl = make_desc_mn<128>(smem_u32(At), 128 * 128);     // LBO: the second 64-slot atom along M
  const uint64_t b_tmpl = make_desc_mn<64>(smem_u32(Bt), 128 * 64);
  const long long hw = (long long)H * W;
  uint32_t phase = 0;
  int it = 0;
  for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int tw = (int)(t % tiles_w);
    const long long row = t / tiles_w;                   // b * H + y
    const int yy0 = (int)(row % H);
    const long long b = row / H;
    const int xo = tw * 128 + tid;
    // ---- patch row of pixel (b, yy0, xo), as in stem_mma_kernel, with a ONE in slot 27
    float v[32];
    const float* pb = x + b * 3 * hw;
    int xoff[3], yoff[3];
    bool okx[3], oky[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int xx = xo + d - 1, yy = yy0 + d - 1;
      okx[d] = xx >= 0 && xx < W;
      oky[d] = yy >= 0 && yy < H;
      xoff[d] = xx;
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
    v[27] = 1.0f;
#pragma unroll
    for (int k = 28; k < 32; ++k) v[k] = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      const float2 hf = __bfloat1622float2(h);
      pk[k] = *reinterpret_cast<const uint32_t*>(&h);
      pk[16 + k] = pack2(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
    }
    {
      unsigned char* rowp = At + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ (tid & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    // ---- gradient rows of the tile: one contiguous run of valid_px * 64 bytes (NHWC, 32 channels); zeros beyond the row end
    {
      const int valid_px = min(128, W - tw * 128);
      const uint4* gp = reinterpret_cast<const uint4*>(g + ((row * W) + tw * 128) * 32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = tid + 128 * k, pp = i >> 2, q = i & 3;
        const uint4 u = pp < valid_px ? __ldg(gp + i) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Bt + pp * 64 + ((q ^ ((pp >> 1) & 3)) << 4)) = u;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)                     // K-step = 16 pixels = two 8-pixel atoms
        umma_bf16(tmem_base, a_tmpl + (uint64_t)((ks * 16 * 128) >> 4), b_tmpl + (uint64_t)((ks * 16 * 64) >> 4), idesc,
                  (it | ks) != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);                               // the instructions have read the tiles: they may be overwritten
    phase ^= 1;
  }
  tc_fence_after();
  if (it > 0) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
    const int k = tid < 27 ? tid : ((tid >= 32 && tid < 59) ? tid - 32 : -1);
    if (k >= 0) {
      float* d = dW + (k % 9) * 3 + k / 9;               // dW [co][kh][kw][c] (KRSC), k = c * 9 + kh * 3 + kw
#pragma unroll
      for (int co = 0; co < 32; ++co) atomicAdd(d + co * 27, __uint_as_float(r[co]));
    } else if (tid == 27 && dbias) {
#pragma unroll
      for (int co = 0; co < 32; ++co) atomicAdd(dbias + co, __uint_as_float(r[co]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

template <int XROWB, int GROWB, bool HALO>
int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mg, const WgradParams& p, int splits, int smem, cudaStream_t st) {
  auto kern = conv3x3_wgrad_kernel<XROWB, GROWB, HALO>;
  static int attr_set = 0;
  if (attr_set < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = smem;
  }
  dim3 grid(splits, p.n_cib * p.n_cob, 1);
  launch_pdl(kern, grid, dim3(192), (size_t)smem, st, mx, mg, p);
  return after_launch("conv3x3_wgrad_kernel");
}

// halo tensor map: box [64 channels, 18, 6, 1], 128B swizzle
bool make_halo_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int XC = 64) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)XC, (cuuint32_t)kHaloW, (cuuint32_t)kHaloH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, XC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// per-(sample, border class, channel) sums of an NHWC bf16 map: out[b][rc][cc][c], rc/cc = 0 first row/col, 1 interior,
// 2 last row/col.  One CTA per (b, row); thread = (column slice, channel), partial sums folded with fp32 atomics.
__global__ void __launch_bounds__(256) border_sum_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ out, int B,
                                                         int H, int W, int C, int Cs) {
  const int row = blockIdx.x;  // b * H + h
  const int b = row / H, h = row % H;
  const int rc = h == 0 ? 0 : (h == H - 1 ? 2 : 1);
  const __nv_bfloat16* src = g + (size_t)row * W * Cs;
  const int parts = C <= 128 ? 256 / C : 1;  // column slices per channel (C = 32, 64, 128 -> 8, 4, 2)
  for (int c = threadIdx.x % (256 / parts); c < C; c += 256 / parts) {
    const int part = threadIdx.x / (256 / parts);
    float mid = 0.f;
    for (int w = 1 + part; w < W - 1; w += parts) mid += __bfloat162float(src[(size_t)w * Cs + c]);
    float* o = out + (((size_t)b * 3 + rc) * 3) * C + c;
    atomicAdd(o + C, mid);
    if (part == 0) {
      atomicAdd(o, __bfloat162float(src[c]));
      atomicAdd(o + 2 * C, __bfloat162float(src[(size_t)(W - 1) * Cs + c]));
    }
  }
}

}  // namespace
}  // namespace faln

using namespace faln;

// g  [B,Hg,Wg,Cg] bf16 NHWC: gradient w.r.t. the conv's pre-activation output (Cg % 32 == 0; channels >= Cout are ignored)
// x  [B,H,W,Cxs]  bf16 NHWC: the conv's input (one source of a concatenated input per call; Cxs % 32 == 0)
// dW [Cout, 3, 3, Cin_tot] fp32 (KRSC): columns [ci_off, ci_off + Cx) are ACCUMULATED into (caller zeroes them once per step)
// flags: bit 0 = never use the halo path (validation: nine boxes per chunk instead)
namespace {
struct PreparedWgrad {
  CUtensorMap mx, mg;
  WgradParams p;
  int splits, smem, key;          // key = template choice: bit 0 halo, bit 1 XC == 64, bit 2 GC == 64
};

// Everything a weight-gradient launch needs (parameters, tensor maps, split count, shared memory), shared by the one-layer and
// the batched entry point.  max_splits > 0 caps the split count (batched launches aim for ONE wave over all their jobs).
int prepare_wgrad(const void* g, const void* x, float* dW, float* dbias, int B, int H, int W, int Cg, int Cxs, int Cout, int Cx,
                  int ci_off, int Cin_tot, int stride, unsigned flags, int ring_kb_cap, PreparedWgrad& out) {
  FALN_REQUIRE(g && x && dW && B > 0 && H > 0 && W > 0, "faln_conv3x3_wgrad: null pointer / bad shape");
  FALN_REQUIRE(stride == 1 || stride == 2, "faln_conv3x3_wgrad: stride must be 1 or 2");
  FALN_REQUIRE(Cg % 32 == 0 && Cxs % 32 == 0 && Cg > 0 && Cxs > 0, "faln_conv3x3_wgrad: channel strides must be multiples of 32");
  FALN_REQUIRE(Cout > 0 && Cout <= Cg && Cx > 0 && Cx <= Cxs && ci_off >= 0 && ci_off + Cx <= Cin_tot,
               "faln_conv3x3_wgrad: channel ranges out of bounds");
  const int Hg = (H - 1) / stride + 1, Wg = (W - 1) / stride + 1;
  const int XC = Cxs % 64 == 0 ? 64 : 32, GC = Cg % 64 == 0 ? 64 : 32;
  FALN_REQUIRE(XC == 64 || Cxs == 32, "faln_conv3x3_wgrad: input channel count must be 32 or a multiple of 64");
  FALN_REQUIRE(GC == 64 || Cg == 32, "faln_conv3x3_wgrad: gradient channel count must be 32 or a multiple of 64");
  WgradParams p{};
  p.B = B; p.Hg = Hg; p.Wg = Wg; p.stride = stride;
  p.tiles_w = (Wg + kTW - 1) / kTW; p.tiles_h = (Hg + kCR - 1) / kCR;
  p.chunks = B * p.tiles_w * p.tiles_h;
  p.NB = GC;
  p.n_cib = (Cx + XC - 1) / XC;       // blocks that contain written channels only
  p.n_cob = (Cout + p.NB - 1) / p.NB;
  const int spr = 128 / XC;
  p.irows = (9 + spr - 1) / spr;
  // halo tile: stride-1 layers; 32-channel blocks too (three instruction rows = tap columns), unless the bias gradient rides
  // along -- the column grouping has no freely addressable spare slot for the ones (only the stem's weight gradient: box path).
  // FALN_WGRAD_NO_HALO32=1: nine boxes per chunk for the 32-channel blocks, as before (A/B switch).
  static const bool no_halo32 = getenv("FALN_WGRAD_NO_HALO32") != nullptr;
  p.halo = (stride == 1 && !(flags & 1u) && (XC == 64 || (!dbias && !no_halo32))) ? 1 : 0;
  const int xbox = kP * XC * 2, gbox = kP * GC * 2;
  if (p.halo) {
    p.tx_x = kHaloW * kHaloH * XC * 2;
    // the spare window of the last instruction row(s) reads up to halo row (3 + 3) * 18 + 2 + 16: keep it inside the stage
    p.x_bytes = ((((kCR - 1 + 3) * kHaloW + 2 + kTW) * XC * 2) + 1023) / 1024 * 1024;
  } else {
    p.tx_x = 9 * xbox;
    p.x_bytes = p.irows * spr * xbox;
  }
  p.Cx = Cx; p.Cout = Cout; p.ci_off = ci_off; p.Cin_tot = Cin_tot; p.dW = dW;
  p.dbias = dbias;
  const int stage_bytes = p.x_bytes + gbox;
  // FALN_WGRAD_SMEM_KB: shared memory the operand ring may take (tuning aid).  The kernel runs beside the data-gradient chain;
  // a CTA that fills the SM's shared memory keeps that chain's CTAs off the SM for as long as it runs.
  static const int ring_kb = getenv("FALN_WGRAD_SMEM_KB") ? atoi(getenv("FALN_WGRAD_SMEM_KB")) : 200;
  const int ring_use = (ring_kb_cap > 0 && ring_kb_cap < ring_kb) ? ring_kb_cap : ring_kb;
  p.stages = ((ring_use > 0 ? ring_use : 200) * 1024) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) p.stages = 2;
  p.ones_off = p.stages * stage_bytes;                       // halo path: 2 KB of bf16 ones behind the ring
  const int smem = 256 + 1024 + p.stages * stage_bytes + (p.halo && dbias ? 2048 : 0);
  // split-K: about two CTAs per SM over the whole grid, at least ~4 chunks per CTA (prologue + epilogue amortisation)
  const int nblk = p.n_cib * p.n_cob;
  // FALN_WGRAD_FILL_PCT: CTAs the split-K aims for, in percent of the SM count.  The kernel runs on the gradient side
  // stream next to the data-gradient chain, which is the critical path of the step: measured on B200 (Stage-1 step,
  // gpurun_out/s7_*) 300 % -> 5.33 ms, 200 % -> 5.07 ms, 100 % -> 4.75 ms.
  static const int fill_pct = getenv("FALN_WGRAD_FILL_PCT") ? atoi(getenv("FALN_WGRAD_FILL_PCT")) : 100;
  int splits = ((fill_pct > 0 ? fill_pct : 100) * sm_count() / 100) / nblk;
  // FALN_WGRAD_MIN_CHUNKS: least number of 64-pixel chunks per CTA.  Every split adds a full set of red.add's for the
  // channel block (9 x XC x NB floats) and one more CTA beside the data-gradient chain.  Measured on B200 (100-step runs,
  // twice each): Stage-1 step 4.125 ms at 4, 4.105 at 12, 4.087 at 16, 4.133 at 24, 4.190 at 32, 4.645 at 64 (each small-map
  // kernel alone is FASTER with more, smaller CTAs: 15.4 us at 4 vs 17.8 at 8 vs 22.7 at 16 -- it is the footprint beside the
  // critical chain that counts); Stage-2 13.46 / 13.48 / 13.46 ms at 4 / 16 / 32.
  static const int min_chunks = getenv("FALN_WGRAD_MIN_CHUNKS") ? atoi(getenv("FALN_WGRAD_MIN_CHUNKS")) : 16;
  if (splits > p.chunks / (min_chunks > 0 ? min_chunks : 16)) splits = p.chunks / (min_chunks > 0 ? min_chunks : 16);
  if (splits < 1) splits = 1;
  CUtensorMap& mx = out.mx;
  CUtensorMap& mg = out.mg;
  const bool ok_x = p.halo ? make_halo_map(&mx, x, B, H, W, Cxs, XC) : make_act_map(&mx, x, B, H, W, Cxs, XC, stride, kCR);
  if (!ok_x || !make_act_map(&mg, g, B, Hg, Wg, Cg, GC, 1, kCR)) {
    set_error("faln_conv3x3_wgrad: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  out.p = p;
  out.splits = splits;
  out.smem = smem;
  out.key = (p.halo ? 1 : 0) | (XC == 64 ? 2 : 0) | (GC == 64 ? 4 : 0);
  return FALN_OK;
}

template <int XROWB, int GROWB, bool HALO>
int launch_wgrad_multi(const WgradBatch& batch, int ctas, int smem, cudaStream_t st) {
  auto kern = conv3x3_wgrad_multi_kernel<XROWB, GROWB, HALO>;
  static int attr_set = 0;
  if (attr_set < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = smem;
  }
  launch_pdl(kern, dim3(ctas), dim3(192), (size_t)smem, st, batch);
  return after_launch("conv3x3_wgrad_multi_kernel");
}

int launch_prepared(const PreparedWgrad& w, cudaStream_t st) {
  switch (w.key) {
    case 7: return launch_wgrad<128, 128, true>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 6: return launch_wgrad<128, 128, false>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 3: return launch_wgrad<128, 64, true>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 2: return launch_wgrad<128, 64, false>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 5: return launch_wgrad<64, 128, true>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 4: return launch_wgrad<64, 128, false>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    case 1: return launch_wgrad<64, 64, true>(w.mx, w.mg, w.p, w.splits, w.smem, st);
    default: return launch_wgrad<64, 64, false>(w.mx, w.mg, w.p, w.splits, w.smem, st);
  }
}
int launch_batch(int key, const WgradBatch& b, int ctas, int smem, cudaStream_t st) {
  switch (key) {
    case 7: return launch_wgrad_multi<128, 128, true>(b, ctas, smem, st);
    case 6: return launch_wgrad_multi<128, 128, false>(b, ctas, smem, st);
    case 3: return launch_wgrad_multi<128, 64, true>(b, ctas, smem, st);
    case 2: return launch_wgrad_multi<128, 64, false>(b, ctas, smem, st);
    case 5: return launch_wgrad_multi<64, 128, true>(b, ctas, smem, st);
    case 4: return launch_wgrad_multi<64, 128, false>(b, ctas, smem, st);
    case 1: return launch_wgrad_multi<64, 64, true>(b, ctas, smem, st);
    default: return launch_wgrad_multi<64, 64, false>(b, ctas, smem, st);
  }
}
}  // namespace

// dbias [Cout] fp32 (may be NULL): += sum over (b, ho, wo) of g[b, ho, wo, co] -- the bias gradient, from the spare tap slot
extern "C" int faln_conv3x3_wgrad_bias(const void* g, const void* x, float* dW, float* dbias, int B, int H, int W, int Cg,
                                       int Cxs, int Cout, int Cx, int ci_off, int Cin_tot, int stride, unsigned flags,
                                       faln_stream_t stream) {
  PreparedWgrad w;
  const int rc = prepare_wgrad(g, x, dW, dbias, B, H, W, Cg, Cxs, Cout, Cx, ci_off, Cin_tot, stride, flags, 0, w);
  if (rc != FALN_OK) return rc;
  return launch_prepared(w, as_stream(stream));
}

// Several layers' weight (and bias) gradients in as few launches as possible: jobs with the same kernel configuration (channel
// block widths, halo / box path) share ONE grid (conv3x3_wgrad_multi_kernel); meant for the small-map layers, whose launches are
// latency chains that hold SMs beside the data-gradient chain.  Every job is what one faln_conv3x3_wgrad_bias call takes.
extern "C" int faln_conv3x3_wgrad_multi(const faln_wgrad_job_t* jobs, int njobs, faln_stream_t stream) {
  FALN_REQUIRE(jobs && njobs > 0 && njobs <= 64, "faln_conv3x3_wgrad_multi: 1..64 jobs");
  static thread_local PreparedWgrad prep[64];      // host scratch (per calling thread: forward and autograd threads may both call)
  static thread_local WgradBatch batch;
  // a batched launch keeps its CTAs short-lived and small: 100 KB of operand ring (4 stages) per CTA
  for (int i = 0; i < njobs; ++i) {
    const faln_wgrad_job_t& j = jobs[i];
    const int rc = prepare_wgrad(j.g, j.x, j.dW, j.dbias, j.B, j.H, j.W, j.Cg, j.Cxs, j.Cout, j.Cx, j.ci_off, j.Cin_tot, j.stride,
                                 j.flags, 100, prep[i]);
    if (rc != FALN_OK) return rc;
  }
  cudaStream_t st = as_stream(stream);
  bool done[64] = {false};
  for (int i = 0; i < njobs; ++i) {
    if (done[i]) continue;
    // group = the not yet launched jobs with job i's configuration, kMaxBatch at a time
    int idx[kMaxBatch], n = 0;
    for (int k = i; k < njobs && n < kMaxBatch; ++k)
      if (!done[k] && prep[k].key == prep[i].key) idx[n++] = k;
    for (int k = 0; k < n; ++k) done[idx[k]] = true;
    if (n == 1) {
      const int rc = launch_prepared(prep[i], st);
      if (rc != FALN_OK) return rc;
      continue;
    }
    // one wave over the group: chunks are dealt so that every CTA gets about total_block_chunks / SMs of them
    long long total = 0;
    for (int k = 0; k < n; ++k) total += (long long)prep[idx[k]].p.chunks * prep[idx[k]].p.n_cib * prep[idx[k]].p.n_cob;
    long long target = (total + sm_count() - 1) / sm_count();
    if (target < 8) target = 8;
    int ctas = 0, smem = 0;
    for (int k = 0; k < n; ++k) {
      PreparedWgrad& w = prep[idx[k]];
      int splits = (int)((w.p.chunks + target - 1) / target);
      if (splits < 1) splits = 1;
      if (splits > w.p.chunks) splits = w.p.chunks;
      batch.job[k].mx = w.mx;
      batch.job[k].mg = w.mg;
      batch.job[k].p = w.p;
      batch.job[k].splits = splits;
      batch.job[k].cta0 = ctas;
      ctas += splits * w.p.n_cib * w.p.n_cob;
      if (w.smem > smem) smem = w.smem;
    }
    batch.njobs = n;
    const int rc = launch_batch(prep[i].key, batch, ctas, smem, st);
    if (rc != FALN_OK) return rc;
  }
  return FALN_OK;
}

// Weight gradient of the folded deconv block (nearest 2x up-sampling + conv3x3, see conv3x3_wgrad_up2_kernel):
// g  [B,2H,2W,Cg] bf16 NHWC: gradient w.r.t. the conv's pre-activation output (on the UP-SAMPLED grid)
// x  [B,H,W,Cxs]  bf16 NHWC: the block's LOW-resolution input; Cg, Cxs multiples of 64
// dW [Cout,3,3,Cin_tot] fp32 (KRSC): columns [ci_off, ci_off + Cx) are accumulated into
namespace {
int prepare_wgrad_up2(const void* g, const void* x, float* dW, int B, int H, int W, int Cg, int Cxs, int Cout, int Cx, int ci_off,
                      int Cin_tot, PreparedWgrad& out) {
  FALN_REQUIRE(g && x && dW && B > 0 && H > 0 && W > 0, "faln_conv3x3_wgrad_up2: null pointer / bad shape");
  FALN_REQUIRE(Cg % 64 == 0 && Cxs % 64 == 0 && Cg > 0 && Cxs > 0, "faln_conv3x3_wgrad_up2: channel strides must be multiples of 64");
  FALN_REQUIRE(Cout > 0 && Cout <= Cg && Cx > 0 && Cx <= Cxs && ci_off >= 0 && ci_off + Cx <= Cin_tot,
               "faln_conv3x3_wgrad_up2: channel ranges out of bounds");
  WgradParams p{};
  p.B = B; p.Hg = H; p.Wg = W; p.stride = 1;
  p.tiles_w = (W + kTW - 1) / kTW; p.tiles_h = (H + kCR - 1) / kCR;
  p.chunks = B * p.tiles_w * p.tiles_h;
  p.NB = 64;
  p.n_cib = (Cx + 63) / 64;
  p.n_cob = (Cout + 63) / 64;
  p.irows = 8; p.halo = 1;
  p.tx_x = kHaloW * kHaloH * 128;
  p.x_bytes = (p.tx_x + 1023) / 1024 * 1024;
  p.Cx = Cx; p.Cout = Cout; p.ci_off = ci_off; p.Cin_tot = Cin_tot; p.dW = dW;
  const int stage_bytes = p.x_bytes + 4 * kP * 128;
  static const int ring_kb = getenv("FALN_WGRAD_SMEM_KB") ? atoi(getenv("FALN_WGRAD_SMEM_KB")) : 200;
  p.stages = ((ring_kb > 0 ? ring_kb : 200) * 1024) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) p.stages = 2;
  const int smem = 256 + 1024 + p.stages * stage_bytes;
  const int nblk = p.n_cib * p.n_cob;
  static const int fill_pct = getenv("FALN_WGRAD_FILL_PCT") ? atoi(getenv("FALN_WGRAD_FILL_PCT")) : 100;
  // a chunk is 4 x 16 low-resolution pixels = 256 output pixels (32 instructions): a quarter of the plain kernel's chunk floor
  static const int min_chunks = getenv("FALN_WGRAD_UP2_MIN_CHUNKS") ? atoi(getenv("FALN_WGRAD_UP2_MIN_CHUNKS")) : 4;
  int splits = ((fill_pct > 0 ? fill_pct : 100) * sm_count() / 100) / nblk;
  if (splits > p.chunks / (min_chunks > 0 ? min_chunks : 4)) splits = p.chunks / (min_chunks > 0 ? min_chunks : 4);
  if (splits < 1) splits = 1;
  CUtensorMap& mx = out.mx;
  CUtensorMap& mg = out.mg;
  if (!make_halo_map(&mx, x, B, H, W, Cxs) || !make_act_map(&mg, g, B, 2 * H, 2 * W, Cg, 64, 2, kCR)) {
    set_error("faln_conv3x3_wgrad_up2: cuTensorMapEncodeTiled failed (driver entry point missing or bad tensor geometry)");
    return FALN_ERR_LAUNCH;
  }
  out.p = p;
  out.splits = splits;
  out.smem = smem;
  out.key = 8;
  return FALN_OK;
}
}  // namespace

extern "C" int faln_conv3x3_wgrad_up2(const void* g, const void* x, float* dW, int B, int H, int W, int Cg, int Cxs, int Cout,
                                      int Cx, int ci_off, int Cin_tot, faln_stream_t stream) {
  PreparedWgrad w;
  const int rc = prepare_wgrad_up2(g, x, dW, B, H, W, Cg, Cxs, Cout, Cx, ci_off, Cin_tot, w);
  if (rc != FALN_OK) return rc;
  auto kern = conv3x3_wgrad_up2_kernel;
  static int attr_set = 0;
  if (attr_set < w.smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, w.smem);
    attr_set = w.smem;
  }
  dim3 grid(w.splits, w.p.n_cib * w.p.n_cob, 1);
  launch_pdl(kern, grid, dim3(192), (size_t)w.smem, as_stream(stream), w.mx, w.mg, w.p);
  return after_launch("conv3x3_wgrad_up2_kernel");
}

// Several folded deconv layers' weight gradients in one grid (the small levels: see faln_conv3x3_wgrad_multi).  Every job is what
// one faln_conv3x3_wgrad_up2 call takes (H, W = the LOW-resolution size; stride, flags and dbias are ignored).
extern "C" int faln_conv3x3_wgrad_up2_multi(const faln_wgrad_job_t* jobs, int njobs, faln_stream_t stream) {
  FALN_REQUIRE(jobs && njobs > 0 && njobs <= kMaxBatch, "faln_conv3x3_wgrad_up2_multi: 1..12 jobs");
  static thread_local PreparedWgrad prep[kMaxBatch];
  static thread_local WgradBatch batch;
  long long total = 0;
  for (int i = 0; i < njobs; ++i) {
    const faln_wgrad_job_t& j = jobs[i];
    const int rc = prepare_wgrad_up2(j.g, j.x, j.dW, j.B, j.H, j.W, j.Cg, j.Cxs, j.Cout, j.Cx, j.ci_off, j.Cin_tot, prep[i]);
    if (rc != FALN_OK) return rc;
    total += (long long)prep[i].p.chunks * prep[i].p.n_cib * prep[i].p.n_cob;
  }
  long long target = (total + sm_count() - 1) / sm_count();
  if (target < 4) target = 4;
  int ctas = 0, smem = 0;
  for (int i = 0; i < njobs; ++i) {
    PreparedWgrad& w = prep[i];
    int splits = (int)((w.p.chunks + target - 1) / target);
    if (splits < 1) splits = 1;
    if (splits > w.p.chunks) splits = w.p.chunks;
    batch.job[i].mx = w.mx;
    batch.job[i].mg = w.mg;
    batch.job[i].p = w.p;
    batch.job[i].splits = splits;
    batch.job[i].cta0 = ctas;
    ctas += splits * w.p.n_cib * w.p.n_cob;
    if (w.smem > smem) smem = w.smem;
  }
  batch.njobs = njobs;
  auto kern = conv3x3_wgrad_up2_multi_kernel;
  static int attr_set = 0;
  if (attr_set < smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = smem;
  }
  launch_pdl(kern, dim3(ctas), dim3(192), (size_t)smem, as_stream(stream), batch);
  return after_launch("conv3x3_wgrad_up2_multi_kernel");
}

// Weight and bias gradient of the 3 -> 32 stem (conv0.0) straight from the fp32 image (stem_wgrad_mma_kernel):
// x [B,3,H,W] fp32 NCHW, g [B,H,W,32] bf16 NHWC (pre-activation gradient), dW [32,3,3,3] fp32 KRSC and dbias [32] (or NULL)
// are accumulated into.
extern "C" int faln_stem_wgrad(const float* x, const void* g, float* dW, float* dbias, int B, int H, int W,
                               faln_stream_t stream) {
  FALN_REQUIRE(x && g && dW && B > 0 && H > 0 && W > 0, "faln_stem_wgrad: bad argument");
  FALN_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "faln_stem_wgrad: g must be 16-byte aligned");
  FALN_REQUIRE((long long)B * 3 * H * W < (1LL << 31), "faln_stem_wgrad: tensor too large");
  const int tiles_w = (W + 127) / 128;
  const long long total = (long long)B * H * tiles_w;
  const int smem = 1024 + 2 * 128 * 128 + 128 * 64 + 64;
  static const int per_sm = getenv("FALN_STEM_WGRAD_CTAS") ? atoi(getenv("FALN_STEM_WGRAD_CTAS")) : 3;
  long long grid = (long long)sm_count() * (per_sm > 0 ? per_sm : 3);
  if (grid > total) grid = total;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(stem_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  stem_wgrad_mma_kernel<<<(int)grid, 128, smem, as_stream(stream)>>>(x, static_cast<const __nv_bfloat16*>(g), dW, dbias, B, H, W,
                                                                     tiles_w, total);
  return after_launch("stem_wgrad_mma_kernel");
}

extern "C" int faln_conv3x3_wgrad(const void* g, const void* x, float* dW, int B, int H, int W, int Cg, int Cxs, int Cout,
                                  int Cx, int ci_off, int Cin_tot, int stride, unsigned flags, faln_stream_t stream) {
  return faln_conv3x3_wgrad_bias(g, x, dW, nullptr, B, H, W, Cg, Cxs, Cout, Cx, ci_off, Cin_tot, stride, flags, stream);
}

// out [B,3,3,C] fp32 += per-sample sums of g [B,H,W,Cs] (bf16 NHWC, channels [0,C)) over the nine border classes
// (first / interior / last row x first / interior / last column).  Used for the weight gradient of a spatially constant
// input channel (the max_disp/100 plane of /root/reference/models/FAL_netB.py:145,208-209).  Needs H, W >= 2.
extern "C" int faln_border_sum_nhwc(const void* g, float* out, int B, int H, int W, int C, int Cs, faln_stream_t stream) {
  FALN_REQUIRE(g && out && B > 0 && H >= 2 && W >= 2 && C > 0 && Cs >= C, "faln_border_sum_nhwc: bad arguments");
  border_sum_kernel<<<B * H, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(g), out, B, H, W, C, Cs);
  return after_launch("border_sum_kernel");
}
