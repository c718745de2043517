// misc.cu -- fused multi-tensor Adam over a flat parameter arena, and layout conversion kernels
// between the reference's fp32 NCHW tensors and the bf16 NHWC activations of the conv family.
#include "common.cuh"

namespace faln {
namespace {

// torch.optim.Adam (no amsgrad), as configured at /root/reference/Train_Stage1_K.py:177-181.
// One pass over the arena: reads p, g, m, v; writes p, m, v and the bf16 shadow weights.
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                   float4* __restrict__ m, float4* __restrict__ v,
                                                   __nv_bfloat162* __restrict__ w16, long long n4, float step_size,
                                                   float beta1, float beta2, float inv_sqrt_bc2, float eps, float wd,
                                                   float gscale, const float* __restrict__ hp) {
  if (hp) {  // CUDA-graph mode: step size and bias correction live on the device (written by adam_tick_kernel)
    step_size = hp[2];
    inv_sqrt_bc2 = hp[3];
  }
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
    float* pp = reinterpret_cast<float*>(&P);
    float* gg = reinterpret_cast<float*>(&G);
    float* mm = reinterpret_cast<float*>(&M);
    float* vv = reinterpret_cast<float*>(&V);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = fmaf(wd, pp[k], gg[k] * gscale);
      mm[k] = fmaf(beta1, mm[k], (1.f - beta1) * gr);
      vv[k] = fmaf(beta2, vv[k], (1.f - beta2) * gr * gr);
      float denom = fmaf(sqrtf(vv[k]), inv_sqrt_bc2, eps);
      pp[k] = pp[k] - step_size * (mm[k] / denom);
    }
    p[i] = P;
    m[i] = M;
    v[i] = V;
    if (w16) {
      w16[2 * i] = __floats2bfloat162_rn(pp[0], pp[1]);
      w16[2 * i + 1] = __floats2bfloat162_rn(pp[2], pp[3]);
    }
  }
}

// hp = {lr, step (as float), lr / (1 - beta1^step), 1 / sqrt(1 - beta2^step)}: advances the step counter and refreshes
// the derived factors on the device, so a captured CUDA graph replays a correct Adam update at every step.
__global__ void adam_tick_kernel(float* hp, float beta1, float beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double t = (double)hp[1] + 1.0;
    hp[1] = (float)t;
    hp[2] = (float)((double)hp[0] / (1.0 - pow((double)beta1, t)));
    hp[3] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, t)));
  }
}

// fp32 NCHW -> bf16 NHWC, one thread per pixel (C small: the 3-channel input image).
__global__ void __launch_bounds__(256) nchw_to_nhwc_small(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                          int B, int C, int H, int W, int Cp, int flip_x) {
  const long long npx = (long long)B * H * W;
  const long long hw = (long long)H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < npx; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const long long b = i / hw;
    const long long rem = i % hw;
    const long long so = flip_x ? rem - x + (W - 1 - x) : rem;
    __nv_bfloat16* d = dst + i * Cp;
    for (int c = 0; c < Cp; ++c) d[c] = __float2bfloat16(c < C ? __ldg(src + (b * C + c) * hw + so) : 0.f);
  }
}

// Tiled transposes between fp32 planar [B,C,H,pitch] and bf16 NHWC [B,H,W,Cp]: a block owns 32 pixels
// of one row and walks the channels in groups of 32 through a padded shared tile, so both the planar
// side (x fastest) and the NHWC side (c fastest) are accessed in full 128-byte / 64-byte segments.
__global__ void __launch_bounds__(256) planar_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                             int C, int H, int W, int Cp, long long pitch) {
  __shared__ float tile[32][33];
  const int xt = blockIdx.x * 32;
  const int y = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  for (int c0 = 0; c0 < Cp; c0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, x = xt + tx;
      tile[ty + 8 * k][tx] = (c < C && x < W) ? __ldg(src + (((long long)b * C + c) * H + y) * pitch + x) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = ty + 8 * k, c = c0 + tx, x = xt + px;
      if (x < W && c < Cp) dst[(((long long)b * H + y) * W + x) * Cp + c] = __float2bfloat16(tile[tx][px]);
    }
    __syncthreads();
  }
}

// Wide variant for Cp <= 64, 16-byte aligned planar rows: a block owns 128 pixels of one row and all channels.  Every
// thread issues its 8 (Cp = 64) 128-bit loads before the one barrier, then writes 16-byte chunks of 8 bf16 channels; a
// warp's store covers 4 pixels x 128 bytes contiguously.  (The 32-pixel kernel above keeps 4 scalar loads in flight per
// thread and synchronises twice per 32 channels: 93 us for the 8x49x192x640 gradient, against ~55 us of DRAM time.)
template <int CP>
__global__ void __launch_bounds__(256) planar_to_nhwc_wide_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                  int C, int H, int W, long long pitch) {
  constexpr int kPX = 128, kStride = kPX + 4;            // +4 floats: 16-byte aligned rows, chunk reads 2-way conflicted at most
  __shared__ __align__(16) float tile[CP * kStride];
  const int xt = blockIdx.x * kPX;
  const int y = blockIdx.y, b = blockIdx.z;
  constexpr int kLoads = CP * (kPX / 4) / 256;           // float4 loads per thread
  float4 v[kLoads];
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    const int i = threadIdx.x + 256 * k;
    const int c = i / (kPX / 4), x = xt + 4 * (i % (kPX / 4));
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C && x < W) {
      const float* sp = src + (((long long)b * C + c) * H + y) * pitch + x;
      if (x + 3 < W) {
        v[k] = __ldg(reinterpret_cast<const float4*>(sp));
      } else {
        v[k].x = __ldg(sp);
        if (x + 1 < W) v[k].y = __ldg(sp + 1);
        if (x + 2 < W) v[k].z = __ldg(sp + 2);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kLoads; ++k) {
    const int i = threadIdx.x + 256 * k;
    *reinterpret_cast<float4*>(tile + (i / (kPX / 4)) * kStride + 4 * (i % (kPX / 4))) = v[k];
  }
  __syncthreads();
  constexpr int kChunks = CP / 8;                         // 16-byte chunks per pixel
#pragma unroll
  for (int k = 0; k < kPX * kChunks / 256; ++k) {
    const int i = threadIdx.x + 256 * k;
    const int ch = i % kChunks, px = i / kChunks, x = xt + px;
    if (x >= W) continue;
    uint4 o;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      h2[e] = __floats2bfloat162_rn(tile[(ch * 8 + 2 * e) * kStride + px], tile[(ch * 8 + 2 * e + 1) * kStride + px]);
    *reinterpret_cast<uint4*>(dst + (((long long)b * H + y) * W + x) * CP + ch * 8) = o;
  }
}

__global__ void __launch_bounds__(256) nhwc_to_planar_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                                             int C, int H, int W, int Cp, long long pitch) {
  __shared__ float tile[32][33];
  const int xt = blockIdx.x * 32;
  const int y = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = ty + 8 * k, c = c0 + tx, x = xt + px;
      tile[px][tx] = (x < W && c < Cp) ? __bfloat162float(src[(((long long)b * H + y) * W + x) * Cp + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k, x = xt + tx;
      if (c < C && x < W) dst[(((long long)b * C + c) * H + y) * pitch + x] = tile[tx][ty + 8 * k];
    }
    __syncthreads();
  }
}


// Batched re-packing of the bf16 shadow weights for the data-gradient kernels: for every job (one 3x3 conv weight)
//   src [Cout][9][Cin_tot]  (KRSC, the layout of the parameter arena)  ->  dst [Cin_used][9][Cout]  (one row per INPUT channel)
// i.e. nine [Cout x Cin_used] -> [Cin_used x Cout] transposes through a padded 32x32 shared tile.  One launch per step
// instead of ~4 ATen launches per layer.  jobs[j] = {src_off, dst_off, Cout, Cin_tot, Cin_used, tiles} (elements).
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const __nv_bfloat16* __restrict__ w16, __nv_bfloat16* __restrict__ wd16,
                                                         const long long* __restrict__ jobs) {
  __shared__ __nv_bfloat16 tile[32][34];
  const long long* jb = jobs + (size_t)blockIdx.y * 6;
  const int Cout = (int)jb[2], Cin_tot = (int)jb[3], Cin_used = (int)jb[4];
  const int tco = Cout / 32, tci = Cin_used / 32;
  const int ntiles = 9 * tco * tci;
  if ((int)blockIdx.x >= ntiles) return;
  const int t = blockIdx.x / (tco * tci), r = blockIdx.x % (tco * tci);
  const int co0 = (r / tci) * 32, ci0 = (r % tci) * 32;
  const __nv_bfloat16* src = w16 + jb[0];
  __nv_bfloat16* dst = wd16 + jb[1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co0 + ty + 8 * k;
    tile[ty + 8 * k][tx] = src[((size_t)co * 9 + t) * Cin_tot + ci0 + tx];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + ty + 8 * k;
    dst[((size_t)ci * 9 + t) * Cout + co0 + tx] = tile[tx][ty + 8 * k];
  }
}

// The same re-pack over a FLAT tile list: jobs[j][5] holds the index of job j's first 32 x 32 tile in the concatenation of all
// jobs' tiles, `total` their number.  A warp owns one tile per turn (no CTA barrier), moves bf16 PAIRS (4-byte accesses, 64-byte
// row segments) and the grid holds exactly the work: the 2-D launch above starts max_tiles x njobs = 76 k blocks for 16.5 k tiles
// of work (most blocks of the small layers exit at once) and took 55 us per step on the critical path right behind Adam.
__global__ void __launch_bounds__(256) pack_dgrad_flat_kernel(const __nv_bfloat16* __restrict__ w16, __nv_bfloat16* __restrict__ wd16,
                                                              const long long* __restrict__ jobs, int njobs, int total) {
  __shared__ uint32_t tile[8][32][17];                      // per warp: 32 output-channel rows x 16 input-channel pairs (+1 pad)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, col = lane & 15;
  for (int ft = blockIdx.x * 8 + warp; ft < total; ft += gridDim.x * 8) {
    // the job whose tile range contains ft: the last one whose first tile is <= ft (prefix starts are ascending)
    int j = 0;
    for (int base = 0; base < njobs; base += 32) {
      const int cand = base + lane;
      const bool le = cand < njobs && (int)jobs[(size_t)cand * 6 + 5] <= ft;
      const unsigned m = __ballot_sync(0xffffffffu, le);
      if (m) j = base + 31 - __clz(m);
    }
    const long long* jb = jobs + (size_t)j * 6;
    const int Cout = (int)jb[2], Cin_tot = (int)jb[3], Cin_used = (int)jb[4];
    const int tci = Cin_used / 32, per_tap = (Cout / 32) * tci;
    const int lt = ft - (int)jb[5];
    const int t = lt / per_tap, r = lt % per_tap;
    const int co0 = (r / tci) * 32, ci0 = (r % tci) * 32;
    const __nv_bfloat16* src = w16 + jb[0];
    __nv_bfloat16* dst = wd16 + jb[1];
    if ((Cin_tot & 1) == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int co = 2 * i + half;
        tile[warp][co][col] = *reinterpret_cast<const uint32_t*>(src + ((size_t)(co0 + co) * 9 + t) * Cin_tot + ci0 + 2 * col);
      }
    } else {                                                // odd row length (the 33-channel conv1.0): rows are not 4-byte aligned
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int co = 2 * i + half;
        const __nv_bfloat16* q = src + ((size_t)(co0 + co) * 9 + t) * Cin_tot + ci0 + 2 * col;
        const uint32_t lo = *reinterpret_cast<const unsigned short*>(q), hi = *reinterpret_cast<const unsigned short*>(q + 1);
        tile[warp][co][col] = lo | (hi << 16);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int ci = 2 * i + half;                          // output row; the pair (co = 2 col, 2 col + 1) of it
      const uint32_t a = tile[warp][2 * col][i], b = tile[warp][2 * col + 1][i];
      const uint32_t v = half ? __byte_perm(a, b, 0x7632) : __byte_perm(a, b, 0x5410);
      *reinterpret_cast<uint32_t*>(dst + ((size_t)(ci0 + ci) * 9 + t) * Cout + co0 + 2 * col) = v;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float4* __restrict__ src, __nv_bfloat162* __restrict__ dst,
                                                          long long n4) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = src[i];
    dst[2 * i] = __floats2bfloat162_rn(v.x, v.y);
    dst[2 * i + 1] = __floats2bfloat162_rn(v.z, v.w);
  }
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_adam(float* p, const float* g, float* m, float* v, void* w16, long long n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, float grad_scale, faln_stream_t stream) {
  FALN_REQUIRE(p && g && m && v && n > 0 && (n & 3) == 0 && step >= 1, "faln_adam: n must be a positive multiple of 4");
  FALN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "faln_adam: arenas must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  adam_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), static_cast<__nv_bfloat162*>(w16), n4, (float)(lr / bc1), beta1, beta2,
      (float)(1.0 / sqrt(bc2)), eps, weight_decay, grad_scale, nullptr);
  return after_launch("adam_kernel");
}

extern "C" int faln_adam_dev_range(float* p, const float* g, float* m, float* v, void* w16, long long n, float* hp,
                                   float beta1, float beta2, float eps, float weight_decay, float grad_scale, int tick,
                                   faln_stream_t stream) {
  FALN_REQUIRE(p && g && m && v && hp && n > 0 && (n & 3) == 0, "faln_adam_dev: n must be a positive multiple of 4");
  FALN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "faln_adam_dev: arenas must be 16-byte aligned");
  FALN_REQUIRE(w16 == nullptr || (reinterpret_cast<uintptr_t>(w16) & 7) == 0, "faln_adam_dev: w16 must be 8-byte aligned");
  if (tick) adam_tick_kernel<<<1, 32, 0, as_stream(stream)>>>(hp, beta1, beta2);
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  adam_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), static_cast<__nv_bfloat162*>(w16), n4, 0.f, beta1, beta2, 0.f, eps, weight_decay,
      grad_scale, hp);
  return after_launch("adam_kernel");
}

extern "C" int faln_adam_dev(float* p, const float* g, float* m, float* v, void* w16, long long n, float* hp, float beta1,
                             float beta2, float eps, float weight_decay, float grad_scale, faln_stream_t stream) {
  return faln_adam_dev_range(p, g, m, v, w16, n, hp, beta1, beta2, eps, weight_decay, grad_scale, 1, stream);
}

extern "C" int faln_nchw_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int Cp, int flip_x,
                                      faln_stream_t stream) {
  FALN_REQUIRE(src && dst && B > 0 && C > 0 && Cp >= C && Cp <= 16, "faln_nchw_to_nhwc_bf16: need C <= Cp <= 16");
  const long long npx = (long long)B * H * W;
  long long grid = (npx + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  nchw_to_nhwc_small<<<(int)grid, 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), B, C, H, W, Cp,
                                                               flip_x);
  return after_launch("nchw_to_nhwc_small");
}

extern "C" int faln_planar_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int Cp, long long pitch,
                                        faln_stream_t stream) {
  FALN_REQUIRE(src && dst && B > 0 && C > 0 && Cp >= C && pitch >= W && H <= 65535 && B <= 65535,
               "faln_planar_to_nhwc_bf16: bad argument");
  if ((Cp == 64 || Cp == 32) && pitch % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    dim3 gridw((W + 127) / 128, H, B);
    if (Cp == 64)
      planar_to_nhwc_wide_kernel<64><<<gridw, 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), C, H, W, pitch);
    else
      planar_to_nhwc_wide_kernel<32><<<gridw, 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), C, H, W, pitch);
    return after_launch("planar_to_nhwc_wide_kernel");
  }
  dim3 grid((W + 31) / 32, H, B);
  planar_to_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), C, H, W, Cp, pitch);
  return after_launch("planar_to_nhwc_kernel");
}

extern "C" int faln_nhwc_bf16_to_planar(const void* src, float* dst, int B, int C, int H, int W, int Cp, long long pitch,
                                        faln_stream_t stream) {
  FALN_REQUIRE(src && dst && B > 0 && C > 0 && Cp >= C && pitch >= W && H <= 65535 && B <= 65535,
               "faln_nhwc_bf16_to_planar: bad argument");
  dim3 grid((W + 31) / 32, H, B);
  nhwc_to_planar_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(src), dst, C, H, W, Cp,
                                                             pitch);
  return after_launch("nhwc_to_planar_kernel");
}

extern "C" int faln_pack_dgrad_batched(const void* w16, void* wd16, const long long* jobs, int njobs, int max_tiles,
                                       faln_stream_t stream) {
  FALN_REQUIRE(w16 && wd16 && jobs && njobs > 0 && max_tiles > 0 && njobs <= 65535, "faln_pack_dgrad_batched: bad arguments");
  dim3 grid(max_tiles, njobs);
  pack_dgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(w16),
                                                          static_cast<__nv_bfloat16*>(wd16), jobs);
  return after_launch("pack_dgrad_kernel");
}

extern "C" int faln_pack_dgrad_flat(const void* w16, void* wd16, const long long* jobs, int njobs, int total_tiles,
                                    faln_stream_t stream) {
  FALN_REQUIRE(w16 && wd16 && jobs && njobs > 0 && total_tiles > 0, "faln_pack_dgrad_flat: bad arguments");
  long long grid = (total_tiles + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  pack_dgrad_flat_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(w16),
                                                                   static_cast<__nv_bfloat16*>(wd16), jobs, njobs, total_tiles);
  return after_launch("pack_dgrad_flat_kernel");
}

extern "C" int faln_f32_to_bf16(const float* src, void* dst, long long n, faln_stream_t stream) {
  FALN_REQUIRE(src && dst && n > 0 && (n & 3) == 0, "faln_f32_to_bf16: n must be a positive multiple of 4");
  FALN_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
               "faln_f32_to_bf16: buffers must be 16-byte aligned");
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  f32_to_bf16_kernel<<<(int)grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src),
                                                               static_cast<__nv_bfloat162*>(dst), n4);
  return after_launch("f32_to_bf16_kernel");
}
