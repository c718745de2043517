// input_pipe.cu -- the training input pipeline on the device.
//
// Replaces, per stereo pair, the CPU work of /root/reference/data_transforms.py:46-157 (RandomResizeCrop -- PIL bicubic
// resize of the WHOLE image, then a crop --, RandomHorizontalFlip, RandomGamma, RandomBrightness, RandomCBrightness) and
// of the input transform at /root/reference/Train_Stage1_K.py:124-128 (ArrayToTensor, /255, - mean): from the decoded
// uint8 HWC image to the normalised fp32 NCHW crop in two batched launches, computing only the pixels of the crop.
//
// This is byte / integer work and bit-exact with the reference:
//   * the resize is Pillow's two-pass 8-bit resampler (horizontal, then vertical, 22-bit fixed-point coefficients,
//     round-half-up accumulators, clamp to uint8 after EACH pass); the coefficient tables are built on the host in
//     double precision by `faln_pil_bicubic_coeffs` exactly as Pillow does, so the kernels only do integer MACs;
//   * gamma / brightness / per-channel brightness / normalisation act on a uint8 value, i.e. they are a 3 x 256 table
//     per image, built on the host with the reference's own numpy expressions (fal_net_b200/input_pipeline.py) --
//     including their float64 -> float32 rounding -- and looked up here.
#include <math.h>

#include "common.cuh"

namespace faln {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift, like Pillow's table lookup on (in >> PRECISION_BITS)
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: source rows [row0, row0+rows) x crop columns -> inter [img][rows_max][tw][3] (uint8)
__global__ void __launch_bounds__(128) aug_resample_h_kernel(const faln_aug_desc* __restrict__ descs,
                                                             const int* __restrict__ tabs, unsigned char* __restrict__ inter,
                                                             long long inter_stride, int tw) {
  const faln_aug_desc d = descs[blockIdx.z];
  const int r = blockIdx.y;
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (r >= d.rows || x >= tw) return;
  const int* bounds = tabs + d.x_tab;
  const int xmin = bounds[2 * x], cnt = bounds[2 * x + 1];
  const int* k = bounds + 2 * tw + (long long)x * d.ksx;
  const unsigned char* src = static_cast<const unsigned char*>(d.src) + ((long long)(d.row0 + r) * d.W + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < cnt; ++t) {
    const int c = k[t];
    s0 += (int)src[3 * t + 0] * c;
    s1 += (int)src[3 * t + 1] * c;
    s2 += (int)src[3 * t + 2] * c;
  }
  unsigned char* o = inter + blockIdx.z * inter_stride + ((long long)r * tw + x) * 3;
  o[0] = clip8(s0);
  o[1] = clip8(s1);
  o[2] = clip8(s2);
}

// vertical pass + value table + optional mirror: inter -> out [n,3,th,tw] fp32
__global__ void __launch_bounds__(128) aug_resample_v_kernel(const faln_aug_desc* __restrict__ descs,
                                                             const int* __restrict__ tabs,
                                                             const unsigned char* __restrict__ inter, long long inter_stride,
                                                             const float* __restrict__ luts, float* __restrict__ out, int th,
                                                             int tw) {
  const faln_aug_desc d = descs[blockIdx.z];
  const int y = blockIdx.y;
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= tw) return;
  const int* bounds = tabs + d.y_tab;
  const int ymin = bounds[2 * y] - d.row0, cnt = bounds[2 * y + 1];
  const int* k = bounds + 2 * th + (long long)y * d.ksy;
  const unsigned char* src = inter + blockIdx.z * inter_stride + ((long long)ymin * tw + x) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < cnt; ++t) {
    const int c = k[t];
    const unsigned char* p = src + (long long)t * tw * 3;
    s0 += (int)p[0] * c;
    s1 += (int)p[1] * c;
    s2 += (int)p[2] * c;
  }
  const float* lut = luts + d.lut;
  const int xo = d.flip ? tw - 1 - x : x;
  float* o = out + ((long long)d.dst * 3 * th + y) * tw + xo;
  const long long plane = (long long)th * tw;
  o[0] = __ldg(lut + clip8(s0));
  o[plane] = __ldg(lut + 256 + clip8(s1));
  o[2 * plane] = __ldg(lut + 512 + clip8(s2));
}

inline double bicubic_filter(double x) {
  const double a = -0.5;  // Pillow's bicubic
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

}  // namespace
}  // namespace faln

using namespace faln;

extern "C" int faln_pil_bicubic_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = (double)((float)in_size - 0.0f) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

// HOST function (no CUDA call): Pillow's precompute_coeffs + normalize_coeffs_8bpc for output pixels
// [out_lo, out_lo + out_n) of an in_size -> out_size bicubic resize over the whole input (box = [0, in_size)).
// bounds [2*out_n] = (first input index, tap count); coeffs [out_n * ksize] fixed point (22 fractional bits).
extern "C" int faln_pil_bicubic_coeffs(int in_size, int out_size, int out_lo, int out_n, int* bounds, int* coeffs) {
  FALN_REQUIRE(in_size > 0 && out_size > 0 && out_lo >= 0 && out_n > 0 && out_lo + out_n <= out_size && bounds && coeffs,
               "faln_pil_bicubic_coeffs: bad argument");
  const float in0 = 0.0f, in1 = (float)in_size;
  double scale = (double)(in1 - in0) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  double kbuf[4096];
  FALN_REQUIRE(ksize <= 4096, "faln_pil_bicubic_coeffs: kernel too long");
  for (int i = 0; i < out_n; ++i) {
    const int xx = out_lo + i;
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      kbuf[x] = w;
      ww += w;
    }
    int* k = coeffs + (long long)i * ksize;
    for (int x = 0; x < ksize; ++x) {
      double v = 0.0;
      if (x < xmax) v = ww != 0.0 ? kbuf[x] / ww : kbuf[x];
      k[x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[2 * i] = xmin;
    bounds[2 * i + 1] = xmax;
  }
  return FALN_OK;
}

extern "C" int faln_augment_crops_u8(const faln_aug_desc* descs, int n_img, const int* tabs, const float* luts,
                                     unsigned char* inter, long long inter_stride, int max_rows, float* out, int th, int tw,
                                     faln_stream_t stream) {
  FALN_REQUIRE(descs && tabs && luts && inter && out && n_img > 0 && n_img <= 65535 && th > 0 && th <= 65535 && tw > 0 &&
                   max_rows > 0 && max_rows <= 65535 && inter_stride >= (long long)max_rows * tw * 3,
               "faln_augment_crops_u8: bad argument");
  const dim3 gh((tw + 127) / 128, max_rows, n_img);
  aug_resample_h_kernel<<<gh, 128, 0, as_stream(stream)>>>(descs, tabs, inter, inter_stride, tw);
  int rc = after_launch("aug_resample_h_kernel");
  if (rc != FALN_OK) return rc;
  const dim3 gv((tw + 127) / 128, th, n_img);
  aug_resample_v_kernel<<<gv, 128, 0, as_stream(stream)>>>(descs, tabs, inter, inter_stride, luts, out, th, tw);
  return after_launch("aug_resample_v_kernel");
}
