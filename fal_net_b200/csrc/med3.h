// med3.h -- host-side interface between the C-ABI entry points (med.cu) and the third-generation MED kernels (med3.cu).
#pragma once
#include <cuda_runtime.h>

namespace faln {
namespace m3 {

struct M3Params {
  const float* logits;
  const float* image;
  const float* g0x;
  const float* x_of;
  const float* d_lvl;
  float* pan;
  float* disp;
  float* maskL;
  float* maskR;
  float* lse0;
  float* lsew;
  // backward only
  const float* pan_in;
  const float* disp_in;
  const float* lse0_in;
  const float* lsew_in;
  const float* g_pan;
  const float* g_disp;
  float* g_logits;
  long long g_pitch;
  long long pitch;   // logits row pitch, elements (multiple of 4)
  int B, N, H, W;
  int S, G;          // ring: S groups of G plane rows (filled in by the launcher)
  int nbuf;          // per-row staging buffers (1 or 2; filled in by the launcher)
  int force_generic; // every plane on the per-pixel generic code (testing)
};

// Return 1 when the kernel was launched, 0 when the shape does not fit (caller uses the second-generation kernel),
// < 0 on a launch error.
int med3_launch_fwd(M3Params p, bool masks, cudaStream_t stream);
int med3_launch_bwd(M3Params p, cudaStream_t stream);

}  // namespace m3
}  // namespace faln
