// common.cuh -- shared host/device helpers of libfalnet_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/falnet_b200.h"

namespace faln {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// Records a kernel launch and converts a launch error to a return code.
int after_launch(const char* what);

inline cudaStream_t as_stream(faln_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define FALN_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::faln::set_error(__VA_ARGS__);      \
      return FALN_ERR_ARG;                 \
    }                                      \
  } while (0)

int sm_count();

// Programmatic dependent launch (FALN_PDL=1; default OFF): a kernel launched with `launch_pdl` may begin -- barrier init,
// TMEM allocation, tensor-map prefetch -- while the previous kernel on the stream is still draining; it calls `pdl_wait()`
// before it touches anything the previous kernel wrote, and `pdl_launch_dependents()` right after, so the NEXT kernel's CTAs
// are admitted as soon as SM resources free up.  In a captured CUDA graph these become programmatic edges.
// Measured on B200 (round 2, 100-step runs, A/B/A/B on one box): Stage-1 step 4.54 ms off vs 4.63 ms on, Stage-2 14.92 vs
// 15.02 ms, Test flip-PP 7.55 vs 7.52 ms -- the early-admitted CTAs of the parameter-gradient kernels on the side stream
// take shared memory / TMEM away from the data-gradient chain, which is the critical path, and the graph's plain edges were
// already cheap.  Parity is unaffected (93 conv / model / MED tests green with it on).  Kept as a switch, off by default.
bool pdl_enabled();

// ---------------------------------------------------------------------------------------------
// device-side PTX helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA unit, SASS UBLKCP), completion on an mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// consumer-only named barrier (id 1..15), nthreads a multiple of 32
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// host: launch `kern` with the programmatic-stream-serialization attribute (when enabled)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

}  // namespace faln
