// Shared host-side helpers: error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/interactron_b200.h"

namespace itn {

int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(ITN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return ITN_OK;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute
// and starts with pdl_wait() (prerequisite grids complete + their writes visible) followed by
// pdl_trigger() (the next grid in the stream may be scheduled: its CTAs become resident on idle
// SMs and run their own prologue up to pdl_wait()).  A step is ~1300 short launches; this hides
// the grid-to-grid launch latency and the GEMM prologue (barrier init, TMEM allocation, tensor-map
// fetch) behind the tail of the previous kernel.  Opt-in with ITN_PDL=1 (without the attribute the
// device instructions are no-ops); see pdl_enabled() for the measurement that keeps it off.
bool pdl_enabled();

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline void launch_cluster(int cluster, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                           cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster > 1) {          // CTA pairs (cta_group::2 kernels): consecutive CTAs share a TPC
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                   Args&&... args) {
  launch_cluster(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

#define ITN_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) return itn::set_error(ITN_ERR_ARG, __VA_ARGS__); \
  } while (0)

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

// Round-to-nearest to TF32 (10 mantissa bits); the result is still an fp32 bit pattern.
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// The residual operand of the tf32x3 products: x - trunc_tf32(x) (exact in fp32, 13 significant bits; the tensor
// core truncates it to 11).  ITN_LO_RN=1 rounds it to nearest TF32 first so that the truncation is exact.  Measured
// on B200 (profiles/README.md): fused-attention outputs 2.0e-6 -> 1.1e-6 vs fp64 and small-K GEMMs 6.3e-7 -> 4.7e-7,
// but the two extra integer ops per element cost the attention kernels 6 % and the GEMM splitter warps (the
// critical path of the tf32x3 main loop) 16 %: off by default, every parity bound holds either way.
#ifndef ITN_LO_RN
#define ITN_LO_RN 0
#endif
__device__ __forceinline__ float tf32_lo_rn(float lo) {
#if ITN_LO_RN
  // round-half-away to 10 mantissa bits with two full-rate integer ops (cvt.rna.tf32.f32 gives the same bits but
  // issues at 1/8 rate: it cost the tf32x3 main loop 35 % when the splitter warps used it)
  return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
#else
  return lo;
#endif
}
__device__ __forceinline__ float tf32_lo_exact(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_lo(float x) { return tf32_lo_rn(tf32_lo_exact(x)); }

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

}  // namespace itn
