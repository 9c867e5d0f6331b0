// Shared device/host helpers for the X3D sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/x3d_b200.h"

namespace x3d {

void set_error(const char* fmt, ...);          // host_util.cpp
int check_launch(const char* what);            // host_util.cpp: cudaGetLastError -> status

#define X3D_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::x3d::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

using bf16 = __nv_bfloat16;

// cudaFuncSetAttribute is a per-DEVICE setting: the opt-in shared-memory size a launcher has
// configured is remembered per (kernel instance, device), so a second GPU driven from the same
// process gets its own call (the C ABI allows one context per GPU in one process).
constexpr int kMaxDevices = 64;
struct SmemOptIn {
  size_t configured[kMaxDevices] = {};
};
template <typename Kern>
inline cudaError_t ensure_dynamic_smem(Kern kern, SmemOptIn& state, size_t bytes, bool max_carveout = true) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool tracked = dev >= 0 && dev < kMaxDevices;
  if (tracked && bytes <= state.configured[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return e;
  if (max_carveout)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (tracked) state.configured[dev] = bytes;
  return cudaSuccess;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A kernel launched through launch_pdl may start while its predecessor in the stream is still
// running: its CTAs become resident as the predecessor's retire, run their set-up (barriers, TMEM,
// tensor-map prefetch, weight loads -- nothing the predecessor writes), and block in pdl_wait()
// until the predecessor grid has completed and its memory is visible.  pdl_trigger() lets the NEXT
// kernel's CTAs be scheduled.  Rule for every kernel launched this way: no access to activations
// (reads of inputs, writes of outputs) before pdl_wait().  X3D_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // host_util.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- two-channel (the unit a channelwise thread owns) and vector loads/stores ----
__device__ __forceinline__ float2 ld2(const float* p) {
  return __ldg(reinterpret_cast<const float2*>(p));
}
__device__ __forceinline__ float2 ld2(const bf16* p) {
  const uint32_t u = __ldg(reinterpret_cast<const unsigned int*>(p));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(bf16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __float22bfloat162_rn(v);
}

__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const bf16* p) {
  return __uint_as_float(static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}

// 8 consecutive channels -> 8 floats
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}

// packed fp32 FMA (Blackwell FFMA2): two independent fp32 FMAs in one issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#if defined(X3D_NO_FFMA2)
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#else
  return __ffma2_rn(a, b, c);
#endif
}

template <bool kFast>
__device__ __forceinline__ float sigmoidf_(float x) {
  if (kFast) return __fdividef(1.f, 1.f + __expf(-x));
  return 1.f / (1.f + expf(-x));
}

}  // namespace x3d
