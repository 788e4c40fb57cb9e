// Shared device/host helpers for the starcop_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/starcop_b200.h"

namespace sc {

extern thread_local cudaError_t g_last_error;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_error = e;
    return SC_ERR_CUDA;
  }
  return SC_OK;
}

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL)
// Every kernel of the train / eval step is launched with cudaLaunchAttributeProgrammaticStreamSerialization and opens with
// pdl_wait(): griddepcontrol.wait blocks until the preceding kernel of the stream has COMPLETED and flushed its writes, so
// the data dependencies are exactly those of a plain stream launch, while the launch itself (grid setup, CTA scheduling,
// parameter / descriptor fetch) overlaps the predecessor's tail; launch_dependents right after it lets the successor do the
// same once all of this kernel's CTAs are running.  The step is ~580 launches, most of them 5-15 us long, so the
// per-boundary bubble matters.  STARCOP_PDL=0 falls back to plain launches (A/B measurement).
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SC_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// for tiny single-wave kernels only: lets the successor's CTAs be dispatched while this kernel runs (they block in
// their own pdl_wait until this grid has completed).  Measured harmful as a blanket policy, see DESIGN.md.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // elementwise.cu: reads STARCOP_PDL once

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through check_launch()
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Division of a non-negative 31-bit integer by a run-time constant as multiply-high + add + shift (Granlund &
// Montgomery round-up method: q = (mulhi(m, x) + x) >> s with s = ceil(log2 d), m = floor(2^32 (2^s - d) / d) + 1; the
// sum cannot overflow for x < 2^31).  A hardware-less 32-bit division is ~20 instructions; the persistent tile loops
// decode (image, tile row, tile column) from the tile index once per tile per thread.
struct FastDiv {
  uint32_t m, s, d;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.s = 0;
  while ((1ull << f.s) < d) ++f.s;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << f.s) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t x, const FastDiv& f) { return (__umulhi(f.m, x) + x) >> f.s; }

// ---- 8-wide channel vectors: the unit every NHWC bandwidth kernel moves (16 B bf16 / 32 B f32)
struct f8 {
  float v[8];
};

template <typename T>
__device__ __forceinline__ f8 load8(const T* p);
template <>
__device__ __forceinline__ f8 load8<float>(const float* p) {
  f8 r;
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <>
__device__ __forceinline__ f8 load8<__nv_bfloat16>(const __nv_bfloat16* p) {
  f8 r;
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const f8& r);
template <>
__device__ __forceinline__ void store8<float>(float* p, const f8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const f8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SC_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SC_ACT_RELU6) return v < 0.f ? 0.f : (v > 6.f ? 6.f : v);
  return v;
}
// derivative mask, torch semantics: relu' = (v > 0); relu6 (hardtanh) ' = (0 < v < 6)
__device__ __forceinline__ float act_mask(float v, int act) {
  if (act == SC_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (act == SC_ACT_RELU6) return (v > 0.f && v < 6.f) ? 1.f : 0.f;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dispatch on storage dtype
#define SC_DISPATCH_DTYPE(dtype, ...)                         \
  do {                                                        \
    if ((dtype) == SC_F32) {                                  \
      using T = float;                                        \
      __VA_ARGS__;                                            \
    } else if ((dtype) == SC_BF16) {                          \
      using T = __nv_bfloat16;                                \
      __VA_ARGS__;                                            \
    } else {                                                  \
      return SC_ERR_BAD_ARG;                                  \
    }                                                         \
  } while (0)

}  // namespace sc
