// Shared device/host helpers for the avdf sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

#include "../../include/avdf.h"

namespace avdf {

// ---- error plumbing (C-ABI never throws; returns codes, message kept per thread) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define AVDF_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      avdf::set_error("%s: invalid argument: %s", __func__, msg);   \
      return AVDF_ERR_INVALID;                                      \
    }                                                               \
  } while (0)

#define AVDF_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      avdf::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_));     \
      return AVDF_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

// ---- small device helpers ----
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

// ---- packed fp32 pairs: sm_100 executes add / mul / fma on two fp32 values per instruction (FFMA2 etc.), which halves
// the issue slots of the FP-heavy, instruction-bound epilogues. A pair lives in one 64-bit register.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 pk2(float v) { return pk2(v, v); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// exact-erf GELU (blocks.py:1239 nn.GELU default)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// Load 8 consecutive channels as fp32 from an fp32 or bf16 row (16-byte aligned for bf16,
// 32-byte aligned for fp32).
template <typename T> struct Row8;
template <> struct Row8<float> {
  __device__ __forceinline__ static void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Row8<__nv_bfloat16> {
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <> struct Row8<__half> {
  __device__ __forceinline__ static void load(const __half* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ static void store(__half* p, const float (&v)[8]) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

__device__ __forceinline__ void store1(float* p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store1(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float load1(const float* p) { return *p; }
__device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float load1(const __half* p) { return __half2float(*p); }

// Run `...` with `T` bound to the C++ type of a runtime AVDF_DTYPE_* value.
#define AVDF_DISPATCH_DTYPE(dt, T, ...)                                              \
  do {                                                                               \
    if ((dt) == AVDF_DTYPE_F32) { using T = float; __VA_ARGS__; }                    \
    else if ((dt) == AVDF_DTYPE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }      \
    else { using T = __half; __VA_ARGS__; }                                          \
  } while (0)
#define AVDF_CHECK_DTYPE(dt, what) \
  AVDF_CHECK_ARG((dt) == AVDF_DTYPE_F32 || (dt) == AVDF_DTYPE_BF16 || (dt) == AVDF_DTYPE_F16, what)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL). A kernel launched through launch_pdl() may start while the previous kernel of
// the stream is still running - as soon as every CTA of that kernel has executed pdl_trigger() or exited - and must
// execute pdl_wait() before it touches anything the previous kernel wrote (the wait returns when the previous grid has
// completed and its memory is visible). What runs before pdl_wait() - barrier initialisation, TMEM allocation,
// tensor-map prefetch, loads of weights - overlaps the previous kernel's tail. Kernels launched normally are unaffected
// by pdl_trigger(). AVDF_PDL=0 switches the launch attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// the same for a kernel that runs as clusters of `cluster_x` CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- per-device host state (runtime.cu). A process may drive several GPUs (the reference yaml ships devices:
// ['cuda:3']): the SM count and the "opt-in shared-memory attribute already set" flags are properties of the CURRENT
// device, not of the process.
int current_device();          // cudaGetDevice, -1 on error
int device_sm_count();         // SM count of the current device (cached per device)
struct DeviceOnce {            // one bit per device ordinal; marked only AFTER the guarded setup succeeded
  unsigned long long bits[4] = {0, 0, 0, 0};
  bool done(int dev) const { return dev >= 0 && dev < 256 && ((__atomic_load_n(&bits[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull); }
  void mark(int dev) { if (dev >= 0 && dev < 256) __atomic_fetch_or(&bits[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE); }
};
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device); idempotent, so a race between two host
// threads is harmless; a failure returns the error and leaves the flag clear.
#define AVDF_SMEM_ATTR_ONCE(kernel, bytes)                                                                   \
  do {                                                                                                       \
    static avdf::DeviceOnce once_;                                                                           \
    const int dev_ = avdf::current_device();                                                                 \
    if (!once_.done(dev_)) {                                                                                 \
      AVDF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
      once_.mark(dev_);                                                                                      \
    }                                                                                                        \
  } while (0)

}  // namespace avdf
