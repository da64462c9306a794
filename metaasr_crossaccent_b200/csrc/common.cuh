// Shared device/host helpers for the metaasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <algorithm>

#include "../../include/metaasr_b200.h"

namespace masr {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MASR_CHECK_CUDA(expr)                                              \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return ::masr::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define MASR_LAUNCH_CHECK() MASR_CHECK_CUDA(cudaGetLastError())

#define MASR_REQUIRE(cond, msg)                                            \
  do {                                                                     \
    if (!(cond)) {                                                         \
      ::masr::set_error(std::string("invalid argument: ") + (msg) + " [" #cond "]"); \
      return MASR_E_INVALID;                                               \
    }                                                                      \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ dtype helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// dispatch a templated functor on a runtime dtype code
#define MASR_DISPATCH_DTYPE(code, T, ...)                                   \
  do {                                                                      \
    if ((code) == MASR_F32) { using T = float; __VA_ARGS__; }               \
    else if ((code) == MASR_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { ::masr::set_error("bad dtype code"); return MASR_E_INVALID; }    \
  } while (0)

// ------------------------------------------------------------------ counter-based dropout RNG
// keep(seed, site, idx) is a pure function, so the backward pass regenerates the forward mask.
// 32-bit integer hash ("lowbias32" multiply-xorshift finaliser, ~8 ALU instructions) of the element index
// under a per-(seed, site) key -> 24-bit uniform.  The key is loop invariant (seed and site are kernel
// arguments), so the per-element cost is one finaliser: the soft-max warps of the attention kernels and the
// elementwise dropout kernels are ALU-bound on this function.
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t mix_hash(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint32_t key = lowbias32(uint32_t(seed) ^ lowbias32(uint32_t(seed >> 32) + site * 0x9E3779B9u + 0x85EBCA6Bu));
  const uint32_t x = uint32_t(idx) ^ (uint32_t(idx >> 32) * 0xC2B2AE35u);
  return lowbias32(x ^ key) >> 8;   // 24 bits
}
// Seed of a dropout site = base (by value) + a device-resident per-step offset.  The offset lives in device
// memory so that a captured CUDA graph of the step replays with fresh masks (masr_seed_bump advances it).
struct SeedArg { uint64_t base; const uint64_t* bump; };
__device__ __forceinline__ uint64_t resolve_seed(SeedArg s) { return s.base + (s.bump != nullptr ? *s.bump : 0ull); }
extern const uint64_t* g_seed_dev_ptr;     // host-side global set by masr_set_seed_ptr (abi.cu)

// returns the multiplier applied to a kept/dropped element: 1/(1-p) or 0
__device__ __forceinline__ float drop_scale(float p, float inv_keep, uint64_t seed, uint32_t site, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  float u = float(mix_hash(seed, site, idx)) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

// Attention-probability dropout.  All attention kernels (CUDA-core and tcgen05, forward and backward) share this
// function, so the backward of a site replays the forward's mask whichever kernel ran it.  One 32-bit hash per
// PAIR of keys, 16 bits each (the soft-max warps are ALU-bound on this); keep-probability = 1 - thr16 / 65536.
__host__ __device__ __forceinline__ uint32_t attn_drop_thr16(float p) {
  const float t = p * 65536.f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : uint32_t(t));
}
__host__ __device__ __forceinline__ float attn_drop_inv_keep(uint32_t thr16) { return 65536.f / float(65536u - thr16); }
// row = (batch * heads + head) * Lq + query index
__device__ __forceinline__ uint32_t attn_drop_rowkey(uint64_t seed, uint32_t site, uint32_t row) {
  const uint32_t key = lowbias32(uint32_t(seed) ^ lowbias32(uint32_t(seed >> 32) + site * 0x9E3779B9u + 0x85EBCA6Bu));
  return lowbias32((row * 0x9E3779B9u) ^ key);
}
__device__ __forceinline__ uint32_t attn_drop_pair(uint32_t rowkey, uint32_t kpair) { return lowbias32(rowkey + kpair * 0x85EBCA77u); }
__device__ __forceinline__ bool attn_drop_keep(uint32_t bits, int kj, uint32_t thr16) {
  return ((bits >> ((kj & 1) << 4)) & 0xffffu) >= thr16;
}
__device__ __forceinline__ float attn_drop_scale(float p, uint64_t seed, uint32_t site, uint32_t row, int kj) {
  if (p <= 0.f) return 1.f;
  const uint32_t thr = attn_drop_thr16(p);
  return attn_drop_keep(attn_drop_pair(attn_drop_rowkey(seed, site, row), uint32_t(kj) >> 1), kj, thr) ? attn_drop_inv_keep(thr) : 0.f;
}

// ------------------------------------------------------------------ warp / block reductions
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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

int sm_count();

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Kernels of the step form one long dependent chain of small launches.  With the programmatic-stream-
// serialization attribute the NEXT kernel's CTAs may become resident while this one still runs, so its
// prologue (barrier init, TMEM allocation, tensor-map prefetch, index math) overlaps our tail; it must call
// pdl_wait() before touching global memory -- that returns once every prerequisite grid has completed and
// flushed.  Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();            // MASR_PDL=0 in the environment disables the attribute (A/B measurements)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace masr

namespace masr {
// 8 consecutive elements <-> 8 floats with 128-bit accesses (pointer must be 16 B aligned for bf16,
// 32 B-contiguous / 16 B aligned for fp32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* v);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* v);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint4 pk;
  __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0], v[1]), a1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 a2 = __floats2bfloat162_rn(v[4], v[5]), a3 = __floats2bfloat162_rn(v[6], v[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
  pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
  *reinterpret_cast<uint4*>(p) = pk;
}
}  // namespace masr
