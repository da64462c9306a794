// Kernel family 4: meta-learning parameter updates over FLAT fp32 arenas.
//
// The reference keeps 112-114 separate tensors and issues one small launch per tensor for
// load_state_dict(_original), clip_grad_norm_, SGD, `_updates[n] += p.grad`, `div_` and Adam
// (src/fo_meta_interface.py:180-250).  Here every parameter of the model lives back to back in one
// arena (same order as the state dict), so each of those steps is ONE streaming pass with 128-bit
// accesses over 24.9 M elements and no host synchronisation: the gradient norm stays on the device
// and the NaN guard (math.isnan(grad_norm), :245) is evaluated by the consuming kernel.
#include "common.cuh"

namespace masr {

constexpr int MT_THREADS = 256;

static inline int mt_grid(int64_t n4) {
  int64_t blocks = ceil_div64(n4, MT_THREADS);
  int64_t cap = int64_t(sm_count()) * 8;           // 8 resident CTAs of 256 threads per SM
  return int(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

__device__ __forceinline__ float clip_coef(const double* sumsq, float max_norm) {
  // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
  const float total = float(sqrt(*sumsq));
  const float c = max_norm / (total + 1e-6f);
  // torch.clamp(NaN, max=1) is NaN: a NaN norm multiplies EVERY gradient by NaN in the reference
  // (fo_meta_interface.py:148-154: the inner-test NaN is only warned about and still accumulated)
  return (c < 1.f || c != c) ? c : 1.f;
}

__global__ void __launch_bounds__(MT_THREADS) mt_sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[MT_THREADS / 32];
  const int64_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    a0 = fmaf(v.x, v.x, a0); a1 = fmaf(v.y, v.y, a1); a2 = fmaf(v.z, v.z, a2); a3 = fmaf(v.w, v.w, a3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; a0 = fmaf(v, v, a0); }
  double s = warp_sum_d(double(a0) + double(a1) + double(a2) + double(a3));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < MT_THREADS / 32 ? red[threadIdx.x] : 0.0;
    s = warp_sum_d(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

__global__ void mt_zero_double_kernel(double* p) {
  pdl_launch_dependents();
  pdl_wait(); *p = 0.0; }

__global__ void __launch_bounds__(MT_THREADS)
mt_clip_sgd_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf, int64_t n,
                   const double* __restrict__ sumsq, float max_norm, float lr, float mu, int nesterov, int first,
                   __nv_bfloat16* __restrict__ shadow, int flags) {
  // shadow (may be NULL): the bf16 copy of the parameter arena the GEMMs read, written in the same pass (the engine then
  // skips its cast pass).  flags & MASR_SGD_LAST_STEP: the task takes no further inner step and nothing reads the
  // scaled gradient or the momentum afterwards -- their write-backs (200 MB of the 500 MB this pass moves) are dropped.
  const bool keep = !(flags & MASR_SGD_LAST_STEP);
  pdl_launch_dependents();
  pdl_wait();
  const double ss = *sumsq;
  const int64_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* b4 = reinterpret_cast<float4*>(buf);
  if (ss != ss) {                                  // NaN gradient norm: skip the step (:245-248)
    // a skipped FIRST step of a task leaves the momentum of the previous task behind while the host clears its
    // `first` flag: zero it, so that the next step's mu * buf + g is the fresh buffer torch.optim.SGD would create
    if (shadow != nullptr) {                       // the parameters did not move, but the shadow must still mirror them
      uint2* s2 = reinterpret_cast<uint2*>(shadow);
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 pv = p4[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
        s2[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
      }
      if (blockIdx.x == 0 && threadIdx.x < (n & 3)) shadow[n4 * 4 + threadIdx.x] = __float2bfloat16(p[n4 * 4 + threadIdx.x]);
    }
    if (first && mu != 0.f && keep) {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x)
        b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blockIdx.x == 0 && threadIdx.x < (n & 3)) buf[n4 * 4 + threadIdx.x] = 0.f;
    }
    return;
  }
  const float coef = clip_coef(sumsq, max_norm);
  auto upd = [&](float& pv, float& gv, float& bv) {
    gv = gv * coef;
    float d = gv;
    if (mu != 0.f) {
      bv = first ? gv : fmaf(mu, bv, gv);
      d = nesterov ? fmaf(mu, bv, gv) : bv;
    }
    pv = fmaf(-lr, d, pv);
  };
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = p4[i], gv = g4[i], bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : b4[i];
    upd(pv.x, gv.x, bv.x); upd(pv.y, gv.y, bv.y); upd(pv.z, gv.z, bv.z); upd(pv.w, gv.w, bv.w);
    p4[i] = pv;
    if (keep) { g4[i] = gv; b4[i] = bv; }
    if (shadow != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      reinterpret_cast<uint2*>(shadow)[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    float pv = p[i], gv = g[i], bv = first ? 0.f : buf[i];
    upd(pv, gv, bv);
    p[i] = pv;
    if (keep) { g[i] = gv; buf[i] = bv; }
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pv);
  }
}

// dst = src (fp32) and shadow = bf16(src) in one pass: load_state_dict(_original) of run_task (fo_meta_interface.py:226)
// plus the engine's compute-dtype copy of the fresh weights
__global__ void __launch_bounds__(MT_THREADS)
mt_copy_cast_kernel(float* __restrict__ dst, __nv_bfloat16* __restrict__ shadow, const float* __restrict__ src, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = n / 4;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  uint2* h2 = reinterpret_cast<uint2*>(shadow);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = s4[i];
    if (dst != nullptr) d4[i] = v;
    if (shadow != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      h2[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    const float v = src[i];
    if (dst != nullptr) dst[i] = v;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(MT_THREADS)
mt_clip_kernel(float* __restrict__ g, int64_t n, const double* __restrict__ sumsq, float max_norm) {
  pdl_launch_dependents();
  pdl_wait();
  const float coef = clip_coef(sumsq, max_norm);
  const int64_t n4 = n / 4;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 v = g4[i]; v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef; g4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[n4 * 4 + threadIdx.x] *= coef;
}

__global__ void __launch_bounds__(MT_THREADS)
mt_accumulate_kernel(float* __restrict__ upd, const float* __restrict__ g, int64_t n,
                     const double* __restrict__ sumsq, float max_norm) {
  pdl_launch_dependents();
  pdl_wait();
  const float coef = sumsq != nullptr ? clip_coef(sumsq, max_norm) : 1.f;
  const int64_t n4 = n / 4;
  float4* u4 = reinterpret_cast<float4*>(upd);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 u = u4[i]; const float4 v = g4[i];
    u.x += v.x * coef; u.y += v.y * coef; u.z += v.z * coef; u.w += v.w * coef;
    u4[i] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const int64_t i = n4 * 4 + threadIdx.x; upd[i] += g[i] * coef; }
}

__global__ void __launch_bounds__(MT_THREADS)
mt_reptile_delta_kernel(float* __restrict__ upd, const float* __restrict__ theta, const float* __restrict__ phi, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = n / 4;
  float4* u4 = reinterpret_cast<float4*>(upd);
  const float4* t4 = reinterpret_cast<const float4*>(theta);
  const float4* f4 = reinterpret_cast<const float4*>(phi);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 u = u4[i]; const float4 t = t4[i], f = f4[i];
    u.x += t.x - f.x; u.y += t.y - f.y; u.z += t.z - f.z; u.w += t.w - f.w;
    u4[i] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const int64_t i = n4 * 4 + threadIdx.x; upd[i] += theta[i] - phi[i]; }
}

// torch/optim/adam.py _single_tensor_adam (amsgrad=False, maximize=False, weight_decay=0):
//   m.lerp_(g, 1-b1); v = b2 v + (1-b2) g g; denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) m/denom
__global__ void __launch_bounds__(MT_THREADS)
mt_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ upd,
               int64_t n, float count, float step_size, float b1, float b2, float eps, float bc2_sqrt,
               const double* __restrict__ skip_if_nan, const double* __restrict__ clip_sumsq, float max_norm) {
  pdl_launch_dependents();
  pdl_wait();
  if (skip_if_nan != nullptr) { const double ss = *skip_if_nan; if (ss != ss) return; }
  const float coef = (clip_sumsq != nullptr ? clip_coef(clip_sumsq, max_norm) : 1.f);
  const int64_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* u4 = reinterpret_cast<const float4*>(upd);
  auto one = [&](float& pv, float& mv, float& vv, float uv) {
    // `_updates /= _counter` is a true division in the reference (fo_meta_interface.py:201-202)
    const float g = (uv / count) * coef;
    mv = mv + (g - mv) * (1.f - b1);
    vv = fmaf(vv, b2, (1.f - b2) * g * g);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - step_size * (mv / denom);
  };
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = p4[i], mv = m4[i], vv = v4[i]; const float4 uv = u4[i];
    one(pv.x, mv.x, vv.x, uv.x); one(pv.y, mv.y, vv.y, uv.y); one(pv.z, mv.z, vv.z, uv.z); one(pv.w, mv.w, vv.w, uv.w);
    p4[i] = pv; m4[i] = mv; v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    float pv = p[i], mv = m[i], vv = v[i];
    one(pv, mv, vv, upd[i]);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

__global__ void __launch_bounds__(MT_THREADS)
mt_axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = n / 4;
  float4* y4 = reinterpret_cast<float4*>(y);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 yv = y4[i]; const float4 xv = x4[i];
    yv.x = fmaf(a, xv.x, yv.x); yv.y = fmaf(a, xv.y, yv.y); yv.z = fmaf(a, xv.z, yv.z); yv.w = fmaf(a, xv.w, yv.w);
    y4[i] = yv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const int64_t i = n4 * 4 + threadIdx.x; y[i] = fmaf(a, x[i], y[i]); }
}

// ------------------------------------------------------------------ the one collective of the path, in the switch
// All-reduce (SUM) of the flat update arena over the NVSwitch multicast mapping of a symmetric allocation: every rank owns a
// slice; multimem.ld_reduce pulls the slice from ALL ranks with the addition done inside the switch, multimem.st
// broadcasts the sum back into every rank's copy.  Per GPU ~(1 + 1/W) arena sizes cross its links once (a ring all-reduce
// moves 2 (W-1)/W of it twice through every GPU's memory).  Ordering against the producers / consumers of the arena on
// the other GPUs is the caller's (a cross-GPU barrier before and after: symmetric-memory signal pads).
__device__ __forceinline__ float4 mm_ld_reduce(const float* a) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* a, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(512) nvls_allreduce_f32_kernel(float* __restrict__ mc, int64_t begin4, int64_t end4) {
  pdl_launch_dependents();
  pdl_wait();
  // four independent switch round trips in flight per thread (a multimem.ld_reduce takes microseconds)
  constexpr int U = 4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = begin4 + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  for (; i + (U - 1) * stride < end4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = mm_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) mm_st(mc + 4 * (i + u * stride), v[u]);
  }
  for (; i < end4; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
}

// The whole outer update of a multi-GPU meta-step in ONE kernel (fo_meta_interface.py:200-221 + the collective): this rank
// owns elements [begin, end).  g = sum over ranks of the update arena (multimem.ld_reduce: added inside the NVSwitch) / count;
// Adam on the owner's slice of (theta, m, v) -- the moments live sharded, an eighth of the pass per GPU --; the new theta
// goes to EVERY rank's meta weights with one multimem.st, and the slice of every rank's update arena is cleared for the
// next step the same way.  Replaces all-reduce (2 x 100 MB through every GPU) + a 0.7 GB Adam pass + a memset on each GPU.
// Same arithmetic per element as mt_adam_kernel; every rank receives identical bits.
__global__ void __launch_bounds__(512)
nvls_reduce_adam_kernel(float* __restrict__ mc_upd, float* __restrict__ mc_theta, const float* __restrict__ theta,
                        float* __restrict__ m, float* __restrict__ v, int64_t begin4, int64_t end4, int64_t n4,
                        float count, float step_size, float b1, float b2, float eps, float bc2_sqrt) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int U = 4;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto one = [&](float& pv, float& mv, float& vv, float uv) {
    const float g = uv / count;
    mv = mv + (g - mv) * (1.f - b1);
    vv = fmaf(vv, b2, (1.f - b2) * g * g);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - step_size * (mv / denom);
  };
  auto apply = [&](int64_t i, const float4& uv) {
    if (i < n4) {
      float4 pv = reinterpret_cast<const float4*>(theta)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      one(pv.x, mv.x, vv.x, uv.x); one(pv.y, mv.y, vv.y, uv.y); one(pv.z, mv.z, vv.z, uv.z); one(pv.w, mv.w, vv.w, uv.w);
      reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
      mm_st(mc_theta + 4 * i, pv);
    }
    mm_st(mc_upd + 4 * i, zero);
  };
  int64_t i = begin4 + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  for (; i + (U - 1) * stride < end4; i += U * stride) {
    float4 uv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) uv[u] = mm_ld_reduce(mc_upd + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) apply(i + u * stride, uv[u]);
  }
  for (; i < end4; i += stride) apply(i, mm_ld_reduce(mc_upd + 4 * i));
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace masr

using namespace masr;

#define MT_ALIGN_CHECK(...)                                                              \
  do {                                                                                   \
    const void* _ps[] = {__VA_ARGS__};                                                   \
    for (const void* _p : _ps) MASR_REQUIRE(aligned16(_p), "flat arenas must be 16-byte aligned"); \
  } while (0)

extern "C" int masr_mt_sumsq(const float* g, int64_t n, double* out, int zero_first, void* stream) {
  MT_ALIGN_CHECK(g);
  cudaStream_t st = as_stream(stream);
  if (zero_first) { launch_pdl(mt_zero_double_kernel, dim3(1), dim3(1), 0, st, out); MASR_LAUNCH_CHECK(); }
  if (n == 0) return MASR_OK;
  launch_pdl(mt_sumsq_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, st, g, n, out);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_clip_sgd_ex(float* p, float* g, float* buf, int64_t n, const double* sumsq, float max_norm,
                                   float lr, float momentum, int nesterov, int first_step, void* shadow_bf16, int flags,
                                   void* stream) {
  MT_ALIGN_CHECK(p, g, buf);
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(shadow_bf16) & 7u) == 0, "masr_mt_clip_sgd_ex: shadow must be 8-byte aligned");
  if (n == 0) return MASR_OK;
  launch_pdl(mt_clip_sgd_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), p, g, buf, n, sumsq, max_norm, lr,
             momentum, nesterov, first_step, static_cast<__nv_bfloat16*>(shadow_bf16), flags);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_clip_sgd(float* p, float* g, float* buf, int64_t n, const double* sumsq, float max_norm,
                                float lr, float momentum, int nesterov, int first_step, void* stream) {
  return masr_mt_clip_sgd_ex(p, g, buf, n, sumsq, max_norm, lr, momentum, nesterov, first_step, nullptr, 0, stream);
}

extern "C" int masr_mt_copy_cast(float* dst, void* shadow_bf16, const float* src, int64_t n, void* stream) {
  MT_ALIGN_CHECK(src);
  MASR_REQUIRE(aligned16(dst) && (reinterpret_cast<uintptr_t>(shadow_bf16) & 7u) == 0, "masr_mt_copy_cast: alignment");
  if (n == 0) return MASR_OK;
  launch_pdl(mt_copy_cast_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), dst,
             static_cast<__nv_bfloat16*>(shadow_bf16), src, n);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_clip(float* g, int64_t n, const double* sumsq, float max_norm, void* stream) {
  MT_ALIGN_CHECK(g);
  if (n == 0) return MASR_OK;
  launch_pdl(mt_clip_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), g, n, sumsq, max_norm);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_accumulate(float* upd, const float* g, int64_t n, const double* sumsq, float max_norm, void* stream) {
  MT_ALIGN_CHECK(upd, g);
  if (n == 0) return MASR_OK;
  launch_pdl(mt_accumulate_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), upd, g, n, sumsq, max_norm);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_reptile_delta(float* upd, const float* theta, const float* phi, int64_t n, void* stream) {
  MT_ALIGN_CHECK(upd, theta, phi);
  if (n == 0) return MASR_OK;
  launch_pdl(mt_reptile_delta_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), upd, theta, phi, n);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_adam(float* p, float* m, float* v, const float* upd, int64_t n, float count,
                            float lr, float beta1, float beta2, float eps, double bc1, double bc2,
                            const double* skip_if_nan, const double* clip_sumsq, float max_norm, void* stream) {
  MT_ALIGN_CHECK(p, m, v, upd);
  if (n == 0) return MASR_OK;
  const float step_size = float(double(lr) / bc1);
  const float bc2_sqrt = float(sqrt(bc2));
  launch_pdl(mt_adam_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), 
      p, m, v, upd, n, count, step_size, beta1, beta2, eps, bc2_sqrt, skip_if_nan, clip_sumsq, max_norm);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_nvls_allreduce_f32(void* multicast_ptr, int64_t begin, int64_t end, void* stream) {
  MASR_REQUIRE(multicast_ptr != nullptr, "masr_nvls_allreduce_f32: no multicast mapping (NVLS not available)");
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15u) == 0 && begin % 4 == 0 && end % 4 == 0 && begin <= end,
               "masr_nvls_allreduce_f32: 16-byte aligned base, slice bounds multiple of 4 elements");
  if (begin == end) return MASR_OK;
  const int64_t n4 = (end - begin) / 4;
  const int blocks = int(std::min<int64_t>(ceil_div64(n4, 512 * 4), int64_t(sm_count()) * 4));
  launch_pdl(nvls_allreduce_f32_kernel, dim3(blocks), dim3(512), 0, as_stream(stream), static_cast<float*>(multicast_ptr),
             begin / 4, end / 4);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_nvls_reduce_adam(void* mc_upd, void* mc_theta, const float* theta, float* m, float* v,
                                     int64_t begin, int64_t end, int64_t n, float count, float lr, float beta1, float beta2,
                                     float eps, double bc1, double bc2, void* stream) {
  MASR_REQUIRE(mc_upd != nullptr && mc_theta != nullptr, "masr_nvls_reduce_adam: no multicast mapping (NVLS not available)");
  MT_ALIGN_CHECK(mc_upd, mc_theta, theta, m, v);
  MASR_REQUIRE(begin % 4 == 0 && end % 4 == 0 && n % 4 == 0 && begin <= end, "masr_nvls_reduce_adam: bounds multiple of 4 elements");
  if (begin == end) return MASR_OK;
  const int64_t n4 = (end - begin) / 4;
  const int blocks = int(std::min<int64_t>(ceil_div64(n4, 512 * 4), int64_t(sm_count()) * 4));
  launch_pdl(nvls_reduce_adam_kernel, dim3(blocks), dim3(512), 0, as_stream(stream), static_cast<float*>(mc_upd),
             static_cast<float*>(mc_theta), theta, m, v, begin / 4, end / 4, n / 4, count, float(double(lr) / bc1), beta1, beta2,
             eps, float(sqrt(bc2)));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_mt_axpy(float* y, const float* x, float a, int64_t n, void* stream) {
  MT_ALIGN_CHECK(y, x);
  if (n == 0) return MASR_OK;
  launch_pdl(mt_axpy_kernel, dim3(mt_grid(n / 4)), dim3(MT_THREADS), 0, as_stream(stream), y, x, a, n);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}
