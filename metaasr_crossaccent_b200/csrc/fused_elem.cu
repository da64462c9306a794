// Kernel family 3: fused elementwise / row-reduction kernels -- residual+dropout+LayerNorm
// (forward and backward), positional encoding, embedding, label-smoothed CE with accuracy and
// gradient, bias-gradient column sums, casts.  All of them are HBM-bound: one pass over the data,
// 128-bit accesses where the row length allows, fp32 math.
#include "common.cuh"

namespace masr {

static inline int grid_for(int64_t total, int threads) {
  int64_t blocks = ceil_div64(total, threads);
  int64_t cap = int64_t(sm_count()) * 16;
  return int(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ------------------------------------------------------------------ residual + dropout + LayerNorm
// One warp per row; the row (d <= 32 * LN_MAX_PER_LANE) lives in registers between the two passes.
constexpr int LN_MAX_PER_LANE = 32;     // d <= 1024

template <typename T>
__global__ void add_layernorm_fwd_kernel(T* __restrict__ x, const T* __restrict__ res,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                         int rows, int d, float eps, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int per = (d + 31) / 32;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    float v[LN_MAX_PER_LANE];
    float sum = 0.f;
    const int64_t base = int64_t(row) * d;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      if (i < per) {
        const int c = lane + i * 32;
        float s = 0.f;
        if (c < d) {
          s = to_f<T>(x[base + c]) * drop_scale(p, inv_keep, seed, site, uint64_t(base + c));
          if (res != nullptr) s += to_f<T>(res[base + c]);
          s = to_f<T>(from_f<T>(s));       // the stored (possibly bf16-rounded) value is what backward sees
          x[base + c] = from_f<T>(s);
        }
        v[i] = s;
        sum += s;
      }
    }
    const float mean = warp_sum(sum) / float(d);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i)
      if (i < per) { const int c = lane + i * 32; if (c < d) { const float t = v[i] - mean; var += t * t; } }
    const float rstd = rsqrtf(warp_sum(var) / float(d) + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i)
      if (i < per) {
        const int c = lane + i * 32;
        if (c < d) y[base + c] = from_f<T>((v[i] - mean) * rstd * gamma[c] + beta[c]);
      }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// Backward.  ds = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// dgamma/dbeta: per-thread column partials across the block's rows, combined through shared
// memory, then one atomicAdd per column per block.
template <typename T>
__global__ void add_layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s,
                                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                         const float* __restrict__ gamma, T* __restrict__ ds, int ds_accum,
                                         T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         int rows, int d, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  extern __shared__ float red[];      // [warps][2][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int per = (d + 31) / 32;
  float dg[LN_MAX_PER_LANE], db[LN_MAX_PER_LANE];
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  for (int row = blockIdx.x * warps_per_block + warp; row < rows; row += gridDim.x * warps_per_block) {
    const int64_t base = int64_t(row) * d;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i)
      if (i < per) {
        const int c = lane + i * 32;
        g[i] = 0.f; xh[i] = 0.f;
        if (c < d) {
          const float dyv = to_f<T>(dy[base + c]);
          xh[i] = (to_f<T>(s[base + c]) - mean) * rstd;
          g[i] = dyv * gamma[c];
          dg[i] += dyv * xh[i];
          db[i] += dyv;
          s1 += g[i];
          s2 += g[i] * xh[i];
        }
      }
    s1 = warp_sum(s1) / float(d);
    s2 = warp_sum(s2) / float(d);
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i)
      if (i < per) {
        const int c = lane + i * 32;
        if (c < d) {
          const float v = rstd * (g[i] - s1 - xh[i] * s2);
          if (dx != nullptr) dx[base + c] = from_f<T>(v * drop_scale(p, inv_keep, seed, site, uint64_t(base + c)));
          float o = v;
          if (ds_accum) o += to_f<T>(ds[base + c]);
          ds[base + c] = from_f<T>(o);
        }
      }
  }
  float* rg = red + size_t(warp) * 2 * d;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i)
    if (i < per) { const int c = lane + i * 32; if (c < d) { rg[c] = dg[i]; rg[d + c] = db[i]; } }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < warps_per_block; ++w) t += red[size_t(w) * 2 * d + c];
    if (c < d) atomicAdd(&dgamma[c], t); else atomicAdd(&dbeta[c - d], t);
  }
}

// Vectorised variants (d % 8 == 0, d <= 1024): a lane owns chunks of 8 consecutive elements
// (chunk index = i * 32 + lane), i.e. 128-bit loads/stores and 1/8th of the instructions.
template <typename T, int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_vec_kernel(T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                             const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
                             float* __restrict__ rstd_out, int rows, int d, float eps, float p, float inv_keep,
                             SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nchunk = d / 8;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const int64_t base = int64_t(row) * d;
    float v[LN_MAX_CHUNKS][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int ch = i * 32 + lane;
      if (ch < nchunk) {
        load8<T>(x + base + ch * 8, v[i]);
        if (p > 0.f) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] *= drop_scale(p, inv_keep, seed, site, uint64_t(base + ch * 8 + j));
        }
        if (res != nullptr) {
          float r[8];
          load8<T>(res + base + ch * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = to_f<T>(from_f<T>(v[i][j])); sum += v[i][j]; }
        store8<T>(x + base + ch * 8, v[i]);
      }
    }
    const float mean = warp_sum(sum) / float(d);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i)
      if (i * 32 + lane < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; var += t * t; }
      }
    const float rstd = rsqrtf(warp_sum(var) / float(d) + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int ch = i * 32 + lane;
      if (ch < nchunk) {
        float g[8], bt[8], o[8];
        load8<float>(gamma + ch * 8, g);
        load8<float>(beta + ch * 8, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + bt[j];
        store8<T>(y + base + ch * 8, o);
      }
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

template <typename T, int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ s, const float* __restrict__ mean_in,
                             const float* __restrict__ rstd_in, const float* __restrict__ gamma, T* __restrict__ ds,
                             int ds_accum, T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                             int rows, int d, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  extern __shared__ float red[];      // [warps][2][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nchunk = d / 8;
  float dg[LN_MAX_CHUNKS][8], db[LN_MAX_CHUNKS][8], gm[LN_MAX_CHUNKS][8];
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int ch = i * 32 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; gm[i][j] = 0.f; }
    if (ch < nchunk) load8<float>(gamma + ch * 8, gm[i]);
  }
  for (int row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
    const int64_t base = int64_t(row) * d;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float g[LN_MAX_CHUNKS][8], xh[LN_MAX_CHUNKS][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int ch = i * 32 + lane;
      if (ch < nchunk) {
        float dyv[8], sv[8];
        load8<T>(dy + base + ch * 8, dyv);
        load8<T>(s + base + ch * 8, sv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (sv[j] - mean) * rstd;
          g[i][j] = dyv[j] * gm[i][j];
          dg[i][j] += dyv[j] * xh[i][j];
          db[i][j] += dyv[j];
          s1 += g[i][j];
          s2 += g[i][j] * xh[i][j];
        }
      }
    }
    s1 = warp_sum(s1) / float(d);
    s2 = warp_sum(s2) / float(d);
#pragma unroll
    for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
      const int ch = i * 32 + lane;
      if (ch < nchunk) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2);
        if (dx != nullptr) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = v[j] * drop_scale(p, inv_keep, seed, site, uint64_t(base + ch * 8 + j));
          store8<T>(dx + base + ch * 8, o);
        }
        if (ds_accum) {
          float o[8];
          load8<T>(ds + base + ch * 8, o);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += o[j];
        }
        store8<T>(ds + base + ch * 8, v);
      }
    }
  }
  float* rg = red + size_t(warp) * 2 * d;
#pragma unroll
  for (int i = 0; i < LN_MAX_CHUNKS; ++i) {
    const int ch = i * 32 + lane;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { rg[ch * 8 + j] = dg[i][j]; rg[d + ch * 8 + j] = db[i][j]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += red[size_t(w) * 2 * d + c];
    if (c < d) atomicAdd(&dgamma[c], t); else atomicAdd(&dbeta[c - d], t);
  }
}

// ------------------------------------------------------------------ positional encoding / embedding / dropout
template <typename T>
__global__ void add_pe_dropout_kernel(T* __restrict__ x, const float* __restrict__ pe, int64_t n, int L, int d,
                                      float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % d);
    const int l = int((i / d) % L);
    const float v = (to_f<T>(x[i]) + pe[int64_t(l) * d + c]) * drop_scale(p, inv_keep, seed, site, uint64_t(i));
    x[i] = from_f<T>(v);
  }
}

template <typename T>
__global__ void embed_pe_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ E,
                                    const float* __restrict__ pe, T* __restrict__ out, int64_t n, int L, int d,
                                    float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % d);
    const int64_t r = i / d;
    const int l = int(r % L);
    const float v = (E[ids[r] * d + c] + pe[int64_t(l) * d + c]) * drop_scale(p, inv_keep, seed, site, uint64_t(i));
    out[i] = from_f<T>(v);
  }
}

template <typename T>
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ dout, float* __restrict__ dE,
                                 int64_t n, int d, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % d);
    const int64_t r = i / d;
    const float g = to_f<T>(dout[i]) * drop_scale(p, inv_keep, seed, site, uint64_t(i));
    if (g != 0.f) atomicAdd(&dE[ids[r] * d + c], g);
  }
}

template <typename T>
__global__ void dropout_kernel(T* __restrict__ x, int64_t n, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t seed = resolve_seed(seed_arg);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    x[i] = from_f<T>(to_f<T>(x[i]) * drop_scale(p, inv_keep, seed, site, uint64_t(i)));
}

// out[n] += sum_m x[m, n].  Block (32 x 8): thread column n = blockIdx.x*32 + tx, rows strided by
// (8 * gridDim.y); partials reduced over ty in shared memory; one atomicAdd per column per block.
template <typename T>
__global__ void colsum_add_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out, int M, int N) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + ty; m < M; m += gridDim.y * 8) acc += to_f<T>(x[int64_t(m) * ldx + n]);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(&out[n], t);
  }
}

// Vectorised variant for N % 8 == 0, N <= 2048, 16 B aligned rows: a thread owns 8 consecutive columns
// (one 128-bit load of bf16), the N/8 column groups of a row sit in adjacent threads, so a warp reads
// 512 contiguous bytes; row partials are combined in shared memory, one atomicAdd per column per block.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out, int M, int N) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256 * 8];
  const int ncg = N / 8;
  const int rl_count = 256 / ncg;
  const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < rl_count) {
    for (int64_t m = int64_t(blockIdx.x) * rl_count + rl; m < M; m += int64_t(gridDim.x) * rl_count) {
      const T* src = x + m * ldx + cg * 8;
      if constexpr (sizeof(T) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(src);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = (rl < rl_count) ? acc[j] : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    const int g = c / 8, j = c % 8;
    float t = 0.f;
    for (int r = 0; r < rl_count; ++r) t += red[(r * ncg + g) * 8 + j];
    atomicAdd(&out[c], t);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dst[i] = from_f<TD>(to_f<TS>(src[i]));
}

// dst[r, 0:cols] = src[r, 0:cols], dst[r, cols:ld_dst] = 0
template <typename TS, typename TD>
__global__ void cast_pad2d_kernel(const TS* __restrict__ src, int64_t ld_src, TD* __restrict__ dst, int64_t ld_dst,
                                  int64_t rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n = rows * ld_dst;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = int(i - r * ld_dst);
    dst[i] = c < cols ? from_f<TD>(to_f<TS>(src[r * ld_src + c])) : from_f<TD>(0.f);
  }
}

__global__ void loss_mix_kernel(double* __restrict__ stats, const float* __restrict__ ctc_loss, float w) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) { stats[3] = double(*ctc_loss); stats[4] = double(w); }
}

// forward: dst[r, f*C + c] = src[r, c*F + f]; inverse_add: dst[r, c*F + f] += src[r, f*C + c]
template <typename TS, typename TD>
__global__ void permute_cf_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t rows, int C, int F, int inverse_add) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n = rows * C * F;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / (int64_t(C) * F);
    const int j = int(i % (int64_t(C) * F));
    if (!inverse_add) {
      const int f = j / C, c = j % C;            // i indexes dst (f-major)
      dst[i] = from_f<TD>(to_f<TS>(src[r * C * F + int64_t(c) * F + f]));
    } else {
      const int c = j / F, f = j % F;            // i indexes dst (c-major)
      dst[i] = from_f<TD>(to_f<TD>(dst[i]) + to_f<TS>(src[r * C * F + int64_t(f) * C + c]));
    }
  }
}

// ------------------------------------------------------------------ every derived weight copy in ONE launch
// engine.prep_weights: (a) the compute-dtype shadow of the parameter arena (skipped when an optimizer kernel wrote it
// already), (b) the [Cout][tap][Cin] / [Cin][tap][Cout] layouts of the 3x3 conv weights, (c) the (c,f)->(f,c) column
// permutation of vgg2enc.weight.  Eight dependent ~3 us launches in front of every batch before; block ranges now.
struct PrepParams {
  const float* params; void* shadow; int64_t n;       // (a)
  masr_conv_prep_job jobs[4]; int njobs;              // (b)
  const float* v2e; void* v2e_p; int v2e_rows, C, F;  // (c)
  int blk_end[6];                                     // block range ends: cast | job 0..3 | permute
};

template <typename T>
__global__ void __launch_bounds__(256) prep_weights_kernel(const __grid_constant__ PrepParams p) {
  pdl_launch_dependents();
  pdl_wait();
  int seg = 0;
  while (seg < 5 && int(blockIdx.x) >= p.blk_end[seg]) ++seg;
  const int b0 = seg == 0 ? 0 : p.blk_end[seg - 1];
  const int64_t tid = int64_t(blockIdx.x - b0) * blockDim.x + threadIdx.x;
  const int64_t nth = int64_t(p.blk_end[seg] - b0) * blockDim.x;
  if (seg == 0) {
    T* sh = static_cast<T*>(p.shadow);
    const int64_t n4 = p.n / 4;
    const float4* s4 = reinterpret_cast<const float4*>(p.params);
    for (int64_t i = tid; i < n4; i += nth) {
      const float4 v = s4[i];
      if constexpr (sizeof(T) == 2) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(sh)[i] = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
      } else {
        reinterpret_cast<float4*>(sh)[i] = v;
      }
    }
    if (tid < (p.n & 3)) sh[n4 * 4 + tid] = from_f<T>(p.params[n4 * 4 + tid]);
  } else if (seg <= 4) {
    const masr_conv_prep_job& j = p.jobs[seg - 1];
    const int Cout = j.Cout, Cin = j.Cin, total = Cout * Cin * 9;
    T* wp = static_cast<T*>(j.wp);
    T* wpt = static_cast<T*>(j.wpt);
    for (int64_t idx = tid; idx < total; idx += nth) {
      {
        const int ci = int(idx % Cin), tap = int((idx / Cin) % 9), co = int(idx / (9 * Cin));
        wp[idx] = from_f<T>(j.w[(co * Cin + ci) * 9 + tap]);
      }
      if (wpt != nullptr) {
        const int co = int(idx % Cout), tap = int((idx / Cout) % 9), ci = int(idx / (9 * Cout));
        wpt[idx] = from_f<T>(j.w[(co * Cin + ci) * 9 + tap]);
      }
    }
  } else {
    T* dst = static_cast<T*>(p.v2e_p);
    const int CF = p.C * p.F;
    const int64_t n = int64_t(p.v2e_rows) * CF;
    for (int64_t i = tid; i < n; i += nth) {
      const int64_t r = i / CF;
      const int jj = int(i - r * CF);
      const int f = jj / p.C, c = jj % p.C;
      dst[i] = from_f<T>(p.v2e[r * CF + int64_t(c) * p.F + f]);
    }
  }
}

// ------------------------------------------------------------------ label-smoothed CE + accuracy + grad
// src/transformer_torch_trainer.py:64-84.  One warp per row of C logits.
//   q_c = (1-eps) for the gold class, eps/C otherwise (sums to 1 - eps/C)
//   loss_row = -sum_c q_c (z_c - logZ);   d loss_row / d z_c = (sum q) softmax_c - q_c
template <typename DT>
__global__ void ls_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ gold, int N, int C,
                             float eps, float inv_n, const float* __restrict__ inv_n_dev, double* __restrict__ stats,
                             int64_t* __restrict__ argmax_out, DT* __restrict__ dlogits, int64_t ld_dl) {
  pdl_launch_dependents();
  pdl_wait();
  if (inv_n_dev != nullptr) inv_n = *inv_n_dev;       // device-resident 1/n_non_pad (CUDA-graph replay)
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  double loss_acc = 0.0;
  int correct = 0, nonpad = 0;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < N; row += gridDim.x * warps_per_block) {
    const float* z = logits + int64_t(row) * C;
    const int64_t g = gold[row];
    float mx = -INFINITY; int am = 0;
    for (int c = lane; c < C; c += 32) { const float v = z[c]; if (v > mx) { mx = v; am = c; } }
    // warp arg-max with first-index tie-break (torch.max(1) on CPU/CUDA returns the first maximum)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    float se = 0.f, sz = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = z[c]; se += expf(v - mx); sz += v; }
    se = warp_sum(se); sz = warp_sum(sz);
    const float logZ = mx + logf(se);
    if (argmax_out != nullptr && lane == 0) argmax_out[row] = am;
    const bool valid = (g >= 0);
    if (valid) {
      const float q_on = 1.f - eps, q_off = eps / float(C);
      const float zg = z[g];
      // -[(1-eps)(zg - logZ) + q_off * (sum_{c != g} z_c - (C-1) logZ)]
      const float row_loss = -(q_on * (zg - logZ) + q_off * ((sz - zg) - float(C - 1) * logZ));
      if (lane == 0) { loss_acc += double(row_loss); nonpad += 1; correct += (am == int(g)); }
    }
    if (dlogits != nullptr) {
      DT* dz = dlogits + int64_t(row) * ld_dl;
      if (valid) {
        const float q_on = 1.f - eps, q_off = eps / float(C);
        const float qsum = q_on + float(C - 1) * q_off;
        for (int c = lane; c < C; c += 32) {
          const float sm = expf(z[c] - logZ);
          dz[c] = from_f<DT>((qsum * sm - (c == int(g) ? q_on : q_off)) * inv_n);
        }
      } else {
        for (int c = lane; c < C; c += 32) dz[c] = from_f<DT>(0.f);
      }
    }
  }
  if (lane == 0 && nonpad > 0) {
    atomicAdd(&stats[0], loss_acc);
    atomicAdd(&stats[1], double(correct));
    atomicAdd(&stats[2], double(nonpad));
  }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_add_layernorm_fwd(void* x_inout, const void* res, const float* gamma, const float* beta,
                                      void* y, float* mean, float* rstd, int dtype, int rows, int d, float eps,
                                      float p_drop, uint64_t seed, uint32_t site, void* stream) {
  MASR_REQUIRE(d > 0 && d <= 32 * LN_MAX_PER_LANE, "layernorm: d must be <= 1024");
  if (rows == 0) return MASR_OK;
  const int threads = 256, wpb = threads / 32;
  const int blocks = int(std::min<int64_t>(ceil_div64(rows, wpb), int64_t(sm_count()) * 8));
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const bool vec = (d % 8 == 0) && ((reinterpret_cast<uintptr_t>(x_inout) | reinterpret_cast<uintptr_t>(res) |
                                     reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                                     reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (vec) {
#define LN_FWD_VEC(NCH) MASR_DISPATCH_DTYPE(dtype, T, (launch_pdl(add_layernorm_fwd_vec_kernel<T, NCH>, dim3(blocks), dim3(threads), 0, as_stream(stream),  \
          static_cast<T*>(x_inout), static_cast<const T*>(res), gamma, beta, static_cast<T*>(y), mean, rstd, rows, d, eps, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site)))
    if (d <= 256) LN_FWD_VEC(1); else if (d <= 512) LN_FWD_VEC(2); else LN_FWD_VEC(4);
#undef LN_FWD_VEC
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(add_layernorm_fwd_kernel<T>, dim3(blocks), dim3(threads), 0, as_stream(stream), 
          static_cast<T*>(x_inout), static_cast<const T*>(res), gamma, beta, static_cast<T*>(y), mean, rstd,
          rows, d, eps, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_add_layernorm_bwd(const void* dy, const void* s, const float* mean, const float* rstd,
                                      const float* gamma, void* ds, int ds_accum, void* dx,
                                      float* dgamma, float* dbeta, int dtype, int rows, int d,
                                      float p_drop, uint64_t seed, uint32_t site, void* stream) {
  MASR_REQUIRE(d > 0 && d <= 32 * LN_MAX_PER_LANE, "layernorm: d must be <= 1024");
  if (rows == 0) return MASR_OK;
  const int threads = 256, wpb = threads / 32;
  const int blocks = int(std::min<int64_t>(ceil_div64(rows, wpb), int64_t(sm_count()) * 2));
  const size_t smem = size_t(wpb) * 2 * d * sizeof(float);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  if (smem > 48 * 1024) {       // d > 768: opt in to more than the default 48 KB of dynamic shared memory
    MASR_CHECK_CUDA(cudaFuncSetAttribute(add_layernorm_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    MASR_CHECK_CUDA(cudaFuncSetAttribute(add_layernorm_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  }
  const bool vec = (d % 8 == 0) && smem <= 48 * 1024 &&
                   ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(ds) |
                     reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gamma)) & 15) == 0;
  if (vec) {
#define LN_BWD_VEC(NCH) MASR_DISPATCH_DTYPE(dtype, T, (launch_pdl(add_layernorm_bwd_vec_kernel<T, NCH>, dim3(blocks), dim3(threads), smem, as_stream(stream),  \
          static_cast<const T*>(dy), static_cast<const T*>(s), mean, rstd, gamma, static_cast<T*>(ds), ds_accum, static_cast<T*>(dx), \
          dgamma, dbeta, rows, d, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site)))
    if (d <= 256) LN_BWD_VEC(1); else if (d <= 512) LN_BWD_VEC(2); else LN_BWD_VEC(4);
#undef LN_BWD_VEC
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(add_layernorm_bwd_kernel<T>, dim3(blocks), dim3(threads), smem, as_stream(stream), 
          static_cast<const T*>(dy), static_cast<const T*>(s), mean, rstd, gamma, static_cast<T*>(ds), ds_accum,
          static_cast<T*>(dx), dgamma, dbeta, rows, d, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_add_pe_dropout(void* x, const float* pe, int dtype, int rows, int L, int d,
                                   float p_drop, uint64_t seed, uint32_t site, void* stream) {
  const int64_t n = int64_t(rows) * d;
  if (n == 0) return MASR_OK;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(add_pe_dropout_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), 
          static_cast<T*>(x), pe, n, L, d, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_embed_pe_fwd(const int64_t* ids, const float* E, const float* pe, void* out, int dtype,
                                 int rows, int L, int d, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  const int64_t n = int64_t(rows) * d;
  if (n == 0) return MASR_OK;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(embed_pe_fwd_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), 
          ids, E, pe, static_cast<T*>(out), n, L, d, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_embed_bwd(const int64_t* ids, const void* dout, int dtype, float* dE, int rows, int L, int d,
                              float p_drop, uint64_t seed, uint32_t site, void* stream) {
  (void)L;
  const int64_t n = int64_t(rows) * d;
  if (n == 0) return MASR_OK;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(embed_bwd_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), 
          ids, static_cast<const T*>(dout), dE, n, d, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_dropout(void* x, int dtype, int64_t n, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  if (n == 0 || p_drop <= 0.f) return MASR_OK;
  const float inv_keep = 1.f / (1.f - p_drop);
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(dropout_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), static_cast<T*>(x), n, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_colsum_add(const void* x, int dtype, int64_t ldx, float* out, int M, int N, void* stream) {
  if (M == 0 || N == 0) return MASR_OK;
  const int esz = dtype == MASR_BF16 ? 2 : 4;
  if (N % 8 == 0 && N <= 2048 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx * esz) % 16 == 0) {
    const int rl_count = 256 / (N / 8);
    const int blocks = int(std::min<int64_t>(ceil_div64(M, int64_t(rl_count) * 4), int64_t(sm_count()) * 4));
    MASR_DISPATCH_DTYPE(dtype, T,
        launch_pdl(colsum_vec_kernel<T>, dim3(std::max(blocks, 1)), dim3(256), 0, as_stream(stream), static_cast<const T*>(x), ldx, out, M, N));
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  dim3 block(32, 8);
  const int gy = int(std::min<int64_t>(ceil_div64(M, 8 * 8), 64));
  dim3 grid(unsigned(ceil_div64(N, 32)), unsigned(gy));
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(colsum_add_kernel<T>, dim3(grid), dim3(block), 0, as_stream(stream), static_cast<const T*>(x), ldx, out, M, N));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  if (n == 0) return MASR_OK;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(n, 256);
  if (src_dtype == MASR_F32 && dst_dtype == MASR_BF16)
    launch_pdl(cast_kernel<float, __nv_bfloat16>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), n);
  else if (src_dtype == MASR_BF16 && dst_dtype == MASR_F32)
    launch_pdl(cast_kernel<__nv_bfloat16, float>, dim3(g), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst), n);
  else if (src_dtype == MASR_F32 && dst_dtype == MASR_F32)
    launch_pdl(cast_kernel<float, float>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), static_cast<float*>(dst), n);
  else if (src_dtype == MASR_BF16 && dst_dtype == MASR_BF16)
    launch_pdl(cast_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(g), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), n);
  else { set_error("masr_cast: bad dtype"); return MASR_E_INVALID; }
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_cast_pad2d(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                               int rows, int cols, void* stream) {
  MASR_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= cols, "cast_pad2d: bad sizes");
  const int64_t n = int64_t(rows) * ld_dst;
  if (n == 0) return MASR_OK;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(n, 256);
  if (src_dtype == MASR_F32 && dst_dtype == MASR_BF16)
    launch_pdl(cast_pad2d_kernel<float, __nv_bfloat16>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), ld_src,
               static_cast<__nv_bfloat16*>(dst), ld_dst, int64_t(rows), cols);
  else if (src_dtype == MASR_F32 && dst_dtype == MASR_F32)
    launch_pdl(cast_pad2d_kernel<float, float>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), ld_src,
               static_cast<float*>(dst), ld_dst, int64_t(rows), cols);
  else { set_error("masr_cast_pad2d: fp32 source, fp32 / bf16 destination"); return MASR_E_INVALID; }
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_loss_mix(double* stats, const float* ctc_loss, float w, void* stream) {
  MASR_REQUIRE(stats != nullptr && ctc_loss != nullptr, "loss_mix: null pointer");
  launch_pdl(loss_mix_kernel, dim3(1), dim3(32), 0, as_stream(stream), stats, ctc_loss, w);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_permute_cf(const void* src, int src_dtype, void* dst, int dst_dtype, int rows, int C, int F,
                               int inverse_add, void* stream) {
  const int64_t n = int64_t(rows) * C * F;
  if (n == 0) return MASR_OK;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(n, 256);
  if (src_dtype == MASR_F32 && dst_dtype == MASR_F32)
    launch_pdl(permute_cf_kernel<float, float>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), static_cast<float*>(dst), rows, C, F, inverse_add);
  else if (src_dtype == MASR_F32 && dst_dtype == MASR_BF16 && !inverse_add)
    launch_pdl(permute_cf_kernel<float, __nv_bfloat16>, dim3(g), dim3(256), 0, st, static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), rows, C, F, 0);
  else { set_error("masr_permute_cf: unsupported dtype combination"); return MASR_E_INVALID; }
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_prep_weights(const float* params, void* shadow, int64_t n, const masr_conv_prep_job* jobs, int njobs,
                                 const float* v2e, void* v2e_p, int v2e_rows, int C, int F, int dtype, void* stream) {
  MASR_REQUIRE(njobs >= 0 && njobs <= 4, "masr_prep_weights: at most four conv jobs");
  MASR_REQUIRE(shadow == nullptr || ((reinterpret_cast<uintptr_t>(params) & 15u) == 0 && (reinterpret_cast<uintptr_t>(shadow) & 15u) == 0),
               "masr_prep_weights: arena alignment");
  PrepParams p{};
  p.params = params; p.shadow = shadow; p.n = shadow != nullptr ? n : 0;
  p.njobs = njobs;
  for (int i = 0; i < njobs; ++i) p.jobs[i] = jobs[i];
  p.v2e = v2e; p.v2e_p = v2e_p; p.v2e_rows = v2e_p != nullptr ? v2e_rows : 0; p.C = C; p.F = F;
  int b = 0;
  if (p.n > 0) b += int(std::min<int64_t>(ceil_div64(p.n / 4 + 1, 256), int64_t(sm_count()) * 8));
  p.blk_end[0] = b;
  for (int i = 0; i < 4; ++i) {
    if (i < njobs) b += int(std::min<int64_t>(ceil_div64(int64_t(jobs[i].Cout) * jobs[i].Cin * 9, 256 * 2), 148));
    p.blk_end[1 + i] = b;
  }
  if (p.v2e_rows > 0) b += int(std::min<int64_t>(ceil_div64(int64_t(v2e_rows) * C * F, 256 * 4), int64_t(sm_count()) * 4));
  p.blk_end[5] = b;
  if (b == 0) return MASR_OK;
  if (dtype == MASR_BF16) launch_pdl(prep_weights_kernel<__nv_bfloat16>, dim3(b), dim3(256), 0, as_stream(stream), p);
  else if (dtype == MASR_F32) launch_pdl(prep_weights_kernel<float>, dim3(b), dim3(256), 0, as_stream(stream), p);
  else { set_error("masr_prep_weights: bad dtype"); return MASR_E_INVALID; }
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_ls_ce_fwd_bwd(const float* logits, const int64_t* gold, int N, int C, float eps, float inv_n,
                                  const float* inv_n_dev, double* stats, int64_t* argmax, void* dlogits, int dl_dtype,
                                  int64_t ld_dl, void* stream) {
  MASR_REQUIRE(C > 0, "ls_ce: C must be positive");
  MASR_REQUIRE(dlogits == nullptr || ld_dl >= C, "ls_ce: dlogits leading dimension must be >= C");
  if (N == 0) return MASR_OK;
  const int threads = 128, wpb = threads / 32;
  const int blocks = int(std::min<int64_t>(ceil_div64(N, wpb), int64_t(sm_count()) * 8));
  if (dl_dtype == MASR_BF16) {
    MASR_CHECK_CUDA(launch_pdl(ls_ce_kernel<__nv_bfloat16>, dim3(blocks), dim3(threads), 0, as_stream(stream), logits, gold, N, C,
                               eps, inv_n, inv_n_dev, stats, argmax, static_cast<__nv_bfloat16*>(dlogits), ld_dl));
  } else {
    MASR_CHECK_CUDA(launch_pdl(ls_ce_kernel<float>, dim3(blocks), dim3(threads), 0, as_stream(stream), logits, gold, N, C,
                               eps, inv_n, inv_n_dev, stats, argmax, static_cast<float*>(dlogits), ld_dl));
  }
  return MASR_OK;
}
