// Multi-head attention forward / backward with the masks generated from lengths in-kernel.
//
// Replaces the SDPA call inside torch's multi_head_attention_forward as used by
// nn.TransformerEncoder/Decoder in mono_transformer_torch.py:74-98,200-203:
//   - key-padding mask (make_bool_pad_mask, src/nets_utils.py:85-94)  -> klens[b]
//   - causal mask (generate_square_subsequent_mask, :9-15)            -> causal flag
// No mask tensor is ever materialised.  Attention is < 2 % of the model FLOPs, so this
// is a CUDA-core flash-style kernel: scores never touch HBM; one pass over K/V per query tile.
//
// Work split: a block owns QT query rows of one (batch, head); K/V tiles of KT keys are staged in
// shared memory as fp32; inside a warp, lane j scores key j of a 32-key chunk (full dot product
// against the query row held in shared memory), the soft-max statistics are combined with warp
// shuffles, and for P*V each lane owns output dims {lane, lane+32}.
#include "common.cuh"

namespace masr {

constexpr int AT_QT = 16;        // query rows per block (4 warps x 4 rows)
constexpr int AT_KT = 64;        // keys per shared-memory tile
constexpr int AT_THREADS = 128;
constexpr int AT_RPW = AT_QT / (AT_THREADS / 32);
constexpr int AT_LD = 65;        // padded row stride (floats): conflict-free row-per-lane reads

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* __restrict__ src, int64_t ld, int row0, int nrows_valid,
                                          int nrows_tile, int hd) {
  // dst[r][c] (stride AT_LD) = src[(row0 + r) * ld + c] for r < nrows_valid, 0 otherwise
  for (int e = threadIdx.x; e < nrows_tile * hd; e += blockDim.x) {
    const int r = e / hd, c = e % hd;
    dst[r * AT_LD + c] = (r < nrows_valid) ? to_f<T>(src[int64_t(row0 + r) * ld + c]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                const T* __restrict__ v, int64_t ldv, T* __restrict__ out, int64_t ldo,
                float* __restrict__ lse, int B, int H, int Lq, int Lk, int kv_rows, int hd,
                const int64_t* __restrict__ klens, int causal, float scale,
                float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  const uint64_t seed = resolve_seed(seed_arg);
  __shared__ float Qs[AT_QT * AT_LD];
  __shared__ float Ks[AT_KT * AT_LD];
  __shared__ float Vs[AT_KT * AT_LD];
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * AT_QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qb = q + int64_t(b) * Lq * ldq + h * hd;
  const T* kb = k + int64_t(b) * kv_rows * ldk + h * hd;      // kv_rows = rows per utterance in K / V (Lk, or a cache's capacity)
  const T* vb = v + int64_t(b) * kv_rows * ldv + h * hd;
  int kmax = Lk;
  if (klens != nullptr) kmax = min(kmax, int(klens[b]));
  if (causal) kmax = min(kmax, q0 + AT_QT);
  load_tile<T>(Qs, qb, ldq, q0, min(AT_QT, Lq - q0), AT_QT, hd);

  float m[AT_RPW], l[AT_RPW], acc0[AT_RPW], acc1[AT_RPW];
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) { m[r] = -INFINITY; l[r] = 0.f; acc0[r] = 0.f; acc1[r] = 0.f; }

  for (int k0 = 0; k0 < kmax; k0 += AT_KT) {
    __syncthreads();
    const int nk = min(AT_KT, kmax - k0);
    load_tile<T>(Ks, kb, ldk, k0, nk, AT_KT, hd);
    load_tile<T>(Vs, vb, ldv, k0, nk, AT_KT, hd);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AT_RPW; ++r) {
      const int qi = q0 + warp * AT_RPW + r;
      if (qi >= Lq) continue;
      const float* qrow = Qs + (warp * AT_RPW + r) * AT_LD;
      for (int c0 = 0; c0 < nk; c0 += 32) {
        const int j = c0 + lane;            // key within tile
        const int kj = k0 + j;              // absolute key index
        const bool valid = (j < nk) && (!causal || kj <= qi);
        float s = -INFINITY;
        if (valid) {
          const float* krow = Ks + j * AT_LD;
          float d = 0.f;
          for (int c = 0; c < hd; ++c) d = fmaf(qrow[c], krow[c], d);
          s = d * scale;
        }
        const float cm = warp_max(s);
        if (cm == -INFINITY) continue;      // whole chunk masked
        const float mn = fmaxf(m[r], cm);
        const float corr = expf(m[r] - mn);      // exp(-inf) = 0 on the first chunk
        const float pj = valid ? expf(s - mn) : 0.f;
        l[r] = l[r] * corr + warp_sum(pj);
        float pd = pj;
        if (p > 0.f && valid)
          pd = pj * attn_drop_scale(p, seed, site, uint32_t(bh) * uint32_t(Lq) + uint32_t(qi), kj);
        float a0 = acc0[r] * corr, a1 = acc1[r] * corr;
        const int jn = min(32, nk - c0);
        for (int jj = 0; jj < jn; ++jj) {
          const float pb = __shfl_sync(0xffffffffu, pd, jj);
          const float* vrow = Vs + (c0 + jj) * AT_LD;
          if (lane < hd) a0 = fmaf(pb, vrow[lane], a0);
          if (lane + 32 < hd) a1 = fmaf(pb, vrow[lane + 32], a1);
        }
        acc0[r] = a0; acc1[r] = a1; m[r] = mn;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) {
    const int qi = q0 + warp * AT_RPW + r;
    if (qi >= Lq) continue;
    const float inv_l = l[r] > 0.f ? 1.f / l[r] : 0.f;
    T* orow = out + (int64_t(b) * Lq + qi) * ldo + h * hd;
    if (lane < hd) orow[lane] = from_f<T>(acc0[r] * inv_l);
    if (lane + 32 < hd) orow[lane + 32] = from_f<T>(acc1[r] * inv_l);
    if (lane == 0) lse[int64_t(bh) * Lq + qi] = l[r] > 0.f ? m[r] + logf(l[r]) : -INFINITY;
  }
}

// dQ (and D_i = dO_i . O_i, written to dsum for the dK/dV kernel)
template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                   const T* __restrict__ v, int64_t ldv, const T* __restrict__ out, int64_t ldo,
                   const T* __restrict__ dout, int64_t lddo, const float* __restrict__ lse,
                   float* __restrict__ dsum, T* __restrict__ dq, int64_t lddq,
                   int B, int H, int Lq, int Lk, int hd, const int64_t* __restrict__ klens, int causal,
                   float scale, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  const uint64_t seed = resolve_seed(seed_arg);
  __shared__ float Qs[AT_QT * AT_LD];
  __shared__ float dOs[AT_QT * AT_LD];
  __shared__ float Ks[AT_KT * AT_LD];
  __shared__ float Vs[AT_KT * AT_LD];
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * AT_QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* kb = k + int64_t(b) * Lk * ldk + h * hd;
  const T* vb = v + int64_t(b) * Lk * ldv + h * hd;
  int kmax = Lk;
  if (klens != nullptr) kmax = min(kmax, int(klens[b]));
  if (causal) kmax = min(kmax, q0 + AT_QT);
  const int nq = min(AT_QT, Lq - q0);
  load_tile<T>(Qs, q + int64_t(b) * Lq * ldq + h * hd, ldq, q0, nq, AT_QT, hd);
  load_tile<T>(dOs, dout + int64_t(b) * Lq * lddo + h * hd, lddo, q0, nq, AT_QT, hd);
  __syncthreads();

  float D[AT_RPW], L[AT_RPW], a0[AT_RPW], a1[AT_RPW];
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) {
    const int qi = q0 + warp * AT_RPW + r;
    D[r] = 0.f; L[r] = 0.f; a0[r] = 0.f; a1[r] = 0.f;
    if (qi < Lq) {
      const T* orow = out + (int64_t(b) * Lq + qi) * ldo + h * hd;
      const float* dorow = dOs + (warp * AT_RPW + r) * AT_LD;
      float t = 0.f;
      if (lane < hd) t += to_f<T>(orow[lane]) * dorow[lane];
      if (lane + 32 < hd) t += to_f<T>(orow[lane + 32]) * dorow[lane + 32];
      D[r] = warp_sum(t);
      L[r] = lse[int64_t(bh) * Lq + qi];
      if (lane == 0) dsum[int64_t(bh) * Lq + qi] = D[r];
    }
  }

  for (int k0 = 0; k0 < kmax; k0 += AT_KT) {
    __syncthreads();
    const int nk = min(AT_KT, kmax - k0);
    load_tile<T>(Ks, kb, ldk, k0, nk, AT_KT, hd);
    load_tile<T>(Vs, vb, ldv, k0, nk, AT_KT, hd);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < AT_RPW; ++r) {
      const int qi = q0 + warp * AT_RPW + r;
      if (qi >= Lq) continue;
      const float* qrow = Qs + (warp * AT_RPW + r) * AT_LD;
      const float* dorow = dOs + (warp * AT_RPW + r) * AT_LD;
      for (int c0 = 0; c0 < nk; c0 += 32) {
        const int j = c0 + lane, kj = k0 + j;
        const bool valid = (j < nk) && (!causal || kj <= qi);
        float ds = 0.f;
        if (valid) {
          const float* krow = Ks + j * AT_LD;
          const float* vrow = Vs + j * AT_LD;
          float d = 0.f, dp = 0.f;
          for (int c = 0; c < hd; ++c) { d = fmaf(qrow[c], krow[c], d); dp = fmaf(dorow[c], vrow[c], dp); }
          const float pj = expf(d * scale - L[r]);
          const float dm = attn_drop_scale(p, seed, site, uint32_t(bh) * uint32_t(Lq) + uint32_t(qi), kj);
          ds = pj * (dp * dm - D[r]) * scale;
        }
        const int jn = min(32, nk - c0);
        float x0 = a0[r], x1 = a1[r];
        for (int jj = 0; jj < jn; ++jj) {
          const float dsb = __shfl_sync(0xffffffffu, ds, jj);
          const float* krow = Ks + (c0 + jj) * AT_LD;
          if (lane < hd) x0 = fmaf(dsb, krow[lane], x0);
          if (lane + 32 < hd) x1 = fmaf(dsb, krow[lane + 32], x1);
        }
        a0[r] = x0; a1[r] = x1;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) {
    const int qi = q0 + warp * AT_RPW + r;
    if (qi >= Lq) continue;
    T* drow = dq + (int64_t(b) * Lq + qi) * lddq + h * hd;
    if (lane < hd) drow[lane] = from_f<T>(a0[r]);
    if (lane + 32 < hd) drow[lane + 32] = from_f<T>(a1[r]);
  }
}

// dK, dV: a block owns AT_QT key rows; query tiles (Q, dO) are streamed through shared memory and
// lane i scores query i of a 32-query chunk.
template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dkv_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                    const T* __restrict__ v, int64_t ldv, const T* __restrict__ dout, int64_t lddo,
                    const float* __restrict__ lse, const float* __restrict__ dsum,
                    T* __restrict__ dk, int64_t lddk, T* __restrict__ dv, int64_t lddv,
                    int B, int H, int Lq, int Lk, int hd, const int64_t* __restrict__ klens, int causal,
                    float scale, float p, float inv_keep, SeedArg seed_arg, uint32_t site) {
  const uint64_t seed = resolve_seed(seed_arg);
  __shared__ float Ks[AT_QT * AT_LD];
  __shared__ float Vs[AT_QT * AT_LD];
  __shared__ float Qs[AT_KT * AT_LD];
  __shared__ float dOs[AT_KT * AT_LD];
  __shared__ float Ls[AT_KT], Ds[AT_KT];
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int j0 = blockIdx.x * AT_QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* qb = q + int64_t(b) * Lq * ldq + h * hd;
  const T* dob = dout + int64_t(b) * Lq * lddo + h * hd;
  const int nkeys = min(AT_QT, Lk - j0);
  const int klen = klens != nullptr ? min(Lk, int(klens[b])) : Lk;
  load_tile<T>(Ks, k + int64_t(b) * Lk * ldk + h * hd, ldk, j0, nkeys, AT_QT, hd);
  load_tile<T>(Vs, v + int64_t(b) * Lk * ldv + h * hd, ldv, j0, nkeys, AT_QT, hd);

  float k0a[AT_RPW], k1a[AT_RPW], v0a[AT_RPW], v1a[AT_RPW];
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) { k0a[r] = 0.f; k1a[r] = 0.f; v0a[r] = 0.f; v1a[r] = 0.f; }

  const bool block_active = j0 < klen;               // every key of the block padded -> zeros
  const int i_begin = causal ? (j0 / AT_KT) * AT_KT : 0;   // queries i < j never see key j
  if (block_active) {
    for (int i0 = i_begin; i0 < Lq; i0 += AT_KT) {
      __syncthreads();
      const int nq = min(AT_KT, Lq - i0);
      load_tile<T>(Qs, qb, ldq, i0, nq, AT_KT, hd);
      load_tile<T>(dOs, dob, lddo, i0, nq, AT_KT, hd);
      for (int e = threadIdx.x; e < AT_KT; e += blockDim.x) {
        Ls[e] = (e < nq) ? lse[int64_t(bh) * Lq + i0 + e] : 0.f;
        Ds[e] = (e < nq) ? dsum[int64_t(bh) * Lq + i0 + e] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < AT_RPW; ++r) {
        const int kj = j0 + warp * AT_RPW + r;
        if (kj >= klen) continue;
        const float* krow = Ks + (warp * AT_RPW + r) * AT_LD;
        const float* vrow = Vs + (warp * AT_RPW + r) * AT_LD;
        for (int c0 = 0; c0 < nq; c0 += 32) {
          const int i = c0 + lane, qi = i0 + i;
          const bool valid = (i < nq) && (!causal || kj <= qi);
          float ds = 0.f, pm = 0.f;
          if (valid) {
            const float* qrow = Qs + i * AT_LD;
            const float* dorow = dOs + i * AT_LD;
            float d = 0.f, dp = 0.f;
            for (int c = 0; c < hd; ++c) { d = fmaf(qrow[c], krow[c], d); dp = fmaf(dorow[c], vrow[c], dp); }
            const float pj = expf(d * scale - Ls[i]);
            const float dm = attn_drop_scale(p, seed, site, uint32_t(bh) * uint32_t(Lq) + uint32_t(qi), kj);
            pm = pj * dm;
            ds = pj * (dp * dm - Ds[i]) * scale;
          }
          const int in = min(32, nq - c0);
          float x0 = k0a[r], x1 = k1a[r], y0 = v0a[r], y1 = v1a[r];
          for (int ii = 0; ii < in; ++ii) {
            const float dsb = __shfl_sync(0xffffffffu, ds, ii);
            const float pmb = __shfl_sync(0xffffffffu, pm, ii);
            const float* qrow = Qs + (c0 + ii) * AT_LD;
            const float* dorow = dOs + (c0 + ii) * AT_LD;
            if (lane < hd) { x0 = fmaf(dsb, qrow[lane], x0); y0 = fmaf(pmb, dorow[lane], y0); }
            if (lane + 32 < hd) { x1 = fmaf(dsb, qrow[lane + 32], x1); y1 = fmaf(pmb, dorow[lane + 32], y1); }
          }
          k0a[r] = x0; k1a[r] = x1; v0a[r] = y0; v1a[r] = y1;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < AT_RPW; ++r) {
    const int kj = j0 + warp * AT_RPW + r;
    if (kj >= Lk) continue;
    T* dkr = dk + (int64_t(b) * Lk + kj) * lddk + h * hd;
    T* dvr = dv + (int64_t(b) * Lk + kj) * lddv + h * hd;
    if (lane < hd) { dkr[lane] = from_f<T>(k0a[r]); dvr[lane] = from_f<T>(v0a[r]); }
    if (lane + 32 < hd) { dkr[lane + 32] = from_f<T>(k1a[r]); dvr[lane + 32] = from_f<T>(v1a[r]); }
  }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* out, int64_t ldo, float* lse, int dtype,
                             int B, int H, int Lq, int Lk, int hd, const int64_t* klens, int causal,
                             float p_drop, uint64_t seed, uint32_t site, void* stream) {
  return masr_attn_fwd_cached(q, ldq, k, ldk, v, ldv, out, ldo, lse, dtype, B, H, Lq, Lk, Lk, hd, klens, causal, p_drop, seed,
                              site, stream);
}

extern "C" int masr_attn_fwd_cached(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                    void* out, int64_t ldo, float* lse, int dtype,
                                    int B, int H, int Lq, int Lk, int kv_rows, int hd, const int64_t* klens, int causal,
                                    float p_drop, uint64_t seed, uint32_t site, void* stream) {
  MASR_REQUIRE(hd > 0 && hd <= 64, "attention: head dim must be <= 64");
  MASR_REQUIRE(kv_rows >= Lk, "attention: kv_rows (rows per utterance of K / V) must be >= Lk");
  if (B == 0 || H == 0 || Lq == 0) return MASR_OK;
  dim3 grid(unsigned(ceil_div64(Lq, AT_QT)), unsigned(B * H));
  const float scale = 1.0f / sqrtf(float(hd));
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  MASR_DISPATCH_DTYPE(dtype, T,
      attn_fwd_kernel<T><<<grid, AT_THREADS, 0, as_stream(stream)>>>(
          static_cast<const T*>(q), ldq, static_cast<const T*>(k), ldk, static_cast<const T*>(v), ldv,
          static_cast<T*>(out), ldo, lse, B, H, Lq, Lk, kv_rows, hd, klens, causal, scale, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                             float* dsum_ws, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                             int dtype, int B, int H, int Lq, int Lk, int hd, const int64_t* klens, int causal,
                             float p_drop, uint64_t seed, uint32_t site, void* stream) {
  MASR_REQUIRE(hd > 0 && hd <= 64, "attention: head dim must be <= 64");
  MASR_REQUIRE(dsum_ws != nullptr, "attention backward needs a [B*H*Lq] fp32 workspace");
  if (B == 0 || H == 0) return MASR_OK;
  const float scale = 1.0f / sqrtf(float(hd));
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  cudaStream_t st = as_stream(stream);
  if (Lq > 0) {
    dim3 grid(unsigned(ceil_div64(Lq, AT_QT)), unsigned(B * H));
    MASR_DISPATCH_DTYPE(dtype, T,
        attn_bwd_dq_kernel<T><<<grid, AT_THREADS, 0, st>>>(
            static_cast<const T*>(q), ldq, static_cast<const T*>(k), ldk, static_cast<const T*>(v), ldv,
            static_cast<const T*>(out), ldo, static_cast<const T*>(dout), lddo, lse, dsum_ws,
            static_cast<T*>(dq), lddq, B, H, Lq, Lk, hd, klens, causal, scale, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
    MASR_LAUNCH_CHECK();
  }
  if (Lk > 0) {
    dim3 grid(unsigned(ceil_div64(Lk, AT_QT)), unsigned(B * H));
    MASR_DISPATCH_DTYPE(dtype, T,
        attn_bwd_dkv_kernel<T><<<grid, AT_THREADS, 0, st>>>(
            static_cast<const T*>(q), ldq, static_cast<const T*>(k), ldk, static_cast<const T*>(v), ldv,
            static_cast<const T*>(dout), lddo, lse, dsum_ws, static_cast<T*>(dk), lddk, static_cast<T*>(dv), lddv,
            B, H, Lq, Lk, hd, klens, causal, scale, p_drop, inv_keep, SeedArg{seed, g_seed_dev_ptr}, site));
    MASR_LAUNCH_CHECK();
  }
  return MASR_OK;
}
