// Multi-head attention forward / backward for SHORT query sequences (Lq <= 64, head dim 64, bf16): the decoder's causal
// self-attention and its attention over the encoder memory (mono_transformer_torch.py:200-203: tgt length L+1 = 33 at the
// benchmark shape, against 33 or T' = 128..375 keys).
//
// Same contract as attn_umma.cu (masks from lengths, the shared attention dropout hash, lse in natural log, D = rowsum(dO.O)
// optionally supplied by the out-projection dgrad).  These problems are 0.1 % of the step's FLOPs but were 8 of the 10
// attention launches of a batch: on the tcgen05 kernel a 33-row problem still pays the 128-row tile pipeline (TMA ->
// tcgen05.mma -> tcgen05.ld -> soft-max -> shared memory -> tcgen05.mma -> tcgen05.ld, ~9 us forward / 15-17 us backward
// per launch, one CTA per SM).  Here one CTA of four warps owns one (batch, head); every warp owns 16 query rows and keeps
// S / P / dP / dS in mma.sync (m16n8k16) accumulator registers, FlashAttention-2 style: no TMEM round trips, no barriers
// inside a key tile, 24-48 KB of shared memory (several CTAs per SM).  Longer query sequences (the encoder) stay on tcgen05.
#include "common.cuh"

namespace masr {

namespace {

constexpr int AS_THREADS = 128;
constexpr int AS_QMAX = 64;          // 4 warps x 16 rows
constexpr int AS_KT = 64;            // keys per shared-memory tile
constexpr uint32_t AS_TILE = 64 * 128;

struct AttnSmallParams {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  int64_t ldq, ldk, ldv, ldo, lddo;
  __nv_bfloat16 *out, *dq, *dk, *dv;
  int64_t ldout, lddq, lddk, lddv;
  float* lse;                    // fwd: written; bwd: read
  const float* dsum;             // bwd: D rows (NULL: computed here from o / dout)
  int B, H, Lq, Lk, kv_rows;
  const int64_t* klens;
  int causal;
  float scale, p_drop, inv_keep;
  uint32_t thr16;
  uint64_t seed; uint32_t site;
  const uint64_t* seed_ptr;
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// [rows][64 bf16] tile with 128-byte rows; the 16-byte chunk c of row r lives at chunk c ^ (r & 7) (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t toff(int r, int c) { return uint32_t(r) * 128u + (uint32_t((c ^ r) & 7) << 4); }

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// rows [row0, row0 + 64) of a [*, ld] bf16 matrix (64 columns from col0) -> swizzled tile; rows >= nvalid are zero
__device__ __forceinline__ void load_tile64(unsigned char* dst, const __nv_bfloat16* __restrict__ src, int64_t ld, int nvalid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * AS_THREADS;
    const int r = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < nvalid) v = __ldg(reinterpret_cast<const uint4*>(src + int64_t(r) * ld + c * 8));
    *reinterpret_cast<uint4*>(dst + toff(r, c)) = v;
  }
}

// A fragments (16 rows x 64 k) of rows [r0, r0 + 16) of a tile: four k-steps
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t tile, int r0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldsm4(f[ks], tile + toff(r0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(AS_THREADS) attn_small_fwd_kernel(const AttnSmallParams p) {
  __shared__ __align__(128) unsigned char sQ[AS_TILE];
  __shared__ __align__(128) unsigned char sK[AS_TILE];
  __shared__ __align__(128) unsigned char sV[AS_TILE];
  pdl_launch_dependents();
  pdl_wait();
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const uint64_t seed_eff = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
  int kmax = p.Lk;
  if (p.klens != nullptr) kmax = min(kmax, int(p.klens[b]));
  const int kend = p.causal ? min(kmax, p.Lq) : kmax;             // no query sees a key beyond this
  const int ntiles = (kend + AS_KT - 1) / AS_KT;
  const __nv_bfloat16* qb = p.q + int64_t(b) * p.Lq * p.ldq + h * 64;
  const __nv_bfloat16* kb = p.k + int64_t(b) * p.kv_rows * p.ldk + h * 64;
  const __nv_bfloat16* vb = p.v + int64_t(b) * p.kv_rows * p.ldv + h * 64;

  load_tile64(sQ, qb, p.ldq, p.Lq);
  if (ntiles > 0) {
    load_tile64(sK, kb, p.ldk, min(AS_KT, kend));
    load_tile64(sV, vb, p.ldv, min(AS_KT, kend));
  }
  __syncthreads();
  const int rw = warp * 16;
  const bool active = rw < p.Lq;                                   // warps past the last query row only help loading
  uint32_t qf[4][4];
  load_a_frags(qf, s_u32(sQ), rw, lane);
  const int row[2] = {rw + g, rw + g + 8};
  int lim[2];
  uint32_t rowkey[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    lim[e] = row[e] < p.Lq ? (p.causal ? min(kmax, row[e] + 1) : kmax) : 0;
    rowkey[e] = attn_drop_rowkey(seed_eff, p.site, uint32_t(bh) * uint32_t(p.Lq) + uint32_t(row[e]));
  }
  const float sl2 = p.scale * 1.4426950408889634f;
  const bool drop = p.p_drop > 0.f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int t = 0; t < ntiles; ++t) {
    if (t > 0) {
      __syncthreads();
      const int nv = min(AS_KT, kend - t * AS_KT);
      load_tile64(sK, kb + int64_t(t) * AS_KT * p.ldk, p.ldk, nv);
      load_tile64(sV, vb + int64_t(t) * AS_KT * p.ldv, p.ldv, nv);
      __syncthreads();
    }
    if (!active) continue;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kf[4];
        ldsm4(kf, s_u32(sK) + toff(n2 * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
        mma16816(s[2 * n2], qf[ks], kf[0], kf[1]);
        mma16816(s[2 * n2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // masks + row maxima (log2 domain)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = t * AS_KT + i * 8 + 2 * t4 + (e & 1);
        const float v = key < lim[e >> 1] ? s[i][e] * sl2 : -INFINITY;
        s[i][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 1));
      mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 2));
      const float m_new = fmaxf(m_run[e], mx[e]);
      m_use[e] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[e] = ex2(m_run[e] - m_use[e]);
      m_run[e] = m_new;
      l_run[e] *= corr[e];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        float p0 = ex2(s[i][2 * e2] - m_use[e2]), p1 = ex2(s[i][2 * e2 + 1] - m_use[e2]);
        l_run[e2] += p0 + p1;
        if (drop) {
          const uint32_t bits = attn_drop_pair(rowkey[e2], uint32_t(t * AS_KT + i * 8 + 2 * t4) >> 1);
          p0 = ((bits & 0xffffu) >= p.thr16) ? p0 * p.inv_keep : 0.f;
          p1 = ((bits >> 16) >= p.thr16) ? p1 * p.inv_keep : 0.f;
        }
        s[i][2 * e2] = p0; s[i][2 * e2 + 1] = p1;
      }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                       pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        uint32_t vf[4];
        ldsm4t(vf, s_u32(sV) + toff(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dn * 2 + (lane >> 4)));
        mma16816(o[2 * dn], a, vf[0], vf[1]);
        mma16816(o[2 * dn + 1], a, vf[2], vf[3]);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    l_run[e] += __shfl_xor_sync(0xffffffffu, l_run[e], 1);
    l_run[e] += __shfl_xor_sync(0xffffffffu, l_run[e], 2);
    if (row[e] < p.Lq) {
      const float inv_l = l_run[e] > 0.f ? 1.f / l_run[e] : 0.f;
      __nv_bfloat16* orow = p.out + (int64_t(b) * p.Lq + row[e]) * p.ldout + h * 64 + 2 * t4;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pack2(o[i][2 * e] * inv_l, o[i][2 * e + 1] * inv_l);
      if (t4 == 0)
        p.lse[int64_t(bh) * p.Lq + row[e]] = l_run[e] > 0.f ? (m_run[e] + log2f(l_run[e])) * 0.6931471805599453f : -INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr size_t AS_BWD_SMEM = 6 * AS_TILE + 128;

__global__ void __launch_bounds__(AS_THREADS) attn_small_bwd_kernel(const AttnSmallParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~uintptr_t(127));
  unsigned char* sQ = base;
  unsigned char* sdO = sQ + AS_TILE;
  unsigned char* sK = sdO + AS_TILE;
  unsigned char* sV = sK + AS_TILE;
  unsigned char* sP = sV + AS_TILE;             // [query][key] dropped probabilities of the tile (bf16)
  unsigned char* sdS = sP + AS_TILE;            // [query][key] dS
  pdl_launch_dependents();
  pdl_wait();
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const uint64_t seed_eff = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
  int kmax = p.Lk;
  if (p.klens != nullptr) kmax = min(kmax, int(p.klens[b]));
  const int kend = p.causal ? min(kmax, p.Lq) : kmax;
  const int ntiles_all = (p.Lk + AS_KT - 1) / AS_KT;             // every key row gets a gradient (zeros beyond kend)
  const __nv_bfloat16* qb = p.q + int64_t(b) * p.Lq * p.ldq + h * 64;
  const __nv_bfloat16* dob = p.dout + int64_t(b) * p.Lq * p.lddo + h * 64;
  const __nv_bfloat16* kb = p.k + int64_t(b) * p.Lk * p.ldk + h * 64;
  const __nv_bfloat16* vb = p.v + int64_t(b) * p.Lk * p.ldv + h * 64;

  load_tile64(sQ, qb, p.ldq, p.Lq);
  load_tile64(sdO, dob, p.lddo, p.Lq);
  load_tile64(sK, kb, p.ldk, min(AS_KT, kend));
  load_tile64(sV, vb, p.ldv, min(AS_KT, kend));
  __syncthreads();
  const int rw = warp * 16;
  const bool active = rw < p.Lq;
  const int nqs = (p.Lq + 15) >> 4;                              // 16-query steps of the dK / dV reductions
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(qf, s_u32(sQ), rw, lane);
  load_a_frags(dof, s_u32(sdO), rw, lane);
  const int row[2] = {rw + g, rw + g + 8};
  int lim[2];
  uint32_t rowkey[2];
  float lse2[2], dsum[2];
  const float sl2 = p.scale * 1.4426950408889634f;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const bool rv = row[e] < p.Lq;
    lim[e] = rv ? (p.causal ? min(kmax, row[e] + 1) : kmax) : 0;
    rowkey[e] = attn_drop_rowkey(seed_eff, p.site, uint32_t(bh) * uint32_t(p.Lq) + uint32_t(row[e]));
    const float l = rv ? p.lse[int64_t(bh) * p.Lq + row[e]] : 0.f;
    lse2[e] = l * 1.4426950408889634f;
    if (l == -INFINITY) lim[e] = 0;                               // a row without any visible key
    float d = 0.f;
    if (rv) {
      if (p.dsum != nullptr) {
        d = p.dsum[int64_t(bh) * p.Lq + row[e]];
      } else {                                                     // D = dO . O over the 64 dims: 16 per thread of the quad
        const __nv_bfloat16* orow = p.o + (int64_t(b) * p.Lq + row[e]) * p.ldo + h * 64 + t4 * 16;
        const __nv_bfloat16* drow = dob + int64_t(row[e]) * p.lddo + t4 * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) d = fmaf(__bfloat162float(orow[j]), __bfloat162float(drow[j]), d);
      }
    }
    if (p.dsum == nullptr) {
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
    }
    dsum[e] = d;
  }
  const bool drop = p.p_drop > 0.f;
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  for (int t = 0; t < ntiles_all; ++t) {
    const int k0 = t * AS_KT;
    const bool live = k0 < kend;                                  // tiles past the last visible key: zero gradients only
    if (t > 0) {
      __syncthreads();
      if (live) {
        const int nv = min(AS_KT, kend - k0);
        load_tile64(sK, kb + int64_t(k0) * p.ldk, p.ldk, nv);
        load_tile64(sV, vb + int64_t(k0) * p.ldv, p.ldv, nv);
      }
      __syncthreads();
    }
    if (live && active) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float s[2][4], dp[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kf[4], vf[4];
          const uint32_t off = toff(kk * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1));
          ldsm4(kf, s_u32(sK) + off);
          ldsm4(vf, s_u32(sV) + off);
          mma16816(s[0], qf[ks], kf[0], kf[1]);
          mma16816(s[1], qf[ks], kf[2], kf[3]);
          mma16816(dp[0], dof[ks], vf[0], vf[1]);
          mma16816(dp[1], dof[ks], vf[2], vf[3]);
        }
        uint32_t pa[4], da[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int key = k0 + kk * 16 + i * 8 + 2 * t4;
            float p0 = key < lim[e2] ? ex2(fmaf(s[i][2 * e2], sl2, -lse2[e2])) : 0.f;
            float p1 = key + 1 < lim[e2] ? ex2(fmaf(s[i][2 * e2 + 1], sl2, -lse2[e2])) : 0.f;
            float d0 = dp[i][2 * e2], d1 = dp[i][2 * e2 + 1];
            float pm0 = p0, pm1 = p1;
            if (drop) {
              const uint32_t bits = attn_drop_pair(rowkey[e2], uint32_t(key) >> 1);
              const float k0s = ((bits & 0xffffu) >= p.thr16) ? p.inv_keep : 0.f;
              const float k1s = ((bits >> 16) >= p.thr16) ? p.inv_keep : 0.f;
              pm0 *= k0s; pm1 *= k1s; d0 *= k0s; d1 *= k1s;
            }
            const float ds0 = p0 * (d0 - dsum[e2]) * p.scale, ds1 = p1 * (d1 - dsum[e2]) * p.scale;
            pa[i * 2 + e2] = pack2(pm0, pm1);
            da[i * 2 + e2] = pack2(ds0, ds1);
            const uint32_t so = toff(row[e2], kk * 2 + i) + uint32_t(t4) * 4u;
            *reinterpret_cast<uint32_t*>(sP + so) = pa[i * 2 + e2];
            *reinterpret_cast<uint32_t*>(sdS + so) = da[i * 2 + e2];
          }
        }
        // dQ += dS K  (K read [key][d] as the row-major B operand)
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t kt[4];
          ldsm4t(kt, s_u32(sK) + toff(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dn * 2 + (lane >> 4)));
          mma16816(dq[2 * dn], da, kt[0], kt[1]);
          mma16816(dq[2 * dn + 1], da, kt[2], kt[3]);
        }
      }
    }
    __syncthreads();                                              // P / dS of every query row are in shared memory
    // dK = dS^T Q, dV = P^T dO for the 16 keys of this warp
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    if (live) {
      for (int qs = 0; qs < nqs; ++qs) {
        uint32_t sa[4], pa[4];
        const uint32_t off = toff(qs * 16 + (lane & 7) + (lane >> 4) * 8, warp * 2 + ((lane >> 3) & 1));
        ldsm4t(sa, s_u32(sdS) + off);
        ldsm4t(pa, s_u32(sP) + off);
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t qt[4], dt[4];
          const uint32_t off2 = toff(qs * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dn * 2 + (lane >> 4));
          ldsm4t(qt, s_u32(sQ) + off2);
          ldsm4t(dt, s_u32(sdO) + off2);
          mma16816(dk[2 * dn], sa, qt[0], qt[1]);
          mma16816(dk[2 * dn + 1], sa, qt[2], qt[3]);
          mma16816(dv[2 * dn], pa, dt[0], dt[1]);
          mma16816(dv[2 * dn + 1], pa, dt[2], dt[3]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = k0 + warp * 16 + g + 8 * e;
      if (key < p.Lk) {
        __nv_bfloat16* kr = p.dk + (int64_t(b) * p.Lk + key) * p.lddk + h * 64 + 2 * t4;
        __nv_bfloat16* vr = p.dv + (int64_t(b) * p.Lk + key) * p.lddv + h * 64 + 2 * t4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<uint32_t*>(kr + i * 8) = pack2(dk[i][2 * e], dk[i][2 * e + 1]);
          *reinterpret_cast<uint32_t*>(vr + i * 8) = pack2(dv[i][2 * e], dv[i][2 * e + 1]);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (row[e] < p.Lq) {
      __nv_bfloat16* qr = p.dq + (int64_t(b) * p.Lq + row[e]) * p.lddq + h * 64 + 2 * t4;
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<uint32_t*>(qr + i * 8) = pack2(dq[i][2 * e], dq[i][2 * e + 1]);
    }
  }
}

int g_small_lq = AS_QMAX;

}  // namespace

bool attn_small_applicable(int Lq) { return Lq > 0 && Lq <= g_small_lq; }

int attn_small_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                   float* lse, int B, int H, int Lq, int Lk, int kv_rows, const int64_t* klens, int causal, float p_drop,
                   uint64_t seed, uint32_t site, cudaStream_t st) {
  MASR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 &&
               ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 3) == 0, "short-query attention: operands must be 16 B aligned");
  const uint32_t thr16 = attn_drop_thr16(p_drop);
  AttnSmallParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k); p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldout = ldo; p.lse = lse;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.kv_rows = kv_rows; p.klens = klens; p.causal = causal;
  p.scale = 0.125f; p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? attn_drop_inv_keep(thr16) : 1.f; p.thr16 = thr16;
  p.seed = seed; p.site = site; p.seed_ptr = g_seed_dev_ptr;
  MASR_CHECK_CUDA(launch_pdl(attn_small_fwd_kernel, dim3(unsigned(B * H)), dim3(AS_THREADS), 0, st, p));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

int attn_small_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out, int64_t ldo,
                   const void* dout, int64_t lddo, const float* lse, const float* dsum, void* dq, int64_t lddq, void* dk,
                   int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, const int64_t* klens, int causal,
                   float p_drop, uint64_t seed, uint32_t site, cudaStream_t st) {
  MASR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && lddq % 2 == 0 && lddk % 2 == 0 && lddv % 2 == 0 &&
               ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(dout)) & 15) == 0 &&
               ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 3) == 0,
               "short-query attention backward: operands must be 16 B aligned");
  const uint32_t thr16 = attn_drop_thr16(p_drop);
  AttnSmallParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k); p.v = static_cast<const __nv_bfloat16*>(v);
  p.o = static_cast<const __nv_bfloat16*>(out); p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.lse = const_cast<float*>(lse); p.dsum = dsum;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.kv_rows = Lk; p.klens = klens; p.causal = causal;
  p.scale = 0.125f; p.p_drop = p_drop; p.inv_keep = p_drop > 0.f ? attn_drop_inv_keep(thr16) : 1.f; p.thr16 = thr16;
  p.seed = seed; p.site = site; p.seed_ptr = g_seed_dev_ptr;
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AS_BWD_SMEM))); attr = true; }
  MASR_CHECK_CUDA(launch_pdl(attn_small_bwd_kernel, dim3(unsigned(B * H)), dim3(AS_THREADS), AS_BWD_SMEM, st, p));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

}  // namespace masr

// Longest query sequence served by the warp-MMA kernels (default 64; 0: every problem takes the tcgen05 kernels -- the
// tests use this to pin both families against the same reference)
extern "C" int masr_attn_set_small_lq(int max_lq) {
  masr::g_small_lq = max_lq < 0 ? 0 : (max_lq > masr::AS_QMAX ? masr::AS_QMAX : max_lq);
  return MASR_OK;
}
