// Multi-head attention forward / backward on tcgen05 tensor cores (head dim 64, bf16, fp32 soft-max).
//
// Same contract as attention.cu (masks from lengths, dropout replayed from (seed, site, index)); this is
// the fast path used for the hkust network (8 heads x 64).  One CTA = 128 query rows (fwd) or 128 key rows
// (bwd) of one (batch, head); 320 threads:
//   warp 0   TMA producer (Q / K / V / dO tiles as [128 rows x 64] boxes, 128B swizzle)
//   warp 1   MMA issuer, TMEM owner
//   warps 2-9 soft-max warps: TWO threads per row (TMEM lane r = 32 * (warp % 4) + lane; warps 2-5 own key
//            columns 0-63 of the tile, warps 6-9 columns 64-127): tcgen05.ld of S / dP, exp2, masks, dropout,
//            P / dS written back to shared memory in the UMMA K-major swizzled layout for the second GEMMs.
//            The soft-max is the critical path of these small problems (the GEMMs take < 1 us): one warp per
//            scheduler with 128 keys per thread was latency-bound at ~19 us per tile; two warps per scheduler,
//            a single MUFU.EX2 per element, tile-uniform masking and a dropout hash shared by two keys cut it
//            to ~2 us.  Row maxima / sums are exchanged between the two threads of a row through shared memory.
// Forward, per 128-key tile:  S = Q K^T -> TMEM;  P = softmax-tile -> smem;  O_tile = P V -> TMEM;
//            O (registers) = O * corr + O_tile  (online soft-max over key tiles).
// Backward, per 128-query tile (keys fixed):  S = Q K^T, dP = dO V^T -> TMEM;  P, dS -> smem;
//            dQ = dS K (fresh), dK += dS^T Q, dV += P^T dO (accumulated in TMEM over query tiles).
// The transposed operands (K as [key, d] for dS K; dS^T, P^T, Q, dO reduced over queries) are the SAME
// shared-memory tiles read through MN-major descriptors -- nothing is transposed in memory.
#include "common.cuh"
#include "umma.cuh"

namespace masr {

constexpr int AU_THREADS = 320;
constexpr int AU_SM_THREADS = 256;     // soft-max threads (warps 2..9)
constexpr int AU_TILE = 128;
constexpr uint32_t AU_T64 = 128 * 128;          // bytes of a [128 rows x 64 bf16] tile
constexpr uint32_t AU_T128 = 2 * AU_T64;        // [128 rows x 128 bf16] = two 64-wide k-blocks

// byte offset of the 16-byte chunk holding columns [8*c8, 8*c8+8) of row r in a [128 x 128] bf16 K-major SW128 tile
__device__ __forceinline__ uint32_t sw_chunk_off(int r, int c8) {
  return uint32_t(c8 >> 3) * AU_T64 + uint32_t(r) * 128 + (uint32_t((c8 & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float fast_exp2(float x) {      // single MUFU.EX2 (flush-to-zero; 2^-inf = 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct AttnFwdParams {
  int B, H, Lq, Lk;
  const int64_t* klens;
  int causal;
  float scale, p_drop, inv_keep;
  uint32_t thr16;                // dropout threshold on 16-bit uniforms (common.cuh attn_drop_*)
  uint64_t seed; uint32_t site;
  __nv_bfloat16* out; int64_t ldo;
  float* lse;
  const uint64_t* seed_ptr;      // device-resident per-step seed offset (CUDA-graph replay), may be NULL
  int kv_stages;                 // K/V ring depth: 1 (Lk <= 128) or 2
  int kv_rows;                   // rows per utterance in K / V: Lk, or the capacity of a decode cache
};

__global__ void __launch_bounds__(AU_THREADS, 2)
attn_fwd_umma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, AttnFwdParams p) {
  using namespace umma;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                         // 16 KB
  unsigned char* sKV = sQ + AU_T64;                 // kv_stages x (K 16 KB + V 16 KB)
  const int kvst = p.kv_stages;                     // 1 when all keys fit one tile: 83 KB -> two CTAs per SM
  unsigned char* sP = sKV + kvst * 2 * AU_T64;      // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AU_T128);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                     // [2]
  uint64_t* kv_empty = bars + 3;                    // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xch = reinterpret_cast<float*>(bars + 10);        // [2][128] row maxima, then [2][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, AU_SM_THREADS); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();           // global memory (lengths, seed offset, tiles) is only touched below
  const uint64_t seed_eff = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.x * AU_TILE;
  int kmax = p.Lk;
  if (p.klens != nullptr) kmax = min(kmax, int(p.klens[b]));
  if (p.causal) kmax = min(kmax, q0 + AU_TILE);
  const int ntiles = (kmax + AU_TILE - 1) / AU_TILE;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  // producer / issuer warps run converged; one elected lane issues (umma::elect_one_sync)
  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, AU_T64);
      tma_load_2d(sQ, &map_q, q_full, h * 64, b * p.Lq + q0);
    }
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      const int s = kvst == 2 ? (t & 1) : 0;
      mbar_wait(&kv_empty[s], ((kvst == 2 ? (t >> 1) : t) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[s], 2 * AU_T64);
        tma_load_2d(sKV + s * 2 * AU_T64, &map_k, &kv_full[s], h * 64, b * p.kv_rows + t * AU_TILE);
        tma_load_2d(sKV + s * 2 * AU_T64 + AU_T64, &map_v, &kv_full[s], h * 64, b * p.kv_rows + t * AU_TILE);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);       // S = Q K^T
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);        // O = P V   (V read MN-major)
    mbar_wait(q_full, 0);
    for (int t = 0; t < ntiles; ++t) {
      const int s = kvst == 2 ? (t & 1) : 0;
      mbar_wait(&kv_full[s], (kvst == 2 ? (t >> 1) : t) & 1);
      tc_fence_after();
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sKV + s * 2 * AU_T64), av = ak + AU_T64, ap = smem_u32(sP);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tmem_S, desc_kmajor_sw128(aq + k * 32), desc_kmajor_sw128(ak + k * 32), idesc_s, k > 0 ? 1u : 0u);
        mma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, t & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)      // 128 keys = 8 x K16: P k-block k/4, 32 B steps; V +16 key rows = 2048 B
          mma_f16_ss(tmem_O, desc_kmajor_sw128(ap + (k >> 2) * AU_T64 + (k & 3) * 32),
                     desc_mnmajor_sw128(av + k * 2048, AU_T64), idesc_o, k > 0 ? 1u : 0u);
        mma_commit(o_full);
        mma_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
  } else {
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;                   // which 64 key columns of the tile this thread owns
    const int r = qd * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    // soft-max statistics are kept in the log2 domain: exp(x) = exp2(x * log2 e) is a single MUFU.EX2
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool drop = p.p_drop > 0.f;
    const uint32_t rowkey = attn_drop_rowkey(seed_eff, p.site, uint32_t(bh) * uint32_t(p.Lq) + uint32_t(qi));
    float m_run = -INFINITY, l_run = 0.f;
    float o[32];                                         // O columns [32*half, 32*half+32) of this row
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int c0 = t * AU_TILE + half * 64;            // first key of this thread's 64 columns
      // keys [c0, c0 + lim) are visible to this row: key padding / causal mask as ONE bound
      const int lim = max(0, min(64, min(kmax, p.causal ? qi + 1 : kmax) - c0));
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      float sv[64];
      tmem_ld_32x32(tmem_S + lane_addr + half * 64, sv);
      tmem_ld_32x32(tmem_S + lane_addr + half * 64 + 32, sv + 32);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (lim == 64) {
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, sv[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, j < lim ? sv[j] : -INFINITY);
      }
      mx *= sl2;                                         // sl2 > 0: scaling commutes with the maximum
      xch[half * 128 + r] = mx;
      softmax_bar();
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = fast_exp2(m_run - m_use);       // 2^(-inf) = 0 on the first tile
      float lsum = 0.f;
      // probabilities -> shared memory (bf16, K-major swizzled), row sum; 8 keys = one 16-byte chunk
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float pr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = g * 8 + e;
          const float ex = fast_exp2(fmaf(sv[j], sl2, -m_use));
          pr[e] = (lim == 64 || j < lim) ? ex : 0.f;
          lsum += pr[e];
        }
        if (drop) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const uint32_t bits = attn_drop_pair(rowkey, uint32_t(c0 + g * 8 + e) >> 1);
            pr[e] = ((bits & 0xffffu) >= p.thr16) ? pr[e] * p.inv_keep : 0.f;
            pr[e + 1] = ((bits >> 16) >= p.thr16) ? pr[e + 1] * p.inv_keep : 0.f;
          }
        }
        uint4 pk;
        pk.x = pack_bf16x2(pr[0], pr[1]); pk.y = pack_bf16x2(pr[2], pr[3]);
        pk.z = pack_bf16x2(pr[4], pr[5]); pk.w = pack_bf16x2(pr[6], pr[7]);
        *reinterpret_cast<uint4*>(sP + sw_chunk_off(r, half * 8 + g)) = pk;
      }
      l_run = l_run * corr + lsum;                       // partial over this thread's columns (same m in both halves)
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_full);
      // O = O * corr + P V
      mbar_wait(o_full, t & 1);
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(tmem_O + lane_addr + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = fmaf(o[j], corr, v[j]);
      tc_fence_before();
    }
    xch[256 + half * 128 + r] = l_run;
    softmax_bar();
    l_run += xch[256 + (half ^ 1) * 128 + r];
    if (qi < p.Lq) {
      const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = p.out + (int64_t(b) * p.Lq + qi) * p.ldo + h * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float t8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t8[e] = o[j + e] * inv_l;
        store8<__nv_bfloat16>(orow + j, t8);
      }
      if (half == 0)
        p.lse[int64_t(bh) * p.Lq + qi] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// D[b,h,q] = dO_q . O_q  (one warp per (b, q, h) row of 64)
__global__ void attn_dsum_kernel(const __nv_bfloat16* __restrict__ out, int64_t ldo, const __nv_bfloat16* __restrict__ dout,
                                 int64_t lddo, float* __restrict__ dsum, int B, int H, int Lq) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t total = int64_t(B) * H * Lq;
  for (int64_t idx = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; idx < total;
       idx += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int qi = int(idx % Lq);
    const int h = int((idx / Lq) % H);
    const int64_t b = idx / (int64_t(Lq) * H);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(out + (b * Lq + qi) * ldo + h * 64 + lane * 2);
    const __nv_bfloat162 g = *reinterpret_cast<const __nv_bfloat162*>(dout + (b * Lq + qi) * lddo + h * 64 + lane * 2);
    const float2 fa = __bfloat1622float2(a), fg = __bfloat1622float2(g);
    const float s = warp_sum(fa.x * fg.x + fa.y * fg.y);
    if (lane == 0) dsum[(b * H + h) * Lq + qi] = s;
  }
}

struct AttnBwdParams {
  int B, H, Lq, Lk;
  const int64_t* klens;
  int causal;
  float scale, p_drop, inv_keep;
  uint32_t thr16;
  uint64_t seed; uint32_t site;
  const float* lse; const float* dsum;
  __nv_bfloat16* dq; int64_t lddq;
  __nv_bfloat16* dk; int64_t lddk;
  __nv_bfloat16* dv; int64_t lddv;
  const uint64_t* seed_ptr;
  float* dq_ws;                  // [B*Lq, H*64] fp32, zeroed: dQ partials of the key tiles when Lk > 128 (else NULL)
};

// fp32 dQ partial sums -> bf16 dQ (memories longer than one key tile)
__global__ void attn_dq_cast_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ dq, int64_t lddq, int64_t rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n8 = rows * (cols / 8);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / (cols / 8);
    const int c = int(i - r * (cols / 8)) * 8;
    const float4 a = *reinterpret_cast<const float4*>(ws + r * cols + c), b = *reinterpret_cast<const float4*>(ws + r * cols + c + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    store8<__nv_bfloat16>(dq + r * lddq + c, v);
  }
}

// One CTA = one 128-key tile of one (batch, head), looping over the query tiles: dK / dV accumulate in TMEM.  dQ of a
// query tile is complete when the memory fits ONE key tile (gridDim.x == 1: written directly as bf16); for longer
// memories (utterances up to max_ilen 1500 -> 375 keys) every key tile adds its partial to the fp32 workspace with
// vector reductions and attn_dq_cast_kernel rounds the sum once.
__global__ void __launch_bounds__(AU_THREADS, 1)
attn_bwd_umma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do, AttnBwdParams p) {
  using namespace umma;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sK = smem;                         // 16 KB
  unsigned char* sV = sK + AU_T64;                  // 16 KB
  unsigned char* sQdO = sV + AU_T64;                // 2 stages x (Q 16 KB + dO 16 KB)
  unsigned char* sP = sQdO + 4 * AU_T64;            // 32 KB  (dropped probabilities, bf16)
  unsigned char* sdS = sP + AU_T128;                // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + AU_T128);
  uint64_t* kv_full = bars;
  uint64_t* qd_full = bars + 1;                     // [2]
  uint64_t* qd_empty = bars + 3;                    // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dkv_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v); prefetch_tmap(&map_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 1); }
    mbar_init(sdp_full, 1); mbar_init(pds_full, AU_SM_THREADS); mbar_init(dq_full, 1); mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint64_t seed_eff = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
  const int k0 = blockIdx.x * AU_TILE;
  int klen = p.Lk;
  if (p.klens != nullptr) klen = min(klen, int(p.klens[b]));
  const int nq_tiles = (p.Lq + AU_TILE - 1) / AU_TILE;
  const int qt_begin = p.causal ? (k0 / AU_TILE) : 0;
  const bool active = (k0 < klen) && (qt_begin < nq_tiles);
  const int ntiles = active ? (nq_tiles - qt_begin) : 0;
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 128, tm_dQ = tmem_base + 256, tm_dK = tmem_base + 320, tm_dV = tmem_base + 384;

  // producer / issuer warps run converged (`active` is CTA-uniform); one elected lane issues
  if (warp == 0) {
    if (active) {
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(kv_full, 2 * AU_T64);
        tma_load_2d(sK, &map_k, kv_full, h * 64, b * p.Lk + k0);
        tma_load_2d(sV, &map_v, kv_full, h * 64, b * p.Lk + k0);
      }
      __syncwarp();
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        const int q0 = (qt_begin + t) * AU_TILE;
        mbar_wait(&qd_empty[s], ((t >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&qd_full[s], 2 * AU_T64);
          tma_load_2d(sQdO + s * 2 * AU_T64, &map_q, &qd_full[s], h * 64, b * p.Lq + q0);
          tma_load_2d(sQdO + s * 2 * AU_T64 + AU_T64, &map_do, &qd_full[s], h * 64, b * p.Lq + q0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (active) {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);     // S = Q K^T ; dP = dO V^T
      constexpr uint32_t id_dq = make_idesc_bf16(128, 64, 0, 1);     // dQ = dS K        (K read MN-major)
      constexpr uint32_t id_dkv = make_idesc_bf16(128, 64, 1, 1);    // dK = dS^T Q ; dV = P^T dO
      mbar_wait(kv_full, 0);
      const uint32_t ak = smem_u32(sK), av = smem_u32(sV), ap = smem_u32(sP), ads = smem_u32(sdS);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        mbar_wait(&qd_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t aq = smem_u32(sQdO + s * 2 * AU_T64), ado = aq + AU_T64;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(tm_S, desc_kmajor_sw128(aq + k * 32), desc_kmajor_sw128(ak + k * 32), id_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(tm_dP, desc_kmajor_sw128(ado + k * 32), desc_kmajor_sw128(av + k * 32), id_s, k > 0 ? 1u : 0u);
          mma_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(pds_full, t & 1);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // dQ[q, d] = sum_key dS[q, key] K[key, d]
            mma_f16_ss(tm_dQ, desc_kmajor_sw128(ads + (k >> 2) * AU_T64 + (k & 3) * 32),
                       desc_mnmajor_sw128(ak + k * 2048, AU_T64), id_dq, k > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // dK[key, d] += sum_q dS[q, key] Q[q, d];  dV[key, d] += sum_q P[q, key] dO[q, d]   (16 query rows / step)
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            mma_f16_ss(tm_dK, desc_mnmajor_sw128(ads + k * 2048, AU_T64), desc_mnmajor_sw128(aq + k * 2048, AU_T64), id_dkv, acc);
            mma_f16_ss(tm_dV, desc_mnmajor_sw128(ap + k * 2048, AU_T64), desc_mnmajor_sw128(ado + k * 2048, AU_T64), id_dkv, acc);
          }
          mma_commit(dq_full);
          mma_commit(&qd_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) mma_commit(dkv_full);
      __syncwarp();
    }
  } else {
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;                   // key columns [64*half, 64*half+64) of S / dP; 32 of the 64 output columns
    const int r = qd * 32 + lane;
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool drop = p.p_drop > 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int q0 = (qt_begin + t) * AU_TILE;
      const int qi = q0 + r;
      const bool qok = qi < p.Lq;
      const float lse_r = qok ? p.lse[int64_t(bh) * p.Lq + qi] * 1.4426950408889634f : 0.f;    // log2 domain
      const float d_r = qok ? p.dsum[int64_t(bh) * p.Lq + qi] : 0.f;
      const uint32_t rowkey = attn_drop_rowkey(seed_eff, p.site, uint32_t(bh) * uint32_t(p.Lq) + uint32_t(qi));
      mbar_wait(sdp_full, t & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;                     // 32-column chunk of the key tile
        const int c0 = k0 + c * 32;
        const int lim = qok ? max(0, min(32, min(klen, p.causal ? qi + 1 : klen) - c0)) : 0;
        float sv[32], dp[32];
        tmem_ld_32x32(tm_S + lane_addr + c * 32, sv);
        tmem_ld_32x32(tm_dP + lane_addr + c * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float dm0 = 1.f, dm1 = 1.f;
          if (drop) {
            const uint32_t bits = attn_drop_pair(rowkey, uint32_t(c0 + j) >> 1);
            dm0 = ((bits & 0xffffu) >= p.thr16) ? p.inv_keep : 0.f;
            dm1 = ((bits >> 16) >= p.thr16) ? p.inv_keep : 0.f;
          }
          const float e0 = fast_exp2(fmaf(sv[j], sl2, -lse_r)), e1 = fast_exp2(fmaf(sv[j + 1], sl2, -lse_r));
          const float p0 = (j < lim) ? e0 : 0.f, p1 = (j + 1 < lim) ? e1 : 0.f;
          const float g0 = (j < lim) ? dp[j] : 0.f, g1 = (j + 1 < lim) ? dp[j + 1] : 0.f;   // dP may hold garbage there
          dp[j] = p0 * (g0 * dm0 - d_r) * p.scale;
          dp[j + 1] = p1 * (g1 * dm1 - d_r) * p.scale;
          sv[j] = p0 * dm0;
          sv[j + 1] = p1 * dm1;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk, dk4;
          pk.x = pack_bf16x2(sv[g * 8 + 0], sv[g * 8 + 1]); pk.y = pack_bf16x2(sv[g * 8 + 2], sv[g * 8 + 3]);
          pk.z = pack_bf16x2(sv[g * 8 + 4], sv[g * 8 + 5]); pk.w = pack_bf16x2(sv[g * 8 + 6], sv[g * 8 + 7]);
          dk4.x = pack_bf16x2(dp[g * 8 + 0], dp[g * 8 + 1]); dk4.y = pack_bf16x2(dp[g * 8 + 2], dp[g * 8 + 3]);
          dk4.z = pack_bf16x2(dp[g * 8 + 4], dp[g * 8 + 5]); dk4.w = pack_bf16x2(dp[g * 8 + 6], dp[g * 8 + 7]);
          const uint32_t off = sw_chunk_off(r, c * 4 + g);
          *reinterpret_cast<uint4*>(sP + off) = pk;
          *reinterpret_cast<uint4*>(sdS + off) = dk4;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(pds_full);
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      {
        float v[32];
        tmem_ld_32x32(tm_dQ + lane_addr + half * 32, v);
        tmem_ld_wait();
        if (qok) {
          if (gridDim.x == 1) {
            __nv_bfloat16* drow = p.dq + (int64_t(b) * p.Lq + qi) * p.lddq + h * 64 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) store8<__nv_bfloat16>(drow + j, v + j);
          } else {
            float* wrow = p.dq_ws + (int64_t(b) * p.Lq + qi) * (int64_t(p.H) * 64) + h * 64 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                           "f"(v[j + 3]) : "memory");
          }
        }
      }
      tc_fence_before();
    }
    // dK / dV of this key tile (row r = key k0 + r), output columns [32*half, 32*half+32)
    const int kj = k0 + r;
    float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    {
      float vk[32], vv[32];
      if (active) {
        tmem_ld_32x32(tm_dK + lane_addr + half * 32, vk);
        tmem_ld_32x32(tm_dV + lane_addr + half * 32, vv);
        tmem_ld_wait();
      }
      if (kj < p.Lk) {
        __nv_bfloat16* dkr = p.dk + (int64_t(b) * p.Lk + kj) * p.lddk + h * 64 + half * 32;
        __nv_bfloat16* dvr = p.dv + (int64_t(b) * p.Lk + kj) * p.lddv + h * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          store8<__nv_bfloat16>(dkr + j, active ? vk + j : zero8);
          store8<__nv_bfloat16>(dvr + j, active ? vv + j : zero8);
        }
      }
    }
    // With a single key tile (gridDim.x == 1) every query tile is visited above, so dQ is fully written --
    // except when the whole key tile is padding (klen == 0): then dQ is zero.
    if (!active && gridDim.x == 1) {
      for (int qi = r; qi < p.Lq; qi += AU_TILE) {
        __nv_bfloat16* drow = p.dq + (int64_t(b) * p.Lq + qi) * p.lddq + h * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) store8<__nv_bfloat16>(drow + j, zero8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int rows_map(CUtensorMap* out, const void* base, int64_t ld, int64_t nrows, int ncols) {
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, "umma attention: 16 B aligned rows required");
  uint64_t dims[2] = {uint64_t(ncols), uint64_t(nrows)};
  uint64_t strides[1] = {uint64_t(ld) * 2};
  uint32_t box[2] = {64, 128};
  return make_tmap_bf16(out, base, 2, dims, strides, box, true);
}

// attn_small.cu: warp-MMA kernels for short query sequences (the decoder's 33-row problems)
bool attn_small_applicable(int Lq);
int attn_small_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                   float* lse, int B, int H, int Lq, int Lk, int kv_rows, const int64_t* klens, int causal, float p_drop,
                   uint64_t seed, uint32_t site, cudaStream_t st);
int attn_small_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out, int64_t ldo,
                   const void* dout, int64_t lddo, const float* lse, const float* dsum, void* dq, int64_t lddq, void* dk,
                   int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, const int64_t* klens, int causal,
                   float p_drop, uint64_t seed, uint32_t site, cudaStream_t st);

constexpr size_t AU_FWD_SMEM = AU_T64 + 4 * AU_T64 + AU_T128 + 256 + 2048 + 1024;
constexpr size_t AU_BWD_SMEM = 2 * AU_T64 + 4 * AU_T64 + 2 * AU_T128 + 256 + 1024;

}  // namespace masr

using namespace masr;

extern "C" int masr_umma_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  void* out, int64_t ldo, float* lse, int B, int H, int Lq, int Lk,
                                  const int64_t* klens, int causal, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  return masr_umma_attn_fwd_cached(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, H, Lq, Lk, Lk, klens, causal, p_drop, seed, site, stream);
}

extern "C" int masr_umma_attn_fwd_cached(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                         void* out, int64_t ldo, float* lse, int B, int H, int Lq, int Lk, int kv_rows,
                                         const int64_t* klens, int causal, float p_drop, uint64_t seed, uint32_t site,
                                         void* stream) {
  if (B == 0 || H == 0 || Lq == 0) return MASR_OK;
  MASR_REQUIRE(kv_rows >= Lk, "umma attention: kv_rows (rows per utterance of K / V) must be >= Lk");
  if (attn_small_applicable(Lq))
    return attn_small_fwd(q, ldq, k, ldk, v, ldv, out, ldo, lse, B, H, Lq, Lk, kv_rows, klens, causal, p_drop, seed, site,
                          as_stream(stream));
  MASR_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "umma attention: out must be 16 B aligned");
  CUtensorMap mq, mk, mv;
  int rc = rows_map(&mq, q, ldq, int64_t(B) * Lq, H * 64); if (rc) return rc;
  rc = rows_map(&mk, k, ldk, int64_t(B) * kv_rows, H * 64); if (rc) return rc;
  rc = rows_map(&mv, v, ldv, int64_t(B) * kv_rows, H * 64); if (rc) return rc;
  const uint32_t thr16 = attn_drop_thr16(p_drop);
  AttnFwdParams p{B, H, Lq, Lk, klens, causal, 0.125f, p_drop, p_drop > 0.f ? attn_drop_inv_keep(thr16) : 1.f, thr16, seed, site,
                  static_cast<__nv_bfloat16*>(out), ldo, lse, g_seed_dev_ptr, Lk <= AU_TILE ? 1 : 2, kv_rows};
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AU_FWD_SMEM))); attr = true; }
  dim3 grid(unsigned(ceil_div64(Lq, AU_TILE)), unsigned(B * H));
  const size_t smem = AU_FWD_SMEM - (p.kv_stages == 1 ? 2 * AU_T64 : 0);
  MASR_CHECK_CUDA(launch_pdl(attn_fwd_umma_kernel, grid, dim3(AU_THREADS), smem, as_stream(stream), mq, mk, mv, p));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_umma_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                  float* dsum_ws, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                  int B, int H, int Lq, int Lk, const int64_t* klens, int causal,
                                  float p_drop, uint64_t seed, uint32_t site, int dsum_ready, float* dq_ws, void* stream) {
  if (B == 0 || H == 0) return MASR_OK;
  if (attn_small_applicable(Lq))
    return attn_small_bwd(q, ldq, k, ldk, v, ldv, out, ldo, dout, lddo, lse, dsum_ready ? dsum_ws : nullptr, dq, lddq, dk, lddk,
                          dv, lddv, B, H, Lq, Lk, klens, causal, p_drop, seed, site, as_stream(stream));
  const int nkt = int(ceil_div64(std::max(Lk, 1), AU_TILE));
  MASR_REQUIRE(nkt == 1 || dq_ws != nullptr, "umma attention backward: Lk > 128 needs the [B*Lq, H*64] fp32 dq workspace");
  MASR_REQUIRE(dsum_ws != nullptr, "attention backward needs a [B*H*Lq] fp32 workspace");
  MASR_REQUIRE(lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && ldo % 2 == 0 && lddo % 2 == 0 &&
               ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0,
               "umma attention backward: gradients must be 16 B aligned");
  cudaStream_t st = as_stream(stream);
  CUtensorMap mq, mk, mv, mdo;
  int rc = rows_map(&mq, q, ldq, int64_t(B) * Lq, H * 64); if (rc) return rc;
  rc = rows_map(&mk, k, ldk, int64_t(B) * Lk, H * 64); if (rc) return rc;
  rc = rows_map(&mv, v, ldv, int64_t(B) * Lk, H * 64); if (rc) return rc;
  rc = rows_map(&mdo, dout, lddo, int64_t(B) * Lq, H * 64); if (rc) return rc;
  if (Lq > 0 && !dsum_ready) {
    const int64_t rows = int64_t(B) * H * Lq;
    const int blocks = int(std::min<int64_t>(ceil_div64(rows, 8), int64_t(sm_count()) * 8));
    MASR_CHECK_CUDA(launch_pdl(attn_dsum_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(out), ldo,
                               static_cast<const __nv_bfloat16*>(dout), lddo, dsum_ws, B, H, Lq));
  }
  const uint32_t thr16 = attn_drop_thr16(p_drop);
  AttnBwdParams p{B, H, Lq, Lk, klens, causal, 0.125f, p_drop, p_drop > 0.f ? attn_drop_inv_keep(thr16) : 1.f, thr16, seed, site,
                  lse, dsum_ws, static_cast<__nv_bfloat16*>(dq), lddq, static_cast<__nv_bfloat16*>(dk), lddk,
                  static_cast<__nv_bfloat16*>(dv), lddv, g_seed_dev_ptr, nkt > 1 ? dq_ws : nullptr};
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(AU_BWD_SMEM))); attr = true; }
  if (nkt > 1 && Lq > 0) MASR_CHECK_CUDA(cudaMemsetAsync(dq_ws, 0, sizeof(float) * size_t(B) * Lq * H * 64, st));
  dim3 grid(unsigned(nkt), unsigned(B * H));
  MASR_CHECK_CUDA(launch_pdl(attn_bwd_umma_kernel, grid, dim3(AU_THREADS), AU_BWD_SMEM, st, mq, mk, mv, mdo, p));
  MASR_LAUNCH_CHECK();
  if (nkt > 1 && Lq > 0) {
    const int64_t rows = int64_t(B) * Lq;
    const int blocks = int(std::min<int64_t>(ceil_div64(rows * H * 8, 256), int64_t(sm_count()) * 8));
    MASR_CHECK_CUDA(launch_pdl(attn_dq_cast_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(dq_ws),
                               static_cast<__nv_bfloat16*>(dq), lddq, rows, H * 64));
    MASR_LAUNCH_CHECK();
  }
  return MASR_OK;
}
