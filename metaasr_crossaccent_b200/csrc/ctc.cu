// Kernel 1: log-space CTC alpha-beta forward-backward, one CTA per utterance.
//
// Replaces F.log_softmax + nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True) and its
// backward (src/blstm_trainer.py:22,65-70; ATen native ctc_loss).  ATen runs log_softmax, an alpha
// kernel, a beta kernel and a gradient-collect kernel with alpha/beta tables in HBM; here the
// [T, C] activation slab of an utterance is read once and the gradient written once:
//
//   phase 0  all warps: per-frame log-normaliser (coalesced row reads) and the emission table
//            E[t][j] = logp_t(label_j), E[t][L] = logp_t(blank) gathered into shared memory;
//   phase 1  warps 0-3 run the alpha recursion, warps 4-7 the beta recursion CONCURRENTLY (the two
//            serial chains are independent), one label state per thread (strided when S > 128),
//            neighbours read from the previous table row, one 128-thread named barrier per frame;
//   phase 2  state posteriors gamma_t(s) = exp(alpha+beta-E-ll) summed per class with shared-memory
//            atomics (linear domain: posteriors are in [0,1]);
//   phase 3  all warps: grad[t][c] = (softmax_t(c) - Gamma_t(c)) * scale, one coalesced write per row.
//
// Tables live in shared memory when they fit (2 CTAs/SM at the BASELINE shape T'=128, L=34), else in
// a caller-provided global workspace (L2 resident); the code is identical through generic pointers.
#include "common.cuh"

namespace masr {

constexpr int CTC_THREADS = 256;
constexpr int CTC_HALF = 128;
constexpr float CTC_NEG_INF = -INFINITY;

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == CTC_NEG_INF) return CTC_NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

struct CtcTables {
  float* alpha;   // [T][Sstride]
  float* beta;    // [T][Sstride]
  float* E;       // [T][L1stride]   emissions per label position, blank last
  float* G;       // [T][L1stride]   class posteriors per slot
};

__global__ void __launch_bounds__(CTC_THREADS)
ctc_fwd_bwd_kernel(const float* __restrict__ acts, int T, int B, int C, int is_logprob,
                   const int64_t* __restrict__ targets, const int64_t* __restrict__ tgt_offsets,
                   const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                   int Lmax, int blank, int zero_infinity, float grad_scale,
                   float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ grad,
                   float* __restrict__ ws, int tables_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = CTC_THREADS / 32;
  const int Sstride = 2 * Lmax + 1, L1stride = Lmax + 1;

  // ---- shared-memory carve-up: small arrays first, then (optionally) the tables
  float* logZ = reinterpret_cast<float*>(smem_raw);                 // [T]
  int* tg = reinterpret_cast<int*>(logZ + T);                        // [Lmax]     labels
  int* slot = tg + Lmax;                                             // [Lmax+1]   first position with the same class
  int* cmap = slot + (Lmax + 1);                                     // [C]        class -> slot or -1
  float* red = reinterpret_cast<float*>(cmap + C);                   // [2] scratch (ll)
  float* tbl = red + 4;
  const size_t table_floats = size_t(T) * (2 * Sstride + 2 * L1stride);
  if (!tables_in_smem) tbl = ws + size_t(b) * table_floats;
  CtcTables tb;
  tb.alpha = tbl;
  tb.beta = tb.alpha + size_t(T) * Sstride;
  tb.E = tb.beta + size_t(T) * Sstride;
  tb.G = tb.E + size_t(T) * L1stride;

  const int L = int(tgt_lens[b]);
  const int Tb = min(T, int(in_lens[b]));
  const int S = 2 * L + 1;
  const int64_t toff = tgt_offsets[b];

  for (int j = tid; j < L; j += CTC_THREADS) tg[j] = int(targets[toff + j]);
  for (int c = tid; c < C; c += CTC_THREADS) cmap[c] = -1;
  __syncthreads();
  for (int j = tid; j <= L; j += CTC_THREADS) {
    int sl = j;
    if (j < L) { for (int i = 0; i < j; ++i) if (tg[i] == tg[j]) { sl = i; break; } }
    else       { sl = L; }
    slot[j] = sl;
  }
  __syncthreads();
  // a label equal to `blank` inside the target is legal for ATen; it shares the blank's class row
  for (int j = tid; j <= L; j += CTC_THREADS) {
    if (slot[j] == j) {
      const int cls = j < L ? tg[j] : blank;
      if (j < L && cls == blank) continue;        // merged below
      cmap[cls] = j;
    }
  }
  __syncthreads();
  for (int j = tid; j < L; j += CTC_THREADS) if (tg[j] == blank) slot[j] = L;
  __syncthreads();

  // ---- phase 0: normalisers + emission gather (+ zero G)
  for (int t = warp; t < Tb; t += nwarps) {
    const float* row = acts + (int64_t(t) * B + b) * C;
    float z = 0.f;
    if (!is_logprob) {
      float mx = CTC_NEG_INF;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
      mx = warp_max(mx);
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += __expf(row[c] - mx);
      se = warp_sum(se);
      z = mx + __logf(se);
    }
    if (lane == 0) logZ[t] = z;
    float* Et = tb.E + size_t(t) * L1stride;
    float* Gt = tb.G + size_t(t) * L1stride;
    for (int j = lane; j <= L; j += 32) {
      const int cls = j < L ? tg[j] : blank;
      Et[j] = row[cls] - z;
      Gt[j] = 0.f;
    }
  }
  __syncthreads();

  // ---- phase 1: alpha (threads 0..127) and beta (threads 128..255) concurrently
  if (Tb > 0) {
    if (tid < CTC_HALF) {
      const int g = tid;
      for (int s = g; s < S; s += CTC_HALF) {
        float v = CTC_NEG_INF;
        if (s == 0) v = tb.E[L];                       // blank at t = 0
        else if (s == 1) v = tb.E[0];
        tb.alpha[s] = v;
      }
      bar_sync_named(1, CTC_HALF);
      for (int t = 1; t < Tb; ++t) {
        const float* prev = tb.alpha + size_t(t - 1) * Sstride;
        float* cur = tb.alpha + size_t(t) * Sstride;
        const float* Et = tb.E + size_t(t) * L1stride;
        for (int s = g; s < S; s += CTC_HALF) {
          const int j = s >> 1;
          const bool odd = s & 1;
          const float a = prev[s];
          const float bb = s > 0 ? prev[s - 1] : CTC_NEG_INF;
          const float c = (odd && s > 1 && tg[j] != blank && tg[j] != tg[j - 1]) ? prev[s - 2] : CTC_NEG_INF;
          cur[s] = lse3(a, bb, c) + (odd ? Et[j] : Et[L]);
        }
        bar_sync_named(1, CTC_HALF);
      }
    } else {
      const int g = tid - CTC_HALF;
      {
        float* last = tb.beta + size_t(Tb - 1) * Sstride;
        const float* Et = tb.E + size_t(Tb - 1) * L1stride;
        for (int s = g; s < S; s += CTC_HALF) {
          float v = CTC_NEG_INF;
          if (s == S - 1) v = Et[L];
          else if (s == S - 2) v = Et[L - 1];
          last[s] = v;
        }
      }
      bar_sync_named(2, CTC_HALF);
      for (int t = Tb - 2; t >= 0; --t) {
        const float* nxt = tb.beta + size_t(t + 1) * Sstride;
        float* cur = tb.beta + size_t(t) * Sstride;
        const float* Et = tb.E + size_t(t) * L1stride;
        for (int s = g; s < S; s += CTC_HALF) {
          const int j = s >> 1;
          const bool odd = s & 1;
          const float a = nxt[s];
          const float bb = s + 1 < S ? nxt[s + 1] : CTC_NEG_INF;
          const float c = (odd && s + 2 < S && tg[j] != blank && tg[j] != tg[j + 1]) ? nxt[s + 2] : CTC_NEG_INF;
          cur[s] = lse3(a, bb, c) + (odd ? Et[j] : Et[L]);
        }
        bar_sync_named(2, CTC_HALF);
      }
    }
  }
  __syncthreads();

  // ---- log-likelihood
  if (tid == 0) {
    float ll;
    if (Tb > 0) {
      const float* last = tb.alpha + size_t(Tb - 1) * Sstride;
      ll = lse3(last[S - 1], S > 1 ? last[S - 2] : CTC_NEG_INF, CTC_NEG_INF);
    } else {
      ll = (S == 1) ? 0.f : CTC_NEG_INF;
    }
    red[0] = ll;
  }
  __syncthreads();
  const float ll = red[0];
  const bool feasible = (ll != CTC_NEG_INF);
  if (tid == 0) {
    float nll = -ll;
    if (!feasible && zero_infinity) nll = 0.f;
    nll_out[b] = nll;
    if (loss_out != nullptr) atomicAdd(loss_out, nll / float(max(L, 1)) / float(B));
  }
  if (grad == nullptr) return;
  const float scale = grad_scale / (float(B) * float(max(L, 1)));

  // ---- phase 2: class posteriors
  if (feasible) {
    const int total = Tb * S;
    for (int e = tid; e < total; e += CTC_THREADS) {
      const int t = e / S, s = e - t * S;
      const int j = (s & 1) ? (s >> 1) : L;
      const float ab = tb.alpha[size_t(t) * Sstride + s] + tb.beta[size_t(t) * Sstride + s];
      if (ab == CTC_NEG_INF) continue;
      const float gmm = __expf(ab - tb.E[size_t(t) * L1stride + j] - ll);
      atomicAdd(&tb.G[size_t(t) * L1stride + slot[j]], gmm);
    }
  }
  __syncthreads();

  // ---- phase 3: gradient rows
  for (int t = warp; t < T; t += nwarps) {
    float* grow = grad + (int64_t(t) * B + b) * C;
    if (t >= Tb || (!feasible)) {
      // beyond the input length ATen writes zeros; an infeasible utterance under zero_infinity too
      // (without zero_infinity the loss is inf and the gradient NaN, as in ATen)
      const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
      for (int c = lane; c < C; c += 32) grow[c] = fill;
      continue;
    }
    const float* row = acts + (int64_t(t) * B + b) * C;
    const float z = logZ[t];
    const float* Gt = tb.G + size_t(t) * L1stride;
    for (int c = lane; c < C; c += 32) {
      const float pr = __expf(row[c] - z);
      const int u = cmap[c];
      const float occ = u >= 0 ? Gt[u] : 0.f;
      grow[c] = (pr - occ) * scale;
    }
  }
}

// ====================================================================== v2: warp-shuffle recursions
// Same contract, restructured for throughput (this is the kernel reported as "CTC fwd-bwd GB/s"):
//   * the alpha recursion runs in ONE warp and the beta recursion in ANOTHER, concurrently; each lane owns
//     SPL consecutive label states in registers, the two neighbours across the lane boundary arrive by warp
//     shuffle -> no block barrier and no shared-memory round trip on the 2*T-step dependent chain;
//   * everything is kept in the log2 domain (emissions pre-scaled by log2 e), so the log-sum-exp of the
//     three predecessors is 3 x MUFU.EX2 + 1 x MUFU.LG2;
//   * phase 0 (normaliser + emission gather) and phase 3 (posteriors + gradient row) are frame-parallel over
//     all 8 warps with coalesced row accesses; class posteriors of a frame are combined inside the warp
//     (shuffle sum for the blank, shared-memory atomics only for repeated labels), one write per gradient row.
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lse3_2(float a, float b, float c) {      // log2(2^a + 2^b + 2^c)
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == CTC_NEG_INF) return CTC_NEG_INF;
  return m + lg2f(ex2f(a - m) + ex2f(b - m) + ex2f(c - m));
}
constexpr float CTC_LOG2E = 1.4426950408889634f, CTC_LN2 = 0.6931471805599453f;

template <int SPL>
__global__ void __launch_bounds__(CTC_THREADS)
ctc_fwd_bwd_v2_kernel(const float* __restrict__ acts, int T, int B, int C, int is_logprob,
                      const int64_t* __restrict__ targets, const int64_t* __restrict__ tgt_offsets,
                      const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                      int Lmax, int blank, int zero_infinity, float grad_scale,
                      float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ grad,
                      float* __restrict__ ws, int tables_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CTC_THREADS / 32;
  const int Sstride = 2 * Lmax + 1, L1stride = Lmax + 1;

  float* logZ2 = reinterpret_cast<float*>(smem_raw);                 // [T]
  int* tg = reinterpret_cast<int*>(logZ2 + T);                       // [Lmax]
  int* slot = tg + Lmax;                                             // [Lmax+1]
  int* cmap = slot + (Lmax + 1);                                     // [C]
  float* Gw = reinterpret_cast<float*>(cmap + C);                    // [NW][L1stride] per-warp class posteriors
  float* red = Gw + NW * L1stride;                                   // [4]
  float* tbl = red + 4;
  const size_t table_floats = size_t(T) * (2 * Sstride + L1stride);
  if (!tables_in_smem) tbl = ws + size_t(b) * size_t(T) * (2 * Sstride + 2 * L1stride);
  float* A = tbl;                                 // alpha2 [T][Sstride]
  float* Bt = A + size_t(T) * Sstride;            // beta2  [T][Sstride]
  float* E2 = Bt + size_t(T) * Sstride;           // emissions (log2) [T][L1stride]
  (void)table_floats;

  const int L = int(tgt_lens[b]);
  const int Tb = min(T, int(in_lens[b]));
  const int S = 2 * L + 1;
  const int64_t toff = tgt_offsets[b];

  for (int j = tid; j < L; j += CTC_THREADS) tg[j] = int(targets[toff + j]);
  for (int c = tid; c < C; c += CTC_THREADS) cmap[c] = -1;
  __syncthreads();
  for (int j = tid; j <= L; j += CTC_THREADS) {
    int sl = j;
    if (j < L) {
      if (tg[j] == blank) sl = L;
      else for (int i = 0; i < j; ++i) if (tg[i] == tg[j]) { sl = i; break; }
    } else {
      sl = L;
    }
    slot[j] = sl;
    if (sl == j) cmap[j < L ? tg[j] : blank] = j;
  }
  __syncthreads();

  // ---- phase 0: log2-domain normaliser and emission table, frame-parallel
  for (int t = warp; t < Tb; t += NW) {
    const float* row = acts + (int64_t(t) * B + b) * C;
    float z2 = 0.f;
    if (!is_logprob) {
      float mx = CTC_NEG_INF;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
      mx = warp_max(mx);
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += ex2f((row[c] - mx) * CTC_LOG2E);
      se = warp_sum(se);
      z2 = mx * CTC_LOG2E + lg2f(se);
    }
    if (lane == 0) logZ2[t] = z2;
    float* Et = E2 + size_t(t) * L1stride;
    for (int j = lane; j <= L; j += 32) Et[j] = row[j < L ? tg[j] : blank] * CTC_LOG2E - z2;
  }
  __syncthreads();

  // ---- phase 1: alpha in warp 0, beta in warp 1 (registers + shuffles, no block barrier)
  if (Tb > 0 && warp < 2) {
    const bool fwd = (warp == 0);
    float a[SPL];
    bool skip[SPL];
    int ecol[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const int s = lane * SPL + i;
      const int j = s >> 1;
      const bool odd = s & 1;
      ecol[i] = (odd && j < L) ? j : L;
      if (fwd) skip[i] = odd && s > 1 && s < S && tg[j] != blank && tg[j] != tg[j - 1];
      else     skip[i] = odd && s + 2 < S && tg[j] != blank && tg[j] != tg[j + 1];
    }
    const int t0 = fwd ? 0 : Tb - 1;
    {
      const float* Et = E2 + size_t(t0) * L1stride;
      float* dst = (fwd ? A : Bt) + size_t(t0) * Sstride;
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s = lane * SPL + i;
        float v = CTC_NEG_INF;
        if (fwd) { if (s == 0) v = Et[L]; else if (s == 1 && S > 1) v = Et[0]; }
        else     { if (s == S - 1) v = Et[L]; else if (s == S - 2) v = Et[L - 1]; }
        a[i] = v;
        if (s < S) dst[s] = v;
      }
    }
    float e[SPL];
    if (Tb > 1) {
      const float* En = E2 + size_t(fwd ? 1 : Tb - 2) * L1stride;
#pragma unroll
      for (int i = 0; i < SPL; ++i) e[i] = En[ecol[i]];
    }
    for (int step = 1; step < Tb; ++step) {
      const int t = fwd ? step : Tb - 1 - step;
      float en[SPL];
      if (step + 1 < Tb) {                      // prefetch the next frame's emissions
        const float* En = E2 + size_t(fwd ? t + 1 : t - 1) * L1stride;
#pragma unroll
        for (int i = 0; i < SPL; ++i) en[i] = En[ecol[i]];
      }
      float n1, n2;                             // neighbours across the lane boundary
      if (fwd) {
        if (SPL >= 2) { n1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1); n2 = __shfl_up_sync(0xffffffffu, a[SPL >= 2 ? SPL - 2 : 0], 1); }
        else          { n1 = __shfl_up_sync(0xffffffffu, a[0], 1); n2 = __shfl_up_sync(0xffffffffu, a[0], 2); if (lane < 2) n2 = CTC_NEG_INF; }
        if (lane == 0) { n1 = CTC_NEG_INF; n2 = CTC_NEG_INF; }
      } else {
        if (SPL >= 2) { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[SPL >= 2 ? 1 : 0], 1); }
        else          { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[0], 2); if (lane > 29) n2 = CTC_NEG_INF; }
        if (lane == 31) { n1 = CTC_NEG_INF; n2 = CTC_NEG_INF; }
      }
      float nw[SPL];
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        float x1, x2;
        if (fwd) {
          x1 = (i >= 1) ? a[i >= 1 ? i - 1 : 0] : n1;
          x2 = (i >= 2) ? a[i >= 2 ? i - 2 : 0] : (i == 1 ? n1 : n2);
        } else {
          x1 = (i + 1 < SPL) ? a[i + 1 < SPL ? i + 1 : 0] : n1;
          x2 = (i + 2 < SPL) ? a[i + 2 < SPL ? i + 2 : 0] : (i + 1 < SPL ? n1 : n2);
        }
        if (!skip[i]) x2 = CTC_NEG_INF;
        const int s = lane * SPL + i;
        nw[i] = (s < S) ? lse3_2(a[i], x1, x2) + e[i] : CTC_NEG_INF;
      }
      float* dst = (fwd ? A : Bt) + size_t(t) * Sstride;
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        a[i] = nw[i];
        e[i] = en[i];
        const int s = lane * SPL + i;
        if (s < S) dst[s] = nw[i];
      }
    }
  }
  __syncthreads();
  if (!tables_in_smem) __threadfence_block();

  if (tid == 0) {
    float ll2;
    if (Tb > 0) {
      const float* last = A + size_t(Tb - 1) * Sstride;
      ll2 = lse3_2(last[S - 1], S > 1 ? last[S - 2] : CTC_NEG_INF, CTC_NEG_INF);
    } else {
      ll2 = (S == 1) ? 0.f : CTC_NEG_INF;
    }
    red[0] = ll2;
  }
  __syncthreads();
  const float ll2 = red[0];
  const bool feasible = (ll2 != CTC_NEG_INF);
  if (tid == 0) {
    float nll = -ll2 * CTC_LN2;
    if (!feasible && zero_infinity) nll = 0.f;
    nll_out[b] = nll;
    if (loss_out != nullptr) atomicAdd(loss_out, nll / float(max(L, 1)) / float(B));
  }
  if (grad == nullptr) return;
  const float scale = grad_scale / (float(B) * float(max(L, 1)));

  // ---- phase 3: posteriors + gradient rows, frame-parallel
  float* G = Gw + warp * L1stride;
  for (int t = warp; t < T; t += NW) {
    float* grow = grad + (int64_t(t) * B + b) * C;
    if (t >= Tb || !feasible) {
      const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
      for (int c = lane; c < C; c += 32) grow[c] = fill;
      continue;
    }
    for (int j = lane; j <= L; j += 32) G[j] = 0.f;
    __syncwarp();
    const float* At = A + size_t(t) * Sstride;
    const float* Btt = Bt + size_t(t) * Sstride;
    const float* Et = E2 + size_t(t) * L1stride;
    float blank_sum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const int j = (s & 1) ? (s >> 1) : L;
      const float ab = At[s] + Btt[s];
      const float g = (ab == CTC_NEG_INF) ? 0.f : ex2f(ab - Et[j] - ll2);
      if (s & 1) { if (g != 0.f) atomicAdd(&G[slot[j]], g); }
      else blank_sum += g;
    }
    blank_sum = warp_sum(blank_sum);
    __syncwarp();
    if (lane == 0) G[L] += blank_sum;
    __syncwarp();
    const float* row = acts + (int64_t(t) * B + b) * C;
    const float z2 = logZ2[t];
    for (int c = lane; c < C; c += 32) {
      const float pr = ex2f(row[c] * CTC_LOG2E - z2);
      const int u = cmap[c];
      grow[c] = (pr - (u >= 0 ? G[u] : 0.f)) * scale;
    }
    __syncwarp();
  }
}

static size_t ctc2_small_bytes(int T, int Lmax, int C) {
  return sizeof(float) * size_t(T) + sizeof(int) * (size_t(Lmax) + (Lmax + 1) + C) +
         sizeof(float) * (size_t(CTC_THREADS / 32) * (Lmax + 1) + 4);
}
static size_t ctc2_table_bytes(int T, int Lmax) {
  return sizeof(float) * size_t(T) * (2 * (2 * size_t(Lmax) + 1) + (size_t(Lmax) + 1));
}

template <int SPL>
static int launch_ctc2(const float* acts, int T, int B, int C, int is_logprob, const int64_t* targets,
                       const int64_t* tgt_offsets, const int64_t* in_lens, const int64_t* tgt_lens, int Lmax,
                       int blank, int zero_infinity, float grad_scale, float* nll, float* loss, float* grad,
                       float* ws, bool in_smem, size_t smem, cudaStream_t st) {
  MASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_fwd_bwd_v2_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  ctc_fwd_bwd_v2_kernel<SPL><<<B, CTC_THREADS, smem, st>>>(acts, T, B, C, is_logprob, targets, tgt_offsets, in_lens, tgt_lens,
                                                           Lmax, blank, zero_infinity, grad_scale, nll, loss, grad, ws,
                                                           in_smem ? 1 : 0);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

__global__ void ctc_zero_kernel(float* p) { if (p != nullptr) *p = 0.f; }

static size_t ctc_small_bytes(int T, int Lmax, int C) {
  return sizeof(float) * size_t(T) + sizeof(int) * (size_t(Lmax) + (Lmax + 1) + C) + sizeof(float) * 4;
}
static size_t ctc_table_bytes(int T, int Lmax) {
  return sizeof(float) * size_t(T) * (2 * (2 * size_t(Lmax) + 1) + 2 * (size_t(Lmax) + 1));
}
constexpr size_t CTC_SMEM_LIMIT = 110 * 1024;   // two CTAs per SM

}  // namespace masr

using namespace masr;

extern "C" size_t masr_ctc_workspace_bytes(int T, int B, int C, int max_tgt_len) {
  const size_t small = ctc_small_bytes(T, max_tgt_len, C), tables = ctc_table_bytes(T, max_tgt_len);
  if (small + tables <= CTC_SMEM_LIMIT) return 0;
  return tables * size_t(B);
}

extern "C" int masr_ctc_fwd_bwd(const float* acts, int T, int B, int C, int act_is_logprob,
                                const int64_t* targets, const int64_t* tgt_offsets,
                                const int64_t* in_lens, const int64_t* tgt_lens, int max_tgt_len,
                                int blank, int zero_infinity, float grad_scale,
                                float* nll, float* loss, float* grad,
                                void* workspace, size_t workspace_bytes, void* stream) {
  MASR_REQUIRE(T >= 0 && B >= 0 && C > 0 && max_tgt_len >= 0, "ctc: bad sizes");
  MASR_REQUIRE(blank >= 0 && blank < C, "ctc: blank out of range");
  cudaStream_t st = as_stream(stream);
  if (loss != nullptr) { ctc_zero_kernel<<<1, 1, 0, st>>>(loss); MASR_LAUNCH_CHECK(); }
  if (B == 0) return MASR_OK;
  {
    // v2 (warp-shuffle recursions) whenever the extended label sequence fits 32 lanes x 12 states
    const int S = 2 * max_tgt_len + 1;
    const int spl = (S + 31) / 32;
    const size_t small2 = ctc2_small_bytes(T, max_tgt_len, C), tables2 = ctc2_table_bytes(T, max_tgt_len);
    const bool fits = small2 + tables2 <= CTC_SMEM_LIMIT;
    const bool ws_ok = workspace != nullptr && workspace_bytes >= ctc_table_bytes(T, max_tgt_len) * size_t(B);
    if (spl <= 12 && small2 <= 100 * 1024 && (fits || ws_ok)) {
      const size_t smem2 = fits ? small2 + tables2 : small2;
      float* wsf = static_cast<float*>(workspace);
#define CTC2_CASE(N) return launch_ctc2<N>(acts, T, B, C, act_is_logprob, targets, tgt_offsets, in_lens, tgt_lens, max_tgt_len, \
                                           blank, zero_infinity, grad_scale, nll, loss, grad, wsf, fits, smem2, st)
      if (spl <= 1) CTC2_CASE(1);
      if (spl <= 2) CTC2_CASE(2);
      if (spl <= 3) CTC2_CASE(3);
      if (spl <= 4) CTC2_CASE(4);
      if (spl <= 6) CTC2_CASE(6);
      if (spl <= 8) CTC2_CASE(8);
      CTC2_CASE(12);
#undef CTC2_CASE
    }
  }
  const size_t small = ctc_small_bytes(T, max_tgt_len, C), tables = ctc_table_bytes(T, max_tgt_len);
  const bool in_smem = small + tables <= CTC_SMEM_LIMIT;
  MASR_REQUIRE(small <= 200 * 1024, "ctc: T / C too large for the per-utterance index arrays");
  if (!in_smem) {
    MASR_REQUIRE(workspace != nullptr && workspace_bytes >= tables * size_t(B),
                 "ctc: workspace too small (see masr_ctc_workspace_bytes)");
  }
  const size_t smem = in_smem ? small + tables : small;
  MASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  ctc_fwd_bwd_kernel<<<B, CTC_THREADS, smem, st>>>(acts, T, B, C, act_is_logprob, targets, tgt_offsets, in_lens,
                                                   tgt_lens, max_tgt_len, blank, zero_infinity, grad_scale,
                                                   nll, loss, grad, static_cast<float*>(workspace), in_smem ? 1 : 0);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}
