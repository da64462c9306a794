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
ctc_fwd_bwd_kernel(const float* __restrict__ acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob,
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
    const float* row = acts + (int64_t(t) * st_t + b * st_b);
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
          const float c = (odd && s > 1 && tg[j] != tg[j - 1]) ? prev[s - 2] : CTC_NEG_INF;
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
          const float c = (odd && s + 2 < S && tg[j] != tg[j + 1]) ? nxt[s + 2] : CTC_NEG_INF;
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
    float* grow = grad + (int64_t(t) * st_t + b * st_b);
    if (t >= Tb || (!feasible)) {
      // beyond the input length ATen writes zeros; an infeasible utterance under zero_infinity too
      // (without zero_infinity the loss is inf and the gradient NaN, as in ATen)
      const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
      for (int c = lane; c < C; c += 32) grow[c] = fill;
      continue;
    }
    const float* row = acts + (int64_t(t) * st_t + b * st_b);
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
constexpr int CTC_F = 4;        // frames per warp iteration in the streaming phases
constexpr int CTC_CPL = 12;     // classes per lane held in registers (fast path: C <= 384)


// One warp runs one recursion (FWD: alpha, t = 0..Tb-1; !FWD: beta, t = Tb-1..0).  Lane l owns the SPL
// consecutive states l*SPL .. l*SPL+SPL-1 in registers.  "Impossible" is a large negative finite number
// instead of -inf so the log-sum-exp needs no special cases (2^(x - m) underflows to exactly 0).
constexpr float CTC_NEG = -1.0e30f;
template <int SPL, bool FWD>
__device__ __forceinline__ void ctc_chain(float* __restrict__ tab, const float* __restrict__ E2, const int* __restrict__ tg,
                                          int lane, int L, int S, int Tb, int Sstride, int L1stride, int blank) {
  float a[SPL], e[SPL];
  bool skip[SPL], valid[SPL];
  int ecol[SPL];
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    const int j = s >> 1;
    const bool odd = s & 1;
    valid[i] = s < S;
    ecol[i] = (odd && j < L) ? j : L;
    if (FWD) skip[i] = odd && s > 1 && s < S && tg[j] != tg[j - 1];
    else     skip[i] = odd && s + 2 < S && tg[j] != tg[j + 1];
  }
  const int tstep = FWD ? 1 : -1;
  int t = FWD ? 0 : Tb - 1;
  const float* Et = E2 + size_t(t) * L1stride;
  float* dst = tab + size_t(t) * Sstride + lane * SPL;
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    float v = CTC_NEG;
    if (FWD) { if (s == 0) v = Et[L]; else if (s == 1 && S > 1) v = Et[0]; }
    else     { if (s == S - 1) v = Et[L]; else if (s == S - 2) v = Et[L - 1]; }
    a[i] = v;
    if (valid[i]) dst[i] = v;
  }
  Et += tstep * L1stride;
  if (Tb > 1) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) e[i] = Et[ecol[i]];
  }
  for (int step = 1; step < Tb; ++step) {
    dst += tstep * Sstride;
    Et += tstep * L1stride;
    float en[SPL];
    const bool more = step + 1 < Tb;
#pragma unroll
    for (int i = 0; i < SPL; ++i) en[i] = more ? Et[ecol[i]] : 0.f;      // prefetch the next frame's emissions
    float n1, n2;                                                          // neighbours across the lane boundary
    if (FWD) {
      if (SPL >= 2) { n1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1); n2 = __shfl_up_sync(0xffffffffu, a[SPL >= 2 ? SPL - 2 : 0], 1); }
      else          { n1 = __shfl_up_sync(0xffffffffu, a[0], 1); n2 = __shfl_up_sync(0xffffffffu, a[0], 2); if (lane < 2) n2 = CTC_NEG; }
      if (lane == 0) { n1 = CTC_NEG; n2 = CTC_NEG; }
    } else {
      if (SPL >= 2) { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[SPL >= 2 ? 1 : 0], 1); }
      else          { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[0], 2); if (lane > 29) n2 = CTC_NEG; }
      if (lane == 31) { n1 = CTC_NEG; n2 = CTC_NEG; }
    }
    float nw[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      float x1, x2;
      if (FWD) {
        x1 = (i >= 1) ? a[i >= 1 ? i - 1 : 0] : n1;
        x2 = (i >= 2) ? a[i >= 2 ? i - 2 : 0] : (i == 1 ? n1 : n2);
      } else {
        x1 = (i + 1 < SPL) ? a[i + 1 < SPL ? i + 1 : 0] : n1;
        x2 = (i + 2 < SPL) ? a[i + 2 < SPL ? i + 2 : 0] : (i + 1 < SPL ? n1 : n2);
      }
      x2 = skip[i] ? x2 : CTC_NEG;
      const float m = fmaxf(a[i], fmaxf(x1, x2));
      const float v = m + lg2f(ex2f(a[i] - m) + ex2f(x1 - m) + ex2f(x2 - m)) + e[i];
      nw[i] = valid[i] ? fmaxf(v, CTC_NEG) : CTC_NEG;
    }
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      a[i] = nw[i];
      e[i] = en[i];
      if (valid[i]) dst[i] = nw[i];
    }
  }
}

template <int SPL>
__global__ void __launch_bounds__(CTC_THREADS)
ctc_fwd_bwd_v2_kernel(const float* __restrict__ acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob,
                      const int64_t* __restrict__ targets, const int64_t* __restrict__ tgt_offsets,
                      const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                      int Lmax, int blank, int zero_infinity, float grad_scale,
                      float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ grad,
                      float* __restrict__ ws, int tables_in_smem, long long* __restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CTC_THREADS / 32;
  // optional phase timestamps of CTA 0 (masr_ctc_debug_enable): [start, setup, phase0, chains, ll, end]
#define CTC_STAMP(i) do { if (dbg != nullptr && b == 0 && tid == 0) dbg[i] = clock64(); } while (0)
  CTC_STAMP(0);
  const int Sstride = 2 * Lmax + 1, L1stride = Lmax + 1;

  float* logZ2 = reinterpret_cast<float*>(smem_raw);                 // [T]
  int* tg = reinterpret_cast<int*>(logZ2 + T);                       // [Lmax]
  int* slot = tg + Lmax;                                             // [Lmax+1]
  int* cmap = slot + (Lmax + 1);                                     // [C]
  float* Gw = reinterpret_cast<float*>(cmap + C);                    // [NW][CTC_F][L1stride] per-warp class posteriors
  float* red = Gw + NW * CTC_F * L1stride;                                   // [4]
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

  CTC_STAMP(1);
  // ---- phase 0: log2-domain normaliser and emission table.  A warp handles CTC_F frames at a time and
  // issues all their row loads (CTC_F x ceil(C/32) independent 128 B requests per warp) before reducing:
  // the phase is bounded by memory-level parallelism, not by per-frame round trips.
  const bool fast_c = C <= 32 * CTC_CPL;
  if (fast_c) {
    for (int t0 = warp * CTC_F; t0 < Tb; t0 += NW * CTC_F) {
      float x[CTC_F][CTC_CPL];
#pragma unroll
      for (int f = 0; f < CTC_F; ++f) {
        const float* row = acts + (int64_t(min(t0 + f, Tb - 1)) * st_t + b * st_b);
#pragma unroll
        for (int k = 0; k < CTC_CPL; ++k) { const int c = lane + 32 * k; x[f][k] = c < C ? __ldg(row + c) : CTC_NEG; }
      }
      float z2[CTC_F];
#pragma unroll
      for (int f = 0; f < CTC_F; ++f) {
        z2[f] = 0.f;
        if (!is_logprob) {
          float mx = CTC_NEG;
#pragma unroll
          for (int k = 0; k < CTC_CPL; ++k) mx = fmaxf(mx, x[f][k]);
          mx = warp_max(mx);
          float se = 0.f;
#pragma unroll
          for (int k = 0; k < CTC_CPL; ++k) se += ex2f((x[f][k] - mx) * CTC_LOG2E);
          se = warp_sum(se);
          z2[f] = mx * CTC_LOG2E + lg2f(se);
        }
      }
#pragma unroll
      for (int f = 0; f < CTC_F; ++f) {
        const int t = t0 + f;
        if (t < Tb) {
          const float* row = acts + (int64_t(t) * st_t + b * st_b);
          if (lane == 0) logZ2[t] = z2[f];
          float* Et = E2 + size_t(t) * L1stride;
          for (int j = lane; j <= L; j += 32) Et[j] = __ldg(row + (j < L ? tg[j] : blank)) * CTC_LOG2E - z2[f];
        }
      }
    }
  } else {
    for (int t = warp; t < Tb; t += NW) {
      const float* row = acts + (int64_t(t) * st_t + b * st_b);
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
  }
  __syncthreads();

  CTC_STAMP(2);
  // ---- phase 1: alpha in warp 0, beta in warp 1 (registers + shuffles, no block barrier)
  if (Tb > 0 && warp < 2) {
    if (warp == 0) ctc_chain<SPL, true>(A, E2, tg, lane, L, S, Tb, Sstride, L1stride, blank);
    else           ctc_chain<SPL, false>(Bt, E2, tg, lane, L, S, Tb, Sstride, L1stride, blank);
  }
  __syncthreads();
  if (!tables_in_smem) __threadfence_block();
  CTC_STAMP(3);

  if (tid == 0) {
    float ll2;
    if (Tb > 0) {
      const float* last = A + size_t(Tb - 1) * Sstride;
      ll2 = lse3_2(last[S - 1], S > 1 ? last[S - 2] : CTC_NEG, CTC_NEG);
    } else {
      ll2 = (S == 1) ? 0.f : CTC_NEG;
    }
    red[0] = ll2;
  }
  __syncthreads();
  const float ll2 = red[0];
  const bool feasible = (ll2 > 0.5f * CTC_NEG);
  if (tid == 0) {
    float nll = feasible ? -ll2 * CTC_LN2 : INFINITY;
    if (!feasible && zero_infinity) nll = 0.f;
    nll_out[b] = nll;
    if (loss_out != nullptr) atomicAdd(loss_out, nll / float(max(L, 1)) / float(B));
  }
  if (grad == nullptr) return;
  const float scale = grad_scale / (float(B) * float(max(L, 1)));

  // ---- phase 3: posteriors + gradient rows.  Again CTC_F frames per warp iteration: the activation rows are
  // requested first, the class posteriors of the frames are formed while those loads are in flight
  // (shuffle sum for the blank, shared-memory atomics only where a label repeats), then each gradient row
  // is written once, coalesced.
  float* Gbase = Gw + warp * CTC_F * L1stride;
  int ucls[CTC_CPL];
#pragma unroll
  for (int k = 0; k < CTC_CPL; ++k) { const int c = lane + 32 * k; ucls[k] = (fast_c && c < C) ? cmap[c] : -1; }
  int pcol[SPL], pslot[SPL];          // per strided state: emission column, posterior slot (-1 = blank: shuffle sum)
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s2 = lane + 32 * i;
    const bool odd = s2 & 1;
    pcol[i] = (odd && s2 < S) ? (s2 >> 1) : L;
    pslot[i] = (odd && s2 < S) ? slot[s2 >> 1] : -1;
    if (pslot[i] == L) pslot[i] = -1;  // a label equal to the blank class joins the blank sum
  }
  for (int t0 = warp * CTC_F; t0 < T; t0 += NW * CTC_F) {
    float x[CTC_F][CTC_CPL];
    if (fast_c) {
#pragma unroll
      for (int f = 0; f < CTC_F; ++f) {
        const int t = t0 + f;
        const bool live = feasible && t < Tb;
        const float* row = acts + (int64_t(live ? t : 0) * st_t + b * st_b);
#pragma unroll
        for (int k = 0; k < CTC_CPL; ++k) { const int c = lane + 32 * k; x[f][k] = (live && c < C) ? __ldg(row + c) : 0.f; }
      }
    }
    // class posteriors of the CTC_F frames: states in the outer loop (strided over lanes), frames unrolled
    // inside -> CTC_F independent load/exp chains per state (ILP) instead of one frame at a time
    for (int j = lane; j < CTC_F * L1stride; j += 32) Gbase[j] = 0.f;
    __syncwarp();
    float blank_sum[CTC_F];
#pragma unroll
    for (int f = 0; f < CTC_F; ++f) blank_sum[f] = 0.f;
    if (feasible && Tb > 0) {
#pragma unroll
      for (int i = 0; i < SPL; ++i) {
        const int s2 = lane + 32 * i;
        if (s2 < S) {
#pragma unroll
          for (int f = 0; f < CTC_F; ++f) {
            const int t = min(t0 + f, Tb - 1);
            const float g = ex2f(A[size_t(t) * Sstride + s2] + Bt[size_t(t) * Sstride + s2] - E2[size_t(t) * L1stride + pcol[i]] - ll2);
            if (t0 + f < Tb) {
              if (pslot[i] >= 0) { if (g != 0.f) atomicAdd(&Gbase[f * L1stride + pslot[i]], g); }
              else blank_sum[f] += g;
            }
          }
        }
      }
    }
#pragma unroll
    for (int f = 0; f < CTC_F; ++f) blank_sum[f] = warp_sum(blank_sum[f]);
    __syncwarp();
    if (lane < CTC_F) {
      float bs = blank_sum[0];
#pragma unroll
      for (int f = 1; f < CTC_F; ++f) if (lane == f) bs = blank_sum[f];
      Gbase[lane * L1stride + L] += bs;
    }
    __syncwarp();
#pragma unroll
    for (int f = 0; f < CTC_F; ++f) {
      const int t = t0 + f;
      if (t >= T) continue;
      float* grow = grad + (int64_t(t) * st_t + b * st_b);
      if (t >= Tb || !feasible) {
        const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
        for (int c = lane; c < C; c += 32) grow[c] = fill;
        continue;
      }
      const float* G = Gbase + f * L1stride;
      const float z2 = logZ2[t];
      if (fast_c) {
#pragma unroll
        for (int k = 0; k < CTC_CPL; ++k) {
          const int c = lane + 32 * k;
          if (c < C) {
            const float pr = ex2f(x[f][k] * CTC_LOG2E - z2);
            grow[c] = (pr - (ucls[k] >= 0 ? G[ucls[k]] : 0.f)) * scale;
          }
        }
      } else {
        const float* row = acts + (int64_t(t) * st_t + b * st_b);
        for (int c = lane; c < C; c += 32) {
          const float pr = ex2f(row[c] * CTC_LOG2E - z2);
          const int u = cmap[c];
          grow[c] = (pr - (u >= 0 ? G[u] : 0.f)) * scale;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  CTC_STAMP(5);
#undef CTC_STAMP
}

static long long* g_ctc_dbg = nullptr;

// ctc3.cu: v3 kernel (single posterior table); returns CTC3_NOT_APPLICABLE when the shape does not fit shared memory
constexpr int CTC3_NOT_APPLICABLE = 12345;
int ctc3_try(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets, const int64_t* tgt_offsets,
             const int64_t* in_lens, const int64_t* tgt_lens, int Lmax, int blank, int zero_infinity, float grad_scale,
             float* nll, float* loss, float* grad, long long* dbg, cudaStream_t st);

static size_t ctc2_small_bytes(int T, int Lmax, int C) {
  return sizeof(float) * size_t(T) + sizeof(int) * (size_t(Lmax) + (Lmax + 1) + C) +
         sizeof(float) * (size_t(CTC_THREADS / 32) * CTC_F * (Lmax + 1) + 4);
}
static size_t ctc2_table_bytes(int T, int Lmax) {
  return sizeof(float) * size_t(T) * (2 * (2 * size_t(Lmax) + 1) + (size_t(Lmax) + 1));
}

template <int SPL>
static int launch_ctc2(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets,
                       const int64_t* tgt_offsets, const int64_t* in_lens, const int64_t* tgt_lens, int Lmax,
                       int blank, int zero_infinity, float grad_scale, float* nll, float* loss, float* grad,
                       float* ws, bool in_smem, size_t smem, cudaStream_t st) {
  MASR_CHECK_CUDA(cudaFuncSetAttribute(ctc_fwd_bwd_v2_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  ctc_fwd_bwd_v2_kernel<SPL><<<B, CTC_THREADS, smem, st>>>(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens,
                                                           Lmax, blank, zero_infinity, grad_scale, nll, loss, grad, ws,
                                                           in_smem ? 1 : 0, g_ctc_dbg);
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
  return masr_ctc_fwd_bwd_ex(acts, T, B, C, int64_t(B) * C, int64_t(C), act_is_logprob, targets, tgt_offsets, in_lens, tgt_lens,
                             max_tgt_len, blank, zero_infinity, grad_scale, nll, loss, grad, workspace, workspace_bytes, stream);
}

extern "C" int masr_ctc_fwd_bwd_ex(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int act_is_logprob,
                                   const int64_t* targets, const int64_t* tgt_offsets,
                                   const int64_t* in_lens, const int64_t* tgt_lens, int max_tgt_len,
                                   int blank, int zero_infinity, float grad_scale,
                                   float* nll, float* loss, float* grad,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  MASR_REQUIRE(T >= 0 && B >= 0 && C > 0 && max_tgt_len >= 0, "ctc: bad sizes");
  MASR_REQUIRE(st_t >= C && st_b >= C, "ctc: frame / utterance strides must be >= C (rows are contiguous)");
  MASR_REQUIRE(blank >= 0 && blank < C, "ctc: blank out of range");
  cudaStream_t st = as_stream(stream);
  if (loss != nullptr) { ctc_zero_kernel<<<1, 1, 0, st>>>(loss); MASR_LAUNCH_CHECK(); }
  if (B == 0) return MASR_OK;
  {
    // v3 (single posterior table) when its shared-memory footprint fits one CTA; else v2 (warp-shuffle recursions,
    // tables in the workspace) whenever the extended label sequence fits 32 lanes x 12 states
    const int S = 2 * max_tgt_len + 1;
    const int spl = (S + 31) / 32;
    {
      const int rc3 = ctc3_try(acts, T, B, C, st_t, st_b, act_is_logprob, targets, tgt_offsets, in_lens, tgt_lens, max_tgt_len, blank,
                               zero_infinity, grad_scale, nll, loss, grad, g_ctc_dbg, st);
      if (rc3 != CTC3_NOT_APPLICABLE) return rc3;
    }
    const size_t small2 = ctc2_small_bytes(T, max_tgt_len, C), tables2 = ctc2_table_bytes(T, max_tgt_len);
    const bool fits = small2 + tables2 <= CTC_SMEM_LIMIT;
    const bool ws_ok = workspace != nullptr && workspace_bytes >= ctc_table_bytes(T, max_tgt_len) * size_t(B);
    if (spl <= 12 && small2 <= 100 * 1024 && (fits || ws_ok)) {
      const size_t smem2 = fits ? small2 + tables2 : small2;
      float* wsf = static_cast<float*>(workspace);
#define CTC2_CASE(N) return launch_ctc2<N>(acts, T, B, C, st_t, st_b, act_is_logprob, targets, tgt_offsets, in_lens, tgt_lens, max_tgt_len, \
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
  ctc_fwd_bwd_kernel<<<B, CTC_THREADS, smem, st>>>(acts, T, B, C, st_t, st_b, act_is_logprob, targets, tgt_offsets, in_lens,
                                                   tgt_lens, max_tgt_len, blank, zero_infinity, grad_scale,
                                                   nll, loss, grad, static_cast<float*>(workspace), in_smem ? 1 : 0);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

// Profiling hook: phase timestamps (SM clock) of CTA 0 of the next CTC launches; out[6] on the host.
extern "C" int masr_ctc_debug_enable(int on) {
  if (on && g_ctc_dbg == nullptr) { MASR_CHECK_CUDA(cudaMalloc(&g_ctc_dbg, 8 * sizeof(long long))); }
  if (!on && g_ctc_dbg != nullptr) { cudaFree(g_ctc_dbg); g_ctc_dbg = nullptr; }
  return MASR_OK;
}
extern "C" int masr_ctc_debug_read(long long* out6) {
  MASR_REQUIRE(g_ctc_dbg != nullptr, "ctc debug not enabled");
  MASR_CHECK_CUDA(cudaDeviceSynchronize());
  MASR_CHECK_CUDA(cudaMemcpy(out6, g_ctc_dbg, 6 * sizeof(long long), cudaMemcpyDeviceToHost));
  return MASR_OK;
}
