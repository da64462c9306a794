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
