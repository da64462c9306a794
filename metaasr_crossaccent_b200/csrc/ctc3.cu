// Kernel 1: log-space CTC alpha-beta forward-backward, one CTA per utterance, ONE posterior table, pipelined in the CTA.
//
// Same contract as ctc.cu (replaces F.log_softmax + nn.CTCLoss(blank, 'mean', zero_infinity) and its backward,
// src/blstm_trainer.py:22,65-70).
//   * ONE table.  The alpha warp walks t = 0 .. Tb-1, the beta warp t = Tb-1 .. 0, concurrently.  Until they meet in the
//     middle each stores its own values; afterwards every frame they reach already holds the other recursion's values,
//     so they store the combined log2 posterior P[t][s] = alpha_t(s) + beta_t(s) - E_t(s) in place;
//   * emissions and class posteriors are indexed by SLOT (= first position of the class in the target, L = blank,
//     Lmax+1 = a dummy slot for classes that do not occur).  Every lane keeps the slot of its 12 classes in registers:
//     the emission workers scatter x*log2e - logZ to E[slot][t] with unconditional shared stores (no scattered global
//     gather), the gradient workers read the class posterior G[slot] with unconditional shared loads (the dummy slot
//     holds 0): no branches, no predicates, no shared-memory atomics in the streaming loops;
//   * the lse of the three predecessors is 1 + 2^(lo1-m) + 2^(lo2-m): two MUFU.EX2 + one MUFU.LG2 per state.  (A
//     linear-domain recursion with per-lane block exponents -- adds and multiplies on the dependent chain, logarithms off
//     it -- was built and measured in round 2: same 252 us at 2 048 utterances.  With three co-resident CTAs the SM is
//     bound by the issue slots of the streaming warps (2.2 IPC of 4, 35 % warp occupancy), not by the recursion chain,
//     so the exact log-domain step stays);
//   * six streaming warps produce emission frames from both ends towards the middle and announce them per frame; the two
//     recursion warps poll eight frames ahead; the log-likelihood is taken at the meeting frame, so the streaming warps
//     turn into gradient workers that follow the recursions outward from the middle;
//   * gradient rows are written with streaming stores (evict-first) so they do not push the activation rows out of L2
//     before the gradient workers re-read them.
#include "common.cuh"

namespace masr {

constexpr int CTC3_NOT_APPLICABLE = 12345;

namespace {

constexpr int THREADS = 256;
constexpr int NW = THREADS / 32;
constexpr int CPL = 12;                 // classes per lane in registers (fast path: 32 <= C <= 384)
constexpr float NEG = -1.0e30f;         // "impossible": finite, so the log-sum-exp needs no special cases
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one recursion step for the SPL states of a lane: a <- lse(a[s], a[s -+ 1], skip ? a[s -+ 2]) + e
template <int SPL, bool FWD>
__device__ __forceinline__ void recur(float (&a)[SPL], const float (&e)[SPL], const float (&skipadd)[SPL],
                                      const float (&validadd)[SPL], int lane) {
  float n1, n2;                                                            // neighbours across the lane boundary
  if (FWD) {
    if (SPL >= 2) { n1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1); n2 = __shfl_up_sync(0xffffffffu, a[SPL >= 2 ? SPL - 2 : 0], 1); }
    else          { n1 = __shfl_up_sync(0xffffffffu, a[0], 1); n2 = __shfl_up_sync(0xffffffffu, a[0], 2); if (lane < 2) n2 = NEG; }
    if (lane == 0) { n1 = NEG; n2 = NEG; }
  } else {
    if (SPL >= 2) { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[SPL >= 2 ? 1 : 0], 1); }
    else          { n1 = __shfl_down_sync(0xffffffffu, a[0], 1); n2 = __shfl_down_sync(0xffffffffu, a[0], 2); if (lane > 29) n2 = NEG; }
    if (lane == 31) { n1 = NEG; n2 = NEG; }
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
    x2 += skipadd[i];                    // 0 where the skip transition exists, NEG elsewhere (no predicates in the chain)
    const float hi = fmaxf(a[i], x1), lo = fminf(a[i], x1);
    const float m = fmaxf(hi, x2), lo2 = fminf(hi, x2);
    const float v = m + lg2f(1.f + ex2f(lo - m) + ex2f(lo2 - m)) + (e[i] + validadd[i]);
    nw[i] = fmaxf(v, NEG);               // states past S carry validadd = NEG and stay at NEG
  }
#pragma unroll
  for (int i = 0; i < SPL; ++i) a[i] = nw[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Pipelined variant (ctc3p_kernel): the two recursion warps run CONCURRENTLY with the six streaming warps of the
// same CTA instead of between two block barriers.
//   emission frames are produced from both ends towards the middle (what alpha / beta consume first) and announced
//   per frame (erdy[t]); the recursions poll the flag of the frame they are about to load;
//   at the meeting point the alpha warp knows log2 P(labels | x) already (sum over s of alpha_t(s) beta_t(s) / y_t(s)
//   is the same for every t) and publishes it; from then on every step finalises one frame per recursion warp and
//   bumps a counter (fin[0] alpha side, fin[1] beta side); the streaming warps turn into gradient workers that take
//   finalised frame pairs outward from the middle.
// Cheaper warp reductions for the streaming warps of the pipelined kernel (their shuffles share the MIO pipe with
// the recursion warps' shuffles): max through one integer REDUX (order-preserving float <-> int map), the sums of two
// frames through a transposed butterfly (lanes 0-15 reduce frame 0, lanes 16-31 frame 1: 5 shuffles instead of 10).
__device__ __forceinline__ float warp_max_redux(float v) {
  int i = __float_as_int(v);
  i ^= (i >> 31) & 0x7fffffff;
  i = __reduce_max_sync(0xffffffffu, i);
  i ^= (i >> 31) & 0x7fffffff;
  return __int_as_float(i);
}
// returns the total of a (lanes 0-15) / of b (lanes 16-31)
__device__ __forceinline__ float warp_sum2_split(float a, float b, int lane) {
  const bool hi = (lane & 16) != 0;
  float r = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

struct PipeSync {
  volatile unsigned char* erdy;      // [T]  emission column of frame t is complete
  volatile int* fin;                 // [0] frames finalised by alpha (mid, mid+1, ..), [1] by beta (mid-1, mid-2, ..), [2] ll ready
  float* red;                        // [0] log2 likelihood
};

// acquire / release at CTA scope (lighter than __threadfence_block(), which is a sequentially consistent fence)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

__device__ __forceinline__ void wait_frame(const PipeSync& ps, int t) {
  while (ps.erdy[t] == 0) __nanosleep(20);
  fence_cta();
}
// frames t + tstep .. t + n * tstep (n <= 8) at once: lane l polls one flag, one vote and one fence per eight steps
// keep the flag round trip off the per-frame dependent chain
__device__ __forceinline__ void wait_ahead(const PipeSync& ps, int t, int tstep, int n, int lane) {
  const bool mine = lane < n;
  const int tf = mine ? t + (lane + 1) * tstep : t;
  while (!__all_sync(0xffffffffu, !mine || ps.erdy[tf] != 0)) __nanosleep(20);
  fence_cta();
}

template <int SPL, bool FWD>
__device__ __forceinline__ void chain_pipe(float* __restrict__ P, const float* __restrict__ E2, const int* __restrict__ tg,
                                           const short* __restrict__ cmap, int lane, int L, int S, int Tb, int Sstride,
                                           int TP, const PipeSync ps) {
  float a[SPL], e[SPL], skipadd[SPL], validadd[SPL];
  bool valid[SPL];
  const float* pe[SPL];
  const int t0 = FWD ? 0 : Tb - 1;
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    const int j = s >> 1;
    const bool odd = s & 1;
    valid[i] = s < S;
    pe[i] = E2 + ((odd && j < L) ? int(cmap[tg[j]]) : L) * TP + t0;
    bool skip;
    if (FWD) skip = odd && s > 1 && s < S && tg[j] != tg[j - 1];
    else     skip = odd && s + 2 < S && tg[j] != tg[j + 1];
    skipadd[i] = skip ? 0.f : NEG;
    validadd[i] = valid[i] ? 0.f : NEG;
  }
  const int mid = Tb >> 1;
  const int npre = FWD ? mid : Tb - mid;
  const int tstep = FWD ? 1 : -1, sstep = FWD ? Sstride : -Sstride;
  float* dst = P + t0 * Sstride + lane * SPL;
  int t = t0;
  wait_frame(ps, t);
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    e[i] = *pe[i];
    const bool start = FWD ? (s <= 1) : (s >= S - 2);
    a[i] = (start && valid[i]) ? e[i] : NEG;
  }
  int k = 0;
  for (; k < npre; ++k) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) if (valid[i]) dst[i] = a[i];
    if (k + 1 < Tb) {
      if ((k & 7) == 0) wait_ahead(ps, t, tstep, min(8, Tb - 1 - k), lane);
      dst += sstep;
      t += tstep;
#pragma unroll
      for (int i = 0; i < SPL; ++i) { pe[i] += tstep; e[i] = *pe[i]; }
      recur<SPL, FWD>(a, e, skipadd, validadd, lane);
    }
  }
  bar_sync_named(1, 64);
  for (; k < Tb; ++k) {
    float o[SPL], pc[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) o[i] = valid[i] ? dst[i] : 0.f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) { pc[i] = (a[i] - e[i]) + o[i]; if (valid[i]) dst[i] = pc[i]; }
    if (FWD && k == npre) {
      // log2 P(labels | x) = log2 sum_s 2^(alpha + beta - E) at the meeting frame
      float m = NEG;
#pragma unroll
      for (int i = 0; i < SPL; ++i) if (valid[i]) m = fmaxf(m, pc[i]);
      m = warp_max(m);
      float se = 0.f;
#pragma unroll
      for (int i = 0; i < SPL; ++i) if (valid[i]) se += ex2f(fmaxf(pc[i], 2.f * NEG) - m);
      se = warp_sum(se);
      if (lane == 0) ps.red[0] = se > 0.f ? fmaxf(m + lg2f(se), NEG) : NEG;
    }
    // progress is published per frame PAIR (what a gradient worker takes) and at the last frame
    const int nfin = k - npre + 1;
    if ((FWD && k == npre) || (nfin & 1) == 0 || k + 1 == Tb) {
      __syncwarp();
      if (lane == 0) {
        fence_cta();
        if (FWD && k == npre) ps.fin[2] = 1;
        ps.fin[FWD ? 0 : 1] = nfin;
      }
    }
    if (k + 1 < Tb) {
      if ((k & 7) == 0) wait_ahead(ps, t, tstep, min(8, Tb - 1 - k), lane);
      dst += sstep;
      t += tstep;
#pragma unroll
      for (int i = 0; i < SPL; ++i) { pe[i] += tstep; e[i] = *pe[i]; }
      recur<SPL, FWD>(a, e, skipadd, validadd, lane);
    }
  }
}

constexpr int EXMAX = 32;           // label positions beyond a class's second occurrence (or equal to the blank class)

__host__ __device__ inline int ctc3_tp(int T, int F3) { const int t = T > NW * F3 ? T : NW * F3; return t | 1; }

__host__ __device__ inline size_t smem_bytes_dev(int T, int Lmax, int C, int spl_dispatched) {
  const size_t npos = 32 * size_t(spl_dispatched / 2 + 1);
  const size_t b = sizeof(float) * (size_t(T) + 4 + size_t(Lmax + 3) * ctc3_tp(T, 2) + size_t(T) * (2 * size_t(Lmax) + 2)) +
                   sizeof(int) * size_t(Lmax + 1) + sizeof(short) * (size_t(C) + 3 * npos + 2 * EXMAX);
  return (b + 15) / 16 * 16;
}

template <int F3>
inline size_t smem_bytes(int T, int Lmax, int C, int spl_dispatched) {
  const size_t npos = 32 * size_t(spl_dispatched / 2 + 1);      // 32 * NJ
  const size_t b = sizeof(float) * (size_t(T) + 4 + size_t(Lmax + 3) * ctc3_tp(T, F3) + size_t(T) * (2 * size_t(Lmax) + 2)) +
                   sizeof(int) * size_t(Lmax + 1) + sizeof(short) * (size_t(C) + 3 * npos + 2 * EXMAX);
  return (b + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------------------------------------------
inline size_t smem_bytes_pipe(int T, int Lmax, int C, int spl_dispatched) {
  return smem_bytes<2>(T, Lmax, C, spl_dispatched) + sizeof(float) * size_t(NW * 2 * (Lmax + 3)) + size_t((T + 15) / 16 * 16) + 16;
}

template <int SPL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ctc3p_kernel(const float* __restrict__ acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob,
             const int64_t* __restrict__ targets, const int64_t* __restrict__ tgt_offsets,
             const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
             int Lmax, int blank, int zero_infinity, float grad_scale,
             float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ grad,
             long long* __restrict__ dbg) {
  constexpr int F = 2;               // frames per worker iteration (emission and gradient)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#define CTC_STAMP(i) do { if (dbg != nullptr && b == 0 && tid == 0) dbg[i] = clock64(); } while (0)
  CTC_STAMP(0);
  const int Sstride = 2 * Lmax + 2, NEGCOL = 2 * Lmax + 1;
  const int NSLOT = Lmax + 3, DUMMY = Lmax + 1, TRASH = Lmax + 2;
  const int TP = ctc3_tp(T, F);
  constexpr int NJ = SPL / 2 + 1;

  float* logZ2 = reinterpret_cast<float*>(smem_raw);             // [T]
  float* red = logZ2 + T;                                        // [4]
  float* E2 = red + 4;                                           // [NSLOT][TP]
  float* P = E2 + size_t(NSLOT) * TP;                            // [T][Sstride]
  int* tg = reinterpret_cast<int*>(P + size_t(T) * Sstride);     // [Lmax]
  int* nextra = tg + Lmax;                                       // [1]
  short* cmap = reinterpret_cast<short*>(nextra + 1);            // [C]
  short* colA = cmap + C;
  short* colB = colA + 32 * NJ;
  short* gslot = colB + 32 * NJ;
  short* excol = gslot + 32 * NJ;
  short* exslot = excol + EXMAX;
  // pipeline extras behind the (16 B rounded) v3 layout
  unsigned char* xtra = smem_raw + smem_bytes_dev(T, Lmax, C, SPL);
  float* Gall = reinterpret_cast<float*>(xtra);                  // [NW][F][NSLOT] class posteriors per warp
  int* fin = reinterpret_cast<int*>(Gall + NW * F * NSLOT);      // [4]
  unsigned char* erdy = reinterpret_cast<unsigned char*>(fin + 4);   // [T]

  const int L = int(tgt_lens[b]);
  const int Tb = min(T, int(in_lens[b]));
  const int S = 2 * L + 1;
  const int64_t toff = tgt_offsets[b];

  for (int j = tid; j < L; j += THREADS) tg[j] = int(targets[toff + j]);
  for (int c = tid; c < C; c += THREADS) cmap[c] = short(c == blank ? L : DUMMY);
  for (int t = tid; t < T; t += THREADS) erdy[t] = 0;
  if (tid < 4) fin[tid] = 0;
  if (tid == 0) *nextra = 0;
  __syncthreads();
  for (int j = tid; j < 32 * NJ; j += THREADS) {
    int ca = NEGCOL, cb = NEGCOL, gs = TRASH;
    if (j < L) {
      const int cls = tg[j];
      int firstpos = j, rank = 0, nx = -1;
      if (cls != blank) {
        for (int i = j - 1; i >= 0; --i) if (tg[i] == cls) { firstpos = i; ++rank; }
        if (rank == 0) {
          for (int i = j + 1; i < L; ++i) if (tg[i] == cls) { nx = i; break; }
          cmap[cls] = short(j);
          ca = 2 * j + 1; gs = j;
          if (nx >= 0) cb = 2 * nx + 1;
        }
      } else {
        firstpos = L; rank = 2;
      }
      if (rank >= 2) {
        const int e = atomicAdd(nextra, 1);
        if (e < EXMAX) { excol[e] = short(2 * j + 1); exslot[e] = short(firstpos); }
      }
    }
    colA[j] = short(ca); colB[j] = short(cb); gslot[j] = short(gs);
  }
  for (int t = tid; t < T; t += THREADS) P[t * Sstride + NEGCOL] = NEG;
  __syncthreads();
  const int nex = *nextra;
  const bool fast_c = C >= 32 && C <= 32 * CPL;
  const bool pipe0 = fast_c && Tb >= 4;                          // emission frames from both ends, two at a time
  const bool pipe3 = pipe0 && nex <= EXMAX && grad != nullptr;   // gradient rows by the workers, behind the recursions
  const int mid = Tb >> 1;
  const float scale = grad_scale / (float(B) * float(max(L, 1)));
  const PipeSync ps{erdy, fin, red};
  CTC_STAMP(1);

  const int cw = (__popc(unsigned(b)) & 1) * 2;                  // recursion warps {0,1} or {2,3}
  if (warp == cw || warp == cw + 1) {
    // =========================================================== recursion warps
    if (Tb > 0) {
      if (warp == cw) chain_pipe<SPL, true>(P, E2, tg, cmap, lane, L, S, Tb, Sstride, TP, ps);
      else            chain_pipe<SPL, false>(P, E2, tg, cmap, lane, L, S, Tb, Sstride, TP, ps);
      CTC_STAMP(2);                  // CTA 0: warp 0 is the alpha warp -> [1..2] = recursion, [2..3] = gradient tail
    } else if (warp == cw && lane == 0) {
      red[0] = (S == 1) ? 0.f : NEG;
    }
  } else {
    // =========================================================== streaming warps
    // (measured: leaving the two warps that share a scheduler with a recursion warp idle does not speed the
    // recursion up -- 57 k vs 55 k cycles -- and costs 20 % throughput: the contention is SM-wide, MIO / LSU)
    constexpr int NWK = NW - 2;
    const int wi = warp - (warp > cw ? 2 : 0);                   // 0 .. NWK-1
    // ---- emissions
    if (pipe0) {
      float* ek[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; ek[k] = E2 + (c < C ? int(cmap[c]) : DUMMY) * TP; }
      const int nfront = (mid + 1) / 2, nback = (Tb - mid + 1) / 2;      // nfront <= nback <= nfront + 1
      for (int q = wi; q < nfront + nback; q += NWK) {
        int t0;
        if (q < 2 * nfront && (q & 1) == 0) { t0 = 2 * (q >> 1); if (t0 + 1 >= mid) t0 = mid - 2; }       // front, upward
        else { const int i = q < 2 * nfront ? (q >> 1) : nfront; t0 = Tb - 2 - 2 * i; if (t0 < mid) t0 = mid; }   // back, downward
        float x[F][CPL];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float* row = acts + (int64_t(t0 + f) * st_t + b * st_b) + lane;
#pragma unroll
          for (int k = 0; k < CPL; ++k) x[f][k] = (lane + 32 * k < C) ? __ldg(row + 32 * k) : NEG;
        }
        float z2[F] = {0.f, 0.f};
        if (!is_logprob) {
          float mx[F], se[F];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            mx[f] = x[f][0];
#pragma unroll
            for (int k = 1; k < CPL; ++k) mx[f] = fmaxf(mx[f], x[f][k]);
            mx[f] = warp_max_redux(mx[f]) * LOG2E;
            se[f] = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) se[f] += ex2f(fmaf(x[f][k], LOG2E, -mx[f]));
          }
          const float mine = warp_sum2_split(se[0], se[1], lane);          // frame 0 in lanes 0-15, frame 1 in 16-31
          const float other = __shfl_xor_sync(0xffffffffu, mine, 16);
          const bool hi = (lane & 16) != 0;
          z2[0] = mx[0] + lg2f(hi ? other : mine);
          z2[1] = mx[1] + lg2f(hi ? mine : other);
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          float* d = ek[k] + t0;
#pragma unroll
          for (int f = 0; f < F; ++f) d[f] = fmaf(x[f][k], LOG2E, -z2[f]);
        }
        if (lane < F) logZ2[t0 + lane] = lane == 0 ? z2[0] : z2[1];
        __syncwarp();
        if (lane == 0) { fence_cta(); ps.erdy[t0] = 1; ps.erdy[t0 + 1] = 1; }
      }
    } else {
      for (int t = wi; t < Tb; t += NWK) {
        const float* row = acts + (int64_t(t) * st_t + b * st_b);
        float z2 = 0.f;
        if (!is_logprob) {
          float mx = -INFINITY;
          for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
          mx = warp_max(mx);
          float se = 0.f;
          for (int c = lane; c < C; c += 32) se += ex2f((row[c] - mx) * LOG2E);
          se = warp_sum(se);
          z2 = mx * LOG2E + lg2f(se);
        }
        if (lane == 0) logZ2[t] = z2;
        for (int j = lane; j <= L; j += 32) {
          const int cls = j < L ? tg[j] : blank;
          if (j == L || int(cmap[cls]) == j) E2[j * TP + t] = row[cls] * LOG2E - z2;
        }
        __syncwarp();
        if (lane == 0) { fence_cta(); ps.erdy[t] = 1; }
      }
    }
    // ---- gradient rows of finalised frames, outward from the meeting point
    if (pipe3) {
      while (ps.fin[2] == 0) __nanosleep(100);      // (volatile view: the plain pointer would be hoisted)
      fence_cta();
      const float ll2 = red[0];
      if (ll2 > 0.5f * NEG) {
        float* Gw = Gall + warp * (F * NSLOT);
        if (lane < F) Gw[DUMMY * F + lane] = 0.f;
        const float* gk[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) { const int c = lane + 32 * k; gk[k] = Gw + (c < C ? int(cmap[c]) : DUMMY) * F; }
        __syncwarp();
        const int nA = (Tb - mid + 1) / 2, nB = (mid + 1) / 2;           // nB <= nA <= nB + 1
        for (int q = wi; q < nA + nB; q += NWK) {
          int t0, need, side;
          if (q < 2 * nB && (q & 1) == 1) { side = 1; const int j = q >> 1; t0 = mid - 2 - 2 * j; need = 2 * j + 2; if (t0 < 0) { t0 = 0; need = mid; } }
          else { side = 0; const int j = q < 2 * nB ? (q >> 1) : nB; t0 = mid + 2 * j; need = 2 * j + 2; if (t0 + 1 >= Tb) { t0 = Tb - 2; need = Tb - mid; } }
          float x[F][CPL];
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float* row = acts + (int64_t(t0 + f) * st_t + b * st_b) + lane;
#pragma unroll
            for (int k = 0; k < CPL; ++k) x[f][k] = (lane + 32 * k < C) ? __ldcs(row + 32 * k) : 0.f;
          }
          while (ps.fin[side] < need) __nanosleep(100);
          fence_cta();
          const float* Pt = P + t0 * Sstride;
          float bs[F];
#pragma unroll
          for (int f = 0; f < F; ++f) bs[f] = 0.f;
#pragma unroll
          for (int i = 0; i < NJ; ++i) {
            const int m = lane + 32 * i;
            const int ca = colA[m], cb = colB[m], gs = gslot[m];
            const int ce = m <= L ? 2 * m : NEGCOL;
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const float* Pf = Pt + f * Sstride;
              Gw[gs * F + f] = (ex2f(Pf[ca] - ll2) + ex2f(Pf[cb] - ll2)) * scale;
              bs[f] += ex2f(Pf[ce] - ll2);
            }
          }
          {
            const float tot = warp_sum2_split(bs[0], bs[1], lane);
            if ((lane & 15) == 0) Gw[L * F + (lane >> 4)] = tot * scale;
          }
          __syncwarp();
          if (nex > 0) {
            if (lane < F) {
              const float* Pf = Pt + lane * Sstride;
              for (int e = 0; e < nex; ++e) Gw[int(exslot[e]) * F + lane] += ex2f(Pf[excol[e]] - ll2) * scale;
            }
            __syncwarp();
          }
          float z2[F];
#pragma unroll
          for (int f = 0; f < F; ++f) z2[f] = logZ2[t0 + f];
          float* grow = grad + (int64_t(t0) * st_t + b * st_b) + lane;
          const int64_t fstride = st_t;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            if (lane + 32 * k < C) {
#pragma unroll
              for (int f = 0; f < F; ++f) {
                const float pr = ex2f(fmaf(x[f][k], LOG2E, -z2[f]));
                __stcs(grow + f * fstride + 32 * k, fmaf(pr, scale, -gk[k][f]));
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  CTC_STAMP(3);

  const float ll2 = red[0];
  const bool feasible = (ll2 > 0.5f * NEG);
  if (tid == 0) {
    float nll = feasible ? -ll2 * LN2 : INFINITY;
    if (!feasible && zero_infinity) nll = 0.f;
    nll_out[b] = nll;
    if (loss_out != nullptr) atomicAdd(loss_out, nll / float(max(L, 1)) / float(B));
  }
  if (grad == nullptr) return;
  // ---- what the pipeline did not cover: padded frames; everything for infeasible / odd-shaped utterances
  const bool done_live = pipe3 && feasible;
  for (int t = (done_live ? Tb : 0) + warp; t < T; t += NW) {
    float* grow = grad + (int64_t(t) * st_t + b * st_b);
    if (!feasible || t >= Tb) {
      const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
      for (int c = lane; c < C; c += 32) grow[c] = fill;
      continue;
    }
    const float* row = acts + (int64_t(t) * st_t + b * st_b);
    const float* Pt = P + t * Sstride;
    float bsum = 0.f;
    for (int m = lane; m <= L; m += 32) bsum += ex2f(Pt[2 * m] - ll2);
    for (int j = lane; j < L; j += 32) if (int(cmap[tg[j]]) == L) bsum += ex2f(Pt[2 * j + 1] - ll2);
    bsum = warp_sum(bsum);
    const float z2 = logZ2[t];
    for (int c = lane; c < C; c += 32) {
      const float pr = ex2f(row[c] * LOG2E - z2);
      const int u = int(cmap[c]);
      float occ = (u == L) ? bsum : 0.f;
      if (u < L) for (int j = u; j < L; ++j) if (tg[j] == c) occ += ex2f(Pt[2 * j + 1] - ll2);
      grow[c] = (pr - occ) * scale;
    }
  }
  CTC_STAMP(5);
#undef CTC_STAMP
}

template <int SPL, int MINB>
int launch_pipe(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets, const int64_t* tgt_offsets,
                const int64_t* in_lens, const int64_t* tgt_lens, int Lmax, int blank, int zero_infinity, float grad_scale,
                float* nll, float* loss, float* grad, long long* dbg, size_t smem, cudaStream_t st) {
  auto kern = ctc3p_kernel<SPL, MINB>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr_smem = smem;
  }
  kern<<<B, THREADS, smem, st>>>(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens, Lmax, blank,
                                 zero_infinity, grad_scale, nll, loss, grad, dbg);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

// ===============================================================================================================
// Long utterances (ctc3l_kernel): the posterior table of ONE utterance no longer fits shared memory (T' = 375 / 750
// frames x 2L+2 = 206 / 306 states: 0.3 / 0.9 MB; these are the reference's real lengths, max_ilen 1500 / 3000).
// The frames are processed in CHUNKS of Tc frames whose table does fit, with recomputation instead of a table in HBM:
//   sweep A (chunks 0 .. nc-2, forward):  emissions of the chunk -> alpha through the chunk, only the alpha row at the
//           chunk's last frame is kept (Abound[c], nc x S floats);
//   sweep B (chunks nc-1 .. 0, backward): emissions of the chunk again -> alpha from Abound[c-1] and beta from the carry
//           of the later chunk run through the chunk from both ends, meet in the middle and leave the log2 posterior
//           table of the chunk (same single-table scheme as above) -> gradient rows of the chunk.  The log-likelihood
//           comes out of the alpha row of the last frame, i.e. of the FIRST chunk sweep B visits.
// Activations are read twice from HBM (+ once more out of L2 by the gradient workers), gradients written once; alpha is
// computed twice.  Phases inside a chunk are separated by block barriers (emission | recursions | gradient).
constexpr int EXMAXL = 192;         // long targets repeat classes often: room for more third-and-later occurrences

template <int SPL, bool FWD>
__device__ __forceinline__ void lane_consts(const int* __restrict__ tg, const short* __restrict__ cmap, int lane, int L, int S,
                                            int (&slot)[SPL], float (&skipadd)[SPL], float (&validadd)[SPL], bool (&valid)[SPL]) {
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int s = lane * SPL + i;
    const int j = s >> 1;
    const bool odd = s & 1;
    valid[i] = s < S;
    slot[i] = (odd && j < L) ? int(cmap[tg[j]]) : L;
    bool skip;
    if (FWD) skip = odd && s > 1 && s < S && tg[j] != tg[j - 1];
    else     skip = odd && s + 2 < S && tg[j] != tg[j + 1];
    skipadd[i] = skip ? 0.f : NEG;
    validadd[i] = valid[i] ? 0.f : NEG;
  }
}

// alpha through one chunk without a table (sweep A).  fresh: the chunk starts the utterance; else a[] = alpha of the
// frame before the chunk.  On return a[] = alpha of the chunk's last frame.
template <int SPL>
__device__ __forceinline__ void chain_fwd_chunk(float (&a)[SPL], bool fresh, const float* __restrict__ E2, const int (&slot)[SPL],
                                                const float (&skipadd)[SPL], const float (&validadd)[SPL], const bool (&valid)[SPL],
                                                int lane, int Tn, int TPc) {
  float e[SPL];
  for (int k = 0; k < Tn; ++k) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) e[i] = E2[slot[i] * TPc + k];
    if (fresh && k == 0) {
#pragma unroll
      for (int i = 0; i < SPL; ++i) a[i] = (lane * SPL + i <= 1 && valid[i]) ? e[i] : NEG;
    } else {
      recur<SPL, true>(a, e, skipadd, validadd, lane);
    }
  }
}

// One recursion warp over one chunk of sweep B (alpha forward from the chunk's first frame, beta backward from its
// last), meeting the other warp in the middle; leaves the log2 posteriors of the chunk in P.  fresh: the chunk holds the
// utterance's first (alpha) / last (beta) frame; else a[] = the values of the frame just outside the chunk.  On return
// a[] = the values at the far end of the chunk (the beta warp's carry into the previous chunk).
template <int SPL, bool FWD>
__device__ __forceinline__ void chain_chunk(float (&a)[SPL], bool fresh, float* __restrict__ P, const float* __restrict__ E2,
                                            const int (&slot)[SPL], const float (&skipadd)[SPL], const float (&validadd)[SPL],
                                            const bool (&valid)[SPL], int lane, int S, int Tn, int Sstride, int TPc) {
  float e[SPL];
  const float* pe[SPL];
  const int t0 = FWD ? 0 : Tn - 1;
#pragma unroll
  for (int i = 0; i < SPL; ++i) { pe[i] = E2 + slot[i] * TPc + t0; e[i] = *pe[i]; }
  if (fresh) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const int s = lane * SPL + i;
      const bool start = FWD ? (s <= 1) : (s >= S - 2);
      a[i] = (start && valid[i]) ? e[i] : NEG;
    }
  } else {
    recur<SPL, FWD>(a, e, skipadd, validadd, lane);
  }
  const int mid = Tn >> 1;
  const int npre = FWD ? mid : Tn - mid;
  const int tstep = FWD ? 1 : -1, sstep = FWD ? Sstride : -Sstride;
  float* dst = P + t0 * Sstride + lane * SPL;
  int k = 0;
  for (; k < npre; ++k) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) if (valid[i]) dst[i] = a[i];
    if (k + 1 < Tn) {
      dst += sstep;
#pragma unroll
      for (int i = 0; i < SPL; ++i) { pe[i] += tstep; e[i] = *pe[i]; }
      recur<SPL, FWD>(a, e, skipadd, validadd, lane);
    }
  }
  bar_sync_named(1, 64);                           // everything the other warp stored so far is visible from here on
  for (; k < Tn; ++k) {
    float o[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) o[i] = valid[i] ? dst[i] : 0.f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) if (valid[i]) dst[i] = (a[i] - e[i]) + o[i];
    if (k + 1 < Tn) {
      dst += sstep;
#pragma unroll
      for (int i = 0; i < SPL; ++i) { pe[i] += tstep; e[i] = *pe[i]; }
      recur<SPL, FWD>(a, e, skipadd, validadd, lane);
    }
  }
}

__host__ __device__ inline int ctc3l_nj(int spl) { return spl / 2 + 1; }
inline size_t smem_bytes_long(int Tc, int T, int Lmax, int C, int spl_dispatched) {
  const size_t nc = size_t((T + Tc - 1) / Tc), Sstride = 2 * size_t(Lmax) + 2, npos = 32 * size_t(ctc3l_nj(spl_dispatched));
  const size_t b = sizeof(float) * (size_t(Tc) + 4 + size_t(Lmax + 3) * size_t(Tc | 1) + size_t(Tc) * Sstride + nc * Sstride) +
                   sizeof(int) * size_t(Lmax + 1) + sizeof(short) * (size_t(C) + 3 * npos + 2 * EXMAXL);
  return (b + 15) / 16 * 16;
}

template <int SPL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ctc3l_kernel(const float* __restrict__ acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob,
             const int64_t* __restrict__ targets, const int64_t* __restrict__ tgt_offsets,
             const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
             int Lmax, int blank, int zero_infinity, float grad_scale, int Tc,
             float* __restrict__ nll_out, float* __restrict__ loss_out, float* __restrict__ grad) {
  constexpr int F = 2;               // frames per warp iteration (emission and gradient)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Sstride = 2 * Lmax + 2, NEGCOL = 2 * Lmax + 1;
  const int NSLOT = Lmax + 3, DUMMY = Lmax + 1, TRASH = Lmax + 2;
  const int TPc = Tc | 1;
  const int NCmax = (T + Tc - 1) / Tc;
  constexpr int NJ = SPL / 2 + 1;

  float* logZ2 = reinterpret_cast<float*>(smem_raw);             // [Tc]
  float* red = logZ2 + Tc;                                       // [4]
  float* E2 = red + 4;                                           // [NSLOT][TPc] emissions of the chunk; gradient phase: G
  float* P = E2 + size_t(NSLOT) * TPc;                           // [Tc][Sstride]
  float* Abound = P + size_t(Tc) * Sstride;                      // [NCmax][Sstride] alpha at the last frame of each chunk
  int* tg = reinterpret_cast<int*>(Abound + size_t(NCmax) * Sstride);   // [Lmax]
  int* nextra = tg + Lmax;
  short* cmap = reinterpret_cast<short*>(nextra + 1);            // [C]
  short* colA = cmap + C;
  short* colB = colA + 32 * NJ;
  short* gslot = colB + 32 * NJ;
  short* excol = gslot + 32 * NJ;
  short* exslot = excol + EXMAXL;

  const int L = int(tgt_lens[b]);
  const int Tb = min(T, int(in_lens[b]));
  const int S = 2 * L + 1;
  const int64_t toff = tgt_offsets[b];

  for (int j = tid; j < L; j += THREADS) tg[j] = int(targets[toff + j]);
  for (int c = tid; c < C; c += THREADS) cmap[c] = short(c == blank ? L : DUMMY);
  if (tid == 0) *nextra = 0;
  __syncthreads();
  for (int j = tid; j < 32 * NJ; j += THREADS) {
    int ca = NEGCOL, cb = NEGCOL, gs = TRASH;
    if (j < L) {
      const int cls = tg[j];
      int firstpos = j, rank = 0, nx = -1;
      if (cls != blank) {
        for (int i = j - 1; i >= 0; --i) if (tg[i] == cls) { firstpos = i; ++rank; }
        if (rank == 0) {
          for (int i = j + 1; i < L; ++i) if (tg[i] == cls) { nx = i; break; }
          cmap[cls] = short(j);
          ca = 2 * j + 1; gs = j;
          if (nx >= 0) cb = 2 * nx + 1;
        }
      } else {
        firstpos = L; rank = 2;
      }
      if (rank >= 2) {
        const int e = atomicAdd(nextra, 1);
        if (e < EXMAXL) { excol[e] = short(2 * j + 1); exslot[e] = short(firstpos); }
      }
    }
    colA[j] = short(ca); colB[j] = short(cb); gslot[j] = short(gs);
  }
  for (int t = tid; t < Tc; t += THREADS) P[t * Sstride + NEGCOL] = NEG;
  __syncthreads();
  const int nex = *nextra;
  const bool fast3 = nex <= EXMAXL;
  const float scale = grad_scale / (float(B) * float(max(L, 1)));
  const int nc = (Tb + Tc - 1) / Tc;                             // chunks of THIS utterance

  // ---- emissions of chunk c (frames c*Tc .. c*Tc+Tn-1), all warps, two frames per warp iteration (C in [32, 384])
  auto emissions = [&](int c, int Tn) {
    float* ek[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { const int cc = lane + 32 * k; ek[k] = E2 + (cc < C ? int(cmap[cc]) : DUMMY) * TPc; }
    const int tbase = c * Tc;
    for (int t0 = warp * F; t0 < Tn; t0 += NW * F) {
      float x[F][CPL];
      const bool full = t0 + F <= Tn;       // a last odd frame is computed twice (idempotent stores)
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float* row = acts + (int64_t(tbase + (full ? t0 + f : min(t0 + f, Tn - 1))) * st_t + b * st_b) + lane;
#pragma unroll
        for (int k = 0; k < CPL; ++k) x[f][k] = (lane + 32 * k < C) ? __ldg(row + 32 * k) : NEG;
      }
      float z2[F];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        z2[f] = 0.f;
        if (!is_logprob) {
          float mx = x[f][0];
#pragma unroll
          for (int k = 1; k < CPL; ++k) mx = fmaxf(mx, x[f][k]);
          mx = warp_max(mx) * LOG2E;
          float se = 0.f;
#pragma unroll
          for (int k = 0; k < CPL; ++k) se += ex2f(fmaf(x[f][k], LOG2E, -mx));
          se = warp_sum(se);
          z2[f] = mx + lg2f(se);
        }
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const int tt = full ? t0 + f : min(t0 + f, Tn - 1);
#pragma unroll
        for (int k = 0; k < CPL; ++k) ek[k][tt] = fmaf(x[f][k], LOG2E, -z2[f]);
        if (lane == 0) logZ2[tt] = z2[f];
      }
    }
  };

  const int cw = (__popc(unsigned(b)) & 1) * 2;                  // recursion warps {0,1} or {2,3}
  float a[SPL], skipadd[SPL], validadd[SPL];
  int slot[SPL];
  bool valid[SPL];
  if (warp == cw) lane_consts<SPL, true>(tg, cmap, lane, L, S, slot, skipadd, validadd, valid);
  else            lane_consts<SPL, false>(tg, cmap, lane, L, S, slot, skipadd, validadd, valid);
#pragma unroll
  for (int i = 0; i < SPL; ++i) a[i] = NEG;

  // ---- sweep A: alpha rows at the chunk boundaries
  for (int c = 0; c + 1 < nc; ++c) {
    emissions(c, Tc);
    __syncthreads();
    if (warp == cw) {
      chain_fwd_chunk<SPL>(a, c == 0, E2, slot, skipadd, validadd, valid, lane, Tc, TPc);
#pragma unroll
      for (int i = 0; i < SPL; ++i) if (valid[i]) Abound[c * Sstride + lane * SPL + i] = a[i];
    }
    __syncthreads();
  }

  // ---- sweep B: posteriors and gradient rows, last chunk first
  float ll2 = (S == 1 && Tb == 0) ? 0.f : NEG;
  bool feasible = Tb == 0 ? (S == 1) : true;
  for (int c = nc - 1; c >= 0; --c) {
    const int Tn = min(Tc, Tb - c * Tc);
    emissions(c, Tn);
    __syncthreads();
    if (warp == cw) {
      if (c > 0) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) a[i] = valid[i] ? Abound[(c - 1) * Sstride + lane * SPL + i] : NEG;
      }
      chain_chunk<SPL, true>(a, c == 0, P, E2, slot, skipadd, validadd, valid, lane, S, Tn, Sstride, TPc);
      if (c == nc - 1) {                      // alpha of the last frame: log2 P(labels | x)
        float m = NEG;
#pragma unroll
        for (int i = 0; i < SPL; ++i) { const int s = lane * SPL + i; if (valid[i] && s >= S - 2) m = fmaxf(m, a[i]); }
        m = warp_max(m);
        float se = 0.f;
#pragma unroll
        for (int i = 0; i < SPL; ++i) { const int s = lane * SPL + i; if (valid[i] && s >= S - 2) se += ex2f(a[i] - m); }
        se = warp_sum(se);
        if (lane == 0) red[0] = se > 0.f ? fmaxf(m + lg2f(se), NEG) : NEG;
      }
    } else if (warp == cw + 1) {
      chain_chunk<SPL, false>(a, c == nc - 1, P, E2, slot, skipadd, validadd, valid, lane, S, Tn, Sstride, TPc);
    }
    __syncthreads();
    if (c == nc - 1) {
      ll2 = red[0];
      feasible = ll2 > 0.5f * NEG;
    }
    if (grad == nullptr || !feasible) break;

    // gradient rows of the chunk: grad[t][c] = softmax_t(c) * scale - G_t(slot(c)); the emission table is dead now, each
    // warp keeps its class posteriors G[slot][f] there (dummy slot = 0)
    float* Gw = E2 + warp * (F * NSLOT);
    if (lane < F) Gw[DUMMY * F + lane] = 0.f;
    const float* gk[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { const int cc = lane + 32 * k; gk[k] = Gw + (cc < C ? int(cmap[cc]) : DUMMY) * F; }
    __syncwarp();
    const int tbase = c * Tc;
    for (int t0 = warp * F; t0 < Tn; t0 += NW * F) {
      if (fast3 && t0 + F <= Tn) {
        float x[F][CPL];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float* row = acts + (int64_t(tbase + t0 + f) * st_t + b * st_b) + lane;
#pragma unroll
          for (int k = 0; k < CPL; ++k) x[f][k] = (lane + 32 * k < C) ? __ldcs(row + 32 * k) : 0.f;
        }
        const float* Pt = P + t0 * Sstride;
        float bs[F];
#pragma unroll
        for (int f = 0; f < F; ++f) bs[f] = 0.f;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          const int m = lane + 32 * i;
          const int ca = colA[m], cb = colB[m], gs = gslot[m];
          const int ce = m <= L ? 2 * m : NEGCOL;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float* Pf = Pt + f * Sstride;
            Gw[gs * F + f] = (ex2f(Pf[ca] - ll2) + ex2f(Pf[cb] - ll2)) * scale;
            bs[f] += ex2f(Pf[ce] - ll2);
          }
        }
#pragma unroll
        for (int f = 0; f < F; ++f) bs[f] = warp_sum(bs[f]);
        if (lane == 0) {
#pragma unroll
          for (int f = 0; f < F; ++f) Gw[L * F + f] = bs[f] * scale;
        }
        __syncwarp();
        if (nex > 0) {
          if (lane < F) {
            const float* Pf = Pt + lane * Sstride;
            for (int e = 0; e < nex; ++e) Gw[int(exslot[e]) * F + lane] += ex2f(Pf[excol[e]] - ll2) * scale;
          }
          __syncwarp();
        }
        float z2[F];
#pragma unroll
        for (int f = 0; f < F; ++f) z2[f] = logZ2[t0 + f];
        float* grow = grad + (int64_t(tbase + t0) * st_t + b * st_b) + lane;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          if (lane + 32 * k < C) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const float pr = ex2f(fmaf(x[f][k], LOG2E, -z2[f]));
              __stcs(grow + f * st_t + 32 * k, fmaf(pr, scale, -gk[k][f]));
            }
          }
        }
        __syncwarp();
      } else {
        // odd last frame of a chunk / targets with very many repeats: one frame at a time, classes matched by search
#pragma unroll 1
        for (int f = 0; f < F; ++f) {
          const int tl = t0 + f;
          if (tl >= Tn) break;
          const float* row = acts + (int64_t(tbase + tl) * st_t + b * st_b);
          float* grow = grad + (int64_t(tbase + tl) * st_t + b * st_b);
          const float* Pt = P + tl * Sstride;
          float bsum = 0.f;
          for (int m = lane; m <= L; m += 32) bsum += ex2f(Pt[2 * m] - ll2);
          for (int j = lane; j < L; j += 32) if (int(cmap[tg[j]]) == L) bsum += ex2f(Pt[2 * j + 1] - ll2);
          bsum = warp_sum(bsum);
          const float z2 = logZ2[tl];
          for (int cc = lane; cc < C; cc += 32) {
            const float pr = ex2f(row[cc] * LOG2E - z2);
            const int u = int(cmap[cc]);
            float occ = (u == L) ? bsum : 0.f;
            if (u < L) for (int j = u; j < L; ++j) if (tg[j] == cc) occ += ex2f(Pt[2 * j + 1] - ll2);
            grow[cc] = (pr - occ) * scale;
          }
        }
      }
    }
    __syncthreads();                          // the next chunk's emissions overwrite G and the table
  }

  if (tid == 0) {
    float nll = feasible ? -ll2 * LN2 : INFINITY;
    if (!feasible && zero_infinity) nll = 0.f;
    nll_out[b] = nll;
    if (loss_out != nullptr) atomicAdd(loss_out, nll / float(max(L, 1)) / float(B));
  }
  if (grad == nullptr) return;
  // padded frames get zeros; an infeasible utterance zeros (zero_infinity) or NaN (as ATen) on its live frames
  for (int t = (feasible ? Tb : 0) + warp; t < T; t += NW) {
    float* grow = grad + (int64_t(t) * st_t + b * st_b);
    const float fill = (!feasible && !zero_infinity && t < Tb) ? NAN : 0.f;
    for (int cc = lane; cc < C; cc += 32) grow[cc] = fill;
  }
}

template <int SPL>
int launch_long(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets,
                const int64_t* tgt_offsets, const int64_t* in_lens, const int64_t* tgt_lens, int Lmax, int blank,
                int zero_infinity, float grad_scale, int Tc, float* nll, float* loss, float* grad, size_t smem, cudaStream_t st) {
  auto kern = ctc3l_kernel<SPL, 2>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr_smem = smem;
  }
  kern<<<B, THREADS, smem, st>>>(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens, Lmax, blank,
                                 zero_infinity, grad_scale, Tc, nll, loss, grad);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

// chunk length: the largest even Tc <= 128 that leaves room for two CTAs per SM, else for one
static int pick_chunk(int T, int Lmax, int C, int spld, size_t* smem_out) {
  for (size_t budget : {size_t(113) * 1024, size_t(227) * 1024}) {
    for (int Tc = 128; Tc >= 16; Tc -= 2) {
      const size_t sm = smem_bytes_long(Tc, T, Lmax, C, spld);
      if (sm <= budget) { *smem_out = sm; return Tc; }
    }
  }
  return 0;
}

static int dispatch(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets,
                    const int64_t* tgt_offsets, const int64_t* in_lens, const int64_t* tgt_lens, int Lmax, int blank,
                    int zero_infinity, float grad_scale, float* nll, float* loss, float* grad, long long* dbg, cudaStream_t st) {
  const int S = 2 * Lmax + 1;
  const int spl = (S + 31) / 32;
  if (spl > 12 || Lmax > 30000 || C > 32000) return CTC3_NOT_APPLICABLE;
  const int spld = spl <= 4 ? spl : (spl <= 6 ? 6 : (spl <= 8 ? 8 : 12));
  const size_t smem_p = smem_bytes_pipe(T, Lmax, C, spld);
  if (smem_p > 227 * 1024) {
    // the table of an utterance does not fit shared memory: frame-chunked kernel (recomputation, no table in HBM)
    size_t smem_l = 0;
    const int Tc = (C >= 32 && C <= 32 * CPL && T >= 32) ? pick_chunk(T, Lmax, C, spld, &smem_l) : 0;
    if (Tc == 0) return CTC3_NOT_APPLICABLE;                     // ctc.cu: tables in a global workspace
#define CTC3L_CASE(N) return launch_long<N>(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens, Lmax, blank, \
                                            zero_infinity, grad_scale, Tc, nll, loss, grad, smem_l, st)
    if (spl <= 1) CTC3L_CASE(1);
    if (spl <= 2) CTC3L_CASE(2);
    if (spl <= 3) CTC3L_CASE(3);
    if (spl <= 4) CTC3L_CASE(4);
    if (spl <= 6) CTC3L_CASE(6);
    if (spl <= 8) CTC3L_CASE(8);
    CTC3L_CASE(12);
#undef CTC3L_CASE
  }
#define CTC3P_CASE(N) return launch_pipe<N, 3>(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens, Lmax, blank, \
                                               zero_infinity, grad_scale, nll, loss, grad, dbg, smem_p, st)
  if (spl <= 1) CTC3P_CASE(1);
  if (spl <= 2) CTC3P_CASE(2);
  if (spl <= 3) CTC3P_CASE(3);
  if (spl <= 4) CTC3P_CASE(4);
  if (spl <= 6) CTC3P_CASE(6);
  if (spl <= 8) CTC3P_CASE(8);
  CTC3P_CASE(12);
#undef CTC3P_CASE
}

}  // namespace

int ctc3_try(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int is_logprob, const int64_t* targets, const int64_t* tgt_offsets,
             const int64_t* in_lens, const int64_t* tgt_lens, int Lmax, int blank, int zero_infinity, float grad_scale,
             float* nll, float* loss, float* grad, long long* dbg, cudaStream_t st) {
  return dispatch(acts, T, B, C, st_t, st_b, is_logprob, targets, tgt_offsets, in_lens, tgt_lens, Lmax, blank, zero_infinity,
                  grad_scale, nll, loss, grad, dbg, st);
}

}  // namespace masr
