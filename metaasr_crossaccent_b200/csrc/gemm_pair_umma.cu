// Kernel family 2, large shapes: PERSISTENT CTA-PAIR GEMM (tcgen05.mma.cta_group::2).
//
// Why (profiles/r1_full_gemm_traffic.md): the 128 x 128 one-tile-per-CTA kernel of gemm_umma.cu runs the
// encoder-sized problems (M = 4096, K = 512) at 0.29 of the tensor peak: every CTA pays barrier init + TMEM allocation +
// pipeline fill + a 128 x 128 epilogue for eight k-blocks of MMA work, and it pulls 1 byte of operands per 64 flop through
// L2 (134 MB for 6.4 MB of DRAM reads).  Here
//   * two CTAs of a cluster (one per SM of a TPC) compute ONE 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages
//     its own 128 rows of A and only HALF of B (BN/2 rows), the tensor core reads the other half from the peer's shared
//     memory -- 256 x 256 tiles move half the operand bytes per flop of 128 x 128 ones;
//   * the kernel is persistent (74 pairs), tiles are striped over the pairs; the TMA->MMA ring runs across tile borders
//     and the accumulator is DOUBLE-BUFFERED in TMEM (2 x 256 columns), so the epilogue of tile i (TMEM -> registers ->
//     staging -> coalesced stores) overlaps the MMAs of tile i+1; prologue (barriers, TMEM allocation, tensor-map prefetch)
//     is paid once per SM instead of once per 128 x 128 tile.
// Warp roles per CTA (192 threads): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only; allocates TMEM in
// both), warps 2-5 epilogue (both CTAs, each drains its own 128 accumulator rows).
// Barrier protocol (CTA pair):
//   full[s]       leader CTA only; 1 arrival (leader's arrive.expect_tx of BOTH CTAs' bytes) + complete_tx of both CTAs' TMA
//                 loads (the peer's loads signal the leader's barrier: cp.async.bulk.tensor ... cta_group::2)
//   empty[s]      in each CTA; 1 arrival = tcgen05.commit.multicast of the leader after the MMAs of that stage
//   tmem_full[a]  in each CTA; 1 arrival = commit.multicast after the last MMA of a tile
//   tmem_empty[a] leader CTA only; 8 arrivals = the 4 epilogue warps of both CTAs (the peer arrives remotely)
#include <type_traits>
#include "gemm_common.cuh"

namespace masr {
namespace umma {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cluster address of `smem_addr` (a shared::cta address of THIS CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// TMA load of a CTA pair: destination = this CTA's shared memory, completion bytes go to `bar_cluster_addr`
// (a shared::cluster address: the LEADER's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows per CTA), B rows split between the CTAs
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// Remote arrive WITHOUT a memory fence: what is handed over is a TMEM accumulator stage, ordered by the tcgen05 fences.
// (.release.cluster compiles to MEMBAR.ALL.CTA + ERRBAR, which waits for all of the warp's outstanding global stores:
// a third of the epilogue warps' time in the first version of this kernel.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // barrier with remote arrivals
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace umma

constexpr int UP_THREADS = 192;
constexpr int UP_ACC_STRIDE = 256;          // TMEM columns per accumulator stage (2 stages = the whole TMEM)
constexpr int UP_MAX_STAGES = 8;

struct PairSched {
  int tiles_m;        // 256-row tiles
  int tiles_n;        // BN-column tiles
  int nsplit;         // split-K slices
  int total;          // tiles_m * tiles_n * nsplit work items
  int staging_bytes;  // epilogue staging tile (fp32 or bf16 output rows)
};

template <int BN>
struct PairSmem {
  static constexpr uint32_t A_BYTES = UG_BM * UG_BK * 2;              // 16 KB: this CTA's 128 rows of A
  static constexpr uint32_t B_BYTES = (BN / 2) * UG_BK * 2;           // this CTA's half of B
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // dedicated epilogue staging tile (the TMA ring keeps running under the epilogue); both sizes are multiples of 1024
  static constexpr uint32_t staging(bool f32) { return f32 ? EpiLayout<128, float>::BYTES : EpiLayout<128, __nv_bfloat16>::BYTES; }
  static constexpr uint32_t TAIL = 2048 /*ones*/ + 1024 /*barriers, tmem slot*/ + BN * 4 /*bias*/;
  static constexpr size_t bytes(int stages, bool f32) { return size_t(stages) * STAGE_BYTES + staging(f32) + TAIL + 1024 /*alignment slack*/; }
  static_assert(EpiLayout<128, float>::BYTES % 1024 == 0 && EpiLayout<128, __nv_bfloat16>::BYTES % 1024 == 0, "ones tile alignment");
};

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UP_THREADS, 1)
umma_pair_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_c, UmmaGemmParams p, PairSched sch) {
  using namespace umma;
  using SM = PairSmem<BN>;
  constexpr uint32_t A_BYTES = SM::A_BYTES, B_BYTES = SM::B_BYTES, STAGE_BYTES = SM::STAGE_BYTES;
  static_assert(BN == 128 || BN == 256, "pair tile: 256 x 128 or 256 x 256");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  unsigned char* staging = smem + STAGES * STAGE_BYTES;                  // 16 B aligned (STAGE_BYTES multiple of 1024)
  unsigned char* sones = staging + sch.staging_bytes;                    // 2 KB of bf16 1.0 (row-sum B operand), 1024 B aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sones + 2048);
  uint64_t* empty_bar = full_bar + UP_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + UP_MAX_STAGES;                   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* sbias = reinterpret_cast<float*>(sones + 2048 + 1024);          // BN floats
  pdl_launch_dependents();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (p.K + UG_BK - 1) / UG_BK;
  const bool want_rowsum = p.rowsum != nullptr;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (EPI != GEPI_LEGACY) prefetch_tmap(&map_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
  if (want_rowsum && threadIdx.x >= 64) {
    reinterpret_cast<uint4*>(sones)[threadIdx.x - 64] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
  }
  // both CTAs' barriers are initialised and both TMEM allocations are done before anyone signals the peer
  tc_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // work item w -> (m tile, n tile, k slice); consecutive pairs share the B tile (weights) in L2
  auto decode = [&](int w, int& m0, int& n0, int& kb_begin, int& num_kb) {
    const int mt = w % sch.tiles_m;
    const int r = w / sch.tiles_m;
    const int nt = r % sch.tiles_n;
    const int ks = r / sch.tiles_n;
    m0 = mt * 256;
    n0 = nt * BN;
    kb_begin = ks * p.kb_per_split;
    num_kb = min(total_kb - kb_begin, p.kb_per_split);
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own half of B; bytes are counted on the LEADER's full barrier
    int s = 0; uint32_t ph = 0;
    for (int w = pair; w < sch.total; w += npairs) {
      int m0, n0, kb_begin, num_kb;
      decode(w, m0, n0, kb_begin, num_kb);
      const int ma = m0 + int(rank) * 128;                 // this CTA's A rows
      const int nb = n0 + int(rank) * (BN / 2);            // this CTA's B rows
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        const int k0 = (kb_begin + kb) * UG_BK;
        const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d_pair(sa, &map_a, fb, k0, ma);                          // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_pair(sa + c * (64 * UG_BK * 2), &map_a, fb, ma + c * 64, k0);   // box {64 m, 64 k}
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &map_b, fb, k0, nb);                          // box {64 k, BN/2 n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c) tma_load_2d_pair(sb + c * (64 * UG_BK * 2), &map_b, fb, nb + c * 64, k0);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer (leader CTA; converged warp, elected lane issues) =====
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_ones = make_idesc_bf16(256, 16, A_MN ? 1 : 0, 0);
      const uint64_t dones = desc_kmajor_sw128(smem_u32(sones));
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int w = pair; w < sch.total; w += npairs) {
        int m0, n0, kb_begin, num_kb;
        decode(w, m0, n0, kb_begin, num_kb);
        mbar_wait_cluster(&tmem_empty_bar[as], aph ^ 1);       // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t acc_addr = tmem_base + uint32_t(as * UP_ACC_STRIDE);
        auto kloop = [&](auto with_rowsum) {
          constexpr bool RS = decltype(with_rowsum)::value;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint32_t sb = sa + A_BYTES;
            const uint64_t da0 = A_MN ? desc_mnmajor_sw128(sa, 64 * UG_BK * 2) : desc_kmajor_sw128(sa);
            const uint64_t db0 = B_MN ? desc_mnmajor_sw128(sb, 64 * UG_BK * 2) : desc_kmajor_sw128(sb);
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < UG_BK / 16; ++k) {
                const uint64_t da = da0 + uint64_t(A_MN ? k * 128 : k * 2);
                const uint64_t db = db0 + uint64_t(B_MN ? k * 128 : k * 2);
                const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                mma_f16_ss_pair(acc_addr, da, db, idesc, acc);
                if constexpr (RS) mma_f16_ss_pair(acc_addr + BN, da, dones, idesc_ones, acc);
              }
              mma_commit_pair(&empty_bar[s]);            // frees this stage in BOTH CTAs once the MMAs have read it
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        };
        if (want_rowsum && n0 == 0) kloop(std::true_type{}); else kloop(std::false_type{});
        if (elect_one_sync()) mma_commit_pair(&tmem_full_bar[as]);     // accumulator stage complete (both CTAs)
        __syncwarp();
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 2..5 of both CTAs): TMEM lane quarter = warp % 4, own 128 rows of the 256-row tile =====
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const bool splitk = p.flags & MASR_GEMM_SPLITK;
    const uint32_t te_leader = mapa_shared(smem_u32(&tmem_empty_bar[0]), 0);
    int as = 0; uint32_t aph = 0;
    int boxsel = 0;
    for (int w = pair; w < sch.total; w += npairs) {
      int m0, n0, kb_begin, num_kb;
      decode(w, m0, n0, kb_begin, num_kb);
      const bool use_bias = p.bias != nullptr && (!splitk || kb_begin == 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous tile's readers of sbias are done
      if (use_bias) {
        for (int i = et; i < BN; i += 128) sbias[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");          // bias tile visible to the 4 epilogue warps
      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const uint32_t acc_addr = tmem_base + uint32_t(as * UP_ACC_STRIDE);
      const int m = m0 + int(rank) * 128 + q * 32 + lane;
      if constexpr (EPI == GEPI_LEGACY) {
#pragma unroll 1
        for (int c = 0; c < BN / 128; ++c)
          gemm_epilogue_piece<128>(p, acc_addr + uint32_t(c * 128), q, lane, m, n0 + c * 128, staging, sbias + c * 128, use_bias);
      } else {
        gemm_epilogue_tma_piece<EPI>(p, &map_c, acc_addr, q, lane, m0 + int(rank) * 128 + q * 32, n0, BN,
                                     staging + q * EPT_WARP_BYTES, sbias, use_bias, boxsel);
      }
      if (want_rowsum && n0 == 0) {                            // column 0 of the ones-accumulator = sum_k A(m, k)
        float v[32];
        tmem_ld_32x32(acc_addr + BN + (uint32_t(q * 32) << 16), v);
        tmem_ld_wait();
        if (m < p.M) atomicAdd(p.rowsum + m, v[0]);
      }
      // this warp has read its accumulator rows into registers / staging: hand the TMEM stage back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te_leader + uint32_t(as * 8));
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if constexpr (EPI != GEPI_LEGACY) epilogue_tma_drain(lane);      // bulk stores have left shared memory and landed
  }
  // teardown: nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory / signal its barriers
  tc_fence_before();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, UmmaGemmParams p, PairSched sch,
                       cudaStream_t st) {
  using SM = PairSmem<BN>;
  auto kern = umma_pair_gemm_kernel<BN, A_MN, B_MN, EPI>;
  const uint32_t staging = EPI == GEPI_LEGACY ? SM::staging(p.c_is_f32 != 0) : uint32_t(EPT_CTA_BYTES);
  int stages = int((226 * 1024 - staging - SM::TAIL - 1024) / SM::STAGE_BYTES);
  stages = std::max(2, std::min(stages, UP_MAX_STAGES));
  static bool attr_set = false;
  if (!attr_set) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  p.stages = stages;
  sch.staging_bytes = int(staging);
  const int pairs = std::max(1, std::min(sch.total, sm_count() / 2));
  const size_t smem = size_t(stages) * SM::STAGE_BYTES + staging + SM::TAIL + 1024;
  MASR_CHECK_CUDA(launch_pdl(kern, dim3(unsigned(2 * pairs)), dim3(UP_THREADS), smem, st, ma, mb, mc, p, sch));
  return MASR_OK;
}

template <int BN>
static int launch_pair_bn(int a_mn, int b_mn, int epi, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                          const UmmaGemmParams& p, const PairSched& sch, cudaStream_t st) {
  // forward (K-major x K-major) and dgrad (K-major x MN-major) produce bf16 activations / gradients; wgrad (MN x MN)
  // produces fp32 split-K sums
  if (!a_mn && !b_mn) {
    if (epi == GEPI_TMA_BF16) return launch_pair<BN, false, false, GEPI_TMA_BF16>(ma, mb, mc, p, sch, st);
    if (epi == GEPI_TMA_BF16_X) return launch_pair<BN, false, false, GEPI_TMA_BF16_X>(ma, mb, mc, p, sch, st);
    if (epi == GEPI_TMA_F32) return launch_pair<BN, false, false, GEPI_TMA_F32>(ma, mb, mc, p, sch, st);
    return launch_pair<BN, false, false, GEPI_LEGACY>(ma, mb, mc, p, sch, st);
  }
  if (!a_mn && b_mn) {
    if (epi == GEPI_TMA_BF16) return launch_pair<BN, false, true, GEPI_TMA_BF16>(ma, mb, mc, p, sch, st);
    if (epi == GEPI_TMA_BF16_X) return launch_pair<BN, false, true, GEPI_TMA_BF16_X>(ma, mb, mc, p, sch, st);
    return launch_pair<BN, false, true, GEPI_LEGACY>(ma, mb, mc, p, sch, st);
  }
  if (a_mn && b_mn) {
    if (epi == GEPI_TMA_F32) return launch_pair<BN, true, true, GEPI_TMA_F32>(ma, mb, mc, p, sch, st);
    return launch_pair<BN, true, true, GEPI_LEGACY>(ma, mb, mc, p, sch, st);
  }
  return launch_pair<BN, true, false, GEPI_LEGACY>(ma, mb, mc, p, sch, st);
}

// 256 x 256 tiles halve the operand bytes per flop but leave pairs idle unless the problem has many of them; measured
// in a replayed graph (profiles/r2_gemm_probe.md): N = 2048 9.4 vs 9.8 us, N = 1536 8.8 vs 8.5 us, N <= 1024 worse.
static int pair_pick_bn(int M, int N, int nsplit) {
  (void)M; (void)nsplit;
  return (N >= 2048 && N % 256 == 0) ? 256 : 128;
}

static int g_pair_mode = 1;            // 0 = never, 1 = by problem size (default)
int g_force_legacy_epilogue = 0;       // mode bit 2: staged epilogue everywhere (A/B measurements)
void set_pair_mode(int mode) { g_pair_mode = mode & 1; g_force_legacy_epilogue = (mode >> 1) & 1; }

bool umma_pair_preferred(int M, int N, int K, int flags) {
  if (g_pair_mode == 0) return false;
  // measured crossover (replayed graph, B200): wide outputs of the M = B T/4 = 4096-row problems and the long-reduction
  // weight gradients win 20-35 %; N <= 1024 forward / dgrad problems and everything decoder-sized (M = B (L+1) ~ 1 k rows)
  // are as fast or faster on the one-CTA-per-tile kernel
  if (flags & MASR_GEMM_SPLITK) return K >= 2048 && int64_t(M) * N >= int64_t(512) * 512 && M >= 256 && N >= 256;
  return M >= 2048 && N >= 1536 && K >= 256;
}

int launch_umma_pair(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, UmmaGemmParams p,
                     int splitk, int force_bn, cudaStream_t st) {
  const int total_kb = (p.K + UG_BK - 1) / UG_BK;
  int BN = force_bn > 0 ? force_bn : 0;
  if (p.rowsum != nullptr) BN = 128;                      // the row-sum accumulator lives behind a 128-column tile
  int kb_per_split = total_kb;
  if (p.flags & MASR_GEMM_SPLITK) {
    // weight gradients (long reductions, few output tiles): slice the reduction until the 74 pairs have one or two
    // work items each, at least 4 k-blocks per slice.  splitk < 0 leaves the choice to this function.
    int want = splitk;
    if (splitk <= 0) {
      const int64_t tiles = ceil_div64(p.M, 256) * ceil_div64(p.N, BN > 0 ? BN : 256);
      want = int(std::max<int64_t>(1, std::min<int64_t>(total_kb / 4, (2 * (sm_count() / 2)) / std::max<int64_t>(tiles, 1))));
    }
    if (want > 1) kb_per_split = (total_kb + want - 1) / want;
  }
  p.kb_per_split = kb_per_split;
  const int nsplit = int(ceil_div64(total_kb, kb_per_split));
  if (BN == 0) BN = pair_pick_bn(p.M, p.N, nsplit);
  MASR_REQUIRE(BN == 128 || BN == 256, "pair gemm: BN must be 128 or 256");
  PairSched sch;
  sch.tiles_m = int(ceil_div64(p.M, 256));
  sch.tiles_n = int(ceil_div64(p.N, BN));
  sch.nsplit = nsplit;
  sch.total = sch.tiles_m * sch.tiles_n * sch.nsplit;
  CUtensorMap ma, mb;
  int rc = gemm_operand_map(&ma, A, lda, p.M, p.K, a_mn != 0, UG_BM);
  if (rc != MASR_OK) return rc;
  rc = gemm_operand_map(&mb, B, ldb, p.N, p.K, b_mn != 0, BN / 2);
  if (rc != MASR_OK) return rc;
  CUtensorMap mc = ma;
  int epi = g_force_legacy_epilogue ? GEPI_LEGACY : gemm_pick_epilogue(p);
  if (epi != GEPI_LEGACY) {
    rc = gemm_c_map(&mc, p);
    if (rc != MASR_OK) return rc;
  }
  return BN == 128 ? launch_pair_bn<128>(a_mn, b_mn, epi, ma, mb, mc, p, sch, st)
                   : launch_pair_bn<256>(a_mn, b_mn, epi, ma, mb, mc, p, sch, st);
}

}  // namespace masr
