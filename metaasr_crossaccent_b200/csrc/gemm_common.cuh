// Shared between the tcgen05 GEMM kernels (gemm_umma.cu: one CTA per 128 x BN tile; gemm_pair_umma.cu: persistent
// CTA pairs, cta_group::2, 256 x BN tiles): the kernel parameter block and the per-tile epilogue set-up that turns it
// into epilogue.cuh's EpiOpts.
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "epilogue.cuh"
#include "epilogue_tma.cuh"

namespace masr {

constexpr int UG_BM = 128;         // rows of C per CTA (UMMA M per CTA)
constexpr int UG_BK = 64;          // k-block: 64 bf16 = one 128 B swizzle row

struct UmmaGemmParams {
  int M, N, K;                     // C[M,N] = sum_k A(m,k) B(n,k)
  void* C; int64_t ldc; int c_is_f32;
  const float* bias;
  int flags;
  int kb_per_split;                // k-blocks per split-K slice (split-K: fp32 reductions into C)
  int stages;                      // depth of the TMA->MMA ring
  // fused extras (masr_gemm_epilogue)
  float* rowsum;                   // rowsum[m] += sum_k A(m,k): a second, 16-column accumulator fed by an all-ones B tile
  const __nv_bfloat16* mask; int64_t ldmask; float mask_scale;
  float p_drop, inv_keep; uint64_t seed; const uint64_t* seed_ptr; uint32_t site;
  const __nv_bfloat16* dot_src; int64_t lddot; float* dot_out; int dot_L, dot_H;   // per-head row dots (see EpiOpts)
};

// One epilogue thread's 128-row x BN-column piece of a tile: TMEM accumulator at `tmem_acc` (column base of the piece),
// output row m (= m0 + TMEM lane of this thread), columns n0 .. n0 + BN.  `use_bias`: sbias holds the BN bias values
// of these columns.  split_first: this is k-slice 0 of a split-K problem (or not split at all).
template <int BN>
__device__ __forceinline__ void gemm_epilogue_piece(const UmmaGemmParams& p, uint32_t tmem_acc, int q, int lane, int m, int n0,
                                                    unsigned char* stage, const float* sbias, bool use_bias) {
  const bool relu = p.flags & MASR_GEMM_RELU, accum = p.flags & MASR_GEMM_ACCUM, splitk = p.flags & MASR_GEMM_SPLITK;
  const int ncols = min(BN, p.N - n0);
  if (ncols <= 0) return;
  const int mode = splitk ? EPI_ATOMIC : (accum ? EPI_ACCUM : EPI_STORE);
  EpiOpts o;
  o.sbias = use_bias ? sbias : nullptr;
  o.relu = relu;
  if (p.p_drop > 0.f) {
    o.p_drop = p.p_drop; o.inv_keep = p.inv_keep; o.site = p.site;
    o.seed = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
    o.drop_row_base = int64_t(m) * p.N + n0;
  }
  if (p.dot_src != nullptr && m < p.M) {   // row m = b * L + q; columns n0.. = heads n0 / 64..
    o.dot_row = p.dot_src + int64_t(m) * p.lddot + n0;
    o.dot_out = p.dot_out + (int64_t(m / p.dot_L) * p.dot_H + (n0 >> 6)) * p.dot_L + (m % p.dot_L);
    o.dot_stride = p.dot_L;
  }
  o.mask_scale = p.mask_scale;      // warp-uniform: a lane drains OTHER rows' chunks in phase 2
  if (p.mask != nullptr && m < p.M) o.mask_row = p.mask + int64_t(m) * p.ldmask + n0;
  if (p.c_is_f32) {
    float* row = (m < p.M) ? static_cast<float*>(p.C) + int64_t(m) * p.ldc + n0 : nullptr;
    const bool vec_ok = ncols == BN && (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) + size_t(n0) * 4) & 15) == 0;
    epilogue_tile<BN, float>(tmem_acc, q, lane, stage, row, ncols, vec_ok, mode, o);
  } else {
    __nv_bfloat16* row = (m < p.M) ? static_cast<__nv_bfloat16*>(p.C) + int64_t(m) * p.ldc + n0 : nullptr;
    const bool vec_ok = ncols == BN && (p.ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.C) + size_t(n0) * 2) & 15) == 0 &&
                        (p.mask == nullptr || ((p.ldmask & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.mask) + size_t(n0) * 2) & 15) == 0));
    epilogue_tile<BN, __nv_bfloat16>(tmem_acc, q, lane, stage, row, ncols, vec_ok, mode, o);
  }
}

// Epilogue flavours (kernel template parameter: each instantiation carries only its own code -- the epilogue of the
// K = 512 problems is instruction-cache sensitive)
enum GemmEpi {
  GEPI_LEGACY = 0,      // epilogue.cuh: any alignment / dtype / mode (staged, per-thread global stores)
  GEPI_TMA_BF16 = 1,    // bf16 C through bulk tensor stores; bias / ReLU only
  GEPI_TMA_BF16_X = 2,  // + dropout / backward mask / row dots / += old C
  GEPI_TMA_F32 = 3      // fp32 C: bulk store, or bulk reduce-add for split-K (bias on the first slice)
};

// One warp's 32 rows x `width` columns of a tile through the TMA epilogue.  m_warp0: global row of the warp's lane 0.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_tma_piece(const UmmaGemmParams& p, const CUtensorMap* map_c, uint32_t tmem_acc, int q,
                                                        int lane, int m_warp0, int n0, int width, unsigned char* wstage,
                                                        const float* sbias, bool use_bias, int& boxsel) {
  const int ncols = min(width, p.N - n0);
  if (ncols <= 0) return;
  const int m = m_warp0 + lane;
  EptOpts o;
  o.sbias = use_bias ? sbias : nullptr;
  o.relu = p.flags & MASR_GEMM_RELU;
  if constexpr (EPI == GEPI_TMA_F32) {
    if (p.flags & (MASR_GEMM_SPLITK | MASR_GEMM_ACCUM)) epilogue_tma_f32<true>(map_c, tmem_acc, q, lane, m_warp0, n0, width / 32, ncols, wstage, boxsel, o);
    else epilogue_tma_f32<false>(map_c, tmem_acc, q, lane, m_warp0, n0, width / 32, ncols, wstage, boxsel, o);
  } else {
    const bool row_valid = m < p.M;
    if constexpr (EPI == GEPI_TMA_BF16_X) {
      if (p.p_drop > 0.f) {
        o.p_drop = p.p_drop; o.inv_keep = p.inv_keep; o.site = p.site;
        o.seed = p.seed + (p.seed_ptr != nullptr ? *p.seed_ptr : 0ull);
        o.drop_row_base = int64_t(m) * p.N + n0;
      }
      if (p.dot_src != nullptr && row_valid) {
        o.dot_row = p.dot_src + int64_t(m) * p.lddot + n0;
        o.dot_out = p.dot_out + (int64_t(m / p.dot_L) * p.dot_H + (n0 >> 6)) * p.dot_L + (m % p.dot_L);
        o.dot_stride = p.dot_L;
      }
      o.mask_scale = p.mask_scale;
      if (p.mask != nullptr && row_valid) o.mask_row = p.mask + int64_t(m) * p.ldmask + n0;
      if ((p.flags & MASR_GEMM_ACCUM) && row_valid) o.old_row = static_cast<const __nv_bfloat16*>(p.C) + int64_t(m) * p.ldc + n0;
    }
    epilogue_tma_bf16<EPI == GEPI_TMA_BF16_X>(map_c, tmem_acc, q, lane, m_warp0, n0, width / 64, ncols, row_valid, wstage, boxsel, o);
  }
}

// host: which epilogue flavour serves this problem (GEPI_LEGACY when C cannot be described by a tensor map or a side
// input is not 16-byte addressable)
int gemm_pick_epilogue(const UmmaGemmParams& p);
// host: tensor map of C for the TMA epilogues: box = 32 rows x 128 bytes, 128B swizzle
int gemm_c_map(CUtensorMap* out, const UmmaGemmParams& p);

// host: operand tensor map.  K-major -> dims {K, rows}, box {64, tile_rows};  MN-major -> dims {rows(MN), K}, box {64, 64}
int gemm_operand_map(CUtensorMap* out, const void* base, int64_t ld_elems, int rows_mn, int K, bool mn_major, int tile_rows);

// host: the CTA-pair kernel (gemm_pair_umma.cu).  Returns MASR_OK, or < 0 on error.
int launch_umma_pair(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, UmmaGemmParams p,
                     int splitk, int force_bn, cudaStream_t st);
// host: true when the pair kernel is the better choice for this problem (large M, enough tiles)
bool umma_pair_preferred(int M, int N, int K, int flags);
void set_pair_mode(int mode);

}  // namespace masr
