// Kernel family 2: bf16 GEMMs on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in
// TMEM), operands staged by TMA into 128B-swizzled shared memory through an mbarrier ring.
//
// One CTA computes a 128 x BN tile of C.  Warp roles (192 threads):
//   warp 0    TMA producer   : one elected lane issues cp.async.bulk.tensor loads, STAGES deep
//   warp 1    MMA issuer     : allocates TMEM, one lane issues tcgen05.mma (4 x K=16 per 64-wide k-block),
//                              tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2-5 epilogue       : tcgen05.ld 32 lanes x 32 columns -> registers -> bias / ReLU / accumulate ->
//                              global (each thread owns one output row: 64-128 B contiguous per store burst)
// Operand forms (template): K-major ("TN" forward GEMM: x[M,K] w[N,K]) or MN-major (dgrad: w[N,K] read as
// B[K_out, N_red]; wgrad: dy[M,N], x[M,K] reduced over M) -- the latter only changes the TMA box, the smem
// descriptor (LBO/SBO) and two bits of the instruction descriptor, the data are never transposed in memory.
#include <type_traits>
#include "gemm_common.cuh"
#include <vector>

namespace masr {

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle128);
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MASR_E_CUDA; }
  cuuint64_t gdim[5]; cuuint64_t gstr[5]; cuuint32_t bx[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];       // stride of dim i+1
  CUresult r = enc(out, dt, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu] stride %llu box [%u,%u]", int(r), rank,
             (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    set_error(buf);
    return MASR_E_CUDA;
  }
  return MASR_OK;
}

// ------------------------------------------------------------------ kernel
constexpr int UG_THREADS = 192;

// One CTA of the one-tile GEMM: tile (bx, by) of k-slice bz of the problem described by (maps, p).  Shared by the
// plain kernel (one problem per launch) and the grouped kernel (many small problems -- the decoder's weight gradients --
// in one launch).
template <int BN, int MIN_STAGES, bool A_MN, bool B_MN, int EPI>
__device__ __forceinline__ void umma_gemm_cta(const CUtensorMap* pmap_a, const CUtensorMap* pmap_b, const CUtensorMap* pmap_c,
                                              const UmmaGemmParams& p, const int bx, const int by, const int bz) {
  using namespace umma;
  constexpr uint32_t A_BYTES = UG_BM * UG_BK * 2;          // 16 KB
  constexpr uint32_t B_BYTES = BN * UG_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  // 1024 B alignment is required by the 128 B swizzle atom
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;                                          // >= MIN_STAGES
  unsigned char* sones = smem + STAGES * STAGE_BYTES;                   // 2 KB: [16 x 64] bf16 tile of 1.0 (row sums)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sones + 2048);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* sbias = reinterpret_cast<float*>(tmem_full_bar + 2);          // BN floats, 16 B aligned
  static_assert(MIN_STAGES * STAGE_BYTES >= EpiLayout<BN, float>::BYTES, "staging tile must fit in the pipeline stages");
  static_assert(MIN_STAGES * STAGE_BYTES >= EPT_CTA_BYTES, "TMA epilogue boxes must fit in the pipeline stages");
  pdl_launch_dependents();          // the next kernel's prologue may overlap this kernel (see common.cuh)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = by * UG_BM, n0 = bx * BN;
  const int total_kb = (p.K + UG_BK - 1) / UG_BK;
  const int kb_begin = bz * p.kb_per_split;
  const int num_kb = min(total_kb - kb_begin, p.kb_per_split);      // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    prefetch_tmap(pmap_a);
    prefetch_tmap(pmap_b);
    if (EPI != GEPI_LEGACY) prefetch_tmap(pmap_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  const bool rowsum = p.rowsum != nullptr && bx == 0;          // the n-tile-0 CTAs also sum A's rows
  const uint32_t tmem_cols = p.rowsum != nullptr ? 2 * BN : BN;
  if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  if (rowsum && threadIdx.x >= 64) {
    reinterpret_cast<uint4*>(sones)[threadIdx.x - 64] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();            // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                       // everything above touched no global memory

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the loop, one elected lane issues (see umma::elect_one_sync) =====
    {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        const int k0 = (kb_begin + kb) * UG_BK;
        if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        if (!A_MN) {
          tma_load_2d(sa, pmap_a, &full_bar[s], k0, m0);               // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int c = 0; c < UG_BM / 64; ++c)                          // box {64 m, 64 k} per 64-wide chunk
            tma_load_2d(sa + c * (64 * UG_BK * 2), pmap_a, &full_bar[s], m0 + c * 64, k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, pmap_b, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sb + c * (64 * UG_BK * 2), pmap_b, &full_bar[s], n0 + c * 64, k0);
        }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane issues) =====
    {
      constexpr uint32_t idesc = make_idesc_bf16(UG_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_ones = make_idesc_bf16(UG_BM, 16, A_MN ? 1 : 0, 0);
      const uint64_t dones = desc_kmajor_sw128(smem_u32(sones));
      // Two copies of the k-loop (with / without the row-sum accumulator) so that neither contains a
      // conditionally issued tcgen05.mma: the loop body is straight-line code for the single issuing thread.
      auto kloop = [&](auto with_rowsum) {
        constexpr bool RS = decltype(with_rowsum)::value;
        int s = 0; uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          // descriptors once per k-block, advanced by encoded offsets (start-address field = bytes >> 4):
          // K-major +16 elements = +32 B -> +2; MN-major +16 k-rows = +2048 B -> +128
          const uint64_t da0 = A_MN ? desc_mnmajor_sw128(sa, 64 * UG_BK * 2) : desc_kmajor_sw128(sa);
          const uint64_t db0 = B_MN ? desc_mnmajor_sw128(sb, 64 * UG_BK * 2) : desc_kmajor_sw128(sb);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < UG_BK / 16; ++k) {
              const uint64_t da = da0 + uint64_t(A_MN ? k * 128 : k * 2);
              const uint64_t db = db0 + uint64_t(B_MN ? k * 128 : k * 2);
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              mma_f16_ss(tmem_base, da, db, idesc, acc);
              if constexpr (RS) mma_f16_ss(tmem_base + BN, da, dones, idesc_ones, acc);
            }
            mma_commit(&empty_bar[s]);           // frees the smem stage once these MMAs have read it
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      };
      if (rowsum) kloop(std::true_type{}); else kloop(std::false_type{});
      if (elect_one_sync()) mma_commit(tmem_full_bar);             // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4; staged through shared memory (epilogue.cuh) =====
    const int q = warp & 3;
    const int et = threadIdx.x - 64;                       // 0..127 among the epilogue threads
    const bool splitk = p.flags & MASR_GEMM_SPLITK;
    const bool use_bias = p.bias != nullptr && (!splitk || bz == 0);
    if (use_bias) {
      for (int i = et; i < BN; i += 128) sbias[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");          // bias tile visible to the 4 epilogue warps
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int m = m0 + q * 32 + lane;
    if constexpr (EPI == GEPI_LEGACY) {
      gemm_epilogue_piece<BN>(p, tmem_base, q, lane, m, n0, smem, sbias, use_bias);
    } else {
      // the accumulator is complete, i.e. every MMA has read its operands: the pipeline stages are free for the boxes
      int boxsel = 0;
      gemm_epilogue_tma_piece<EPI>(p, pmap_c, tmem_base, q, lane, m0 + q * 32, n0, BN, smem + q * EPT_WARP_BYTES, sbias,
                                   use_bias, boxsel);
      epilogue_tma_drain(lane);
    }
    if (rowsum) {                   // column 0 of the ones-accumulator = sum_k A(m, k)
      float v[32];
      tmem_ld_32x32(tmem_base + BN + (uint32_t(q * 32) << 16), v);
      tmem_ld_wait();
      if (m < p.M) atomicAdd(p.rowsum + m, v[0]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}


template <int BN, int MIN_STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(UG_THREADS, 2)
umma_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ UmmaGemmParams p) {
  umma_gemm_cta<BN, MIN_STAGES, A_MN, B_MN, EPI>(&map_a, &map_b, &map_c, p, int(blockIdx.x), int(blockIdx.y), int(blockIdx.z));
}

// ---- grouped launch: up to GG_MAX independent problems of the same kernel instantiation in ONE grid.  The ~26
// weight-gradient GEMMs of the decoder (M, N <= 2048, K = 1 056 rows) are 16-64 CTAs each: launched one by one on the
// side stream they fill a fraction of the machine and pay a launch each; together they are one ~1 500-CTA wave.
constexpr int GG_MAX = 24;
struct alignas(64) GemmGroupEntry {
  CUtensorMap ma, mb, mc;
  UmmaGemmParams p;
  int cta_begin, nx, ny, nz;
};
struct GemmGroupTable {
  GemmGroupEntry e[GG_MAX];
  int n, total;
};

template <int BN, int MIN_STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(UG_THREADS, 2)
umma_gemm_group_kernel(const __grid_constant__ GemmGroupTable t) {
  int i = 0;
  while (i + 1 < t.n && int(blockIdx.x) >= t.e[i + 1].cta_begin) ++i;
  const GemmGroupEntry& e = t.e[i];
  const int local = int(blockIdx.x) - e.cta_begin;
  const int bx = local % e.nx, by = (local / e.nx) % e.ny, bz = local / (e.nx * e.ny);
  umma_gemm_cta<BN, MIN_STAGES, A_MN, B_MN, EPI>(&e.ma, &e.mb, &e.mc, e.p, bx, by, bz);
}

// masr_gemm_set_stage_cap: upper bound of the operand ring depth.  A lone CTA per SM wants the whole shared memory as a
// deep ring; when several task lanes run their small GEMMs concurrently (lock-step meta-step) a 192 KB CTA would keep
// every other lane off its SM -- three stages (96 KB) leave room for a second CTA.
static int g_stage_cap = 0;

// masr_gemm_group_begin / _end: between the two, problems that resolve to the one-tile kernel are recorded (tensor maps,
// parameters, grid) instead of launched; _end launches them grouped by kernel instantiation.
struct GroupRec {
  int (*launch)(const GemmGroupTable&, cudaStream_t);
  GemmGroupEntry e;
};
static thread_local bool g_group_on = false;
static thread_local int g_group_recorded = 0, g_group_launched = 0;      // of the last masr_gemm_group_end
static thread_local std::vector<GroupRec>* g_group = nullptr;

template <int BN, int MIN_STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_group(const GemmGroupTable& t, cudaStream_t st) {
  constexpr size_t STAGE = size_t(UG_BM * UG_BK * 2 + BN * UG_BK * 2);
  auto kern = umma_gemm_group_kernel<BN, MIN_STAGES, A_MN, B_MN, EPI>;
  const size_t smem = size_t(MIN_STAGES) * STAGE + 2048 + 1024 + 512 + BN * 4;
  static bool attr_set = false;
  if (!attr_set) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    attr_set = true;
  }
  MASR_CHECK_CUDA(launch_pdl(kern, dim3(unsigned(t.total)), dim3(UG_THREADS), smem, st, t));
  return MASR_OK;
}

template <int BN, int MIN_STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_umma(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, UmmaGemmParams p, cudaStream_t st) {
  constexpr size_t STAGE = size_t(UG_BM * UG_BK * 2 + BN * UG_BK * 2);
  constexpr int MAX_STAGES = int((198 * 1024) / STAGE);
  if (g_group_on) {                     // recorded: many problems share the launch, every CTA gets the shallow ring
    const int total_kb = int(ceil_div64(p.K, UG_BK));
    GroupRec r;
    r.launch = &launch_group<BN, MIN_STAGES, A_MN, B_MN, EPI>;
    r.e.ma = ma; r.e.mb = mb; r.e.mc = mc;
    p.stages = MIN_STAGES;
    r.e.p = p;
    r.e.nx = int(ceil_div64(p.N, BN)); r.e.ny = int(ceil_div64(p.M, UG_BM)); r.e.nz = int(ceil_div64(total_kb, p.kb_per_split));
    r.e.cta_begin = 0;
    g_group->push_back(r);
    return MASR_OK;
  }
  auto kern = umma_gemm_kernel<BN, MIN_STAGES, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(MAX_STAGES * STAGE + 2048 + 1024 + 512 + BN * 4)));
    attr_set = true;
  }
  const int total_kb = int(ceil_div64(p.K, UG_BK));
  const int nsplit = int(ceil_div64(total_kb, p.kb_per_split));
  dim3 grid(unsigned(ceil_div64(p.N, BN)), unsigned(ceil_div64(p.M, UG_BM)), unsigned(nsplit));
  // one CTA per SM (grid fits in a wave): use the whole shared memory as a deep ring -- a lone CTA has to
  // cover the full L2/HBM latency by itself; otherwise keep two CTAs per SM resident (epilogue of one
  // overlaps the main loop of the other)
  const int64_t ctas = int64_t(grid.x) * grid.y * grid.z;
  int stages = (ctas <= sm_count()) ? MAX_STAGES : MIN_STAGES;
  if (g_stage_cap > 0) stages = std::min(stages, g_stage_cap);
  stages = std::max(MIN_STAGES, std::min(stages, std::min(p.kb_per_split, total_kb)));
  p.stages = stages;
  const size_t smem = size_t(stages) * STAGE + 2048 + 1024 + 512 + BN * 4;
  MASR_CHECK_CUDA(launch_pdl(kern, grid, dim3(UG_THREADS), smem, st, ma, mb, mc, p));
  return MASR_OK;
}

// forward (K-major x K-major) and dgrad (K-major x MN-major): bf16 or fp32 C; wgrad (MN x MN): fp32 split-K sums
template <int BN, int MIN_STAGES>
static int launch_umma_bn(int a_mn, int b_mn, int epi, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                          const UmmaGemmParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) {
    if (epi == GEPI_TMA_BF16) return launch_umma<BN, MIN_STAGES, false, false, GEPI_TMA_BF16>(ma, mb, mc, p, st);
    if (epi == GEPI_TMA_BF16_X) return launch_umma<BN, MIN_STAGES, false, false, GEPI_TMA_BF16_X>(ma, mb, mc, p, st);
    if (epi == GEPI_TMA_F32) return launch_umma<BN, MIN_STAGES, false, false, GEPI_TMA_F32>(ma, mb, mc, p, st);
    return launch_umma<BN, MIN_STAGES, false, false, GEPI_LEGACY>(ma, mb, mc, p, st);
  }
  if (!a_mn && b_mn) {
    if (epi == GEPI_TMA_BF16) return launch_umma<BN, MIN_STAGES, false, true, GEPI_TMA_BF16>(ma, mb, mc, p, st);
    if (epi == GEPI_TMA_BF16_X) return launch_umma<BN, MIN_STAGES, false, true, GEPI_TMA_BF16_X>(ma, mb, mc, p, st);
    return launch_umma<BN, MIN_STAGES, false, true, GEPI_LEGACY>(ma, mb, mc, p, st);
  }
  if (a_mn && b_mn) {
    if (epi == GEPI_TMA_F32) return launch_umma<BN, MIN_STAGES, true, true, GEPI_TMA_F32>(ma, mb, mc, p, st);
    return launch_umma<BN, MIN_STAGES, true, true, GEPI_LEGACY>(ma, mb, mc, p, st);
  }
  return launch_umma<BN, MIN_STAGES, true, false, GEPI_LEGACY>(ma, mb, mc, p, st);
}

// operand map: K-major  -> dims {K, rows}, box {64, tile_rows};  MN-major -> dims {rows(MN), K}, box {64, 64}
int gemm_operand_map(CUtensorMap* out, const void* base, int64_t ld_elems, int rows_mn, int K, bool mn_major, int tile_rows) {
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "umma: operand base must be 16-byte aligned");
  MASR_REQUIRE(ld_elems % 8 == 0, "umma: operand leading dimension must be a multiple of 8 elements");
  uint64_t dims[2]; uint64_t strides[1]; uint32_t box[2];
  if (!mn_major) { dims[0] = uint64_t(K); dims[1] = uint64_t(rows_mn); box[0] = 64; box[1] = uint32_t(tile_rows); }
  else           { dims[0] = uint64_t(rows_mn); dims[1] = uint64_t(K); box[0] = 64; box[1] = 64; }
  strides[0] = uint64_t(ld_elems) * 2;
  return make_tmap_bf16(out, base, 2, dims, strides, box, true);
}

int gemm_pick_epilogue(const UmmaGemmParams& p) {
  const size_t es = p.c_is_f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(p.C) & 15) != 0 || (size_t(p.ldc) * es) % 16 != 0) return GEPI_LEGACY;
  // a bulk tensor store clips rows exactly but writes whole 16-byte units along the row (measured: the element after a
  // row end that is not 16-byte aligned is overwritten with zero): such outputs keep the staged epilogue
  if ((size_t(p.N) * es) % 16 != 0) return GEPI_LEGACY;
  const bool extras = p.p_drop > 0.f || p.mask != nullptr || p.dot_src != nullptr;
  if (p.c_is_f32) return extras ? GEPI_LEGACY : GEPI_TMA_F32;          // store, or reduce-add for split-K / accumulate
  const bool accum = p.flags & MASR_GEMM_ACCUM;
  if (!extras && !accum) return GEPI_TMA_BF16;
  if (p.N % 8 != 0) return GEPI_LEGACY;                                // side inputs are read as 8-element vectors
  if (p.mask != nullptr && ((reinterpret_cast<uintptr_t>(p.mask) & 15) != 0 || p.ldmask % 8 != 0)) return GEPI_LEGACY;
  if (p.dot_src != nullptr && ((reinterpret_cast<uintptr_t>(p.dot_src) & 15) != 0 || p.lddot % 8 != 0)) return GEPI_LEGACY;
  if (accum && p.ldc % 8 != 0) return GEPI_LEGACY;
  return GEPI_TMA_BF16_X;
}

int gemm_c_map(CUtensorMap* out, const UmmaGemmParams& p) {
  const uint64_t es = p.c_is_f32 ? 4 : 2;
  uint64_t dims[2] = {uint64_t(p.N), uint64_t(p.M)};
  uint64_t strides[1] = {uint64_t(p.ldc) * es};
  uint32_t box[2] = {uint32_t(128 / es), 32};
  return make_tmap_typed(out, p.c_is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p.C, 2, dims,
                         strides, box, true);
}

extern int g_force_legacy_epilogue;

}  // namespace masr

using namespace masr;

// C[M,N] = A[M,K] B[N,K]^T  (a_mn / b_mn = 0) or with MN-major operands:
//   a_mn: A is stored as At[K, M] (M contiguous, leading dimension lda)
//   b_mn: B is stored as Bt[K, N] (N contiguous, leading dimension ldb)
extern "C" int masr_umma_gemm_ex(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                                 void* C, int c_dtype, int64_t ldc, const float* bias,
                                 int M, int N, int K, int flags, int splitk, const masr_gemm_epilogue* epi, void* stream) {
  MASR_REQUIRE(M > 0 && N > 0 && K > 0, "umma gemm: empty problem");
  MASR_REQUIRE(!(flags & MASR_GEMM_SPLITK) || c_dtype == MASR_F32, "umma gemm: split-K needs an fp32 C");
  MASR_REQUIRE(!((flags & MASR_GEMM_SPLITK) && (flags & MASR_GEMM_RELU)), "umma gemm: split-K cannot fuse ReLU");
  const int total_kb = (K + UG_BK - 1) / UG_BK;
  int kb_per_split = total_kb;
  if ((flags & MASR_GEMM_SPLITK) && splitk > 1) kb_per_split = (total_kb + splitk - 1) / splitk;
  // narrow tiles when 128-wide ones would leave most of the 148 SMs idle AND the k-loop is short (decoder-sized,
  // latency-bound problems).  Long k-loops are bound by the L2->shared-memory operand traffic (measured ~6 TB/s
  // chip-wide for these access patterns), where a 128x64 tile moves 1.5x the bytes per flop of a 128x128 one.
  const int64_t tiles128 = ceil_div64(M, UG_BM) * ceil_div64(N, 128) * ceil_div64(total_kb, kb_per_split);
  const int BN = (N <= 64 || (tiles128 * 3 < sm_count() * 2 && kb_per_split <= 16)) ? 64 : 128;
  CUtensorMap ma, mb;
  int rc = gemm_operand_map(&ma, A, lda, M, K, a_mn != 0, UG_BM);
  if (rc != MASR_OK) return rc;
  rc = gemm_operand_map(&mb, B, ldb, N, K, b_mn != 0, BN);
  if (rc != MASR_OK) return rc;
  UmmaGemmParams p{M, N, K, C, ldc, c_dtype == MASR_F32 ? 1 : 0, bias, flags, kb_per_split, 0,
                   nullptr, nullptr, 0, 1.f, 0.f, 1.f, 0, nullptr, 0, nullptr, 0, nullptr, 1, 1};
  if (epi != nullptr) {
    MASR_REQUIRE(epi->mask == nullptr || c_dtype == MASR_BF16, "umma gemm: the output mask needs a bf16 C");
    MASR_REQUIRE(!(epi->p_drop > 0.f && (flags & MASR_GEMM_SPLITK)), "umma gemm: split-K cannot fuse dropout");
    p.rowsum = epi->rowsum;
    p.mask = static_cast<const __nv_bfloat16*>(epi->mask); p.ldmask = epi->ldmask; p.mask_scale = epi->mask_scale;
    if (epi->p_drop > 0.f) {
      p.p_drop = epi->p_drop; p.inv_keep = 1.f / (1.f - epi->p_drop);
      p.seed = epi->seed; p.seed_ptr = g_seed_dev_ptr; p.site = epi->site;
    }
    if (epi->dot_src != nullptr) {
      MASR_REQUIRE(epi->dot_out != nullptr && epi->dot_L > 0 && epi->dot_H > 0 && N == epi->dot_H * 64 && M % epi->dot_L == 0 &&
                   epi->lddot % 8 == 0 && (reinterpret_cast<uintptr_t>(epi->dot_src) & 15) == 0 && !(flags & MASR_GEMM_SPLITK),
                   "umma gemm: row-dot epilogue needs N = heads * 64, M = batch * L, 16 B aligned rows, no split-K");
      p.dot_src = static_cast<const __nv_bfloat16*>(epi->dot_src); p.lddot = epi->lddot;
      p.dot_out = epi->dot_out; p.dot_L = epi->dot_L; p.dot_H = epi->dot_H;
    }
  }
  cudaStream_t st = as_stream(stream);
  // large problems: persistent CTA-pair kernel (256-row tiles, cta_group::2, double-buffered TMEM), gemm_pair_umma.cu
  if (umma_pair_preferred(M, N, K, flags))
    return launch_umma_pair(A, lda, a_mn, B, ldb, b_mn, p, (flags & MASR_GEMM_SPLITK) ? -1 : 1, 0, st);
  CUtensorMap mc = ma;
  const int epi_kind = g_force_legacy_epilogue ? GEPI_LEGACY : gemm_pick_epilogue(p);
  if (epi_kind != GEPI_LEGACY) {
    rc = gemm_c_map(&mc, p);
    if (rc != MASR_OK) return rc;
  }
  return BN == 64 ? launch_umma_bn<64, 4>(a_mn, b_mn, epi_kind, ma, mb, mc, p, st)
                  : launch_umma_bn<128, 3>(a_mn, b_mn, epi_kind, ma, mb, mc, p, st);
}

extern "C" int masr_umma_gemm(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                              void* C, int c_dtype, int64_t ldc, const float* bias,
                              int M, int N, int K, int flags, int splitk, void* stream) {
  return masr_umma_gemm_ex(A, lda, a_mn, B, ldb, b_mn, C, c_dtype, ldc, bias, M, N, K, flags, splitk, nullptr, stream);
}

extern "C" int masr_umma_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb,
                                 void* C, int c_dtype, int64_t ldc, const float* bias,
                                 int M, int N, int K, int flags, void* stream) {
  return masr_umma_gemm(A, lda, 0, B, ldb, 0, C, c_dtype, ldc, bias, M, N, K, flags, 1, stream);
}

/* Explicit entry to the CTA-pair kernel (tests / probes): bn = 128 or 256 (0 = choose), splitk <= 0 = choose. */
extern "C" int masr_umma_gemm_pair(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                                   void* C, int c_dtype, int64_t ldc, const float* bias,
                                   int M, int N, int K, int flags, int splitk, int bn, const masr_gemm_epilogue* epi, void* stream) {
  MASR_REQUIRE(M > 0 && N > 0 && K > 0, "pair gemm: empty problem");
  MASR_REQUIRE(!(flags & MASR_GEMM_SPLITK) || c_dtype == MASR_F32, "pair gemm: split-K needs an fp32 C");
  MASR_REQUIRE(!((flags & MASR_GEMM_SPLITK) && (flags & MASR_GEMM_RELU)), "pair gemm: split-K cannot fuse ReLU");
  MASR_REQUIRE(bn == 0 || bn == 128 || bn == 256, "pair gemm: bn must be 0, 128 or 256");
  UmmaGemmParams p{M, N, K, C, ldc, c_dtype == MASR_F32 ? 1 : 0, bias, flags, 0, 0,
                   nullptr, nullptr, 0, 1.f, 0.f, 1.f, 0, nullptr, 0, nullptr, 0, nullptr, 1, 1};
  if (epi != nullptr) {
    p.rowsum = epi->rowsum;
    p.mask = static_cast<const __nv_bfloat16*>(epi->mask); p.ldmask = epi->ldmask; p.mask_scale = epi->mask_scale;
    if (epi->p_drop > 0.f) {
      p.p_drop = epi->p_drop; p.inv_keep = 1.f / (1.f - epi->p_drop);
      p.seed = epi->seed; p.seed_ptr = g_seed_dev_ptr; p.site = epi->site;
    }
    if (epi->dot_src != nullptr) {
      p.dot_src = static_cast<const __nv_bfloat16*>(epi->dot_src); p.lddot = epi->lddot;
      p.dot_out = epi->dot_out; p.dot_L = epi->dot_L; p.dot_H = epi->dot_H;
    }
  }
  return launch_umma_pair(A, lda, a_mn, B, ldb, b_mn, p, splitk, bn, as_stream(stream));
}

extern "C" int masr_gemm_group_begin(void) {
  MASR_REQUIRE(!g_group_on, "masr_gemm_group_begin: a group is already open");
  if (g_group == nullptr) g_group = new std::vector<GroupRec>();
  g_group->clear();
  g_group_on = true;
  return MASR_OK;
}

extern "C" int masr_gemm_group_end(void* stream) {
  MASR_REQUIRE(g_group_on, "masr_gemm_group_end: no open group");
  g_group_on = false;
  cudaStream_t st = as_stream(stream);
  std::vector<GroupRec>& recs = *g_group;
  std::vector<char> done(recs.size(), 0);
  g_group_recorded = int(recs.size());
  g_group_launched = 0;
  for (size_t i = 0; i < recs.size(); ++i) {
    if (done[i]) continue;
    GemmGroupTable t;
    t.n = 0; t.total = 0;
    for (size_t j = i; j < recs.size(); ++j) {
      if (done[j] || recs[j].launch != recs[i].launch) continue;
      if (t.n == GG_MAX) {              // table full: launch it and start the next one of the same instantiation
        const int rc = recs[i].launch(t, st);
        if (rc != MASR_OK) { recs.clear(); return rc; }
        ++g_group_launched;
        t.n = 0; t.total = 0;
      }
      GemmGroupEntry& e = t.e[t.n++];
      e = recs[j].e;
      e.cta_begin = t.total;
      t.total += e.nx * e.ny * e.nz;
      done[j] = 1;
    }
    if (t.n > 0) {
      const int rc = recs[i].launch(t, st);
      if (rc != MASR_OK) { recs.clear(); return rc; }
      ++g_group_launched;
    }
  }
  recs.clear();
  return MASR_OK;
}

/* problems recorded by / kernels launched by the last masr_gemm_group_end of this thread (launch accounting) */
extern "C" int masr_gemm_group_last(int* recorded, int* launched) {
  if (recorded != nullptr) *recorded = g_group_recorded;
  if (launched != nullptr) *launched = g_group_launched;
  return MASR_OK;
}

extern "C" int masr_gemm_set_stage_cap(int stages) {
  g_stage_cap = stages > 0 ? stages : 0;
  return MASR_OK;
}

/* 0: masr_umma_gemm* never uses the CTA-pair kernel; 1 (default): by problem size.  For A/B measurements. */
extern "C" int masr_gemm_set_pair_mode(int mode) {
  set_pair_mode(mode);
  return MASR_OK;
}
