// 3x3 convolution forward / dgrad, "band" variant of the implicit GEMM in conv_umma.cu.
//
// conv_umma.cu fetches, for each 128-pixel tile, nine shifted input boxes plus the matching weight k-blocks:
// 9x the input bytes and the whole weight tensor per 128 output pixels cross the L2 -> shared-memory path,
// and that path (not the tensor pipe) bounds the kernel (measured 430-550 TFLOP/s at ~10 TB/s of operand
// traffic).  Here the image is addressed in a PADDED, FLATTENED pixel space
//     f = (h + 1) * (W + 2) + (w + 1),          h in [-1, H], w in [-1, W]   (the ring of zeros is the padding)
// in which a tap (dh, dw) is the constant offset dh * (W + 2) + dw.  A tile = 256 consecutive centres f.  ONE
// 4-D TMA box {64 ch, W + 2, NR rows, 1} (out-of-bounds zero fill = padding) brings in every input pixel the
// tile needs, halo included, as NR * (W + 2) shared-memory rows of 128 B.  The A operand of tap (dh, dw) for
// the m-th group of 128 centres is then just that region read from row
//     off + 128 m + (dh + 1) (W + 2) + (dw + 1)
// i.e. the nine taps are nine UMMA descriptors into the SAME bytes (start addresses at arbitrary multiples of
// 128 B inside the 128-byte-swizzled region: the swizzle is a function of the absolute shared-memory address,
// which TMA used when it wrote the rows).  Centres that fall on the zero ring (2 of W + 2 per row) are computed
// and discarded.  Input traffic drops from 9x to ~(NR W2)/256 = 1.3-1.5x, weight traffic is halved (one weight
// k-block feeds two 128-row MMAs).
//
// Persistent CTAs (one per SM), 224 threads:
//   warp 0  weight (B) producer : ring of SB k-blocks [BN x 64]
//   warp 1  MMA issuer          : per tile 9 * CH k-blocks x 2 M-groups x 4 tcgen05.mma, accumulators
//                                 double-buffered in TMEM (2 x 2 x BN columns) so the epilogue of tile i
//                                 overlaps the main loop of tile i + 1
//   warps 2-5 epilogue          : TMEM -> bias / ReLU (fwd) or ReLU mask (dgrad) -> coalesced NHWC stores
//   warp 6  input (A) producer  : ring of 2 slots, one slot = one 64-channel chunk of one tile's region
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "epilogue.cuh"

namespace masr {

constexpr int CB_THREADS = 256;               // warp 7: second MMA issuer of the resident-weight variant
constexpr int CB_MAX_SB = 12;                 // weight ring depth is chosen at launch from the free shared memory

struct BandParams {
  int B, H, W, W2;
  int Cin;                 // the conv's Cin (column stride of a tap inside Wp)
  int tiles_per_img, ntiles;
  int nr;                  // padded rows per TMA box
  uint32_t slot_bytes;     // bytes of one A slot (multiple of 1024)
  int sb;                  // weight ring depth (RES: 9 * CH resident k-blocks)
  int rows_per_box;        // padded rows per TMA box of the input region (nr is a multiple of it)
  int pf_dist;             // L2 prefetch distance in tiles (0 = off)
  __nv_bfloat16* out;
  const __nv_bfloat16* relu_src;   // dgrad: multiply by (relu_src > 0)
  const float* bias;               // fwd
  int relu;                        // fwd
};

// Note on descriptors into the middle of a swizzled region: the matrix-descriptor "base offset" field stays 0.
// Measured on B200: with base offset = (start address >> 7) & 7 every result is wrong, with 0 every start address
// that is a multiple of 128 B works -- TMA and tcgen05.mma both apply the 128B swizzle to ABSOLUTE address bits.

// CH = reduction channels / 64, BN = output channels, MODE 0 = forward, 1 = dgrad (weights read MN-major from the
// forward layout), 2 = dgrad with the transposed weight layout wpt [Cin][tap][Cout] (K-major B like the forward:
// measurably faster than MN-major B), NM = 128-centre groups per tile,
// RES = the whole weight tensor (9 * CH k-blocks) stays resident in shared memory (loaded once per CTA)
template <int CH, int BN, int MODE, int NM, bool RES>
__global__ void __launch_bounds__(CB_THREADS, 1)
umma_conv_band_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, BandParams p) {
  using namespace umma;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = EpiLayout<BN, __nv_bfloat16>::BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                   // 2 slots
  constexpr int CB_MT = 128 * NM;
  const int SB = p.sb;
  unsigned char* sB = sA + 2 * p.slot_bytes;                  // SB k-blocks
  unsigned char* sStage = sB + SB * B_BYTES;                  // epilogue staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + ((STAGE_BYTES + 15) & ~15u));
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;        // [CB_MAX_SB]
  uint64_t* b_empty = bars + 4 + CB_MAX_SB;
  uint64_t* t_full = bars + 4 + 2 * CB_MAX_SB;    // [2] accumulator buffer complete
  uint64_t* t_empty = t_full + 2;             // [2] accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* sbias = reinterpret_cast<float*>(t_empty + 4);
  // RES: two accumulators per buffer (one per issuing warp)
  constexpr uint32_t TMEM_COLS = RES ? 4 * NM * BN : ((2 * NM * BN < 32) ? 32 : 2 * NM * BN);   // 256 or 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    // RES: two issuing warps (each commits once per chunk / tile)
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], RES ? 2 : 1); mbar_init(&t_full[s], RES ? 2 : 1); mbar_init(&t_empty[s], 4); }
    for (int s = 0; s < CB_MAX_SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int W2 = p.W2;
  const int first = int(blockIdx.x), step = int(gridDim.x);

  if (warp == 6) {
    // ===== input (A) producer (converged warp, elected lane issues) =====
    {
      const uint32_t tx = uint32_t(p.nr) * uint32_t(W2) * 128u;
      int slot = 0; uint32_t ph = 0;
      for (int tile = first; tile < p.ntiles; tile += step) {
        const int b = tile / p.tiles_per_img, ti = tile % p.tiles_per_img;
        const int f0 = W2 + 1 + ti * CB_MT;
        const int hp_lo = (f0 - W2 - 1) / W2;
        // L2 prefetch of the region this CTA will load `pf` tiles from now (only two slots = one load in flight:
        // without it every load pays the full HBM latency and the tile period is latency-bound)
        const int ptile = tile + p.pf_dist * step;
        if (p.pf_dist > 0 && ptile < p.ntiles && elect_one_sync()) {
          const int pb = ptile / p.tiles_per_img, pti = ptile % p.tiles_per_img;
          const int pf0 = W2 + 1 + pti * CB_MT;
          const int php = (pf0 - W2 - 1) / W2;
          for (int cc = 0; cc < CH; ++cc)
            for (int r = 0; r < p.nr; r += p.rows_per_box) tma_prefetch_4d(&map_a, cc * 64, -1, php - 1 + r, pb);
        }
        __syncwarp();
#pragma unroll 1
        for (int cc = 0; cc < CH; ++cc) {
          mbar_wait(&a_empty[slot], ph ^ 1);
          if (elect_one_sync()) {
          mbar_arrive_expect_tx(&a_full[slot], tx);
          // one box per padded row (W2 pixels): a single large box is served serially by the TMA unit (measured
          // ~16 B/cycle), several boxes stream concurrently.  Row r lands at a 128-byte aligned (not 1024-byte
          // aligned) offset: the 128B swizzle is a function of the absolute shared-memory address, so the rows
          // tile the region exactly as one large box would.
          for (int r = 0; r < p.nr; r += p.rows_per_box)
            tma_load_4d(sA + slot * p.slot_bytes + uint32_t(r) * uint32_t(W2) * 128u, &map_a, &a_full[slot], cc * 64, -1,
                        hp_lo - 1 + r, b);
          }
          __syncwarp();
          slot ^= 1; if (slot == 0) ph ^= 1;
        }
      }
    }
  } else if (warp == 0) {
    // ===== weight (B) producer =====
    if (lane == 0) {
      auto load_kb = [&](int kb, unsigned char* sb, uint64_t* bar) {
        const int cc = kb / 9, tap = kb % 9;
        if (MODE != 1) {
          tma_load_2d(sb, &map_w, bar, tap * p.Cin + cc * 64, 0);                          // box {64 k, BN rows}
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)                                                 // box {64 ci, 64 rows(co)}
            tma_load_2d(sb + c * 8192, &map_w, bar, tap * p.Cin + c * 64, cc * 64);
        }
      };
      if (RES) {
        if (first < p.ntiles) {
          mbar_arrive_expect_tx(&b_full[0], 9 * CH * B_BYTES);
          for (int kb = 0; kb < 9 * CH; ++kb) load_kb(kb, sB + kb * B_BYTES, &b_full[0]);
        }
      } else {
        int s = 0; uint32_t ph = 0;
        for (int tile = first; tile < p.ntiles; tile += step) {
#pragma unroll 1
          for (int kb = 0; kb < 9 * CH; ++kb) {
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&b_full[s], B_BYTES);
            load_kb(kb, sB + s * B_BYTES, &b_full[s]);
            if (++s == SB) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || (RES && warp == 7)) {
    // ===== MMA issuer: the whole warp runs the loop (uniform control flow and addresses), one elected lane issues.
    // The single issuing thread is the critical resource of the N = 64 layers (a 128x64x16 MMA occupies the tensor
    // pipe for only 32 cycles): descriptors are built ONCE per (tile, chunk) and advanced by adding encoded byte
    // offsets (start address field = bytes >> 4: +2 per K16 step, +8 per pixel row, +1024 per 128-row group).
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, MODE == 1 ? 1 : 0);   // B major
      int aslot = 0; uint32_t aph = 0;
      int bs = 0; uint32_t bph = 0;
      int it = 0;
      // encoded row offset of tap t relative to the region's first needed row
      int tap_rows[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3 - 1, dw = t % 3 - 1;
        tap_rows[t] = (MODE == 0) ? (dh + 1) * W2 + (dw + 1) : (1 - dh) * W2 + (1 - dw);
      }
      const uint64_t db_res0 = (MODE == 1) ? desc_mnmajor_sw128(smem_u32(sB), 8192) : desc_kmajor_sw128(smem_u32(sB));
      for (int tile = first; tile < p.ntiles; tile += step, ++it) {
        const int ti = tile % p.tiles_per_img;
        const int f0 = W2 + 1 + ti * CB_MT;
        const int hp_lo = (f0 - W2 - 1) / W2;
        const int off = (f0 - W2 - 1) - hp_lo * W2;              // region row of the first needed pixel
        const int ab = it & 1;
        mbar_wait(&t_empty[ab], ((it >> 1) & 1) ^ 1);            // epilogue has drained this accumulator buffer
        tc_fence_after();
        // RES: a single thread dispatches an N = 64 MMA only every ~50 cycles (the MMA occupies the tensor pipe
        // for 32): warp 1 issues taps 0-4 into accumulator 0, warp 7 taps 5-8 into accumulator 1; the epilogue adds them
        const int issuer = (RES && warp == 7) ? 1 : 0;
        const uint32_t acc0 = tmem_base + uint32_t((RES ? 2 * ab + issuer : ab) * NM * BN);
        if (RES && it == 0) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
#pragma unroll 1
        for (int cc = 0; cc < CH; ++cc) {
          mbar_wait(&a_full[aslot], aph);
          tc_fence_after();
          const uint64_t da0 = make_smem_desc(smem_u32(sA + aslot * p.slot_bytes) + uint32_t(off) * 128u, 16, 1024);
          if constexpr (RES) {
            // resident weights: no barrier inside the chunk -> ONE election around all 9 x NM x 4 MMAs (the
            // per-tap elect / reconverge / R2UR sequence costs more issue cycles than four N = 64 MMAs take)
            if (elect_one_sync()) {
              const int t_begin = issuer == 0 ? 0 : 5, t_end = issuer == 0 ? 5 : 9;
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                if (tap < t_begin || tap >= t_end) continue;
                const uint64_t db0 = db_res0 + uint64_t((cc * 9 + tap) * (B_BYTES >> 4));
                const uint64_t da_tap = da0 + uint64_t(tap_rows[tap]) * 8u;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    mma_f16_ss(acc0 + uint32_t(m * BN), da_tap + uint64_t(m * 1024 + k * 2), db0 + uint64_t(MODE == 1 ? k * 128 : k * 2),
                               idesc, (cc > 0 || tap > t_begin || k > 0) ? 1u : 0u);
                }
              }
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t sb = smem_u32(sB + bs * B_BYTES);
              const uint64_t db0 = (MODE == 1) ? desc_mnmajor_sw128(sb, 8192) : desc_kmajor_sw128(sb);
              const uint64_t da_tap = da0 + uint64_t(tap_rows[tap]) * 8u;
              if (elect_one_sync()) {
#pragma unroll
                for (int m = 0; m < NM; ++m) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    // K16 step: A +32 B; B K-major +32 B, MN-major (dgrad) +16 co-rows = +2048 B
                    const uint64_t da = da_tap + uint64_t(m * 1024 + k * 2);
                    const uint64_t db = db0 + uint64_t(MODE == 1 ? k * 128 : k * 2);
                    mma_f16_ss(acc0 + uint32_t(m * BN), da, db, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  }
                }
                mma_commit(&b_empty[bs]);
              }
              __syncwarp();
              if (++bs == SB) { bs = 0; bph ^= 1; }
            }
          }
          if (elect_one_sync()) mma_commit(&a_empty[aslot]);
          __syncwarp();
          aslot ^= 1; if (aslot == 0) aph ^= 1;
        }
        if (elect_one_sync()) mma_commit(&t_full[ab]);
        __syncwarp();
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // ===== epilogue =====
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const bool use_bias = (MODE == 0) && p.bias != nullptr;   // MODE 1 / 2: dgrad
    if (use_bias) {
      for (int i = et; i < BN; i += 128) sbias[i] = p.bias[i];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int last_centre = p.H * W2 + p.W;
    int it = 0;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const int b = tile / p.tiles_per_img, ti = tile % p.tiles_per_img;
      const int f0 = W2 + 1 + ti * CB_MT;
      const int ab = it & 1;
      if (MODE != 0 && p.relu_src != nullptr) {
        // dgrad: pull the ReLU mask rows of the tile two iterations ahead into L2 (one 128 B line per lane and
        // 64 channels): the mask streams from HBM, and the epilogue's own loads would otherwise expose that
        // latency once per 128-centre group, making the epilogue (not the MMAs) the pace setter
        const int ptile = tile + 2 * step;
        if (ptile < p.ntiles) {
          const int pb = ptile / p.tiles_per_img, pti = ptile % p.tiles_per_img;
#pragma unroll 1
          for (int m = 0; m < NM; ++m) {
            const int pf = W2 + 1 + pti * CB_MT + m * 128 + q * 32 + lane;
            const int php = pf / W2, pwp = pf - php * W2;
            if (pf <= last_centre && pwp >= 1 && pwp <= p.W && php >= 1 && php <= p.H) {
              const __nv_bfloat16* mp = p.relu_src + ((int64_t(pb) * p.H + (php - 1)) * p.W + (pwp - 1)) * BN;
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(mp + c * 64));
            }
          }
        }
      }
      mbar_wait(&t_full[ab], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < NM; ++m) {
        const int f = f0 + m * 128 + q * 32 + lane;
        const int hp = f / W2, wp = f - hp * W2;
        const bool valid = (f <= last_centre) && (wp >= 1) && (wp <= p.W) && (hp >= 1) && (hp <= p.H);
        const int64_t pix = (int64_t(b) * p.H + (hp - 1)) * p.W + (wp - 1);
        __nv_bfloat16* orow = valid ? p.out + pix * BN : nullptr;
        EpiOpts o;
        o.sbias = use_bias ? sbias : nullptr;
        o.relu = MODE == 0 && p.relu != 0;
        if (MODE != 0 && p.relu_src != nullptr) {          // mask row = destination row + a constant offset
          o.mask_at_dst_delta = true;
          o.mask_delta = reinterpret_cast<const unsigned char*>(p.relu_src) - reinterpret_cast<const unsigned char*>(p.out);
        }
        if (RES) o.acc2_offset = uint32_t(NM * BN);
        epilogue_tile<BN, __nv_bfloat16>(tmem_base + uint32_t(((RES ? 2 * ab : ab) * NM + m) * BN), q, lane, sStage, orow, BN, true, EPI_STORE, o);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

struct BandGeom { int W2, nr, tiles_per_img, nm, res, sb; uint32_t slot_bytes; size_t smem; };

// Picks (NM, RES, ring depth) for the shape: the whole weight tensor resident when it fits next to a
// double-buffered 128-centre input region (the Cin = Cout = 64 layer), else 256-centre tiles with the
// deepest weight ring the remaining shared memory allows (k-blocks are small: the ring is latency-bound).
static bool band_geometry(int H, int W, int ch, int BN, BandGeom* g) {
  g->W2 = W + 2;
  if (g->W2 > 256) return false;
  const size_t stage = (BN == 64) ? EpiLayout<64, __nv_bfloat16>::BYTES : EpiLayout<128, __nv_bfloat16>::BYTES;
  const size_t fixed = ((stage + 15) & ~size_t(15)) + (4 + 2 * CB_MAX_SB + 6) * 8 + BN * 4 + 1024;
  const size_t limit = 227 * 1024;
  for (int variant = 0; variant < 2; ++variant) {
    const int nm = variant == 0 ? 1 : 2;
    const int res = variant == 0 ? 1 : 0;
    if (res && !(ch == 1 && BN == 64)) continue;          // only instantiated for the Cin = Cout = 64 layer
    const int mt = 128 * nm;
    const int nr = (mt + 3 * g->W2 + 1 + g->W2 - 1) / g->W2;
    if (nr > 256) continue;
    const uint32_t slot = (uint32_t(nr) * uint32_t(g->W2) * 128u + 1023u) & ~1023u;
    const size_t kb_bytes = size_t(BN) * 128;
    int sb;
    if (res) {
      sb = 9 * ch;
      if (2 * size_t(slot) + sb * kb_bytes + fixed > limit) continue;
    } else {
      if (2 * size_t(slot) + 3 * kb_bytes + fixed > limit) continue;
      sb = int(std::min<size_t>(CB_MAX_SB, (limit - fixed - 2 * size_t(slot)) / kb_bytes));
    }
    g->nm = nm; g->res = res; g->sb = sb; g->nr = nr; g->slot_bytes = slot;
    g->tiles_per_img = (H * g->W2 - 2 + mt - 1) / mt;
    g->smem = 2 * size_t(slot) + sb * kb_bytes + fixed;
    return true;
  }
  return false;
}

// rows of the activation band per TMA box: one (measured best of 1 / 2 / all)
static int band_rows_per_box(int nr) { (void)nr; return 1; }

template <int CH, int BN, int MODE, int NM, bool RES>
static int launch_band2(const CUtensorMap& ma, const CUtensorMap& mw, const BandParams& p, const BandGeom& g, cudaStream_t st) {
  auto kern = umma_conv_band_kernel<CH, BN, MODE, NM, RES>;
  static bool attr = false;
  if (!attr) {
    MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(227 * 1024)));
    attr = true;
  }
  const unsigned grid = unsigned(std::min(p.ntiles, sm_count()));
  MASR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(CB_THREADS), g.smem, st, ma, mw, p));
  return MASR_OK;
}
template <int CH, int BN, int MODE>
static int launch_band(const CUtensorMap& ma, const CUtensorMap& mw, const BandParams& p, const BandGeom& g, cudaStream_t st) {
  if (g.res) {
    if constexpr (CH == 1 && BN == 64) return launch_band2<CH, BN, MODE, 1, true>(ma, mw, p, g, st);
    else return 1;
  }
  return launch_band2<CH, BN, MODE, 2, false>(ma, mw, p, g, st);
}

// returns MASR_OK when the band kernel ran, 1 when the caller should use the tile kernel of conv_umma.cu
// mode 0: forward (wp [Cout][tap][Cin]); 1: dgrad reading wp MN-major; 2: dgrad with wpt [Cin][tap][Cout]
int conv_band_try(int mode, const void* act, const void* wp, void* out, const void* relu_src, const float* bias, int relu,
                  int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const int Cred = mode == 0 ? Cin : Cout, Cn = mode == 0 ? Cout : Cin;
  BandGeom g;
  if (!band_geometry(H, W, Cred / 64, Cn, &g)) return 1;
  if ((reinterpret_cast<uintptr_t>(act) & 15) != 0) return 1;
  CUtensorMap ma, mw;
  uint64_t dims[4] = {uint64_t(Cred), uint64_t(W), uint64_t(H), uint64_t(B)};
  uint64_t strides[3] = {uint64_t(Cred) * 2, uint64_t(W) * Cred * 2, uint64_t(H) * W * Cred * 2};
  const int rpb = band_rows_per_box(g.nr);
  uint32_t box[4] = {64, uint32_t(g.W2), uint32_t(rpb), 1};
  int rc = make_tmap_bf16(&ma, act, 4, dims, strides, box, true);
  if (rc != MASR_OK) return rc;
  // weight tensor as a 2-D map: rows = output channels of the GEMM for K-major B (modes 0, 2)
  uint64_t wd[2] = {uint64_t(9 * (mode == 2 ? Cout : Cin)), uint64_t(mode == 2 ? Cin : Cout)};
  uint64_t ws[1] = {wd[0] * 2};
  uint32_t wb[2] = {64, mode == 1 ? 64u : uint32_t(Cn)};
  rc = make_tmap_bf16(&mw, wp, 2, wd, ws, wb, true);
  if (rc != MASR_OK) return rc;
  const int kstride = mode == 2 ? Cout : Cin;              // columns per tap inside the weight rows
  const int pf = 0;                                        // L2 prefetch of the next band: measured, no effect
  BandParams p{B, H, W, g.W2, kstride, g.tiles_per_img, B * g.tiles_per_img, g.nr, g.slot_bytes, g.sb, rpb, pf,
               static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(relu_src), bias, relu};
  const int key = (mode << 2) | ((Cred == 128 ? 1 : 0) << 1) | (Cn == 128 ? 1 : 0);
  switch (key) {
    case 0: return launch_band<1, 64, 0>(ma, mw, p, g, st);
    case 1: return launch_band<1, 128, 0>(ma, mw, p, g, st);
    case 2: return launch_band<2, 64, 0>(ma, mw, p, g, st);
    case 3: return launch_band<2, 128, 0>(ma, mw, p, g, st);
    case 4: return launch_band<1, 64, 1>(ma, mw, p, g, st);
    case 5: return launch_band<1, 128, 1>(ma, mw, p, g, st);
    case 6: return launch_band<2, 64, 1>(ma, mw, p, g, st);
    case 7: return launch_band<2, 128, 1>(ma, mw, p, g, st);
    case 8: return launch_band<1, 64, 2>(ma, mw, p, g, st);
    case 9: return launch_band<1, 128, 2>(ma, mw, p, g, st);
    case 10: return launch_band<2, 64, 2>(ma, mw, p, g, st);
    default: return launch_band<2, 128, 2>(ma, mw, p, g, st);
  }
}

}  // namespace masr
