// 3x3 convolution (stride 1, pad 1) of the VGG front end as IMPLICIT GEMM on tcgen05 tensor cores.
// mono_transformer_torch.py:49-60 (nn.Conv2d + ReLU); 72 % of the model FLOPs live here.
//
// Activations are NHWC bf16.  Nothing is ever unfolded in memory: for each of the 9 taps the TMA engine
// fetches the SHIFTED [th x tw pixels, 64 channels] box of the input with a 4-D tensor map
// {C, W, H, B}; coordinates may be -1 or run past W/H, and the hardware's out-of-bounds zero fill IS
// the convolution's zero padding.  Each box lands in shared memory as th*tw rows of 128 B (one
// swizzle row per pixel) and is consumed directly as a K-major (fwd/dgrad) or MN-major (wgrad) UMMA
// operand.  Weights stay in the [Cout, tap*Cin + ci] layout produced by masr_conv_w_prep:
//   fwd  : D[pix, co] = sum_{tap,ci} X[pix+tap, ci] * Wp[co, tap*Cin+ci]      A K-major (4-D), B K-major
//   dgrad: D[pix, ci] = sum_{tap,co} dY[pix-tap, co] * Wp[co, tap*Cin+ci]     A K-major (4-D), B MN-major
//   wgrad: D_tap[co, ci] = sum_pix dY[pix, co] * X[pix+tap, ci]               A, B MN-major (4-D), split over
//          pixel tiles across CTAs, 3 taps (one kernel row) accumulate side by side in TMEM, fp32 atomics out
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "epilogue.cuh"

namespace masr {

constexpr int CV_THREADS = 192;

struct ConvTile { int tw, th, nw, nh; };

// choose the pixel rectangle (tw x th <= max_rows) maximising useful MMA rows
static ConvTile pick_tile(int H, int W, int max_rows) {
  ConvTile best{0, 0, 0, 0};
  double best_eff = -1.0;
  for (int nw = 1; nw <= 16; ++nw) {
    const int tw = (W + nw - 1) / nw;
    if (tw > max_rows || tw > 256) continue;
    const int th = std::min(std::min(max_rows / tw, 256), H);
    if (th < 1) continue;
    const int nh = (H + th - 1) / th;
    const double eff = (double(W) / (nw * tw)) * (double(tw * th) / max_rows) * (double(H) / (nh * th));
    if (eff > best_eff) { best_eff = eff; best = ConvTile{tw, th, nw, nh}; }
  }
  return best;
}

static int act_map(CUtensorMap* out, const void* base, int B, int H, int W, int C, int tw, int th) {
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 64 == 0, "conv: NHWC tensor must be 16 B aligned, C % 64 == 0");
  uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), uint64_t(B)};
  uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
  uint32_t box[4] = {64, uint32_t(tw), uint32_t(th), 1};
  return make_tmap_bf16(out, base, 4, dims, strides, box, true);
}

struct ConvParams {
  int B, H, W;
  int Cred;          // reduction channels (fwd: Cin, dgrad: Cout)
  int Cn;            // output channels of this GEMM (fwd: Cout, dgrad: Cin) == BN
  int Cin;           // the conv's Cin (column stride of a tap inside Wp)
  int tw, th, nw, nh;
  __nv_bfloat16* out;
  const __nv_bfloat16* relu_src;   // dgrad: multiply by (relu_src > 0)
  const float* bias;               // fwd
  int relu;                        // fwd
};

// MODE 0 = forward, 1 = dgrad
template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(CV_THREADS, 2)
umma_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_m, ConvParams p) {
  using namespace umma;
  constexpr uint32_t A_BYTES = 128 * 128;                 // up to 128 pixel rows of 128 B
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t MASK_BYTES = (MODE == 1) ? 128 * BN * 2 : 0;      // dgrad: ReLU mask tile, fetched up front
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* smask = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smask + MASK_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 2);
  float* sbias = reinterpret_cast<float*>(tmem_full_bar + 4);
  static_assert(STAGES * STAGE_BYTES >= EpiLayout<BN, __nv_bfloat16>::BYTES, "staging tile must fit in the pipeline stages");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int iw = tile % p.nw;
  const int ih = (tile / p.nw) % p.nh;
  const int b = tile / (p.nw * p.nh);
  const int h0 = ih * p.th, w0 = iw * p.tw;
  const int rows = p.th * p.tw;
  const int cpb = p.Cred / 64;
  const int num_kb = 9 * cpb;

  const bool has_mask = (MODE == 1) && p.relu_src != nullptr;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    if (has_mask) prefetch_tmap(&map_m);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_init(mask_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = uint32_t(rows) * 128 + B_BYTES;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (MODE == 1 && kb == STAGES && has_mask) {
          // the ReLU mask tile of the output pixels (same box as an A tile, no shift): needed only by the
          // epilogue, so it is queued behind the first ring fill and lands long before the main loop ends
          mbar_arrive_expect_tx(mask_bar, uint32_t(rows) * 128 * (BN / 64));
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_4d(smask + c * 16384, &map_m, mask_bar, c * 64, w0, h0, b);
        }
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        const int tap = kb / cpb, cc = kb % cpb;
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        if (MODE == 0) {
          tma_load_4d(sa, &map_a, &full_bar[s], cc * 64, w0 + dw, h0 + dh, b);
          tma_load_2d(sb, &map_w, &full_bar[s], tap * p.Cin + cc * 64, 0);          // box {64 k, BN rows(co)}
        } else {
          tma_load_4d(sa, &map_a, &full_bar[s], cc * 64, w0 - dw, h0 - dh, b);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)                                          // box {64 ci, 64 rows(co)}
            tma_load_2d(sb + c * 8192, &map_w, &full_bar[s], tap * p.Cin + c * 64, cc * 64);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, MODE == 1 ? 1 : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = desc_kmajor_sw128(sa + k * 32);
          const uint64_t db = (MODE == 1) ? desc_mnmajor_sw128(sb + k * 2048, 8192) : desc_kmajor_sw128(sb + k * 32);
          mma_f16_ss(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const bool use_bias = (MODE == 0) && p.bias != nullptr;
    if (use_bias) {
      for (int i = et; i < BN; i += 128) sbias[i] = p.bias[i];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int r = q * 32 + lane;
    const int h = h0 + r / p.tw, w = w0 + r % p.tw;
    const bool valid = (r < rows) && (h < p.H) && (w < p.W);
    const int64_t pix = (int64_t(b) * p.H + h) * p.W + w;
    __nv_bfloat16* orow = valid ? p.out + pix * p.Cn : nullptr;
    if (has_mask) mbar_wait(mask_bar, 0);
    EpiOpts o;
    o.sbias = use_bias ? sbias : nullptr;
    o.relu = MODE == 0 && p.relu != 0;
    o.smask = has_mask ? smask : nullptr;
    epilogue_tile<BN, __nv_bfloat16>(tmem_base, q, lane, smem, orow, BN, true, EPI_STORE, o);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BN); }
}

// ------------------------------------------------------------------ wgrad
struct WgradParams {
  int B, H, W, Cout, Cin;
  int tw, th, nw, nh;
  float* dwp;          // [Cout, 9*Cin] fp32, accumulated with atomics
  float* db;           // [Cout] fp32 bias gradient (+= sum over pixels of dy), may be NULL
};

// One CTA: kernel row dh = blockIdx.y - 1, pixel tiles blockIdx.x, +gridDim.x, ...  (split-K over pixels).
// Per stage: dY tile (A, MN-major, rows = pixels) and the three dw-shifted X tiles (B, MN-major).
template <int CI, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
umma_conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, WgradParams p) {
  using namespace umma;
  constexpr uint32_t A_BYTES = 2 * 8192;                  // 2 chunks of 64 co x 64 pixel rows
  constexpr uint32_t B_BYTES = (CI / 64) * 8192;          // per dw
  constexpr uint32_t STAGE_BYTES = A_BYTES + 3 * B_BYTES;
  constexpr uint32_t TMEM_COLS = (3 * CI <= 256) ? 256 : 512;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sones = smem + STAGES * STAGE_BYTES;     // 2 KB of bf16 1.0: B operand of the bias-gradient MMA
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sones + 2048);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  static_assert(3 * CI + 16 <= TMEM_COLS, "bias-gradient accumulator must fit");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dh = int(blockIdx.y) - 1;
  const bool rowsum = p.db != nullptr && blockIdx.y == 0;  // the dh = -1 CTAs also sum dy over their pixels
  const int rows = p.th * p.tw;                            // <= 64 pixel rows per stage
  const int ntiles = p.B * p.nh * p.nw;
  const int my_tiles = (ntiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int co_chunks = p.Cout / 64;

  pdl_launch_dependents();
  // rows [th*tw, 64) of every chunk are never written by TMA: zero the ring once so they contribute 0
  for (uint32_t i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += CV_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 128) reinterpret_cast<uint4*>(sones)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    {   // converged warp, one elected lane issues (umma::elect_one_sync)
      const uint32_t tx_bytes = uint32_t(rows) * 128 * uint32_t(co_chunks + 3 * (CI / 64));
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = int(blockIdx.x) + it * int(gridDim.x);
        const int iw = tile % p.nw, ih = (tile / p.nw) % p.nh, b = tile / (p.nw * p.nh);
        const int h0 = ih * p.th, w0 = iw * p.tw;
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          for (int c = 0; c < co_chunks; ++c) tma_load_4d(sa + c * 8192, &map_dy, &full_bar[s], c * 64, w0, h0, b);
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int c = 0; c < CI / 64; ++c)
              tma_load_4d(sa + A_BYTES + j * B_BYTES + c * 8192, &map_x, &full_bar[s], c * 64, w0 + (j - 1), h0 + dh, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, CI, 1, 1);
      constexpr uint32_t idesc_ones = make_idesc_bf16(128, 16, 1, 0);
      const uint64_t dones = desc_kmajor_sw128(smem_u32(sones));
      // two copies of the loop (with / without the bias-gradient accumulator): no conditionally issued tcgen05.mma
      auto tloop = [&](auto with_rowsum) {
        constexpr bool RS = decltype(with_rowsum)::value;
        for (int it = 0; it < my_tiles; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const uint32_t sb = sa + A_BYTES + j * B_BYTES;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = desc_mnmajor_sw128(sa + k * 2048, 8192);
                const uint64_t db = desc_mnmajor_sw128(sb + k * 2048, 8192);
                mma_f16_ss(tmem_base + uint32_t(j * CI), da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
              }
            }
            if constexpr (RS) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_f16_ss(tmem_base + uint32_t(3 * CI), desc_mnmajor_sw128(sa + k * 2048, 8192), dones, idesc_ones,
                           (it > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&empty_bar[s]);
          }
          __syncwarp();
        }
      };
      if (rowsum) tloop(std::true_type{}); else tloop(std::false_type{});
      if (elect_one_sync()) mma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    if (my_tiles > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int co = q * 32 + lane;
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const int tap = (dh + 1) * 3 + j;
        // staged through shared memory: coalesced fp32 atomics (consecutive lanes -> consecutive addresses)
        float* dst = (co < p.Cout) ? p.dwp + int64_t(co) * (9 * p.Cin) + tap * p.Cin : nullptr;
        epilogue_tile<CI, float>(tmem_base + uint32_t(j * CI), q, lane, smem, dst, CI, true, EPI_ATOMIC, EpiOpts());
      }
      if (rowsum) {                 // column 0 of the ones-accumulator = sum over this CTA's pixels of dy[:, co]
        float v[32];
        tmem_ld_32x32(tmem_base + uint32_t(3 * CI) + (uint32_t(q * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.Cout) atomicAdd(p.db + co, v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------ wgrad, row-group variant
// K-block = r whole image rows in a padded pitch Wp = W + 4 (w = -2 .. W + 1, out-of-bounds zero fill), rounded up
// to a multiple of 16 rows (<= 128).  dY and X use the SAME flattening, so the three dw taps of a kernel row are
// the X region read from row offsets -1 / 0 / +1: ONE X box per k-block instead of three shifted ones, and
// 85-90 % of the MMA rows carry pixels (the tile variant above filled 42 of 64).  Cout = 64 uses M = 64 MMAs
// (its accumulator occupies 16 lanes of each TMEM quadrant) instead of padding co to 128.
struct Wgrad2Params {
  int B, H, W, Cin, Cout;
  int wp, r, kr;        // pitch, image rows per k-block, MMA rows per k-block (multiple of 16)
  int ntiles, tiles_per_img;
  float* dwp; float* db;
};

template <int CO, int CI, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
umma_conv_wgrad2_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, Wgrad2Params p) {
  using namespace umma;
  constexpr uint32_t DY_CHUNK = 128 * 128;                 // up to 128 k-rows of 128 B per 64-co chunk
  constexpr uint32_t X_CHUNK = 17 * 1024;                  // 1 margin row + 128 + 1 margin row (130 x 128 B), padded to 17 KB
  static_assert(X_CHUNK % 1024 == 0, "X chunk must keep 1024 B alignment");
  constexpr uint32_t A_BYTES = (CO / 64) * DY_CHUNK;
  constexpr uint32_t B_BYTES = (CI / 64) * X_CHUNK;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (3 * CI + 16 <= 256) ? 256 : 512;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sones = smem + STAGES * STAGE_BYTES;
  unsigned char* sStage = smem;                            // fp32 epilogue staging aliases the (by then idle) ring
  static_assert(STAGES * STAGE_BYTES >= EpiLayout<CI, float>::BYTES, "staging tile must fit in the ring");
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sones + 2048);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dh = int(blockIdx.y) - 1;
  const bool rowsum = p.db != nullptr && blockIdx.y == 0;
  const int my_tiles = (p.ntiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int ksteps = p.kr / 16;

  pdl_launch_dependents();
  // rows never written by TMA (k-rows beyond wp * r, the X margin rows) must read as zeros: clear the ring once
  for (uint32_t i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += CV_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 128) reinterpret_cast<uint4*>(sones)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    const uint32_t tx_bytes = uint32_t(p.wp * p.r) * 128u * uint32_t(CO / 64 + CI / 64);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = int(blockIdx.x) + it * int(gridDim.x);
      const int b = tile / p.tiles_per_img, h0 = (tile % p.tiles_per_img) * p.r;
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      unsigned char* sa = smem + s * STAGE_BYTES;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
#pragma unroll
        for (int c = 0; c < CO / 64; ++c) tma_load_4d(sa + c * DY_CHUNK, &map_dy, &full_bar[s], c * 64, -2, h0, b);
#pragma unroll
        for (int c = 0; c < CI / 64; ++c)               // lands one row (128 B) into the chunk: row -1 is the margin
          tma_load_4d(sa + A_BYTES + c * X_CHUNK + 128, &map_x, &full_bar[s], c * 64, -2, h0 + dh, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(CO, CI, 1, 1);
    constexpr uint32_t idesc_ones = make_idesc_bf16(CO, 16, 1, 0);
    const uint64_t dones = desc_kmajor_sw128(smem_u32(sones));
    auto tloop = [&](auto with_rowsum) {
      constexpr bool RS = decltype(with_rowsum)::value;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sx = sa + A_BYTES + 128;           // X row 0 of the box
        // descriptors once per k-block; a K16 step advances both operands by 16 rows = 2048 B (+128 encoded),
        // the dw taps are one pixel row (128 B, +8 encoded) apart
        const uint64_t da0 = desc_mnmajor_sw128(sa, DY_CHUNK);
        const uint64_t dx0 = desc_mnmajor_sw128(sx - 128, X_CHUNK);
        if (elect_one_sync()) {
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = da0 + uint64_t(k * 128);
            const uint64_t dx = dx0 + uint64_t(k * 128);
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
#pragma unroll
            for (int j = 0; j < 3; ++j)                   // tap dw = j - 1: the same X bytes, one row earlier / later
              mma_f16_ss(tmem_base + uint32_t(j * CI), da, dx + uint64_t(j * 8), idesc, acc);
            if constexpr (RS) mma_f16_ss(tmem_base + uint32_t(3 * CI), da, dones, idesc_ones, acc);
          }
          mma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
    };
    if (rowsum) tloop(std::true_type{}); else tloop(std::false_type{});
    if (elect_one_sync()) mma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    if (my_tiles > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      // M = 128: TMEM lane = co.  M = 64: rows 16 q .. 16 q + 15 sit in lanes 0..15 of quadrant q.
      const int co = (CO == 128) ? q * 32 + lane : q * 16 + lane;
      const bool row_ok = (CO == 128) ? true : lane < 16;
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const int tap = (dh + 1) * 3 + j;
        float* dst = row_ok ? p.dwp + int64_t(co) * (9 * p.Cin) + tap * p.Cin : nullptr;
        epilogue_tile<CI, float>(tmem_base + uint32_t(j * CI), q, lane, sStage, dst, CI, true, EPI_ATOMIC, EpiOpts());
      }
      if (rowsum) {
        float v[32];
        tmem_ld_32x32(tmem_base + uint32_t(3 * CI) + (uint32_t(q * 32) << 16), v);
        tmem_ld_wait();
        if (row_ok) atomicAdd(p.db + co, v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int CO, int CI, int STAGES>
static int launch_wgrad2(const CUtensorMap& mdy, const CUtensorMap& mx, const Wgrad2Params& p, cudaStream_t st) {
  constexpr size_t STAGE = size_t(CO / 64) * 128 * 128 + size_t(CI / 64) * 17 * 1024;
  const size_t smem = STAGES * STAGE + 2048 + 256 + 1024;
  auto kern = umma_conv_wgrad2_kernel<CO, CI, STAGES>;
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  const int gx = std::max(1, std::min(p.ntiles, sm_count() / 3));
  MASR_CHECK_CUDA(launch_pdl(kern, dim3(unsigned(gx), 3, 1), dim3(CV_THREADS), smem, st, mdy, mx, p));
  return MASR_OK;
}

// ------------------------------------------------------------------ wgrad for Cin = Cout = 64: tap pairs as M = 128
// With 64 output channels the [co x ci] accumulators above are M = 64 MMAs, which occupy the tensor pipe exactly
// as long as M = 128 ones.  Here the product is transposed and TWO TAPS share one MMA:
//     D[(tap, ci), co] = sum_pix X[pix + tap, ci] . dY[pix, co]
// A = the X region (MN-major, K rows = pixels).  The second 64-row half of M is the SAME bytes shifted by one tap:
// the descriptor's leading-dimension byte offset (distance between 64-element M chunks) is set to 128 B (next dw)
// or Wp * 128 B (next dh).  Nine taps = four M = 128 groups + one M = 64 group: 5 MMAs per K16 step instead of 9,
// all nine taps in ONE CTA (X region = three image rows), so X and dY cross the L2->smem path once instead of three
// times.  The bias gradient is an all-ones A tile against the same dY operand.
struct Wgrad3Params {
  int B, H, W;
  int wp, kr;           // pitch W + 4, MMA rows per k-block (multiple of 16)
  int ntiles;           // B * H image rows
  float* dwp; float* db;
};

template <int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
umma_conv_wgrad3_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, Wgrad3Params p) {
  using namespace umma;
  constexpr uint32_t DY_BYTES = 128 * 128;                 // 16 KB: up to 128 k-rows
  constexpr uint32_t X_BYTES = 48 * 1024;                  // margin row + 3 image rows (<= 3 * 124) + tail
  constexpr uint32_t STAGE_BYTES = DY_BYTES + X_BYTES;
  constexpr uint32_t TMEM_COLS = 512;                      // 4 x 64 (tap pairs) + 64 (last tap | bias) = 320
  constexpr int PITCH = 65;                                // staging row pitch (floats): conflict-free transposition
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sones = smem + STAGES * STAGE_BYTES;
  float* sStage = reinterpret_cast<float*>(smem);          // epilogue staging aliases the (by then idle) ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sones + 2048);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  static_assert(STAGES * STAGE_BYTES >= 128 * PITCH * 4, "staging must fit in the ring");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.ntiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int ksteps = p.kr / 16;
  const int wp = p.wp;

  pdl_launch_dependents();
  for (uint32_t i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += CV_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 128) reinterpret_cast<uint4*>(sones)[threadIdx.x] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    const uint32_t tx_bytes = uint32_t(wp) * 128u * 4u;   // 1 row of dY + 3 rows of X
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = int(blockIdx.x) + it * int(gridDim.x);
      const int b = tile / p.H, h = tile % p.H;
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      unsigned char* sa = smem + s * STAGE_BYTES;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(sa, &map_dy, &full_bar[s], 0, -2, h, b);                       // box {64, wp, 1, 1}
        tma_load_4d(sa + DY_BYTES + 128, &map_x, &full_bar[s], 0, -2, h - 1, b);   // box {64, wp, 3, 1}, after the margin row
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 64, 1, 1);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sdy = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t sx = sdy + DY_BYTES;                 // smem row 0 = margin, box row r at smem row r + 1
      // A descriptors: start row (dh + 1) * wp + dw + 1, second M chunk one tap further (LBO)
      const uint64_t db0 = make_smem_desc(sdy, 0, 1024);
      const uint64_t a_g0 = make_smem_desc(sx + uint32_t(0 * wp + 0) * 128u, 128, 1024);            // (-1,-1) & (-1, 0)
      const uint64_t a_g1 = make_smem_desc(sx + uint32_t(1 * wp + 0) * 128u, 128, 1024);            // ( 0,-1) & ( 0, 0)
      const uint64_t a_g2 = make_smem_desc(sx + uint32_t(2 * wp + 0) * 128u, 128, 1024);            // ( 1,-1) & ( 1, 0)
      const uint64_t a_g3 = make_smem_desc(sx + uint32_t(0 * wp + 2) * 128u, uint32_t(wp) * 128u, 1024);   // (-1, 1) & ( 0, 1)
      // ( 1, 1) | all-ones rows: the second M chunk is the 2 KB tile of 1.0 (its distance shrinks as the start
      // address advances), so rows 64..127 of this accumulator are sum_pix dY[pix, co] = the bias gradient
      const uint32_t g4_start = sx + uint32_t(2 * wp + 2) * 128u;
      const uint64_t a_g4 = make_smem_desc(g4_start, smem_u32(sones) - g4_start, 1024);
      if (elect_one_sync()) {
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ko = uint64_t(k * 128);           // 16 rows = 2048 B
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          const uint64_t db = db0 + ko;
          mma_f16_ss(tmem_base + 0, a_g0 + ko, db, idesc128, acc);
          mma_f16_ss(tmem_base + 64, a_g1 + ko, db, idesc128, acc);
          mma_f16_ss(tmem_base + 128, a_g2 + ko, db, idesc128, acc);
          mma_f16_ss(tmem_base + 192, a_g3 + ko, db, idesc128, acc);
          mma_f16_ss(tmem_base + 256, a_g4 + ko - (ko << 16), db, idesc128, acc);   // start += 2048 B, LBO -= 2048 B
        }
        mma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(tmem_full_bar);
    __syncwarp();
  } else if (my_tiles > 0) {
    // ===== epilogue: transpose [(tap, ci), co] -> dwp[co][tap * 64 + ci] through shared memory, fp32 reductions =====
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;                // 0..127
    const int r = q * 32 + lane;                          // TMEM lane
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // taps of group g: first half, second half (-1 = none)
    const int tapA[5] = {0, 3, 6, 2, 8};
    const int tapB[5] = {1, 4, 7, 5, -1};
#pragma unroll 1
    for (int g = 0; g < 5; ++g) {
      float v[64];
      tmem_ld_32x32(tmem_base + uint32_t(g * 64) + (uint32_t(q * 32) << 16), v);
      tmem_ld_32x32(tmem_base + uint32_t(g * 64 + 32) + (uint32_t(q * 32) << 16), v + 32);
      tmem_ld_wait();
      // lane r = row r = (tap half, ci); the last group's second half is the bias gradient (all rows equal)
      if (g == 4 && r == 64 && p.db != nullptr) {
#pragma unroll
        for (int c = 0; c < 64; ++c) atomicAdd(p.db + c, v[c]);
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) sStage[r * PITCH + c] = v[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // 128 threads: thread -> (half, 4 consecutive ci), loop over co
      const int half = et >> 6, l4 = (et & 63) >> 2, cosub = et & 3;
      const int tap = half == 0 ? tapA[g] : tapB[g];
      if (tap >= 0) {
#pragma unroll 1
        for (int co = cosub; co < 64; co += 4) {
          const int rb = half * 64 + l4 * 4;
          const float x0 = sStage[(rb + 0) * PITCH + co], x1 = sStage[(rb + 1) * PITCH + co];
          const float x2 = sStage[(rb + 2) * PITCH + co], x3 = sStage[(rb + 3) * PITCH + co];
          float* d4 = p.dwp + co * 576 + tap * 64 + l4 * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// conv_band.cu: halo-reuse variant; returns 1 when it does not apply (geometry / disabled)
int conv_band_try(int mode, const void* act, const void* wp, void* out, const void* relu_src, const float* bias, int relu,
                  int B, int H, int W, int Cin, int Cout, cudaStream_t st);

template <typename K>
static int set_smem(K kern, size_t smem) {
  MASR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  return MASR_OK;
}

}  // namespace masr

using namespace masr;

// y[B,H,W,Cout] = relu?(conv3x3(x[B,H,W,Cin]) + bias); wp [Cout, 9*Cin] bf16 (masr_conv_w_prep layout)
extern "C" int masr_umma_conv3x3_fwd(const void* x, const void* wp, const float* bias, void* y,
                                     int B, int H, int W, int Cin, int Cout, int relu, void* stream) {
  MASR_REQUIRE((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "umma conv: channels must be 64 or 128");
  if (B * H * W == 0) return MASR_OK;
  {
    const int brc = conv_band_try(0, x, wp, y, nullptr, bias, relu, B, H, W, Cin, Cout, as_stream(stream));
    if (brc != 1) return brc;
  }
  const ConvTile t = pick_tile(H, W, 128);
  CUtensorMap ma, mw;
  int rc = act_map(&ma, x, B, H, W, Cin, t.tw, t.th);
  if (rc != MASR_OK) return rc;
  uint64_t wd[2] = {uint64_t(9 * Cin), uint64_t(Cout)};
  uint64_t ws[1] = {uint64_t(9 * Cin) * 2};
  uint32_t wb[2] = {64, uint32_t(Cout)};
  rc = make_tmap_bf16(&mw, wp, 2, wd, ws, wb, true);
  if (rc != MASR_OK) return rc;
  ConvParams p{B, H, W, Cin, Cout, Cin, t.tw, t.th, t.nw, t.nh, static_cast<__nv_bfloat16*>(y), nullptr, bias, relu};
  const unsigned grid = unsigned(B * t.nh * t.nw);
  cudaStream_t st = as_stream(stream);
  constexpr int ST = 4;
  if (Cout == 64) {
    const size_t smem = ST * (16384 + 64 * 128) + 1024 + 256 + 64 * 4;
    rc = set_smem(umma_conv_kernel<64, ST, 0>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_kernel<64, ST, 0>, dim3(grid), dim3(CV_THREADS), smem, st, ma, mw, ma, p));
  } else {
    const size_t smem = ST * (16384 + 128 * 128) + 1024 + 256 + 128 * 4;
    rc = set_smem(umma_conv_kernel<128, ST, 0>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_kernel<128, ST, 0>, dim3(grid), dim3(CV_THREADS), smem, st, ma, mw, ma, p));
  }
  return MASR_OK;
}

// dx[B,H,W,Cin] = conv3x3^T(dy[B,H,W,Cout]) (times (relu_src > 0) when relu_src != NULL)
extern "C" int masr_umma_conv3x3_dgrad(const void* dy, const void* wp, const void* wpt, void* dx, const void* relu_src,
                                       int B, int H, int W, int Cin, int Cout, void* stream) {
  MASR_REQUIRE((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "umma conv: channels must be 64 or 128");
  if (B * H * W == 0) return MASR_OK;
  {
    const int brc = wpt != nullptr ? conv_band_try(2, dy, wpt, dx, relu_src, nullptr, 0, B, H, W, Cin, Cout, as_stream(stream))
                                   : conv_band_try(1, dy, wp, dx, relu_src, nullptr, 0, B, H, W, Cin, Cout, as_stream(stream));
    if (brc != 1) return brc;
  }
  const ConvTile t = pick_tile(H, W, 128);
  CUtensorMap ma, mw;
  int rc = act_map(&ma, dy, B, H, W, Cout, t.tw, t.th);
  if (rc != MASR_OK) return rc;
  uint64_t wd[2] = {uint64_t(9 * Cin), uint64_t(Cout)};
  uint64_t ws[1] = {uint64_t(9 * Cin) * 2};
  uint32_t wb[2] = {64, 64};
  rc = make_tmap_bf16(&mw, wp, 2, wd, ws, wb, true);
  if (rc != MASR_OK) return rc;
  ConvParams p{B, H, W, Cout, Cin, Cin, t.tw, t.th, t.nw, t.nh, static_cast<__nv_bfloat16*>(dx),
               static_cast<const __nv_bfloat16*>(relu_src), nullptr, 0};
  const unsigned grid = unsigned(B * t.nh * t.nw);
  cudaStream_t st = as_stream(stream);
  CUtensorMap mm = ma;                       // ReLU mask tile of the OUTPUT pixels (Cin channels), same box
  if (relu_src != nullptr) {
    rc = act_map(&mm, relu_src, B, H, W, Cin, t.tw, t.th);
    if (rc != MASR_OK) return rc;
  }
  if (Cin == 64) {
    constexpr int ST = 3;                    // 3 x 24 KB ring + 16 KB mask tile: two CTAs per SM
    const size_t smem = ST * (16384 + 64 * 128) + 128 * 64 * 2 + 1024 + 256 + 64 * 4;
    rc = set_smem(umma_conv_kernel<64, ST, 1>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_kernel<64, ST, 1>, dim3(grid), dim3(CV_THREADS), smem, st, ma, mw, mm, p));
  } else {
    constexpr int ST = 4;
    const size_t smem = ST * (16384 + 128 * 128) + 128 * 128 * 2 + 1024 + 256 + 128 * 4;
    rc = set_smem(umma_conv_kernel<128, ST, 1>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_kernel<128, ST, 1>, dim3(grid), dim3(CV_THREADS), smem, st, ma, mw, mm, p));
  }
  return MASR_OK;
}

// dwp[Cout, 9*Cin] (fp32) += sum_pix dy[pix, co] * x[pix+tap, ci]
extern "C" int masr_umma_conv3x3_wgrad(const void* x, const void* dy, float* dwp, float* db,
                                       int B, int H, int W, int Cin, int Cout, void* stream) {
  MASR_REQUIRE((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128), "umma conv: channels must be 64 or 128");
  if (B * H * W == 0) return MASR_OK;
  {
    // row-group variant: whole image rows in a (W + 4) pitch; needs the k-block to fit 128 MMA rows
    const int wp = W + 4;
    const int r = std::max(1, std::min(H, 96 / wp));
    const int kr = (wp * r + 15) / 16 * 16;
    if (kr <= 128 && wp <= 256) {
      CUtensorMap mdy2, mx2;
      uint32_t box[4] = {64, uint32_t(wp), uint32_t(r), 1};
      uint64_t dd[4] = {uint64_t(Cout), uint64_t(W), uint64_t(H), uint64_t(B)};
      uint64_t ds[3] = {uint64_t(Cout) * 2, uint64_t(W) * Cout * 2, uint64_t(H) * W * Cout * 2};
      int rc2 = make_tmap_bf16(&mdy2, dy, 4, dd, ds, box, true);
      if (rc2 != MASR_OK) return rc2;
      uint64_t xd[4] = {uint64_t(Cin), uint64_t(W), uint64_t(H), uint64_t(B)};
      uint64_t xs[3] = {uint64_t(Cin) * 2, uint64_t(W) * Cin * 2, uint64_t(H) * W * Cin * 2};
      rc2 = make_tmap_bf16(&mx2, x, 4, xd, xs, box, true);
      if (rc2 != MASR_OK) return rc2;
      const int tpi = (H + r - 1) / r;
      Wgrad2Params p2{B, H, W, Cin, Cout, wp, r, kr, B * tpi, tpi, dwp, db};
      cudaStream_t st2 = as_stream(stream);
      if (Cout == 64 && Cin == 64) {
        if (wp <= 124) {
          // tap-pair variant: one image row of dY and three of X per k-block
          CUtensorMap mdy3, mx3;
          uint32_t by[4] = {64, uint32_t(wp), 1, 1}, bx[4] = {64, uint32_t(wp), 3, 1};
          rc2 = make_tmap_bf16(&mdy3, dy, 4, dd, ds, by, true);
          if (rc2 != MASR_OK) return rc2;
          rc2 = make_tmap_bf16(&mx3, x, 4, xd, xs, bx, true);
          if (rc2 != MASR_OK) return rc2;
          Wgrad3Params p3{B, H, W, wp, (wp + 15) / 16 * 16, B * H, dwp, db};
          constexpr int ST3 = 3;
          const size_t smem3 = ST3 * (16384 + 48 * 1024) + 2048 + 256 + 1024;
          auto kern3 = umma_conv_wgrad3_kernel<ST3>;
          static bool attr3 = false;
          if (!attr3) { MASR_CHECK_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem3))); attr3 = true; }
          const int gx3 = std::max(1, std::min(p3.ntiles, sm_count()));
          MASR_CHECK_CUDA(launch_pdl(kern3, dim3(unsigned(gx3)), dim3(CV_THREADS), smem3, st2, mdy3, mx3, p3));
          return MASR_OK;
        }
        return launch_wgrad2<64, 64, 6>(mdy2, mx2, p2, st2);
      }
      if (Cout == 128 && Cin == 64) return launch_wgrad2<128, 64, 4>(mdy2, mx2, p2, st2);
      if (Cout == 128 && Cin == 128) return launch_wgrad2<128, 128, 3>(mdy2, mx2, p2, st2);
      if (Cout == 64 && Cin == 128) return launch_wgrad2<64, 128, 4>(mdy2, mx2, p2, st2);
    }
  }
  const ConvTile t = pick_tile(H, W, 64);
  CUtensorMap mdy, mx;
  int rc = act_map(&mdy, dy, B, H, W, Cout, t.tw, t.th);
  if (rc != MASR_OK) return rc;
  rc = act_map(&mx, x, B, H, W, Cin, t.tw, t.th);
  if (rc != MASR_OK) return rc;
  WgradParams p{B, H, W, Cout, Cin, t.tw, t.th, t.nw, t.nh, dwp, db};
  const int ntiles = B * t.nh * t.nw;
  const int gx = std::max(1, std::min(ntiles, sm_count() / 3));
  dim3 grid(unsigned(gx), 3, 1);
  cudaStream_t st = as_stream(stream);
  if (Cin == 64) {
    constexpr int ST = 4;
    const size_t smem = ST * (16384 + 3 * 8192) + 2048 + 1024 + 256;
    rc = set_smem(umma_conv_wgrad_kernel<64, ST>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_wgrad_kernel<64, ST>, grid, dim3(CV_THREADS), smem, st, mdy, mx, p));
  } else {
    constexpr int ST = 3;
    const size_t smem = ST * (16384 + 3 * 16384) + 2048 + 1024 + 256;
    rc = set_smem(umma_conv_wgrad_kernel<128, ST>, smem); if (rc) return rc;
    MASR_CHECK_CUDA(launch_pdl(umma_conv_wgrad_kernel<128, ST>, grid, dim3(CV_THREADS), smem, st, mdy, mx, p));
  }
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}
