// First VGG convolution (Cin = 1 -> 64, 3x3, pad 1; mono_transformer_torch.py:50) on the tensor cores.
//
// With Cin = 1 the layer is a [pixels x 9] x [9 x 64] product: far too thin to be compute-bound, but the
// CUDA-core kernels of conv.cu spend ~110 instructions per 16 output bytes (forward, 103 us) or re-read dy
// through the LSU (weight gradient, 220 us) for tensors that only take ~30 us to stream at HBM speed.  Here the
// 3x3 patches (9 taps, zero-padded to K = 16) are built in shared memory by CUDA threads as a regular K-major
// 128B-swizzled UMMA operand and the tensor core does the arithmetic:
//   forward : D[128 pixels, 64 co]  = patches[128, 16] . W[64, 16]^T            (one tcgen05.mma per tile)
//   wgrad   : D[64 co, 16]         += dY[pixels, 64 co]^T . patchesT[16, pixels]  (dY streamed by TMA as an MN-major
//             operand straight from its NHWC layout, K = pixels; tap row 9 of patchesT is all ones, so
//             column 9 of D is the bias gradient)
// x is rounded to bf16 on the way into the patches (the rest of the bf16 network sees bf16 activations anyway).
#include "common.cuh"
#include "umma.cuh"
#include "epilogue.cuh"
#include "epilogue_tma.cuh"

namespace masr {

// byte offset of K-element j (0..63) of row t inside a K-major 128B-swizzled tile (rows of 128 B, 8-row atoms)
__device__ __forceinline__ uint32_t kmajor_off(int t, int j) {
  return uint32_t(t >> 3) * 1024u + uint32_t(t & 7) * 128u + (uint32_t((j >> 3) ^ (t & 7)) << 4) + uint32_t(j & 7) * 2u;
}

// Position of a builder thread's pixel, advanced by a constant stride from tile to tile without divisions (the 64-bit
// div / mod of the flattened index was ~100 instructions per tile, and the tap loads that depended on it sat on the
// thread's critical path).
struct PixCursor {
  int64_t pix, row;        // flattened pixel, image row b * H + h
  int h, w;
  int d_w, d_h; int64_t d_row, d_pix;
  __device__ __forceinline__ void init(int64_t p0, int64_t stride, int H, int W) {
    pix = p0; row = p0 / W; w = int(p0 - row * W); h = int(row % H);
    d_pix = stride; d_row = stride / W; d_w = int(stride - d_row * W); d_h = int(d_row % H);
  }
  __device__ __forceinline__ void advance(int H, int W) {
    pix += d_pix; row += d_row; w += d_w; h += d_h;
    if (w >= W) { w -= W; ++row; ++h; }
    if (h >= H) h -= H;
    if (h >= H) h -= H;
  }
};

// 3x3 neighbourhood of the cursor's pixel of x [B, H, W] fp32, zero outside the image / past the last pixel
__device__ __forceinline__ void load_taps(const float* __restrict__ x, const PixCursor& c, int64_t P, int H, int W, float* tap) {
#pragma unroll
  for (int t = 0; t < 9; ++t) tap[t] = 0.f;
  if (c.pix >= P) return;
#pragma unroll
  for (int dh = -1; dh <= 1; ++dh) {
    const int hh = c.h + dh;
    if (hh < 0 || hh >= H) continue;
    const float* xr = x + (c.row + dh) * W;
#pragma unroll
    for (int dw = -1; dw <= 1; ++dw) {
      const int ww = c.w + dw;
      if (ww >= 0 && ww < W) tap[(dh + 1) * 3 + (dw + 1)] = __ldg(xr + ww);
    }
  }
}

// ================================================================================================ forward
constexpr int C1F_THREADS = 320;          // warp 0 -, warp 1 MMA, warps 2-5 patch builders, warps 6-9 epilogue
constexpr int C1F_STAGES = 3;
constexpr int C1_AHEAD = 2;              // tiles of input taps a builder thread keeps in flight

struct C1FwdParams {
  const float* x; const float* w; const float* bias; __nv_bfloat16* y;
  int H, W; int64_t P; int ntiles;
};

__global__ void __launch_bounds__(C1F_THREADS, 2)
conv1_fwd_umma_kernel(const __grid_constant__ CUtensorMap map_y, C1FwdParams p) {
  using namespace umma;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                                      // C1F_STAGES x [128 rows x 128 B] (K = 16 used)
  unsigned char* sW = sA + C1F_STAGES * 16384;                   // [64 co x 128 B]
  unsigned char* sStage = sW + 8192;                             // epilogue boxes (epilogue_tma.cuh), 1024 B aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + EPT_CTA_BYTES);
  uint64_t* a_full = bars;                    // [STAGES] count 4 (builder warps)
  uint64_t* a_empty = bars + C1F_STAGES;      // [STAGES] count 1 (MMA commit)
  uint64_t* t_full = bars + 2 * C1F_STAGES;   // [2]
  uint64_t* t_empty = t_full + 2;             // [2] count 4 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* sbias = reinterpret_cast<float*>(t_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_y);
    for (int s = 0; s < C1F_STAGES; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  pdl_wait();
  // weights [64, 9] fp32 -> K-major bf16 tile (K padded to 16 with zeros); bias -> shared memory
  if (threadIdx.x < 64) {
    const int co = threadIdx.x;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float a = j < 9 ? p.w[co * 9 + j] : 0.f, b = j + 1 < 9 ? p.w[co * 9 + j + 1] : 0.f;
      __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
      pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(sW + kmajor_off(co, 0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(sW + kmajor_off(co, 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    sbias[co] = p.bias[co];
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first = int(blockIdx.x), step = int(gridDim.x);

  if (warp >= 2 && warp <= 5) {
    // ===== patch builders: thread = pixel row of the tile; the taps of the next two tiles are in flight while the
    // current one is packed (the loads were the builders' -- and the kernel's -- critical path) =====
    const int r = (warp - 2) * 32 + lane;
    int s = 0; uint32_t ph = 0;
    PixCursor cur;
    cur.init(int64_t(first) * 128 + r, int64_t(step) * 128, p.H, p.W);
    // C1_AHEAD tiles of taps in flight per thread, in a register ring indexed statically (the loop is unrolled by the
    // ring size: a register-to-register hand-over would wait on the pending loads).  x is re-fetched from HBM under the
    // streaming output (5 MB of input against 174 MB written), so the load latency is microseconds, not an L2 hit.
    float tp[C1_AHEAD][9];
#pragma unroll
    for (int d = 0; d < C1_AHEAD; ++d) { load_taps(p.x, cur, p.P, p.H, p.W, tp[d]); cur.advance(p.H, p.W); }
    for (int tile = first; tile < p.ntiles; tile += step * C1_AHEAD) {
#pragma unroll
      for (int d = 0; d < C1_AHEAD; ++d) {
        if (tile + d * step >= p.ntiles) break;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(j < 9 ? tp[d][j] : 0.f, j + 1 < 9 ? tp[d][j + 1] : 0.f);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        mbar_wait(&a_empty[s], ph ^ 1);
        unsigned char* a = sA + s * 16384;
        *reinterpret_cast<uint4*>(a + kmajor_off(r, 0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(a + kmajor_off(r, 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
        if (++s == C1F_STAGES) { s = 0; ph ^= 1; }
        // C1_AHEAD tiles ahead (zeros past the last pixel); issued behind the fences above, so that the pack of the next
        // slot never shares a scoreboard with these younger loads
        load_taps(p.x, cur, p.P, p.H, p.W, tp[d]);
        cur.advance(p.H, p.W);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane issues) =====
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint64_t dw = desc_kmajor_sw128(smem_u32(sW));
    int s = 0; uint32_t ph = 0;
    int it = 0;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const int ab = it & 1;
      mbar_wait(&t_empty[ab], ((it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
        mma_f16_ss(tmem_base + uint32_t(ab * 64), desc_kmajor_sw128(smem_u32(sA + s * 16384)), dw, idesc, 0u);
        mma_commit(&a_empty[s]);
        mma_commit(&t_full[ab]);
      }
      __syncwarp();
      if (++s == C1F_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 6) {
    // ===== epilogue: bias + ReLU -> bf16 -> swizzled box -> one bulk tensor store per warp and tile (rows past the last
    // pixel are clipped by the TMA) =====
    const int q = warp & 3;
    int it = 0, boxsel = 0;
    EptOpts o;
    o.sbias = sbias;
    o.relu = true;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const int ab = it & 1;
      mbar_wait(&t_full[ab], (it >> 1) & 1);
      tc_fence_after();
      epilogue_tma_bf16<false>(&map_y, tmem_base + uint32_t(ab * 64), q, lane, tile * 128 + q * 32, 0, 1, 64, true,
                               sStage + q * EPT_WARP_BYTES, boxsel, o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[ab]);
    }
    epilogue_tma_drain(lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ================================================================================================ wgrad
constexpr int C1W_THREADS = 192;          // warp 0 TMA (dY), warp 1 MMA, warps 2-5 patch builders + final epilogue
constexpr int C1W_STAGES = 5;
constexpr uint32_t C1W_A = 128 * 128;     // dY tile: 128 pixels x 64 co bf16
constexpr uint32_t C1W_B = 2 * 2048;      // patchesT: two k-blocks of [16 taps x 64 pixels]
constexpr uint32_t C1W_STAGE = C1W_A + C1W_B;

struct C1WgradParams {
  const float* x; float* dw; float* db;
  int H, W; int64_t P; int nblk;
};

__global__ void __launch_bounds__(C1W_THREADS, 2)
conv1_wgrad_umma_kernel(const __grid_constant__ CUtensorMap map_dy, C1WgradParams p) {
  using namespace umma;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C1W_STAGES * C1W_STAGE);   // count 1 (TMA) + 4 (builder warps)
  uint64_t* empty_bar = full_bar + C1W_STAGES;                                        // count 1 (MMA commit)
  uint64_t* tmem_full_bar = empty_bar + C1W_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = int(blockIdx.x), step = int(gridDim.x);
  const int my_blocks = first < p.nblk ? (p.nblk - first + step - 1) / step : 0;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_dy);
    for (int s = 0; s < C1W_STAGES; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===== dY producer: box {64 co, 128 pixels} = the MN-major A operand (K = pixels) as it lies in memory =====
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < my_blocks; ++i) {
      const int blk = first + i * step;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full_bar[s], C1W_A);
        tma_load_2d(smem + s * C1W_STAGE, &map_dy, &full_bar[s], 0, blk * 128);
      }
      __syncwarp();
      if (++s == C1W_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: 8 x (M = 64 co, N = 16, K = 16 pixels) per stage =====
    constexpr uint32_t idesc = make_idesc_bf16(64, 16, 1, 0);
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < my_blocks; ++i) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * C1W_STAGE), sb = sa + C1W_A;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_f16_ss(tmem_base, desc_mnmajor_sw128(sa + k * 2048, 8192), desc_kmajor_sw128(sb + (k >> 2) * 2048 + (k & 3) * 32),
                     idesc, (i > 0 || k > 0) ? 1u : 0u);
        mma_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == C1W_STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) mma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ===== patch builders: thread = pixel; writes column `pixel` of the [16 taps x 64 pixels] K-major tiles =====
    const int r = (warp - 2) * 32 + lane;            // pixel within the 128-pixel block
    const int kbk = r >> 6, j = r & 63;
    int s = 0; uint32_t ph = 0;
    PixCursor cur;
    cur.init(int64_t(first) * 128 + r, int64_t(step) * 128, p.H, p.W);
    float tp[C1_AHEAD][9];                             // taps of the next C1_AHEAD blocks (see the forward kernel)
    bool tv[C1_AHEAD];
#pragma unroll
    for (int d = 0; d < C1_AHEAD; ++d) { tv[d] = cur.pix < p.P; load_taps(p.x, cur, p.P, p.H, p.W, tp[d]); cur.advance(p.H, p.W); }
    for (int i = 0; i < my_blocks; i += C1_AHEAD) {
#pragma unroll
      for (int d = 0; d < C1_AHEAD; ++d) {
        if (i + d >= my_blocks) break;
        __nv_bfloat16 bv[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) bv[t] = __float2bfloat16_rn(t < 9 ? tp[d][t] : ((t == 9 && tv[d]) ? 1.f : 0.f));   // row 9 = ones: bias gradient
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* sb = smem + s * C1W_STAGE + C1W_A + kbk * 2048;
#pragma unroll
        for (int t = 0; t < 16; ++t) *reinterpret_cast<__nv_bfloat16*>(sb + kmajor_off(t, j)) = bv[t];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (++s == C1W_STAGES) { s = 0; ph ^= 1; }
        tv[d] = cur.pix < p.P;
        load_taps(p.x, cur, p.P, p.H, p.W, tp[d]);
        cur.advance(p.H, p.W);
      }
    }
    // ===== final epilogue: an M = 64 accumulator occupies 16 lanes of each 32-lane TMEM quadrant =====
    if (my_blocks > 0) {
      const int q = warp & 3;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16), v);
      tmem_ld_wait();
      if (lane < 16) {
        const int co = q * 16 + lane;
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(p.dw + co * 9 + t, v[t]);
        atomicAdd(p.db + co, v[9]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

}  // namespace masr

using namespace masr;

// y[B,H,W,64] (bf16, NHWC) = relu(conv3x3(x[B,H,W] fp32, w[64,1,3,3] fp32) + bias)
extern "C" int masr_umma_conv1_fwd(const float* x, const float* w, const float* bias, void* y,
                                   int B, int H, int W, int Cout, void* stream) {
  MASR_REQUIRE(Cout == 64, "umma conv1: Cout must be 64");
  MASR_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "umma conv1: y must be 16 B aligned");
  const int64_t P = int64_t(B) * H * W;
  if (P == 0) return MASR_OK;
  MASR_REQUIRE(P < (int64_t(1) << 31) - 256, "umma conv1: too many pixels");
  C1FwdParams p{x, w, bias, static_cast<__nv_bfloat16*>(y), H, W, P, int(ceil_div64(P, 128))};
  CUtensorMap my;
  {
    uint64_t dims[2] = {64, uint64_t(P)};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, 32};
    const int rc = make_tmap_bf16(&my, y, 2, dims, strides, box, true);
    if (rc != MASR_OK) return rc;
  }
  const size_t smem = C1F_STAGES * 16384 + 8192 + EPT_CTA_BYTES + 256 + 64 * 4 + 1024;
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(conv1_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  const unsigned grid = unsigned(std::min(p.ntiles, 2 * sm_count()));
  MASR_CHECK_CUDA(launch_pdl(conv1_fwd_umma_kernel, dim3(grid), dim3(C1F_THREADS), smem, as_stream(stream), my, p));
  return MASR_OK;
}

// dw[64,1,3,3] += sum_pix dy[pix, co] x[pix + tap];  db[64] += sum_pix dy[pix, co]   (dy bf16 NHWC)
extern "C" int masr_umma_conv1_wgrad(const float* x, const void* dy, float* dw, float* db,
                                     int B, int H, int W, int Cout, void* stream) {
  MASR_REQUIRE(Cout == 64, "umma conv1: Cout must be 64");
  const int64_t P = int64_t(B) * H * W;
  if (P == 0) return MASR_OK;
  MASR_REQUIRE(P < (int64_t(1) << 31), "umma conv1 wgrad: too many pixels");
  CUtensorMap mdy;
  uint64_t dims[2] = {64, uint64_t(P)};
  uint64_t strides[1] = {128};
  uint32_t box[2] = {64, 128};
  int rc = make_tmap_bf16(&mdy, dy, 2, dims, strides, box, true);
  if (rc != MASR_OK) return rc;
  C1WgradParams p{x, dw, db, H, W, P, int(ceil_div64(P, 128))};
  const size_t smem = C1W_STAGES * C1W_STAGE + 256 + 1024;
  static bool attr = false;
  if (!attr) { MASR_CHECK_CUDA(cudaFuncSetAttribute(conv1_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr = true; }
  const unsigned grid = unsigned(std::min(p.nblk, 2 * sm_count()));
  MASR_CHECK_CUDA(launch_pdl(conv1_wgrad_umma_kernel, dim3(grid), dim3(C1W_THREADS), smem, as_stream(stream), mdy, p));
  return MASR_OK;
}
