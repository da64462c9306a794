// Lean epilogue of the tcgen05 GEMM kernels: TMEM accumulator -> registers -> (bias, ReLU, dropout, backward mask, row
// dots, += old value) -> 128B-swizzled shared-memory box -> ONE bulk tensor store (or fp32 bulk reduce-add for split-K)
// per 32-row x 128-byte box, issued by one lane of the warp.
//
// Why (profiles/r2_pair_gemm_epilogue.md): on the K = 512 problems of the encoder the staged epilogue of epilogue.cuh was
// the bottleneck of the whole GEMM -- ~780 executed instructions per warp and 128 x 128 tile spread over ~6000 SASS lines
// of mode branches (19 % of the stall samples were instruction-cache misses), a second pass over shared memory with
// 64-bit row pointers travelling by shuffle, per-thread predicated 16-byte global stores; the MMA warp sat waiting for the
// epilogue to hand the TMEM stage back.  Here a thread touches its accumulator row once, the TMA engine clips partial
// tiles (rows >= M, columns >= N) and writes full 128-byte lines.
//
// Thread <-> data: warp q of the 4 epilogue warps owns TMEM lanes 32 q .. 32 q + 31 = output rows m_warp0 + lane.
// Box: 32 rows x 128 bytes (64 bf16 / 32 fp32 columns), row r at box + r * 128, 16-byte chunk j at ((j ^ (r & 7)) << 4)
// (= CU_TENSOR_MAP_SWIZZLE_128B; conflict-free 128-bit shared stores).  Two boxes per warp alternate.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace masr {

constexpr int EPT_BOX_BYTES = 32 * 128;                 // one box
constexpr int EPT_WARP_BYTES = 2 * EPT_BOX_BYTES;       // double-buffered per warp
constexpr int EPT_CTA_BYTES = 4 * EPT_WARP_BYTES;       // 32 KB for the 4 epilogue warps; 1024 B aligned base required

namespace umma {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups still READING shared memory
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace umma

// Side inputs / fused element-wise work of one row (all optional)
struct EptOpts {
  const float* sbias = nullptr;              // bias of the piece's columns in shared memory (zero beyond N)
  bool relu = false;
  float p_drop = 0.f, inv_keep = 1.f;        // forward dropout, element index drop_row_base + column-in-piece
  uint64_t seed = 0; uint32_t site = 0; int64_t drop_row_base = 0;
  const __nv_bfloat16* mask_row = nullptr;   // backward mask row (this thread's row, piece's first column); see epilogue.cuh
  float mask_scale = 1.f;
  const __nv_bfloat16* old_row = nullptr;    // bf16 C row to accumulate onto (MASR_GEMM_ACCUM)
  const __nv_bfloat16* dot_row = nullptr;    // per-64-column row dots with a second matrix (attention backward's D)
  float* dot_out = nullptr; int dot_stride = 0;
};

// bf16 output: piece of 32 rows (this warp) x ncols64 * 64 columns; `col0` = global column of the piece, `row0` = global
// row of the warp's first lane, ncols_valid = N - col0 (columns beyond are clipped by the TMA; side inputs are not read there)
// boxsel: the warp's box toggle, carried across calls (a box is reused only after the bulk store issued two groups ago
// has finished reading it: lane 0 waits until at most ONE of its bulk groups is still reading).
template <bool EXTRAS>
__device__ __forceinline__ void epilogue_tma_bf16(const CUtensorMap* map_c, uint32_t tmem_acc, int q, int lane, int row0, int col0,
                                                  int ngroups, int ncols_valid, bool row_valid, unsigned char* wstage,
                                                  int& boxsel, const EptOpts& o) {
  using namespace umma;
#pragma unroll 1
  for (int g = 0; g < ngroups; ++g) {
    if (g * 64 >= ncols_valid) break;                                     // whole box beyond N
    unsigned char* box = wstage + boxsel * EPT_BOX_BYTES;
    boxsel ^= 1;
    float v[64];
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(g * 64), v);
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(g * 64 + 32), v + 32);
    // the box written two groups ago must have been read by its bulk store (lane 0 owns the bulk groups)
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    tmem_ld_wait();
    if (o.sbias != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(o.sbias + g * 64 + 4 * j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    if (o.relu) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr (EXTRAS) {
      const int nv = min(64, ncols_valid - g * 64);                       // valid columns of this group (multiple of 8 required)
      if (o.p_drop > 0.f) {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          v[j] *= drop_scale(o.p_drop, o.inv_keep, o.seed, o.site, uint64_t(o.drop_row_base + g * 64 + j));
      }
      if (o.dot_row != nullptr && row_valid) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          if (j < nv) {
            float sv[8];
            load8<__nv_bfloat16>(o.dot_row + g * 64 + j, sv);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(v[j + e], sv[e], acc);
          }
        }
        o.dot_out[g * o.dot_stride] = acc;
      }
      if (o.mask_row != nullptr && row_valid) {
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          if (j < nv) {
            float mk[8];
            load8<__nv_bfloat16>(o.mask_row + g * 64 + j, mk);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j + e] = mk[e] > 0.f ? v[j + e] * o.mask_scale : 0.f;
          }
        }
      }
      if (o.old_row != nullptr && row_valid) {
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          if (j < nv) {
            float ov[8];
            load8<__nv_bfloat16>(o.old_row + g * 64 + j, ov);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j + e] += ov[e];
          }
        }
      }
    }
    unsigned char* rowp = box + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 pk;
      __nv_bfloat162 a0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), a1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
      __nv_bfloat162 a2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), a3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
      pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
      *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pk;
    }
    fence_proxy_async();              // generic-proxy writes of this thread -> visible to the bulk (async-proxy) read
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map_c, box, col0 + g * 64, row0);
      bulk_commit();
    }
  }
}

// fp32 output: plain store or reduce-add (split-K) of 32-column boxes; bias optional (first k-slice only: caller)
template <bool ADD>
__device__ __forceinline__ void epilogue_tma_f32(const CUtensorMap* map_c, uint32_t tmem_acc, int q, int lane, int row0, int col0,
                                                 int ngroups, int ncols_valid, unsigned char* wstage, int& boxsel, const EptOpts& o) {
  using namespace umma;
#pragma unroll 1
  for (int g = 0; g < ngroups; ++g) {
    if (g * 32 >= ncols_valid) break;
    unsigned char* box = wstage + boxsel * EPT_BOX_BYTES;
    boxsel ^= 1;
    float v[32];
    tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(g * 32), v);
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    tmem_ld_wait();
    if (o.sbias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(o.sbias + g * 32 + 4 * j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
    if (o.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    unsigned char* rowp = box + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (ADD) tma_reduce_add_2d(map_c, box, col0 + g * 32, row0);
      else tma_store_2d(map_c, box, col0 + g * 32, row0);
      bulk_commit();
    }
  }
}

// all bulk stores of this thread have been written (call by lane 0 before the CTA exits / before the output is consumed
// by a later phase of the same kernel)
__device__ __forceinline__ void epilogue_tma_drain(int lane) {
  if (lane == 0) umma::bulk_wait<0>();
  __syncwarp();
}

}  // namespace masr
