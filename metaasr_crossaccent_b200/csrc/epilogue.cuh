// Shared epilogue of the tcgen05 kernels: TMEM accumulator tile [128 rows x BN fp32] -> registers ->
// (bias, ReLU, dropout) -> shared-memory staging tile -> (ReLU/dropout-backward mask) -> COALESCED global
// stores / accumulates / fp32 vector reductions.
//
// Phase 1: thread r (= TMEM lane r) converts its row and parks it in the staging tile (rows padded by 16 B:
//          conflict-free 128-bit shared stores) together with the row's global destination pointer.
// Phase 2: each warp drains its own 32 rows; consecutive lanes write consecutive 16-byte chunks of a row,
//          so a store instruction covers whole 128 B lines (the naive "one thread = one row" epilogue
//          touched 32 different lines per instruction and fetched the bias with 128 scalar global loads).
// The staging tile aliases the (by then idle) TMA pipeline stages.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace masr {

enum EpiMode { EPI_STORE = 0, EPI_ACCUM = 1, EPI_ATOMIC = 2 };

template <int BN, typename OutT>
struct EpiLayout {
  static constexpr int ROWB = BN * int(sizeof(OutT)) + 16;          // padded staging row (bytes)
  static constexpr int PTR_OFF = 128 * ROWB;                        // then 128 x {dst, mask} pointers
  static constexpr int BYTES = PTR_OFF + 128 * 16;
};

// Optional fused element-wise work of one tile
struct EpiOpts {
  const float* sbias = nullptr;          // BN floats in shared memory
  bool relu = false;                     // max(x, 0) after the bias
  // forward dropout applied after bias / ReLU: element (row, col) of the tile has the linear index
  // drop_row_base + col  (same index as masr_dropout on the contiguous [M, N] tensor)
  float p_drop = 0.f, inv_keep = 1.f;
  uint64_t seed = 0; uint32_t site = 0;
  int64_t drop_row_base = 0;
  // backward mask: output is zeroed where mask <= 0 and multiplied by mask_scale elsewhere.  With
  // mask = the forward output of ReLU(+dropout) and mask_scale = 1/(1-p) this IS the backward of both
  // (an element is positive iff it passed the ReLU and was kept)
  const __nv_bfloat16* mask_row = nullptr;     // global, this thread's row (bf16 outputs only)
  const unsigned char* smask = nullptr;        // or: mask tile resident in shared memory ([128 x 64] halves of
                                               // 16 KB in the TMA 128B-swizzled layout, row = TMEM lane)
  float mask_scale = 1.f;
  // or: the mask row lives at a constant byte offset from the destination row (same [pixels, channels] layout):
  // no second pointer per row has to travel through the warp
  bool mask_at_dst_delta = false;
  long long mask_delta = 0;
  // second accumulator (TMEM column offset from the first, 0 = none) added in phase 1: two MMA-issuing warps each
  // own one accumulator and half of the k-blocks
  uint32_t acc2_offset = 0;
  // per-64-column row dot products with a second bf16 matrix (same row, same columns): dot_out[j * dot_stride] =
  // sum over columns [64 j, 64 j + 64) of acc * dot_row.  The attention backward needs D = rowsum(dO . O) per head;
  // dO is the output of the out-projection dgrad GEMM, so its epilogue produces D for free (no extra kernel)
  const __nv_bfloat16* dot_row = nullptr;
  float* dot_out = nullptr;
  int dot_stride = 0;
};

// packed variant for scale == 1: v * (mask > 0 ? 1 : 0) on bf16x2 words (no fp32 round trip)
__device__ __forceinline__ uint4 apply_mask8_packed(const uint4& val, const uint4& mraw) {
  uint4 out;
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
  const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&val);
  const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mraw);
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
  for (int e = 0; e < 4; ++e) o2[e] = __hmul2(v2[e], __hgt2(m2[e], zero));
  return out;
}
__device__ __forceinline__ void apply_mask8(float* f, const uint4& mraw, float scale) {
  float mk[8];
  load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(&mraw), mk);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = (mk[e] > 0.f) ? f[e] * scale : 0.f;
}

// stage      : >= EpiLayout::BYTES bytes of shared memory, 16 B aligned, private to the 4 epilogue warps
// dst_row    : global pointer of this thread's output row (BN contiguous OutT), nullptr = row not stored
// ncols      : number of valid columns (<= BN); vec_ok: rows are 16 B aligned and ncols == BN
template <int BN, typename OutT>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, int q, int lane, unsigned char* stage, OutT* dst_row,
                                              int ncols, bool vec_ok, int mode, const EpiOpts& o) {
  using L = EpiLayout<BN, OutT>;
  const int r = q * 32 + lane;
  unsigned char* my = stage + r * L::ROWB;
  __syncwarp();                       // a previous call's phase 2 (same warp, same rows) has finished reading
  // ---- phase 1 (two 32-column TMEM loads in flight; the bias chunk is fetched while they complete)
  constexpr int CB = BN >= 64 ? 64 : 32;                // columns per batch
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += CB) {
    float v[CB];
#pragma unroll
    for (int h = 0; h < CB; h += 32) umma::tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c0 + h), v + h);
    float4 b4[CB / 4];
    if (o.sbias != nullptr) {
#pragma unroll
      for (int j = 0; j < CB / 4; ++j) b4[j] = *reinterpret_cast<const float4*>(o.sbias + c0 + 4 * j);
    }
    umma::tmem_ld_wait();
    if (o.acc2_offset != 0) {
      float v2[CB];
#pragma unroll
      for (int h = 0; h < CB; h += 32)
        umma::tmem_ld_32x32(tmem_acc + o.acc2_offset + (uint32_t(q * 32) << 16) + uint32_t(c0 + h), v2 + h);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CB; ++j) v[j] += v2[j];
    }
    if (o.sbias != nullptr) {
#pragma unroll
      for (int j = 0; j < CB / 4; ++j) { v[4 * j] += b4[j].x; v[4 * j + 1] += b4[j].y; v[4 * j + 2] += b4[j].z; v[4 * j + 3] += b4[j].w; }
    }
    if (o.relu) {
#pragma unroll
      for (int j = 0; j < CB; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (o.p_drop > 0.f) {
#pragma unroll
      for (int j = 0; j < CB; ++j)
        v[j] *= drop_scale(o.p_drop, o.inv_keep, o.seed, o.site, uint64_t(o.drop_row_base + c0 + j));
    }
    if constexpr (CB == 64) {
      if (o.dot_row != nullptr) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          float sv[8];
          load8<__nv_bfloat16>(o.dot_row + c0 + j, sv);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(v[j + e], sv[e], acc);
        }
        o.dot_out[(c0 >> 6) * o.dot_stride] = acc;
      }
    }
    if constexpr (sizeof(OutT) == 2) {
#pragma unroll
      for (int j = 0; j < CB; j += 8) store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(my) + c0 + j, v + j);
    } else {
#pragma unroll
      for (int j = 0; j < CB; j += 4)
        *reinterpret_cast<float4*>(my + (c0 + j) * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  void** ptrs = reinterpret_cast<void**>(stage + L::PTR_OFF);
  if (!(sizeof(OutT) == 2 && vec_ok)) {            // the bf16 vector path passes row pointers by shuffle
    ptrs[2 * r] = dst_row;
    ptrs[2 * r + 1] = const_cast<__nv_bfloat16*>(o.mask_row);
  }
  __syncwarp();
  // ---- phase 2: this warp's rows q*32 .. q*32+31
  if (vec_ok) {
    constexpr int CPR = BN * int(sizeof(OutT)) / 16;               // 16 B chunks per row
    constexpr int RPI = 32 / CPR < 1 ? 1 : 32 / CPR;               // rows per iteration
    constexpr int IPR = CPR > 32 ? CPR / 32 : 1;                   // iterations per row (BN*size > 512 B)
    const int sub = lane / (CPR < 32 ? CPR : 32);
    const int ch0 = lane % (CPR < 32 ? CPR : 32);
    if constexpr (sizeof(OutT) == 2) {
      // Row pointers travel by warp shuffle (the row's owner is lane `row % 32` of this warp), all shared-memory
      // and mask loads of the warp's 32 rows are issued before the first store: no dependent round trips.
      constexpr int NIT = 32 / RPI;                                // 8 (BN = 64) or 16 (BN = 128)
      const unsigned long long my_dst = reinterpret_cast<unsigned long long>(dst_row);
      const unsigned long long my_msk = reinterpret_cast<unsigned long long>(o.mask_row);
      const bool any_gmask = __any_sync(0xffffffffu, o.mask_row != nullptr);
      uint4 val[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        val[i] = *reinterpret_cast<const uint4*>(stage + (q * 32 + i * RPI + sub) * L::ROWB + ch0 * 16);
      if (o.mask_at_dst_delta && mode == EPI_STORE) {
        unsigned long long dd[NIT];
        uint4 mraw[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          dd[i] = __shfl_sync(0xffffffffu, my_dst, i * RPI + sub);
          mraw[i] = make_uint4(0, 0, 0, 0);
          if (dd[i] != 0ull) mraw[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(dd[i]) + o.mask_delta + ch0 * 16);
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          if (dd[i] == 0ull) continue;
          if (o.mask_scale == 1.f) {
            *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dd[i]) + ch0 * 16) = apply_mask8_packed(val[i], mraw[i]);
          } else {
            float f[8];
            load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(&val[i]), f);
            apply_mask8(f, mraw[i], o.mask_scale);
            store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(dd[i]) + ch0 * 8, f);
          }
        }
      } else if (o.smask == nullptr && !any_gmask && mode == EPI_STORE) {
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const unsigned long long d = __shfl_sync(0xffffffffu, my_dst, i * RPI + sub);
          if (d != 0ull) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(d) + ch0 * 16) = val[i];
        }
      } else {
        constexpr int UNR = NIT >= 8 ? 8 : NIT;                    // mask loads in flight per batch
#pragma unroll
        for (int i0 = 0; i0 < NIT; i0 += UNR) {
          uint4 mraw[UNR];
          unsigned long long dd[UNR];
          bool mk[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int lrow = (i0 + u) * RPI + sub;                 // row within this warp's 32
            const int row = q * 32 + lrow;
            dd[u] = __shfl_sync(0xffffffffu, my_dst, lrow);
            const unsigned long long gm = any_gmask ? __shfl_sync(0xffffffffu, my_msk, lrow) : 0ull;
            mraw[u] = make_uint4(0, 0, 0, 0);
            mk[u] = false;
            if (dd[u] == 0ull) continue;
            if (o.smask != nullptr) {
              mraw[u] = *reinterpret_cast<const uint4*>(o.smask + (ch0 >> 3) * 16384 + row * 128 + (((ch0 & 7) ^ (row & 7)) << 4));
              mk[u] = true;
            } else if (gm != 0ull) {
              mraw[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(gm) + ch0 * 8);
              mk[u] = true;
            }
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            if (dd[u] == 0ull) continue;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(dd[u]);
            float f[8];
            load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(&val[i0 + u]), f);
            if (mk[u]) apply_mask8(f, mraw[u], o.mask_scale);
            if (mode == EPI_ACCUM) {
              float old[8];
              load8<__nv_bfloat16>(dst + ch0 * 8, old);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += old[e];
            }
            store8<__nv_bfloat16>(dst + ch0 * 8, f);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += RPI) {
        const int row = q * 32 + rr + sub;
        OutT* dst = static_cast<OutT*>(ptrs[2 * row]);
        if (dst == nullptr) continue;
#pragma unroll
        for (int it = 0; it < IPR; ++it) {
          const int ch = ch0 + it * 32;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + row * L::ROWB + ch * 16);
          float* d4 = reinterpret_cast<float*>(dst) + ch * 4;
          const float4 f = *reinterpret_cast<const float4*>(&val);
          if (mode == EPI_ATOMIC) {
            // one 128-bit reduction (REDG.E.ADD.F32x4) instead of four scalar atomics
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
          } else if (mode == EPI_ACCUM) {
            const float4 old = *reinterpret_cast<const float4*>(d4);
            *reinterpret_cast<float4*>(d4) = make_float4(f.x + old.x, f.y + old.y, f.z + old.z, f.w + old.w);
          } else {
            *reinterpret_cast<float4*>(d4) = f;
          }
        }
      }
    }
  } else {
    // generic path: lanes sweep the row element-wise (still coalesced), any alignment / partial width
#pragma unroll 1
    for (int rr = 0; rr < 32; ++rr) {
      const int row = q * 32 + rr;
      OutT* dst = static_cast<OutT*>(ptrs[2 * row]);
      const __nv_bfloat16* msk = static_cast<const __nv_bfloat16*>(ptrs[2 * row + 1]);
      if (dst == nullptr) continue;
      const OutT* src = reinterpret_cast<const OutT*>(stage + row * L::ROWB);
      for (int c = lane; c < ncols; c += 32) {
        float f = to_f<OutT>(src[c]);
        if (msk != nullptr) f = (__bfloat162float(msk[c]) > 0.f) ? f * o.mask_scale : 0.f;
        if constexpr (sizeof(OutT) == 4) {
          if (mode == EPI_ATOMIC) { atomicAdd(reinterpret_cast<float*>(dst) + c, f); continue; }
        }
        if (mode == EPI_ACCUM) f += to_f<OutT>(dst[c]);
        dst[c] = from_f<OutT>(f);
      }
    }
  }
}

}  // namespace masr
