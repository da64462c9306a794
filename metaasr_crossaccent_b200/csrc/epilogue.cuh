// Shared epilogue of the tcgen05 kernels: TMEM accumulator tile [128 rows x BN fp32] -> registers ->
// (bias, ReLU) -> shared-memory staging tile -> COALESCED global stores / accumulates / fp32 atomics.
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

// stage      : >= EpiLayout::BYTES bytes of shared memory, 16 B aligned, private to the 4 epilogue warps
// sbias      : BN floats in shared memory or nullptr
// dst_row    : global pointer of this thread's output row (BN contiguous OutT), nullptr = row not stored
// mask_row   : optional bf16 row (same shape); output is zeroed where mask <= 0 (ReLU backward)
// ncols      : number of valid columns (<= BN); vec_ok: rows are 16 B aligned and ncols == BN
// smask      : optional mask tile already resident in shared memory (bf16 output, vec_ok only): [128 rows x
//              64 ch] halves of 16 KB in the TMA 128B-swizzled layout, row = TMEM lane; replaces mask_row
template <int BN, typename OutT>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, int q, int lane, unsigned char* stage,
                                              const float* sbias, OutT* dst_row, const __nv_bfloat16* mask_row,
                                              int ncols, bool vec_ok, int mode, bool relu,
                                              const unsigned char* smask = nullptr) {
  using L = EpiLayout<BN, OutT>;
  const int r = q * 32 + lane;
  unsigned char* my = stage + r * L::ROWB;
  __syncwarp();                       // a previous call's phase 2 (same warp, same rows) has finished reading
  // ---- phase 1
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    float v[32];
    umma::tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
    umma::tmem_ld_wait();
    if (sbias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + j);
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr (sizeof(OutT) == 2) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(my) + c0 + j, v + j);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(my + (c0 + j) * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  void** ptrs = reinterpret_cast<void**>(stage + L::PTR_OFF);
  ptrs[2 * r] = dst_row;
  ptrs[2 * r + 1] = const_cast<__nv_bfloat16*>(mask_row);
  __syncwarp();
  // ---- phase 2: this warp's rows q*32 .. q*32+31
  if (vec_ok) {
    constexpr int CPR = BN * int(sizeof(OutT)) / 16;               // 16 B chunks per row
    constexpr int RPI = 32 / CPR < 1 ? 1 : 32 / CPR;               // rows per iteration
    constexpr int IPR = CPR > 32 ? CPR / 32 : 1;                   // iterations per row (BN*size > 512 B)
    const int sub = lane / (CPR < 32 ? CPR : 32);
    const int ch0 = lane % (CPR < 32 ? CPR : 32);
#pragma unroll 1
    for (int rr = 0; rr < 32; rr += RPI) {
      const int row = q * 32 + rr + sub;
      OutT* dst = static_cast<OutT*>(ptrs[2 * row]);
      const __nv_bfloat16* msk = static_cast<const __nv_bfloat16*>(ptrs[2 * row + 1]);
      if (dst == nullptr) continue;
#pragma unroll
      for (int it = 0; it < IPR; ++it) {
        const int ch = ch0 + it * 32;
        uint4 val = *reinterpret_cast<const uint4*>(stage + row * L::ROWB + ch * 16);
        if constexpr (sizeof(OutT) == 2) {
          if (msk != nullptr || smask != nullptr || mode == EPI_ACCUM) {
            float f[8];
            load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(&val), f);
            if (smask != nullptr) {
              float mk[8];
              load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(
                                       smask + (ch >> 3) * 16384 + row * 128 + (((ch & 7) ^ (row & 7)) << 4)), mk);
#pragma unroll
              for (int e = 0; e < 8; ++e) if (!(mk[e] > 0.f)) f[e] = 0.f;
            } else if (msk != nullptr) {
              float mk[8];
              load8<__nv_bfloat16>(msk + ch * 8, mk);
#pragma unroll
              for (int e = 0; e < 8; ++e) if (!(mk[e] > 0.f)) f[e] = 0.f;
            }
            if (mode == EPI_ACCUM) {
              float old[8];
              load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(dst) + ch * 8, old);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += old[e];
            }
            store8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(dst) + ch * 8, f);
          } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + ch * 16) = val;
          }
        } else {
          float* d4 = reinterpret_cast<float*>(dst) + ch * 4;
          const float4 f = *reinterpret_cast<const float4*>(&val);
          if (mode == EPI_ATOMIC) {
            atomicAdd(d4, f.x); atomicAdd(d4 + 1, f.y); atomicAdd(d4 + 2, f.z); atomicAdd(d4 + 3, f.w);
          } else if (mode == EPI_ACCUM) {
            const float4 o = *reinterpret_cast<const float4*>(d4);
            *reinterpret_cast<float4*>(d4) = make_float4(f.x + o.x, f.y + o.y, f.z + o.z, f.w + o.w);
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
        if (msk != nullptr && !(__bfloat162float(msk[c]) > 0.f)) f = 0.f;
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
