// VGG front end helpers (mono_transformer_torch.py:49-60): the Cin=1 first conv (HBM-bound,
// direct), im2col/col2im for the fp32 SIMT path, weight layout prep, 2x2 max-pool and ReLU
// backward.  All activation tensors are NHWC (H = time, W = frequency).
#include "common.cuh"

namespace masr {

// ------------------------------------------------------------------ conv1 (Cin = 1)
// One thread computes 8 output channels of one pixel: 8 threads cover a pixel's 64 channels, so a
// warp writes 4 pixels x 64 channels contiguously (coalesced); the 9 input taps come from L1.
template <typename T>
__global__ void conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, T* __restrict__ y,
                                 int B, int H, int W, int Cout) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sw[];            // [Cout][9] weights then [Cout] bias
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[Cout * 9 + i] = bias[i];
  __syncthreads();
  const int groups = Cout / 8;
  const int64_t total = int64_t(B) * H * W * groups;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int cg = int(idx % groups);
    const int64_t p = idx / groups;
    const int wv = int(p % W);
    const int hv = int((p / W) % H);
    const int64_t b = p / (int64_t(W) * H);
    float tap[9];
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        const int hh = hv + dh, ww = wv + dw;
        tap[(dh + 1) * 3 + (dw + 1)] =
            (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(b * H + hh) * W + ww] : 0.f;
      }
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float* wc = sw + (cg * 8 + c) * 9;
      float acc = sw[Cout * 9 + cg * 8 + c];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(tap[t], wc[t], acc);
      o[c] = fmaxf(acc, 0.f);
    }
    store8<T>(y + p * Cout + cg * 8, o);
  }
}

// conv1 forward, register-weight variant: the block stride is a multiple of Cout/8, so a thread keeps the same
// 8-channel group for every pixel it visits and holds its 72 weights + 8 biases in registers; index
// arithmetic is 32-bit.  HBM-bound on the NHWC bf16 output (8 channels = one 128-bit store per thread).
template <typename T>
__global__ void __launch_bounds__(256) conv1_fwd_v2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ y,
                                                           int B, int H, int W, int Cout) {
  pdl_launch_dependents();
  pdl_wait();
  const int groups = Cout / 8;
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = int(gtid % unsigned(groups));
  float wr[8][9], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias[cg * 8 + c];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = w[(cg * 8 + c) * 9 + t];
  }
  const unsigned P = unsigned(B) * H * W;
  const unsigned pstride = (gridDim.x * blockDim.x) / unsigned(groups);
  constexpr int U = 3;                                      // pixels in flight per thread (latency hiding)
  for (unsigned p0 = gtid / unsigned(groups); p0 < P; p0 += pstride * U) {
    float tap[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = p0 + unsigned(u) * pstride;
      const bool pok = p < P;
      const int wv = int(p % unsigned(W));
      const unsigned row = p / unsigned(W);                // b * H + h
      const int hv = int(row % unsigned(H));
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh) {
        const int hh = hv + dh;
        const bool hok = pok && hh >= 0 && hh < H;
        const float* xr = x + (int64_t(row) + dh) * W;
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int ww = wv + dw;
          tap[u][(dh + 1) * 3 + (dw + 1)] = (hok && ww >= 0 && ww < W) ? __ldg(xr + ww) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = p0 + unsigned(u) * pstride;
      if (p >= P) break;
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float acc = br[c];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc = fmaf(tap[u][t], wr[c][t], acc);
        o[c] = fmaxf(acc, 0.f);
      }
      store8<T>(y + int64_t(p) * Cout + cg * 8, o);
    }
  }
}

// dw[co][tap] += sum_p dy[p][co] x[p+tap], db[co] += sum_p dy[p][co].  Block = Cout x PL threads;
// thread (co, lane) walks pixels lane, lane+PL, ... of the block's slice, so reads of dy are
// contiguous over co.  Partials are combined in shared memory, then one atomicAdd per output.
template <typename T>
__global__ void conv1_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                   float* __restrict__ dw, float* __restrict__ db,
                                   int B, int H, int W, int Cout, int64_t pix_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];           // [PL][Cout][10]
  const int co = threadIdx.x % Cout;
  const int lane = threadIdx.x / Cout;
  const int PL = blockDim.x / Cout;
  const int64_t P = int64_t(B) * H * W;
  const int64_t p0 = blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (int64_t p = p0 + lane; p < p1; p += PL) {
    const float g = to_f<T>(dy[p * Cout + co]);
    if (g == 0.f) continue;                // ReLU-masked gradients are mostly exact zeros
    const int wv = int(p % W);
    const int hv = int((p / W) % H);
    const int64_t b = p / (int64_t(W) * H);
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
      for (int dw_ = -1; dw_ <= 1; ++dw_) {
        const int hh = hv + dh, ww = wv + dw_;
        const float xv = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(b * H + hh) * W + ww] : 0.f;
        acc[(dh + 1) * 3 + (dw_ + 1)] = fmaf(g, xv, acc[(dh + 1) * 3 + (dw_ + 1)]);
      }
    acc[9] += g;
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) red[(lane * Cout + co) * 10 + i] = acc[i];
  __syncthreads();
  for (int o = threadIdx.x; o < Cout * 10; o += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += red[l * Cout * 10 + o];
    const int c = o / 10, t = o % 10;
    if (t < 9) atomicAdd(&dw[c * 9 + t], s); else atomicAdd(&db[c], s);
  }
}

// ------------------------------------------------------------------ im2col / col2im (SIMT path)
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ col,
                                 int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(B) * H * W * 9 * C;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int tap = int((idx / C) % 9);
    const int64_t p = idx / (9 * int64_t(C));
    const int wv = int(p % W);
    const int hv = int((p / W) % H);
    const int64_t b = p / (int64_t(W) * H);
    const int hh = hv + tap / 3 - 1, ww = wv + tap % 3 - 1;
    T v = from_f<T>(0.f);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((b * H + hh) * W + ww) * C + c];
    col[idx] = v;
  }
}

template <typename T>
__global__ void col2im3x3_kernel(const T* __restrict__ dcol, T* __restrict__ dx,
                                 const T* __restrict__ relu_src, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = int64_t(B) * H * W * C;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int64_t q = idx / C;
    const int wv = int(q % W);
    const int hv = int((q / W) % H);
    const int64_t b = q / (int64_t(W) * H);
    float s = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      // col[p][tap] = x[p + (dh,dw)]  =>  x[q] feeds col[q - (dh,dw)][tap]
      const int hh = hv - (tap / 3 - 1), ww = wv - (tap % 3 - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W)
        s += to_f<T>(dcol[(((b * H + hh) * W + ww) * 9 + tap) * C + c]);
    }
    if (relu_src != nullptr && !(to_f<T>(relu_src[idx]) > 0.f)) s = 0.f;
    dx[idx] = from_f<T>(s);
  }
}

// w [Cout][Cin][3][3] fp32 -> wp [Cout][tap][Cin]
template <typename T>
__global__ void conv_w_prep_kernel(const float* __restrict__ w, T* __restrict__ wp, int Cout, int Cin) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = Cout * Cin * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % Cin;
    const int tap = (idx / Cin) % 9;
    const int co = idx / (9 * Cin);
    wp[idx] = from_f<T>(w[(co * Cin + ci) * 9 + tap]);
  }
}
// w [Cout][Cin][3][3] fp32 -> wpt [Cin][tap][Cout]: the K-major B operand of the dgrad implicit GEMM
template <typename T>
__global__ void conv_w_prep_t_kernel(const float* __restrict__ w, T* __restrict__ wpt, int Cout, int Cin) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = Cout * Cin * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % Cout;
    const int tap = (idx / Cout) % 9;
    const int ci = idx / (9 * Cout);
    wpt[idx] = from_f<T>(w[(co * Cin + ci) * 9 + tap]);
  }
}
__global__ void conv_w_unprep_add_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = Cout * Cin * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % Cin;
    const int tap = (idx / Cin) % 9;
    const int co = idx / (9 * Cin);
    dw[(co * Cin + ci) * 9 + tap] += dwp[idx];
  }
}

// ------------------------------------------------------------------ max pool 2x2 / 2, floor mode
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = int64_t(B) * Ho * Wo * C;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int wo = int((idx / C) % Wo);
    const int ho = int((idx / (int64_t(C) * Wo)) % Ho);
    const int64_t b = idx / (int64_t(C) * Wo * Ho);
    const T* base = x + ((b * H + 2 * ho) * W + 2 * wo) * C + c;
    float m = to_f<T>(base[0]);
    m = fmaxf(m, to_f<T>(base[C]));
    m = fmaxf(m, to_f<T>(base[int64_t(W) * C]));
    m = fmaxf(m, to_f<T>(base[int64_t(W) * C + C]));
    y[idx] = from_f<T>(m);
  }
}

// One thread per INPUT element: gradient goes to the first maximum of its window in (h, w) scan
// order (ATen's strict '>' update), zero elsewhere and in the floor-dropped last row/column.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                   int relu_mask, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = int64_t(B) * H * W * C;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % C);
    const int wv = int((idx / C) % W);
    const int hv = int((idx / (int64_t(C) * W)) % H);
    const int64_t b = idx / (int64_t(C) * W * H);
    const int ho = hv / 2, wo = wv / 2;
    float g = 0.f;
    if (ho < Ho && wo < Wo) {
      const T* base = x + ((b * H + 2 * ho) * W + 2 * wo) * C + c;
      const float v0 = to_f<T>(base[0]), v1 = to_f<T>(base[C]);
      const float v2 = to_f<T>(base[int64_t(W) * C]), v3 = to_f<T>(base[int64_t(W) * C + C]);
      int arg = 0; float m = v0;
      if (v1 > m) { m = v1; arg = 1; }
      if (v2 > m) { m = v2; arg = 2; }
      if (v3 > m) { m = v3; arg = 3; }
      const int mine = (hv & 1) * 2 + (wv & 1);
      if (arg == mine && (!relu_mask || m > 0.f))
        g = to_f<T>(dy[((b * Ho + ho) * Wo + wo) * C + c]);
    }
    dx[idx] = from_f<T>(g);
  }
}

template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ y, T* __restrict__ dx, int64_t n, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dx[i] = (to_f<T>(y[i]) > 0.f) ? from_f<T>(to_f<T>(dx[i]) * scale) : from_f<T>(0.f);
}

// ------------------------------------------------------------------ vectorised pool kernels (C % 8 == 0)
// One thread owns 8 consecutive channels of one 2x2 window: four 128-bit loads, one (fwd) or four (bwd)
// 128-bit stores; a warp covers 256 contiguous channels-bytes per pixel -> fully coalesced NHWC traffic.
template <typename T>
__global__ void maxpool_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, ncg = C / 8;
  const int64_t total = int64_t(B) * Ho * Wo * ncg;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int cg = int(idx % ncg);
    const int wo = int((idx / ncg) % Wo);
    const int ho = int((idx / (int64_t(ncg) * Wo)) % Ho);
    const int64_t b = idx / (int64_t(ncg) * Wo * Ho);
    const T* base = x + ((b * H + 2 * ho) * W + 2 * wo) * C + cg * 8;
    float v0[8], v1[8], v2[8], v3[8];
    load8<T>(base, v0); load8<T>(base + C, v1); load8<T>(base + int64_t(W) * C, v2); load8<T>(base + int64_t(W) * C + C, v3);
#pragma unroll
    for (int j = 0; j < 8; ++j) v0[j] = fmaxf(fmaxf(v0[j], v1[j]), fmaxf(v2[j], v3[j]));
    store8<T>(y + ((b * Ho + ho) * Wo + wo) * C + cg * 8, v0);
  }
}

template <typename T>
__global__ void maxpool_bwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                       int relu_mask, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, ncg = C / 8;
  const int64_t total = int64_t(B) * Ho * Wo * ncg;
  const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int cg = int(idx % ncg);
    const int wo = int((idx / ncg) % Wo);
    const int ho = int((idx / (int64_t(ncg) * Wo)) % Ho);
    const int64_t b = idx / (int64_t(ncg) * Wo * Ho);
    const int64_t o00 = ((b * H + 2 * ho) * W + 2 * wo) * C + cg * 8;
    const int64_t rowstep = int64_t(W) * C;
    float v0[8], v1[8], v2[8], v3[8], g[8];
    load8<T>(x + o00, v0); load8<T>(x + o00 + C, v1); load8<T>(x + o00 + rowstep, v2); load8<T>(x + o00 + rowstep + C, v3);
    load8<T>(dy + ((b * Ho + ho) * Wo + wo) * C + cg * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int arg = 0; float m = v0[j];
      if (v1[j] > m) { m = v1[j]; arg = 1; }
      if (v2[j] > m) { m = v2[j]; arg = 2; }
      if (v3[j] > m) { m = v3[j]; arg = 3; }
      const float gg = (!relu_mask || m > 0.f) ? g[j] : 0.f;
      v0[j] = arg == 0 ? gg : 0.f; v1[j] = arg == 1 ? gg : 0.f; v2[j] = arg == 2 ? gg : 0.f; v3[j] = arg == 3 ? gg : 0.f;
    }
    store8<T>(dx + o00, v0); store8<T>(dx + o00 + C, v1); store8<T>(dx + o00 + rowstep, v2); store8<T>(dx + o00 + rowstep + C, v3);
    // floor-mode pooling drops an odd last column / row: their gradient is zero
    if ((W & 1) && wo == Wo - 1) { store8<T>(dx + o00 + 2 * C, zero8); store8<T>(dx + o00 + rowstep + 2 * C, zero8); }
    if ((H & 1) && ho == Ho - 1) {
      store8<T>(dx + o00 + 2 * rowstep, zero8); store8<T>(dx + o00 + 2 * rowstep + C, zero8);
      if ((W & 1) && wo == Wo - 1) store8<T>(dx + o00 + 2 * rowstep + 2 * C, zero8);
    }
  }
}

// Pool forward that also records, per pooled element, WHICH input won (bits 0-1: position in (h, w) scan order, ATen's
// first maximum) and whether the maximum is positive (bit 2: the ReLU in front of the pool passes a gradient).  The backward
// then scatters dy from the codes alone -- it no longer re-reads the full-resolution activation to recompute the arg-max
// (391 -> 239 MB per launch at 32 x 512 x 83 x 64).
template <typename T>
__global__ void maxpool_fwd_code_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ code,
                                        int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, ncg = C / 8;
  const int64_t total = int64_t(B) * Ho * Wo * ncg;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int cg = int(idx % ncg);
    const int wo = int((idx / ncg) % Wo);
    const int ho = int((idx / (int64_t(ncg) * Wo)) % Ho);
    const int64_t b = idx / (int64_t(ncg) * Wo * Ho);
    const T* base = x + ((b * H + 2 * ho) * W + 2 * wo) * C + cg * 8;
    float v0[8], v1[8], v2[8], v3[8];
    load8<T>(base, v0); load8<T>(base + C, v1); load8<T>(base + int64_t(W) * C, v2); load8<T>(base + int64_t(W) * C + C, v3);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t arg = 0; float m = v0[j];
      if (v1[j] > m) { m = v1[j]; arg = 1; }
      if (v2[j] > m) { m = v2[j]; arg = 2; }
      if (v3[j] > m) { m = v3[j]; arg = 3; }
      v0[j] = m;
      const uint32_t cd = arg | (m > 0.f ? 4u : 0u);
      if (j < 4) lo |= cd << (8 * j); else hi |= cd << (8 * (j - 4));
    }
    const int64_t o = ((b * Ho + ho) * Wo + wo) * C + cg * 8;
    store8<T>(y + o, v0);
    *reinterpret_cast<uint2*>(code + o) = make_uint2(lo, hi);
  }
}

template <typename T>
__global__ void maxpool_bwd_code_kernel(const unsigned char* __restrict__ code, const T* __restrict__ dy, T* __restrict__ dx,
                                        int relu_mask, int B, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, ncg = C / 8;
  const int64_t total = int64_t(B) * Ho * Wo * ncg;
  const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int cg = int(idx % ncg);
    const int wo = int((idx / ncg) % Wo);
    const int ho = int((idx / (int64_t(ncg) * Wo)) % Ho);
    const int64_t b = idx / (int64_t(ncg) * Wo * Ho);
    const int64_t o00 = ((b * H + 2 * ho) * W + 2 * wo) * C + cg * 8;
    const int64_t rowstep = int64_t(W) * C;
    const int64_t o = ((b * Ho + ho) * Wo + wo) * C + cg * 8;
    float v0[8], v1[8], v2[8], v3[8], g[8];
    load8<T>(dy + o, g);
    const uint2 cd2 = *reinterpret_cast<const uint2*>(code + o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t cd = ((j < 4 ? cd2.x : cd2.y) >> (8 * (j & 3))) & 0xffu;
      const uint32_t arg = cd & 3u;
      const float gg = (!relu_mask || (cd & 4u)) ? g[j] : 0.f;
      v0[j] = arg == 0 ? gg : 0.f; v1[j] = arg == 1 ? gg : 0.f; v2[j] = arg == 2 ? gg : 0.f; v3[j] = arg == 3 ? gg : 0.f;
    }
    store8<T>(dx + o00, v0); store8<T>(dx + o00 + C, v1); store8<T>(dx + o00 + rowstep, v2); store8<T>(dx + o00 + rowstep + C, v3);
    if ((W & 1) && wo == Wo - 1) { store8<T>(dx + o00 + 2 * C, zero8); store8<T>(dx + o00 + rowstep + 2 * C, zero8); }
    if ((H & 1) && ho == Ho - 1) {
      store8<T>(dx + o00 + 2 * rowstep, zero8); store8<T>(dx + o00 + 2 * rowstep + C, zero8);
      if ((W & 1) && wo == Wo - 1) store8<T>(dx + o00 + 2 * rowstep + 2 * C, zero8);
    }
  }
}

// conv1 wgrad, latency-hiding variant: 4 pixel lanes x Cout threads per block, each thread walks its pixels
// four at a time (independent loads in flight), many blocks, partials combined in shared memory.
template <typename T>
__global__ void __launch_bounds__(256) conv1_wgrad_v2_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                             float* __restrict__ dw, float* __restrict__ db,
                                                             int B, int H, int W, int Cout, int64_t pix_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];           // [PL][Cout][10]
  const int co = threadIdx.x % Cout;
  const int lane = threadIdx.x / Cout;
  const int PL = blockDim.x / Cout;
  const int64_t P = int64_t(B) * H * W;
  const int64_t p0 = blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  constexpr int U = 4;
  for (int64_t pb = p0 + lane; pb < p1; pb += int64_t(PL) * U) {
    float g[U], xv[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = pb + int64_t(u) * PL;
      const bool ok = p < p1;
      g[u] = ok ? to_f<T>(dy[p * Cout + co]) : 0.f;
      const int wv = int(p % W);
      const int hv = int((p / W) % H);
      const int64_t b = p / (int64_t(W) * H);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = hv + t / 3 - 1, ww = wv + t % 3 - 1;
        xv[u][t] = (ok && hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(&x[(b * H + hh) * W + ww]) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[t] = fmaf(g[u], xv[u][t], acc[t]);
      acc[9] += g[u];
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) red[(lane * Cout + co) * 10 + i] = acc[i];
  __syncthreads();
  for (int o = threadIdx.x; o < Cout * 10; o += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += red[l * Cout * 10 + o];
    const int c = o / 10, t = o % 10;
    if (t < 9) atomicAdd(&dw[c * 9 + t], s); else atomicAdd(&db[c], s);
  }
}

// conv1 wgrad, row-segment variant: a thread (co, lane) handles SEG = 8 consecutive pixels of one image row
// per iteration: the 3 x 10 input window is loaded once and slid over the 8 pixels (30 loads instead of
// 72), index arithmetic is 32-bit and done once per segment.
template <typename T>
__global__ void __launch_bounds__(256) conv1_wgrad_v3_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                             float* __restrict__ dw, float* __restrict__ db,
                                                             int B, int H, int W, int Cout, int segs_per_row, int total_segs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];           // [PL][Cout][10]
  constexpr int SEG = 8;
  const int co = threadIdx.x % Cout;
  const int lane = threadIdx.x / Cout;
  const int PL = blockDim.x / Cout;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (int sidx = blockIdx.x * PL + lane; sidx < total_segs; sidx += gridDim.x * PL) {
    const unsigned us = unsigned(sidx);
    const int sw = int(us % unsigned(segs_per_row));
    const unsigned row = us / unsigned(segs_per_row);        // b * H + h
    const int h = int(row % unsigned(H));
    const int w0 = sw * SEG;
    float win[3][SEG + 2];
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hh = h + dh - 1;
      const bool hok = hh >= 0 && hh < H;
      const float* xr = x + (int64_t(row) + (dh - 1)) * W;
#pragma unroll
      for (int j = 0; j < SEG + 2; ++j) {
        const int ww = w0 + j - 1;
        win[dh][j] = (hok && ww >= 0 && ww < W) ? __ldg(xr + ww) : 0.f;
      }
    }
    const T* dyr = dy + (int64_t(row) * W + w0) * Cout + co;
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      const float g = (w0 + j < W) ? to_f<T>(dyr[int64_t(j) * Cout]) : 0.f;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw_ = 0; dw_ < 3; ++dw_) acc[dh * 3 + dw_] = fmaf(g, win[dh][j + dw_], acc[dh * 3 + dw_]);
      acc[9] += g;
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) red[(lane * Cout + co) * 10 + i] = acc[i];
  __syncthreads();
  for (int o = threadIdx.x; o < Cout * 10; o += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += red[l * Cout * 10 + o];
    const int c = o / 10, t = o % 10;
    if (t < 9) atomicAdd(&dw[c * 9 + t], s); else atomicAdd(&db[c], s);
  }
}

static inline int grid_for(int64_t total, int threads) {
  int64_t blocks = ceil_div64(total, threads);
  int64_t cap = int64_t(sm_count()) * 16;
  return int(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace masr

using namespace masr;

extern "C" int masr_conv1_fwd(const float* x, const float* w, const float* bias, void* y, int y_dtype,
                              int B, int H, int W, int Cout, void* stream) {
  MASR_REQUIRE(Cout % 8 == 0 && Cout <= 256, "conv1: Cout must be a multiple of 8");
  const int64_t total = int64_t(B) * H * W * (Cout / 8);
  if (total == 0) return MASR_OK;
  const size_t smem = size_t(Cout) * 10 * sizeof(float);
  if (int64_t(B) * H * W < (int64_t(1) << 31) && 256 % (Cout / 8) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const int blocks = int(std::min<int64_t>(ceil_div64(total, 256), int64_t(sm_count()) * 8));
    MASR_DISPATCH_DTYPE(y_dtype, T,
        launch_pdl(conv1_fwd_v2_kernel<T>, dim3(blocks), dim3(256), 0, as_stream(stream), x, w, bias, static_cast<T*>(y), B, H, W, Cout));
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  MASR_DISPATCH_DTYPE(y_dtype, T,
      launch_pdl(conv1_fwd_kernel<T>, dim3(grid_for(total, 256)), dim3(256), smem, as_stream(stream), 
          x, w, bias, static_cast<T*>(y), B, H, W, Cout));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_conv1_wgrad(const float* x, const void* dy, int dy_dtype, float* dw, float* db,
                                int B, int H, int W, int Cout, void* stream) {
  MASR_REQUIRE(Cout > 0 && Cout <= 256 && 256 % Cout == 0, "conv1_wgrad: Cout must divide 256");
  const int64_t P = int64_t(B) * H * W;
  if (P == 0) return MASR_OK;
  const int threads = 256, PL = threads / Cout;
  if (P * 2 < (int64_t(1) << 31)) {
    const int segs_per_row = (W + 7) / 8;
    const int total_segs = B * H * segs_per_row;
    const int blocks3 = std::max(1, std::min((total_segs + PL - 1) / PL, sm_count() * 8));
    const size_t smem3 = size_t(PL) * Cout * 10 * sizeof(float);
    MASR_DISPATCH_DTYPE(dy_dtype, T,
        launch_pdl(conv1_wgrad_v3_kernel<T>, dim3(blocks3), dim3(threads), smem3, as_stream(stream), 
            x, static_cast<const T*>(dy), dw, db, B, H, W, Cout, segs_per_row, total_segs));
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  int blocks = int(std::min<int64_t>(ceil_div64(P, 64), int64_t(sm_count()) * 8));
  const int64_t ppb = ceil_div64(P, blocks);
  blocks = int(ceil_div64(P, ppb));
  const size_t smem = size_t(PL) * Cout * 10 * sizeof(float);
  MASR_DISPATCH_DTYPE(dy_dtype, T,
      launch_pdl(conv1_wgrad_v2_kernel<T>, dim3(blocks), dim3(threads), smem, as_stream(stream), 
          x, static_cast<const T*>(dy), dw, db, B, H, W, Cout, ppb));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_im2col3x3(const void* x, void* col, int dtype, int B, int H, int W, int Cin, void* stream) {
  const int64_t total = int64_t(B) * H * W * 9 * Cin;
  if (total == 0) return MASR_OK;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(im2col3x3_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), 
          static_cast<const T*>(x), static_cast<T*>(col), B, H, W, Cin));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_col2im3x3(const void* dcol, void* dx, int dtype, const void* relu_src,
                              int B, int H, int W, int Cin, void* stream) {
  const int64_t total = int64_t(B) * H * W * Cin;
  if (total == 0) return MASR_OK;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(col2im3x3_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), 
          static_cast<const T*>(dcol), static_cast<T*>(dx), static_cast<const T*>(relu_src), B, H, W, Cin));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_conv_w_prep(const float* w, void* wp, int dtype, int Cout, int Cin, void* stream) {
  const int total = Cout * Cin * 9;
  if (total == 0) return MASR_OK;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(conv_w_prep_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), w, static_cast<T*>(wp), Cout, Cin));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_conv_w_prep_t(const float* w, void* wpt, int dtype, int Cout, int Cin, void* stream) {
  const int total = Cout * Cin * 9;
  if (total == 0) return MASR_OK;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(conv_w_prep_t_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), w, static_cast<T*>(wpt), Cout, Cin));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_conv_w_unprep_add(const float* dwp, float* dw, int Cout, int Cin, void* stream) {
  const int total = Cout * Cin * 9;
  if (total == 0) return MASR_OK;
  launch_pdl(conv_w_unprep_add_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), dwp, dw, Cout, Cin);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_maxpool2x2_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream) {
  const int64_t total = int64_t(B) * (H / 2) * (W / 2) * C;
  if (total == 0) return MASR_OK;
  if (C % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    MASR_DISPATCH_DTYPE(dtype, T,
        launch_pdl(maxpool_fwd_vec_kernel<T>, dim3(grid_for(total / 8, 256)), dim3(256), 0, as_stream(stream), 
            static_cast<const T*>(x), static_cast<T*>(y), B, H, W, C));
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(maxpool_fwd_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), 
          static_cast<const T*>(x), static_cast<T*>(y), B, H, W, C));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int dtype, int relu_mask,
                                   int B, int H, int W, int C, void* stream) {
  const int64_t total = int64_t(B) * H * W * C;
  if (total == 0) return MASR_OK;
  const int64_t wins = int64_t(B) * (H / 2) * (W / 2) * (C / 8);
  if (C % 8 == 0 && H >= 2 && W >= 2 && wins > 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    MASR_DISPATCH_DTYPE(dtype, T,
        launch_pdl(maxpool_bwd_vec_kernel<T>, dim3(grid_for(wins, 256)), dim3(256), 0, as_stream(stream), 
            static_cast<const T*>(x), static_cast<const T*>(dy), static_cast<T*>(dx), relu_mask, B, H, W, C));
    MASR_LAUNCH_CHECK();
    return MASR_OK;
  }
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(maxpool_bwd_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), 
          static_cast<const T*>(x), static_cast<const T*>(dy), static_cast<T*>(dx), relu_mask, B, H, W, C));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_maxpool2x2_fwd_code(const void* x, void* y, void* code, int dtype, int B, int H, int W, int C, void* stream) {
  const int64_t wins = int64_t(B) * (H / 2) * (W / 2) * (C / 8);
  if (wins == 0) return MASR_OK;
  MASR_REQUIRE(C % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(code) & 7) == 0, "maxpool2x2_fwd_code: C % 8 == 0 and 16 B aligned tensors");
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(maxpool_fwd_code_kernel<T>, dim3(grid_for(wins, 256)), dim3(256), 0, as_stream(stream),
          static_cast<const T*>(x), static_cast<T*>(y), static_cast<unsigned char*>(code), B, H, W, C));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_maxpool2x2_bwd_code(const void* code, const void* dy, void* dx, int dtype, int relu_mask,
                                        int B, int H, int W, int C, void* stream) {
  const int64_t wins = int64_t(B) * (H / 2) * (W / 2) * (C / 8);
  MASR_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2 && wins > 0 &&
               ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(code) & 7) == 0, "maxpool2x2_bwd_code: C % 8 == 0 and 16 B aligned tensors");
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(maxpool_bwd_code_kernel<T>, dim3(grid_for(wins, 256)), dim3(256), 0, as_stream(stream),
          static_cast<const unsigned char*>(code), static_cast<const T*>(dy), static_cast<T*>(dx), relu_mask, B, H, W, C));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

extern "C" int masr_relu_bwd(const void* y, void* dx, int dtype, int64_t n, float scale, void* stream) {
  if (n == 0) return MASR_OK;
  MASR_DISPATCH_DTYPE(dtype, T,
      launch_pdl(relu_bwd_kernel<T>, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), static_cast<const T*>(y), static_cast<T*>(dx), n, scale));
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}
