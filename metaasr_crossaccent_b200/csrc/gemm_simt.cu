// Generic strided GEMM with fp32 accumulation on the CUDA cores.
//
// This is the exact-arithmetic path (fp32 mode of north_star: losses within 1e-5 of the
// reference) and the shape-agnostic fallback for operands the tcgen05 kernels do not cover.
// C[M,N] = op(sum_k A(m,k) B(n,k) + bias[n]) with arbitrary element strides, so one kernel serves
// nn.Linear forward (TN), dgrad (A K-major, B N-major) and wgrad (both M-major, split-K).
#include "common.cuh"

namespace masr {

constexpr int GS_BM = 128, GS_BN = 128, GS_BK = 16, GS_THREADS = 256, GS_PAD = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(GS_THREADS)
gemm_simt_kernel(const TA* __restrict__ A, int64_t sam, int64_t sak,
                 const TB* __restrict__ B, int64_t sbn, int64_t sbk,
                 TC* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                 int M, int N, int K, int flags, int k_per_split) {
  __shared__ __align__(16) float As[GS_BK][GS_BM + GS_PAD];
  __shared__ __align__(16) float Bs[GS_BK][GS_BN + GS_PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid % 16, ty = tid / 16;
  const bool a_kc = (sak == 1), b_kc = (sbk == 1);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GS_BK) {
#pragma unroll
    for (int i = 0; i < (GS_BM * GS_BK) / GS_THREADS; ++i) {
      int e = tid + i * GS_THREADS;
      int m, k;
      if (a_kc) { k = e % GS_BK; m = e / GS_BK; } else { m = e % GS_BM; k = e / GS_BM; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = to_f<TA>(A[int64_t(gm) * sam + int64_t(gk) * sak]);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (GS_BN * GS_BK) / GS_THREADS; ++i) {
      int e = tid + i * GS_THREADS;
      int n, k;
      if (b_kc) { k = e % GS_BK; n = e / GS_BK; } else { n = e % GS_BN; k = e / GS_BN; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = to_f<TB>(B[int64_t(gn) * sbn + int64_t(gk) * sbk]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GS_BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool relu = flags & MASR_GEMM_RELU, accum = flags & MASR_GEMM_ACCUM, splitk = flags & MASR_GEMM_SPLITK;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      TC* dst = C + int64_t(m) * ldc + n;
      if (splitk) {
        if (bias != nullptr && blockIdx.z == 0) v += bias[n];
        if constexpr (sizeof(TC) == 4) atomicAdd(reinterpret_cast<float*>(dst), v);
      } else {
        if (bias != nullptr) v += bias[n];
        if (accum) v += to_f<TC>(*dst);
        if (relu) v = fmaxf(v, 0.f);
        *dst = from_f<TC>(v);
      }
    }
  }
}

template <typename TA, typename TB, typename TC>
static int launch_gemm(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk,
                       void* C, int64_t ldc, const float* bias, int M, int N, int K, int flags,
                       int splitk, cudaStream_t st) {
  int nsplit = 1, kper = K;
  if (flags & MASR_GEMM_SPLITK) {
    nsplit = splitk > 0 ? splitk : 1;
    kper = int(ceil_div64(ceil_div64(K, nsplit), GS_BK) * GS_BK);
    nsplit = int(ceil_div64(K, kper));
  }
  dim3 grid(unsigned(ceil_div64(N, GS_BN)), unsigned(ceil_div64(M, GS_BM)), unsigned(nsplit));
  gemm_simt_kernel<TA, TB, TC><<<grid, GS_THREADS, 0, st>>>(
      static_cast<const TA*>(A), sam, sak, static_cast<const TB*>(B), sbn, sbk,
      static_cast<TC*>(C), ldc, bias, M, N, K, flags, kper);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}

}  // namespace masr

using namespace masr;

extern "C" int masr_gemm(const void* A, int a_dtype, int64_t sam, int64_t sak,
                         const void* B, int b_dtype, int64_t sbn, int64_t sbk,
                         void* C, int c_dtype, int64_t ldc, const float* bias,
                         int M, int N, int K, int flags, int splitk, void* stream) {
  MASR_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative GEMM size");
  if (M == 0 || N == 0) return MASR_OK;
  MASR_REQUIRE(!(flags & MASR_GEMM_SPLITK) || c_dtype == MASR_F32, "split-K needs an fp32 C");
  MASR_REQUIRE(!((flags & MASR_GEMM_SPLITK) && (flags & MASR_GEMM_RELU)), "split-K cannot fuse ReLU");
  cudaStream_t st = as_stream(stream);
  const int key = a_dtype * 4 + b_dtype * 2 + c_dtype;
  switch (key) {
#define GEMM_CASE(k, TA, TB, TC) \
    case k: return launch_gemm<TA, TB, TC>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, flags, splitk, st);
    GEMM_CASE(0, float, float, float)
    GEMM_CASE(1, float, float, __nv_bfloat16)
    GEMM_CASE(2, float, __nv_bfloat16, float)
    GEMM_CASE(3, float, __nv_bfloat16, __nv_bfloat16)
    GEMM_CASE(4, __nv_bfloat16, float, float)
    GEMM_CASE(5, __nv_bfloat16, float, __nv_bfloat16)
    GEMM_CASE(6, __nv_bfloat16, __nv_bfloat16, float)
    GEMM_CASE(7, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16)
#undef GEMM_CASE
    default: set_error("masr_gemm: bad dtype combination"); return MASR_E_INVALID;
  }
}
