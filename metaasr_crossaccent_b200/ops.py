"""Thin Python binding of the C ABI: each method forwards torch CUDA tensors (data_ptr, sizes,
strides) and the current CUDA stream to one `masr_*` entry point.  PyTorch is used for device
memory and streams only -- no torch op runs on this path, and there is no fallback: constructing
the backend fails unless libmetaasr_b200.so is built and an sm_100 GPU is present.
"""
from __future__ import annotations

import torch

from . import _lib

F32, BF16 = 0, 1
GEMM_RELU, GEMM_ACCUM, GEMM_SPLITK = 1, 2, 4
_SEED_TENSORS = {}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return None if t is None else t.data_ptr()


class CudaBackend:
    """Kernel dispatch object used by engine.TransformerEngine (the test suite swaps in a torch
    implementation of the same interface on CPU to check the orchestration against the oracle)."""

    name = "cuda"

    def __init__(self, device, act_dtype=torch.float32, gemm="simt"):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.MetaASRLibraryError("metaasr_b200 kernels run on an sm_100 GPU only (no CPU fallback)")
        self.device = device
        self.lib = _lib.init(device.index if device.index is not None else torch.cuda.current_device())
        self.act_dtype = act_dtype
        self.gemm_path = gemm          # "simt" | "umma"
        self.launches = 0              # our kernels launched through this backend (bench: gpu_launches)
        self.prof = None               # dict -> per-launch CUDA events of the tensor-core GEMMs (bench.py)
        self.prof_ops = None           # dict -> per-entry-point CUDA events (bench.py --profile)
        self.prof_tag = ""
        self._scratch = {}
        self.scratch_tag = ""         # set by the engine while it launches on its side stream
        # gemm="umma" promises the tcgen05 kernels.  An operand that cannot take them (dtype / alignment / head dim /
        # channel count) is never silent: the miss is counted in self.fallbacks, reported once per (op, reason) through
        # `warnings`, and raises when strict_umma is set (asr_model.strict_tcgen05: bench.py and the hkust-size tests)
        self.fallbacks = {}
        self.strict_umma = False
        # one device-resident dropout seed offset per device, alive for the whole process (the library
        # keeps its address; kernels add it to every dropout seed so CUDA-graph replays get fresh masks)
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in _SEED_TENSORS:
            _SEED_TENSORS[key] = torch.zeros(1, dtype=torch.int64, device=device)
        self.set_seed_ptr(_SEED_TENSORS[key])

    # -------------------------------------------------------------- plumbing
    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name, *args, n_kernels=1):
        po = self.prof_ops
        if po is not None:          # bench.py --profile: CUDA events around every entry point
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            _lib.check(rc, name)
        self.launches += n_kernels
        if po is not None:
            e1.record()
            po.setdefault(name + self.prof_tag, []).append((e0, e1))

    def scratch(self, key, numel, dtype):
        """Scratch buffers are keyed by (owner lane, size, dtype) and never re-allocated: a captured CUDA graph
        keeps their addresses, and engines that run concurrently (task lanes) must not share them."""
        key = (key, int(numel), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(int(numel), dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    # -------------------------------------------------------------- GEMM family
    def gemm(self, A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, flags=0, splitk=1):
        self._call("masr_gemm", _p(A), _dt(A), sam, sak, _p(B), _dt(B), sbn, sbk, _p(C), _dt(C), ldc, _p(bias),
                   M, N, K, flags, splitk, self.stream)

    def gemm_stage_cap(self, stages):
        """Operand-ring depth bound of the tcgen05 GEMMs (masr_gemm_set_stage_cap); 0 restores the default policy."""
        if getattr(self, "_stage_cap", 0) != stages:
            self.lib.masr_gemm_set_stage_cap(int(stages))
            self._stage_cap = stages

    def _miss(self, op, reason):
        """A launch in gemm='umma' mode that cannot use the tcgen05 kernel: count, warn once, raise if strict."""
        key = (op, reason)
        n = self.fallbacks.get(key, 0)
        self.fallbacks[key] = n + 1
        if self.strict_umma:
            raise _lib.MetaASRLibraryError(f"{op}: tcgen05 path not applicable ({reason}) and strict_tcgen05 is set")
        if n == 0:
            import warnings
            warnings.warn(f"metaasr_b200: {op} runs on the CUDA-core kernel ({reason}); further misses are only counted "
                          f"in backend.fallbacks", RuntimeWarning, stacklevel=3)
        return False

    def _umma_ok(self, *ts, op="gemm"):
        """tcgen05 path: bf16 operands, 16 B aligned bases, leading dimensions multiple of 8."""
        if self.gemm_path != "umma":
            return False
        for t in ts:
            if t.dtype != torch.bfloat16:
                return self._miss(op, f"operand dtype {t.dtype}")
            if t.data_ptr() % 16 or t.stride(0) % 8 or t.stride(1) != 1:
                return self._miss(op, f"operand alignment: stride {tuple(t.stride())}, base % 16 = {t.data_ptr() % 16}")
        return True

    def umma_gemm(self, A, a_mn, B, b_mn, C, bias, M, N, K, flags=0, splitk=1, rowsum=None, mask=None, mask_scale=1.0,
                  p_drop=0.0, seed=0, site=0, rowdot=None):
        prof = self.prof
        if prof is not None:        # bench.py: CUDA events around every launch, on the launching stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        epi = None
        if rowsum is not None or mask is not None or p_drop > 0.0 or rowdot is not None:
            dsrc, dout, dL, dH = rowdot if rowdot is not None else (None, None, 0, 0)
            epi = _lib.GemmEpilogue(_p(rowsum), _p(mask), mask.stride(0) if mask is not None else 0, float(mask_scale),
                                    float(p_drop), int(seed), int(site), _p(dsrc), dsrc.stride(0) if dsrc is not None else 0,
                                    _p(dout), int(dL), int(dH))
            epi = _lib.C.byref(epi)
        self._call("masr_umma_gemm_ex", _p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C), _dt(C),
                   C.stride(0), _p(bias), M, N, K, flags, int(splitk), epi, self.stream)
        if prof is not None:
            e1.record()
            kind = ("fwd", "dgrad", "a_mn", "wgrad")[int(a_mn) * 2 + int(b_mn)]
            prof.setdefault((kind, M, N, K), []).append((e0, e1))

    def umma_gemm_pair(self, A, a_mn, B, b_mn, C, bias, M, N, K, flags=0, splitk=0, bn=0, rowsum=None, mask=None,
                       mask_scale=1.0, p_drop=0.0, seed=0, site=0, rowdot=None):
        """Explicit entry to the persistent CTA-pair kernel (masr_umma_gemm_pair); umma_gemm reaches the same kernel
        through the size-based dispatch of masr_umma_gemm_ex."""
        epi = None
        if rowsum is not None or mask is not None or p_drop > 0.0 or rowdot is not None:
            dsrc, dout, dL, dH = rowdot if rowdot is not None else (None, None, 0, 0)
            epi = _lib.GemmEpilogue(_p(rowsum), _p(mask), mask.stride(0) if mask is not None else 0, float(mask_scale),
                                    float(p_drop), int(seed), int(site), _p(dsrc), dsrc.stride(0) if dsrc is not None else 0,
                                    _p(dout), int(dL), int(dH))
            epi = _lib.C.byref(epi)
        self._call("masr_umma_gemm_pair", _p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), _p(C), _dt(C),
                   C.stride(0), _p(bias), M, N, K, flags, int(splitk), int(bn), epi, self.stream)

    def _conv_umma_ok(self, *ts):
        """implicit-GEMM tcgen05 convolution: bf16, contiguous, 16 B aligned, channels 64 / 128."""
        if self.gemm_path != "umma":
            return False
        for t in ts:
            if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.data_ptr() % 16:
                return self._miss("conv3x3", f"operand dtype {t.dtype} / layout")
            if t.dim() == 4 and t.shape[3] not in (64, 128):
                return self._miss("conv3x3", f"{t.shape[3]} channels (64 or 128 supported)")
        return True

    def _timed_call(self, key, name, *args, n_kernels=1):
        prof = self.prof
        if prof is None:
            return self._call(name, *args, n_kernels=n_kernels)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._call(name, *args, n_kernels=n_kernels)
        e1.record()
        prof.setdefault(key, []).append((e0, e1))

    @staticmethod
    def _wgrad_splitk(rows_out, cols_out, k_red):
        """Weight-gradient GEMMs run on the engine's side stream, next to the dgrad chain.  They are bound by
        the L2->shared-memory operand traffic: one CTA sustains only ~50-60 GB/s, the chip ~6 TB/s, so split
        the reduction until 128-192 CTAs pull operands (fp32 vector reductions combine the slices), but keep
        at least 8 k-blocks per slice."""
        tiles = ((rows_out + 127) // 128) * ((cols_out + 127) // 128)
        kb = (k_red + 63) // 64
        return max(1, min(kb // 8, 192 // max(tiles, 1)))

    def linear_fwd(self, x, w, bias, y, relu=False, dropout=None):
        """y[M,N] = x[M,K] @ w[N,K]^T + bias (nn.Linear forward); optional fused ReLU and dropout
        (dropout = (p, seed, site), identical to a following self.dropout(y, p, seed, site))."""
        M, K = x.shape
        N = w.shape[0]
        assert w.shape[1] == K and y.shape == (M, N) and x.stride(1) == 1 and w.stride(1) == 1 and y.stride(1) == 1
        p, seed, site = dropout if dropout is not None else (0.0, 0, 0)
        if self._umma_ok(x, w):
            if p > 0.0 and y.is_contiguous() and y.dtype == torch.bfloat16:
                return self.umma_gemm(x, 0, w, 0, y, bias, M, N, K, GEMM_RELU if relu else 0, p_drop=p, seed=seed, site=site)
            self.umma_gemm(x, 0, w, 0, y, bias, M, N, K, GEMM_RELU if relu else 0)
        else:
            self.gemm(x, x.stride(0), 1, w, w.stride(0), 1, y, y.stride(0), bias, M, N, K, GEMM_RELU if relu else 0)
        if p > 0.0:
            self.dropout(y, p, seed, site)

    def linear_dgrad(self, dy, w, dx, accumulate=False, relu_drop_mask=None, p=0.0, rowdot=None):
        """dx[M,K] (+)= dy[M,N] @ w[N,K].  relu_drop_mask = the stored forward output f of ReLU followed by
        dropout(p) whose gradient dx is: the backward of both is fused as dx = f > 0 ? dx / (1-p) : 0
        (an element of f is positive iff it passed the ReLU and was kept)."""
        M, N = dy.shape
        K = w.shape[1]
        assert dx.shape == (M, K) and dy.stride(1) == 1 and w.stride(1) == 1 and dx.stride(1) == 1
        assert relu_drop_mask is None or not accumulate
        scale = 1.0 / (1.0 - p) if p > 0.0 else 1.0
        if rowdot is not None:
            # rowdot = (o [M, H*64] bf16, dsum [B*H*L] fp32, L, H): D = rowsum(dx . o) per head, written by the GEMM
            # epilogue (attention backward); returns True when it was produced, False when the caller must compute it
            o_, ds_, L_, H_ = rowdot
            if (self._umma_ok(dy, w) and dx.dtype == torch.bfloat16 and not accumulate and relu_drop_mask is None
                    and o_.dtype == torch.bfloat16 and o_.stride(1) == 1 and o_.stride(0) % 8 == 0 and o_.data_ptr() % 16 == 0
                    and K == H_ * 64 and M % L_ == 0):
                self.umma_gemm(dy, 0, w, 1, dx, None, M, K, N, 0, rowdot=rowdot)
                return True
            self.linear_dgrad(dy, w, dx)
            return False
        if self._umma_ok(dy, w):
            if relu_drop_mask is not None and dx.dtype == torch.bfloat16 and relu_drop_mask.dtype == torch.bfloat16:
                return self.umma_gemm(dy, 0, w, 1, dx, None, M, K, N, 0, mask=relu_drop_mask, mask_scale=scale)
            self.umma_gemm(dy, 0, w, 1, dx, None, M, K, N, GEMM_ACCUM if accumulate else 0)
        else:
            self.gemm(dy, dy.stride(0), 1, w, 1, w.stride(0), dx, dx.stride(0), None, M, K, N,
                      GEMM_ACCUM if accumulate else 0)
        if relu_drop_mask is not None:
            self.relu_bwd(relu_drop_mask, dx, scale)

    def gemm_group(self, calls):
        """Run the callables (each issuing linear_wgrad-style GEMMs) as ONE grouped launch per kernel instantiation
        (masr_gemm_group_begin / _end): the operands of every recorded problem must stay untouched until this returns."""
        rc = self.lib.masr_gemm_group_begin()
        if rc != 0:
            _lib.check(rc, "masr_gemm_group_begin")
        before = self.launches
        try:
            for fn in calls:
                fn()
        finally:
            rc = self.lib.masr_gemm_group_end(self.stream)
        if rc != 0:
            _lib.check(rc, "masr_gemm_group_end")
        # launch accounting: the recorded problems were counted one launch each by _call; replace by what was launched
        rec, lau = _lib.C.c_int(0), _lib.C.c_int(0)
        self.lib.masr_gemm_group_last(_lib.C.byref(rec), _lib.C.byref(lau))
        self.launches += lau.value - rec.value

    def linear_wgrad(self, x, dy, dw, db):
        """dw[N,K] += dy[M,N]^T @ x[M,K] (fp32); db[N] += column sums of dy."""
        M, K = x.shape
        N = dy.shape[1]
        assert dw.shape == (N, K) and dw.dtype == torch.float32 and dw.stride(1) == 1
        if self._umma_ok(dy, x):
            sk = self._wgrad_splitk(N, K, M)
            # bias gradient = row sums of dy^T: a second tensor-core accumulator of the same kernel
            return self.umma_gemm(dy, 1, x, 1, dw, None, N, K, M, GEMM_SPLITK if sk > 1 else GEMM_ACCUM, sk, rowsum=db)
        else:
            tiles = ((N + 127) // 128) * ((K + 127) // 128)
            splitk = max(1, min((M + 255) // 256, (4 * 148 + tiles - 1) // tiles))
            self.gemm(dy, 1, dy.stride(0), x, 1, x.stride(0), dw, dw.stride(0), None, N, K, M, GEMM_SPLITK, splitk)
        if db is not None:
            self.colsum_add(dy, db)

    # -------------------------------------------------------------- conv front end
    def _conv1_umma_ok(self, x, act, Cout):
        if self.gemm_path != "umma":
            return False
        ok = (Cout == 64 and act.dtype == torch.bfloat16 and act.is_contiguous()
              and x.dtype == torch.float32 and x.is_contiguous() and act.data_ptr() % 16 == 0)
        return ok or self._miss("conv1", f"Cout {Cout}, activation dtype {act.dtype}")

    def conv1_fwd(self, x, w, bias, y):
        B, H, W = x.shape
        if self._conv1_umma_ok(x, y, w.shape[0]):
            return self._call("masr_umma_conv1_fwd", _p(x), _p(w), _p(bias), _p(y), B, H, W, w.shape[0], self.stream)
        self._call("masr_conv1_fwd", _p(x), _p(w), _p(bias), _p(y), _dt(y), B, H, W, w.shape[0], self.stream)

    def conv1_wgrad(self, x, dy, dw, db):
        B, H, W = x.shape
        if self._conv1_umma_ok(x, dy, dw.shape[0]):
            return self._call("masr_umma_conv1_wgrad", _p(x), _p(dy), _p(dw), _p(db), B, H, W, dw.shape[0], self.stream)
        self._call("masr_conv1_wgrad", _p(x), _p(dy), _dt(dy), _p(dw), _p(db), B, H, W, dw.shape[0], self.stream)

    def conv_w_prep(self, w, wp):
        self._call("masr_conv_w_prep", _p(w), _p(wp), _dt(wp), w.shape[0], w.shape[1], self.stream)

    def conv_w_prep_t(self, w, wpt):
        """wpt [Cin, 9*Cout]: transposed operand layout for the dgrad implicit GEMM."""
        self._call("masr_conv_w_prep_t", _p(w), _p(wpt), _dt(wpt), w.shape[0], w.shape[1], self.stream)

    def conv_w_unprep_add(self, dwp, dw):
        self._call("masr_conv_w_unprep_add", _p(dwp), _p(dw), dw.shape[0], dw.shape[1], self.stream)

    def _im2col(self, x):
        B, H, W, Cin = x.shape
        # scratch is private to the launching lane (the engine forks weight-gradient work to a side stream)
        col = self.scratch(("col", self.scratch_tag), B * H * W * 9 * Cin, x.dtype).view(B * H * W, 9 * Cin)
        self._call("masr_im2col3x3", _p(x), _p(col), _dt(x), B, H, W, Cin, self.stream)
        return col

    def conv3x3_fwd(self, x, wp, bias, y):
        """y = relu(conv3x3(x) + bias), NHWC; wp [Cout, 9*Cin]."""
        B, H, W, Cin = x.shape
        Cout = wp.shape[0]
        if self._conv_umma_ok(x, wp, y):
            return self._timed_call(("conv_fwd", B * H * W, Cout, 9 * Cin), "masr_umma_conv3x3_fwd", _p(x), _p(wp),
                                    _p(bias), _p(y), B, H, W, Cin, Cout, 1, self.stream)
        col = self._im2col(x)
        y2 = y.view(B * H * W, Cout)
        if self._umma_ok(col, wp):
            return self.umma_gemm(col, 0, wp, 0, y2, bias, B * H * W, Cout, 9 * Cin, GEMM_RELU)
        self.gemm(col, 9 * Cin, 1, wp, 9 * Cin, 1, y, Cout, bias, B * H * W, Cout, 9 * Cin, GEMM_RELU)

    def conv3x3_dgrad(self, dy, wp, dx, relu_src=None, wpt=None):
        """dx = conv3x3^T(dy) (times (relu_src > 0) when given); wpt: optional conv_w_prep_t layout."""
        B, H, W, Cout = dy.shape
        Cin = dx.shape[3]
        P = B * H * W
        if self._conv_umma_ok(dy, wp, dx) and (relu_src is None or relu_src.is_contiguous()):
            return self._timed_call(("conv_dgrad", P, Cin, 9 * Cout), "masr_umma_conv3x3_dgrad", _p(dy), _p(wp), _p(wpt),
                                    _p(dx), _p(relu_src), B, H, W, Cin, Cout, self.stream)
        dcol = self.scratch(("col", self.scratch_tag), P * 9 * Cin, dy.dtype).view(P, 9 * Cin)
        dy2 = dy.view(P, Cout)
        if self._umma_ok(dy2, wp):
            self.umma_gemm(dy2, 0, wp, 1, dcol, None, P, 9 * Cin, Cout, 0)
        else:
            self.gemm(dy, Cout, 1, wp, 1, 9 * Cin, dcol, 9 * Cin, None, P, 9 * Cin, Cout, 0)
        self._call("masr_col2im3x3", _p(dcol), _p(dx), _dt(dx), _p(relu_src), B, H, W, Cin, self.stream)

    def conv3x3_wgrad(self, x, dy, dwp, db):
        """dwp[Cout, 9*Cin] += dy^T @ im2col(x) (fp32); db += column sums of dy."""
        B, H, W, Cin = x.shape
        Cout = dy.shape[3]
        P = B * H * W
        if self._conv_umma_ok(x, dy) and dwp.is_contiguous():
            return self._timed_call(("conv_wgrad", Cout, 9 * Cin, P), "masr_umma_conv3x3_wgrad", _p(x), _p(dy), _p(dwp),
                                    _p(db), B, H, W, Cin, Cout, self.stream)
        col = self._im2col(x)
        dy2 = dy.view(P, Cout)
        if self._umma_ok(dy2, col):
            sk = self._wgrad_splitk(Cout, 9 * Cin, P)
            self.umma_gemm(dy2, 1, col, 1, dwp, None, Cout, 9 * Cin, P, GEMM_SPLITK if sk > 1 else GEMM_ACCUM, sk)
        else:
            tiles = ((Cout + 127) // 128) * ((9 * Cin + 127) // 128)
            splitk = max(1, min((P + 511) // 512, (4 * 148 + tiles - 1) // tiles))
            self.gemm(dy, 1, Cout, col, 1, 9 * Cin, dwp, 9 * Cin, None, Cout, 9 * Cin, P, GEMM_SPLITK, splitk)
        self.colsum_add(dy.view(P, Cout), db)

    pool_codes = True         # the engine may ask for the arg-max code bytes (maxpool_fwd(code=) / maxpool_bwd(code=))

    @staticmethod
    def _pool_code_ok(x, *ts):
        return (x.shape[3] % 8 == 0 and x.shape[1] >= 2 and x.shape[2] >= 2 and x.shape[0] > 0
                and all(t.data_ptr() % 16 == 0 for t in (x,) + ts))

    def maxpool_fwd(self, x, y, code=None):
        """code (optional, uint8 [B, H/2, W/2, C]): per-output arg-max / positive-maximum byte for maxpool_bwd(code=)."""
        B, H, W, Cc = x.shape
        if code is not None and self._pool_code_ok(x, y, code):
            return self._call("masr_maxpool2x2_fwd_code", _p(x), _p(y), _p(code), _dt(x), B, H, W, Cc, self.stream)
        self._call("masr_maxpool2x2_fwd", _p(x), _p(y), _dt(x), B, H, W, Cc, self.stream)

    def maxpool_bwd(self, x, dy, dx, relu_mask=True, code=None):
        B, H, W, Cc = x.shape
        if code is not None and self._pool_code_ok(x, dy, dx, code):
            return self._call("masr_maxpool2x2_bwd_code", _p(code), _p(dy), _p(dx), _dt(x), int(relu_mask), B, H, W, Cc, self.stream)
        self._call("masr_maxpool2x2_bwd", _p(x), _p(dy), _p(dx), _dt(x), int(relu_mask), B, H, W, Cc, self.stream)

    def relu_bwd(self, y, dx, scale=1.0):
        self._call("masr_relu_bwd", _p(y), _p(dx), _dt(y), y.numel(), float(scale), self.stream)

    # -------------------------------------------------------------- attention
    def _attn_umma_ok(self, hd, *ts):
        """tcgen05 attention: bf16, head dim 64, 16 B aligned rows."""
        if self.gemm_path != "umma":
            return False
        if hd != 64:
            return self._miss("attention", f"head dim {hd} (64 supported)")
        ok = all(t.dtype == torch.bfloat16 and t.data_ptr() % 16 == 0 and t.stride(0) % 8 == 0 and t.stride(1) == 1
                 for t in ts)
        return ok or self._miss("attention", "operand dtype / alignment")

    attn_small_lq = 64        # query sequences up to this length run on the warp-MMA attention kernels (attn_small.cu)

    def set_attn_small_lq(self, max_lq):
        """Process-wide (library global): 0 sends every bf16 attention problem to the tcgen05 kernels."""
        self.lib.masr_attn_set_small_lq(int(max_lq))
        CudaBackend.attn_small_lq = max(0, min(64, int(max_lq)))

    def attn_fwd(self, q, k, v, out, lse, B, H, Lq, Lk, klens, causal, p=0.0, seed=0, site=0, kv_rows=None):
        """kv_rows: rows per utterance in k / v when they are a decode cache of fixed capacity (default Lk)."""
        hd = out.shape[1] // H
        kv_rows = Lk if kv_rows is None else int(kv_rows)
        if self._attn_umma_ok(hd, q, k, v, out):
            return self._timed_call(("attn_fwd", B * H, Lq, Lk), "masr_umma_attn_fwd_cached", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                              out.stride(0), _p(lse), B, H, Lq, Lk, kv_rows, _p(klens), int(causal), float(p), seed, site,
                              self.stream)
        self._call("masr_attn_fwd_cached", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                   _p(lse), _dt(q), B, H, Lq, Lk, kv_rows, hd, _p(klens), int(causal), float(p), seed, site, self.stream)

    def attn_bwd(self, q, k, v, out, dout, lse, dsum, dq, dk, dv, B, H, Lq, Lk, klens, causal, p=0.0, seed=0, site=0,
                 dsum_ready=False):
        hd = out.shape[1] // H
        if self._attn_umma_ok(hd, q, k, v, out, dout, dq, dk, dv):
            # memories longer than one 128-key tile: the key tiles reduce their dQ partials in an fp32 workspace
            dq_ws = self.scratch(("attn_dq", self.scratch_tag), B * Lq * H * 64, torch.float32) if Lk > 128 else None
            return self._timed_call(("attn_bwd", B * H, Lq, Lk), "masr_umma_attn_bwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                              out.stride(0), _p(dout), dout.stride(0), _p(lse), _p(dsum), _p(dq), dq.stride(0),
                              _p(dk), dk.stride(0), _p(dv), dv.stride(0), B, H, Lq, Lk, _p(klens), int(causal),
                              float(p), seed, site, int(dsum_ready), _p(dq_ws), self.stream,
                              n_kernels=1 if Lq <= self.attn_small_lq else (1 if dsum_ready else 2) + (2 if Lk > 128 else 0))
        self._call("masr_attn_bwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                   _p(dout), dout.stride(0), _p(lse), _p(dsum), _p(dq), dq.stride(0), _p(dk), dk.stride(0),
                   _p(dv), dv.stride(0), _dt(q), B, H, Lq, Lk, hd, _p(klens), int(causal), float(p), seed, site,
                   self.stream, n_kernels=2)

    # -------------------------------------------------------------- fused elementwise
    def add_layernorm_fwd(self, x, res, gamma, beta, y, mean, rstd, p=0.0, seed=0, site=0, eps=1e-5):
        rows, d = x.shape
        self._call("masr_add_layernorm_fwd", _p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), _dt(x),
                   rows, d, eps, float(p), seed, site, self.stream)

    def add_layernorm_bwd(self, dy, s, mean, rstd, gamma, ds, ds_accum, dx, dgamma, dbeta, p=0.0, seed=0, site=0):
        rows, d = dy.shape
        self._call("masr_add_layernorm_bwd", _p(dy), _p(s), _p(mean), _p(rstd), _p(gamma), _p(ds), int(ds_accum),
                   _p(dx), _p(dgamma), _p(dbeta), _dt(dy), rows, d, float(p), seed, site, self.stream)

    def add_pe_dropout(self, x, pe, L, p=0.0, seed=0, site=0):
        rows, d = x.shape
        self._call("masr_add_pe_dropout", _p(x), _p(pe), _dt(x), rows, L, d, float(p), seed, site, self.stream)

    def embed_pe_fwd(self, ids, E, pe, out, L, p=0.0, seed=0, site=0):
        rows, d = out.shape
        self._call("masr_embed_pe_fwd", _p(ids), _p(E), _p(pe), _p(out), _dt(out), rows, L, d, float(p), seed, site,
                   self.stream)

    def embed_bwd(self, ids, dout, dE, L, p=0.0, seed=0, site=0):
        rows, d = dout.shape
        self._call("masr_embed_bwd", _p(ids), _p(dout), _dt(dout), _p(dE), rows, L, d, float(p), seed, site, self.stream)

    def dropout(self, x, p, seed, site):
        if p > 0.0:
            self._call("masr_dropout", _p(x), _dt(x), x.numel(), float(p), seed, site, self.stream)

    def colsum_add(self, x, out):
        M, N = x.shape
        self._call("masr_colsum_add", _p(x), _dt(x), x.stride(0), _p(out), M, N, self.stream)

    def cast(self, src, dst):
        self._call("masr_cast", _p(src), _dt(src), _p(dst), _dt(dst), src.numel(), self.stream)

    def permute_cf(self, src, dst, Cc, Fq, inverse_add=False):
        rows = src.shape[0]
        self._call("masr_permute_cf", _p(src), _dt(src), _p(dst), _dt(dst), rows, Cc, Fq, int(inverse_add), self.stream)

    def ls_ce(self, logits, gold, eps, inv_n, stats, argmax, dlogits, inv_n_dev=None):
        N, Cc = logits.shape
        self._call("masr_ls_ce_fwd_bwd", _p(logits), _p(gold), N, Cc, float(eps), float(inv_n), _p(inv_n_dev), _p(stats),
                   _p(argmax), _p(dlogits), _dt(dlogits) if dlogits is not None else F32,
                   dlogits.stride(0) if dlogits is not None else Cc, self.stream)

    # -------------------------------------------------------------- joint CTC / attention (ctc_weight extension)
    def ctc_joint(self, logits, B, T, Cc, targets, offs, in_lens, tgt_lens, lmax, w, nll, loss, grad):
        """Kernel 1 on the batch-first CTC-head output logits [B*T, C] fp32 (row b*T + t), log-softmax fused, blank 0,
        reduction 'mean', zero_infinity; grad (or None) = w * d loss / d logits in the same layout."""
        wsb = self.lib.masr_ctc_workspace_bytes(T, B, Cc, lmax)
        ws = self.scratch(("ctc_ws", self.scratch_tag), wsb // 4 + 1, torch.float32) if wsb else None
        self._call("masr_ctc_fwd_bwd_ex", _p(logits), T, B, Cc, logits.stride(0), T * logits.stride(0), 0, _p(targets),
                   _p(offs), _p(in_lens), _p(tgt_lens), int(lmax), 0, 1, float(w), _p(nll), _p(loss), _p(grad), _p(ws), wsb,
                   self.stream, n_kernels=2)

    def cast_pad2d(self, src, dst, cols):
        self._call("masr_cast_pad2d", _p(src), _dt(src), src.stride(0), _p(dst), _dt(dst), dst.stride(0), src.shape[0],
                   int(cols), self.stream)

    def loss_mix(self, stats, ctc_loss, w):
        self._call("masr_loss_mix", _p(stats), _p(ctc_loss), float(w), self.stream)

    def set_seed_ptr(self, t):
        """Device-resident dropout seed offset (uint64 stored in an int64 tensor); see masr_set_seed_ptr."""
        self._seed_t = t
        rc = self.lib.masr_set_seed_ptr(_p(t))
        if rc != 0:
            _lib.check(rc, "masr_set_seed_ptr")

    def seed_bump(self, inc=1):
        self._call("masr_seed_bump", _p(self._seed_t), int(inc), self.stream)

    def zero_(self, t):
        """Device memset through the CUDA runtime (stream-ordered); counts as plumbing, not a kernel of ours."""
        t.zero_()

    # -------------------------------------------------------------- kernel 4: flat arena ops
    def mt_sumsq(self, g, out, zero_first=True):
        self._call("masr_mt_sumsq", _p(g), g.numel(), _p(out), int(zero_first), self.stream, n_kernels=2 if zero_first else 1)

    def mt_clip_sgd(self, p, g, buf, sumsq, max_norm, lr, momentum, nesterov, first_step):
        self._call("masr_mt_clip_sgd", _p(p), _p(g), _p(buf), p.numel(), _p(sumsq), float(max_norm), float(lr),
                   float(momentum), int(nesterov), int(first_step), self.stream)

    def mt_clip_sgd_ex(self, p, g, buf, sumsq, max_norm, lr, momentum, nesterov, first_step, shadow=None, last_step=False):
        """clip + SGD that also writes the bf16 shadow of the arena (and, on a task's last inner step, skips the dead
        write-backs of the scaled gradient and the momentum)."""
        if shadow is not None and shadow.dtype != torch.bfloat16:
            shadow = None
        self._call("masr_mt_clip_sgd_ex", _p(p), _p(g), _p(buf), p.numel(), _p(sumsq), float(max_norm), float(lr),
                   float(momentum), int(nesterov), int(first_step), _p(shadow), 1 if last_step else 0, self.stream)
        return shadow is not None

    def mt_copy_cast(self, dst, shadow, src):
        """dst = src (fp32 arenas) and shadow = bf16(src) in one pass; returns True when the shadow was written."""
        if shadow is not None and shadow.dtype != torch.bfloat16:
            shadow = None
        self._call("masr_mt_copy_cast", _p(dst), _p(shadow), _p(src), src.numel(), self.stream)
        return shadow is not None

    def prep_weights(self, params, shadow, jobs, v2e, v2e_p, Cc, Fq, dtype_of):
        """All derived weight copies in one launch (masr_prep_weights).  jobs: [(w fp32 [Cout,Cin,3,3], wp, wpt or None)];
        shadow None = the arena's compute-dtype copy is fresh already."""
        arr = (_lib.ConvPrepJob * max(len(jobs), 1))()
        for i, (w, wp, wpt) in enumerate(jobs):
            arr[i].w, arr[i].wp, arr[i].wpt = _p(w), _p(wp), _p(wpt)
            arr[i].Cout, arr[i].Cin = w.shape[0], w.shape[1]
        self._call("masr_prep_weights", _p(params), _p(shadow), params.numel(), arr, len(jobs), _p(v2e), _p(v2e_p),
                   v2e.shape[0], int(Cc), int(Fq), _dt(dtype_of), self.stream)

    def mt_clip(self, g, sumsq, max_norm):
        self._call("masr_mt_clip", _p(g), g.numel(), _p(sumsq), float(max_norm), self.stream)

    def mt_accumulate(self, upd, g, sumsq=None, max_norm=0.0):
        self._call("masr_mt_accumulate", _p(upd), _p(g), g.numel(), _p(sumsq), float(max_norm), self.stream)

    def mt_reptile_delta(self, upd, theta, phi):
        self._call("masr_mt_reptile_delta", _p(upd), _p(theta), _p(phi), upd.numel(), self.stream)

    def mt_adam(self, p, m, v, upd, count, lr, beta1, beta2, eps, bc1, bc2, skip_if_nan=None, clip_sumsq=None, max_norm=0.0):
        self._call("masr_mt_adam", _p(p), _p(m), _p(v), _p(upd), p.numel(), float(count), float(lr), float(beta1),
                   float(beta2), float(eps), float(bc1), float(bc2), _p(skip_if_nan), _p(clip_sumsq), float(max_norm),
                   self.stream)

    def mt_axpy(self, y, x, a):
        self._call("masr_mt_axpy", _p(y), _p(x), float(a), y.numel(), self.stream)

    def copy_(self, dst, src):
        """Flat device-to-device copy (cudaMemcpyAsync under torch); replaces the 114 per-tensor
        copies of load_state_dict(_original), fo_meta_interface.py:226."""
        dst.copy_(src, non_blocking=True)
