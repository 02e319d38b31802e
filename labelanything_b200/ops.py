"""Thin torch-tensor wrappers over the C ABI (include/labelanything_b200.h), registered as torch custom ops.

PyTorch is used for device memory and the current stream only; every function here validates shapes/dtypes,
hands raw pointers to the native library and raises RuntimeError on failure.  Nothing falls back to torch math.
Each entry point is also a `torch.library` custom op with a fake implementation (`torch.ops.labelanything_b200.la_*`),
which is what `torch.compile(model)` records (label_anything/experiment/run.py:167-169 compiles the model on request).
"""
from __future__ import annotations

import torch

from . import _native

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
DT_BF16, DT_F32, DT_F16 = 0, 1, 2
_DTC = {torch.bfloat16: DT_BF16, torch.float32: DT_F32, torch.float16: DT_F16}


class OpProfiler:
    """Per-launch CUDA-event timing of the native ops on the launching (= torch's current) stream.

    bench.py installs one over the timed region (`with ops.profile() as prof`) to get each kernel family's share of
    the step and its achieved FLOP/s / GB/s from ALGORITHMIC work (the `cost` the wrappers declare), which is
    what the roofline block of the bench line reports.  Off by default; costs two event records per launch."""

    def __init__(self) -> None:
        self.records: list = []   # (family, flops, bytes, start_event, end_event)
        self.launches = 0

    def summary(self) -> dict:
        fam: dict = {}
        for name, flops, nbytes, e0, e1 in self.records:
            f = fam.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            f["launches"] += 1
            f["ms"] += e0.elapsed_time(e1)
            f["flops"] += flops
            f["bytes"] += nbytes
        return fam


_PROF: OpProfiler | None = None
_COST = (0.0, 0.0)  # (flops, bytes) of the next native call, declared by the wrapper


class profile:
    def __enter__(self) -> OpProfiler:
        global _PROF
        self.prev = _PROF
        _PROF = OpProfiler()
        return _PROF

    def __exit__(self, *exc) -> None:
        global _PROF
        _PROF = self.prev


def _cost(flops: float = 0.0, nbytes: float = 0.0) -> None:
    global _COST
    if _PROF is not None:
        _COST = (float(flops), float(nbytes))


# ------------------------------------------------------------------------------------------------------------------
# The C ABI as torch custom ops.
#
# Every `int la_xxx(void* stream, ...)` entry point of include/labelanything_b200.h is registered as the custom op
# `labelanything_b200::la_xxx` (torch.library): pointer parameters become `Tensor?` arguments (non-const pointers are
# declared as mutated, `Tensor(a!)?`), integers / floats stay scalars, the stream parameter disappears (it is torch's
# current stream on the device of the first CUDA tensor argument).  All outputs are allocated by the Python wrappers
# below and mutated in place, so every op returns nothing and its fake (meta) implementation is a no-op: FakeTensor
# tracing, `torch.compile` (dynamo + AOT functionalisation re-inplaces the mutable ops) and `torch.export` see the
# launch sequence as ordinary graph nodes instead of opaque ctypes calls.  The schema is derived from the header, so the
# header stays the single source of truth for the boundary.
# ------------------------------------------------------------------------------------------------------------------
_OPS_NS = "labelanything_b200"
_OP_PARAMS: dict = {}     # op name -> [(kind, c type)], kind in {"tensor", "int", "float"}; stream parameter dropped
_op_lib = None


def _schema_of(name: str, types: list) -> tuple:
    params, parts, alias = [], [], 0
    for i, t in enumerate(types[1:]):          # types[0] is `void* stream`
        if "*" in t:
            if t.startswith("const"):
                parts.append(f"Tensor? a{i}")
            else:
                parts.append(f"Tensor({chr(ord('a') + alias)}!)? a{i}")
                alias += 1
            params.append(("tensor", t))
        elif t in ("float", "double"):
            parts.append(f"float a{i}")
            params.append(("float", t))
        else:
            parts.append(f"int a{i}")
            params.append(("int", t))
    return f"{name}({', '.join(parts)}) -> ()", params


def _launch(name: str, args) -> int:
    """Raw launch: tensors -> device pointers, torch's current stream of the first CUDA tensor's device, that device made
    current for the call (the C ABI launches on, and queries, the calling thread's current device)."""
    fn = getattr(_native.lib(), name)
    dev = None
    cargs = []
    for (kind, _), a in zip(_OP_PARAMS[name], args):
        if kind == "tensor":
            if a is None:
                cargs.append(None)
            else:
                if dev is None and a.is_cuda:
                    dev = a.device
                cargs.append(a.data_ptr())
        else:
            cargs.append(a)
    if dev is None:
        raise RuntimeError(f"labelanything_b200.{name}: no CUDA tensor among the arguments; there is no CPU fallback")
    stream = torch.cuda.current_stream(dev).cuda_stream
    if dev.index != torch.cuda.current_device():
        # tensors on a device that is not the thread's current one (model.to("cuda:1") in a single process): the stream
        # handle belongs to that device, so the launch, sm_count() and cudaFuncSetAttribute must run there too
        with torch.cuda.device(dev):
            return fn(stream, *cargs)
    return fn(stream, *cargs)


def _register_ops() -> None:
    global _op_lib
    if _op_lib is not None:
        return
    lib = torch.library.Library(_OPS_NS, "DEF")
    for name, (ret, types) in _native.declared_functions().items():
        if ret != "int" or not types or types[0] != "void*" or name == "la_attention_set_trace":
            continue                            # queries (la_version, *_workspace_bytes, la_last_error) are not launches
        schema, params = _schema_of(name, types)
        _OP_PARAMS[name] = params
        lib.define(schema)

        def impl(*args, _name=name):
            _native.check(_launch(_name, args), _name)

        lib.impl(name, impl, "CompositeExplicitAutograd")
        torch.library.register_fake(f"{_OPS_NS}::{name}", lambda *args: None, lib=lib)
    _op_lib = lib


_register_ops()


def _call(what: str, fn: str, *args) -> None:
    """Invoke one C-ABI entry point (see the block comment above); raise RuntimeError with the library's message on
    failure.  Eager calls go straight to the launch; under torch.compile tracing the registered custom op is called, so
    the launch becomes a graph node."""
    global _COST
    if torch.compiler.is_compiling():
        getattr(getattr(torch.ops, _OPS_NS), fn)(*args)
        return
    prof = _PROF
    if prof is None:
        rc = _launch(fn, args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _launch(fn, args)
        e1.record()
        prof.records.append((what, _COST[0], _COST[1], e0, e1))
        prof.launches += 1
        _COST = (0.0, 0.0)
    _native.check(rc, what)


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "labelanything_b200 runs on CUDA (sm_100a) tensors only; got a tensor on "
                f"{t.device}. There is no CPU fallback."
            )


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act: int = ACT_NONE,
         out_dtype: torch.dtype = torch.bfloat16, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = act(a @ w.T + bias).  a [M,K] bf16 (row stride free), w [N,K] bf16, bias [N] fp32."""
    _require_cuda(a, w, bias, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, (a.dtype, w.dtype)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in _DTC
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    _cost(2.0 * M * N * K, 2.0 * (M * K + N * K) + M * N * out.element_size())
    _call(f"gemm.n{N}.k{K}" if _PROF is not None else "gemm", "la_gemm_bf16", a, a.stride(0), w, w.stride(0),
        bias if bias is not None else None, out, out.stride(0),
        _DTC[out.dtype], M, N, K, act)
    return out


def gemm_splitk(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 [M, N] = a @ w.T + bias with the contraction split over the SMs (few output tiles, long K: the weight
    gradients of the training step).  a [M, K] bf16, w [N, K] bf16; partial tiles are added by TMA reduce stores."""
    _require_cuda(a, w, bias)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous())
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _cost(2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N)
    _call(f"gemm_splitk.n{N}.k{K}" if _PROF is not None else "gemm_splitk", "la_gemm_bf16_splitk", a, a.stride(0), w,
          w.stride(0), bias, out, out.stride(0), M, N, K)
    return out


def gemm_to_grid(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, grid: int, padded: int) -> torch.Tensor:
    """bf16 [images * padded^2, N]: row (image, y, x) of a @ w.T + bias at position (image, y, x) of a padded x padded
    grid, the positions outside the grid x grid part = the bias row (the projection of the reference's zero padding)."""
    _require_cuda(a, w, bias)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert grid % 32 == 0 and padded >= grid and M % (grid * grid) == 0, (M, grid, padded)
    n_img = M // (grid * grid)
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous())
    out = torch.empty((n_img * padded * padded, N), dtype=torch.bfloat16, device=a.device)
    _cost(2.0 * M * N * K, 2.0 * (M * K + N * K) + 2.0 * out.numel())
    _call(f"gemm.n{N}.k{K}" if _PROF is not None else "gemm", "la_gemm_bf16_to_grid", a, a.stride(0), w, w.stride(0),
          bias, out, out.stride(0), M, N, K, grid, padded)
    return out


_NO_GEMM_ACCUMULATE = bool(__import__("os").environ.get("LA_NO_GEMM_ACCUMULATE"))   # experiment switch


def gemm_accumulate_supported(m: int, n: int) -> bool:
    return m >= 2048 and n >= 256 and n % 8 == 0 and not _NO_GEMM_ACCUMULATE


def gemm_accumulate(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, x: torch.Tensor) -> torch.Tensor:
    """x (fp32, in place) += a @ w.T + bias, the add done in fp32 in the GEMM epilogue."""
    _require_cuda(a, w, bias, x)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert x.dtype == torch.float32 and x.shape == (M, N) and x.stride(1) == 1 and gemm_accumulate_supported(M, N)
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous())
    _cost(2.0 * M * N * K, 2.0 * (M * K + N * K) + 8.0 * M * N)
    _call(f"gemm_acc.n{N}.k{K}" if _PROF is not None else "gemm", "la_gemm_bf16_accumulate", a,
          a.stride(0), w, w.stride(0), bias, x, x.stride(0), M, N, K)
    return x


def conv3x3_supported(h: int, w: int, c: int, n: int) -> bool:
    """Shapes the implicit-GEMM 3x3 convolution is built for (anything else goes through im2col_3x3 + gemm)."""
    return w == 64 and h % 4 == 0 and c % 64 == 0 and n >= 256 and n % 8 == 0


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, n_img: int, h: int, wd: int, c: int,
            act: int = ACT_NONE, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """3x3 / stride 1 / zero-pad 1 convolution of a token-major bf16 map [n_img*h*wd, c] with w [N, 9c] (column =
    (ky*3+kx)*c + ci) as an implicit GEMM (4-D TMA loads at shifted coordinates; no im2col matrix)."""
    _require_cuda(x, w, bias)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == n_img * h * wd * c
    assert w.dtype == torch.bfloat16 and w.dim() == 2 and w.shape[1] == 9 * c and w.stride(1) == 1
    N = w.shape[0]
    assert conv3x3_supported(h, wd, c, N), (h, wd, c, N)
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous())
    out = torch.empty((n_img * h * wd, N), dtype=out_dtype, device=x.device)
    M = n_img * h * wd
    _cost(2.0 * M * N * 9 * c, 2.0 * (M * c + N * 9 * c) + M * N * out.element_size())
    _call(f"conv3x3.n{N}.c{c}" if _PROF is not None else "conv3x3", "la_conv3x3_bf16", x, n_img, h,
          wd, c, w, w.stride(0), bias, out, out.stride(0), _DTC[out.dtype], N, act)
    return out


def attention_window(q: torch.Tensor, kv: torch.Tensor, n_seq: int, n_heads: int, scale: float, out: torch.Tensor,
                     q_off: int, k_off: int, v_off: int, rel_table: torch.Tensor, rel_pad: int, out_mode: int = 0,
                     nwin: int = 0, img_hw: int = 0, in_pad: int = 0) -> torch.Tensor:
    """Fused MHSA over 14x14 windows (196 tokens, head_dim 64) with the decomposed rel-pos bias formed in-kernel from
    rel_table = bf16 [2 * rel_pad, 64] (reversed rel_pos_h rows, then reversed rel_pos_w rows; see the C header).
    in_pad > 0: q / kv are the padded-grid projections of `gemm_to_grid` ([images * in_pad^2, ld])."""
    _require_cuda(q, kv, out, rel_table)
    for t in (q, kv, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    assert q.shape[0] == kv.shape[0]
    assert rel_table.dtype == torch.bfloat16 and rel_table.is_contiguous() and rel_table.shape == (2 * rel_pad, 64)
    L = 196
    _cost(4.0 * n_seq * n_heads * L * L * 64 + 2.0 * n_seq * n_heads * L * 64 * 2 * rel_pad,
          2.0 * 4 * n_seq * L * n_heads * 64)
    _call(f"attention.L{L}" if _PROF is not None else "attention", "la_attention_window_bf16", q,
          q.stride(0), q_off, kv, kv.stride(0), k_off, v_off, q.shape[0], n_seq, n_heads, float(scale),
          rel_table, rel_pad, out, out.stride(0), out_mode, nwin, img_hw, in_pad)
    return out


def attention(q: torch.Tensor, kv: torch.Tensor, n_seq: int, seq_len: int, n_heads: int, scale: float,
              out: torch.Tensor, q_off: int, k_off: int, v_off: int, bias_h: torch.Tensor | None = None,
              bias_w: torch.Tensor | None = None, grid_hw: int = 0, out_mode: int = 0, nwin: int = 0,
              img_hw: int = 0) -> torch.Tensor:
    """Fused MHSA (head_dim 64); q [rows, ld_q], kv [rows, ld_kv] bf16 (may be one packed buffer).
    bias_h / bias_w: fp32 or fp16 views [rows, heads, >=2g-1] sharing one row stride (see the C header)."""
    _require_cuda(q, kv, out, bias_h, bias_w)
    for t in (q, kv, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    assert q.shape[0] == kv.shape[0]
    ldb = 0
    if bias_h is not None:
        assert bias_h.dtype in (torch.float32, torch.float16) and bias_w.dtype == bias_h.dtype
        assert bias_h.dim() == 3 and bias_h.shape[:2] == (q.shape[0], n_heads) and bias_w.shape[:2] == bias_h.shape[:2]
        assert bias_h.stride(2) == 1 and bias_w.stride(2) == 1 and bias_h.stride() == bias_w.stride()
        ldb = bias_h.stride(1)
        assert bias_h.stride(0) == ldb * n_heads
    _cost(4.0 * n_seq * n_heads * seq_len * seq_len * 64, 2.0 * 4 * n_seq * seq_len * n_heads * 64)
    _call(f"attention.L{seq_len}" if _PROF is not None else "attention", "la_attention_bf16", q, q.stride(0), q_off, kv, kv.stride(0), k_off, v_off, q.shape[0], n_seq,
        seq_len, n_heads, float(scale), bias_h, bias_w,
        _DTC[bias_h.dtype] if bias_h is not None else DT_F32, ldb, grid_hw, out, out.stride(0),
        out_mode, nwin, img_hw)
    return out


def add_layernorm(x_in: torch.Tensor | None, delta: torch.Tensor | None, gamma: torch.Tensor | None,
                  beta: torch.Tensor | None, eps: float, *, rows: int, d: int, x_out: torch.Tensor | None = None,
                  y_out: torch.Tensor | None = None, y2_out: torch.Tensor | None = None,
                  pe: torch.Tensor | None = None, ype_out: torch.Tensor | None = None, act: int = ACT_NONE,
                  delta2: torch.Tensor | None = None, seq_add: torch.Tensor | None = None, seq_rows: int = 0,
                  x_mod: int = 0, map_mode: int = 0, seq_len: int = 0, win: int = 0, nwin: int = 0,
                  hw: int = 0) -> None:
    """x = x_in + delta + delta2 + seq_add[row // seq_rows] (-> x_out); y = act(LN(x)) -> y_out (bf16/fp32),
    y2_out (fp32), ype_out (bf16, y + pe)."""
    _require_cuda(x_in, delta, delta2, seq_add, gamma, beta, x_out, y_out, y2_out, pe, ype_out)
    ref = x_in if x_in is not None else delta
    for t in (x_in, x_out, gamma, beta, y2_out, pe, seq_add):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    for t in (delta, delta2, ype_out):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())
    assert y_out is None or (y_out.is_contiguous() and y_out.dtype in (torch.bfloat16, torch.float32))
    pe_mod = 0
    if ype_out is not None:
        assert pe is not None and pe.shape[-1] == d
        pe_mod = pe.numel() // d
    _cost(0.0, float(rows) * d * sum(sz for t, sz in ((x_in, 4 if x_mod == 0 else 0), (delta, 2), (delta2, 2), (x_out, 4),
                                                        (y_out, y_out.element_size() if y_out is not None else 0),
                                                        (y2_out, 4), (ype_out, 2)) if t is not None))
    _call(f"add_layernorm.d{d}.map{map_mode}" if _PROF is not None else "add_layernorm", "la_add_layernorm", x_in, x_mod, delta, delta2, seq_add, seq_rows, x_out,
        gamma, beta, float(eps), act, y_out,
        DT_F32 if (y_out is not None and y_out.dtype == torch.float32) else DT_BF16, y2_out, pe, pe_mod,
        ype_out, rows, d, map_mode, seq_len, win, nwin, hw)


def add_layernorm_meanpool(x_in: torch.Tensor | None, delta: torch.Tensor | None, gamma: torch.Tensor,
                           beta: torch.Tensor, eps: float, n_seq: int, rows_per_seq: int, d: int,
                           delta2: torch.Tensor | None = None, seq_add: torch.Tensor | None = None,
                           slices: int = 8) -> torch.Tensor:
    """mean over each sequence's rows of LN(x_in + delta + delta2 + seq_add[seq]) -> fp32 [n_seq, d]."""
    _require_cuda(x_in, delta, delta2, seq_add, gamma, beta)
    ref = x_in if x_in is not None else delta
    assert x_in is None or (x_in.dtype == torch.float32 and x_in.is_contiguous())
    for t in (delta, delta2):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())
    assert seq_add is None or (seq_add.dtype == torch.float32 and seq_add.is_contiguous())
    ws = torch.empty((n_seq * slices, d), dtype=torch.float32, device=ref.device)
    out = torch.empty((n_seq, d), dtype=torch.float32, device=ref.device)
    _cost(0.0, float(n_seq) * rows_per_seq * d * ((4 if x_in is not None else 0) + (2 if delta is not None else 0) +
                                                    (2 if delta2 is not None else 0)))
    _call("add_layernorm_meanpool", "la_add_layernorm_meanpool", x_in, delta, delta2, seq_add,
                                                 gamma, beta, float(eps), n_seq, rows_per_seq,
                                                 d, ws, slices, out)
    return out


def embed_tokens(patch: torch.Tensor, cls: torch.Tensor | None, pos: torch.Tensor | None, x: torch.Tensor,
                 n_img: int, tokens_per_img: int, n_cls: int, d: int) -> torch.Tensor:
    _require_cuda(patch, cls, pos, x)
    assert patch.dtype == torch.bfloat16 and patch.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous()
    _call("embed_tokens", "la_embed_tokens", patch, cls, pos, x, n_img,
                                       tokens_per_img, n_cls, d)
    return x


def im2col_patch16(images: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """images [I, C, S, S] fp32 -> [I*(S/16)^2, C*256] bf16."""
    _require_cuda(images, out)
    assert images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
    I, C, S, S2 = images.shape
    assert S == S2 and S % 16 == 0
    if out is None:
        out = torch.empty((I * (S // 16) ** 2, C * 256), dtype=torch.bfloat16, device=images.device)
    _cost(0.0, float(I) * C * S * S * 6)
    _call("im2col_patch16", "la_im2col_patch16", images, out, I, C, S)
    return out


def im2col_3x3(x: torch.Tensor, n_img: int, h: int, w: int, c: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """token-major bf16 [n_img*h*w, c] -> [n_img*h*w, 9c] bf16 (zero pad 1)."""
    _require_cuda(x, out)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == n_img * h * w * c
    if out is None:
        out = torch.empty((n_img * h * w, 9 * c), dtype=torch.bfloat16, device=x.device)
    _cost(0.0, float(n_img) * h * w * c * 2 * 10)
    _call("im2col_3x3", "la_im2col_3x3", x, out, n_img, h, w, c)
    return out


@torch.compiler.assume_constant_result
def _attention_tokens_workspace_bytes(n_seq: int, nq: int, nk: int, n_heads: int, head_dim: int) -> int:
    return int(_native.lib().la_attention_tokens_workspace_bytes(n_seq, nq, nk, n_heads, head_dim))


@torch.compiler.assume_constant_result
def _focal_loss_workspace_bytes() -> int:
    return int(_native.lib().la_focal_loss_workspace_bytes())


def attention_tokens(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_seq: int, nq: int, nk: int, n_heads: int,
                     head_dim: int, q_add: torch.Tensor | None = None, k_add: torch.Tensor | None = None,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """softmax((q + q_add)(k + k_add)^T / sqrt(head_dim)) v per (sequence, head).  q/k/v: bf16 2-D views (column
    slices of packed projection buffers are fine) with n_heads*head_dim columns; *_add: fp32 [nq|nk, heads*dh]."""
    _require_cuda(q, k, v, q_add, k_add, out)
    w = n_heads * head_dim
    for t, rows in ((q, n_seq * nq), (k, n_seq * nk), (v, n_seq * nk)):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.shape == (rows, w) and t.stride(1) == 1, (t.shape, rows, w)
    for t, rows in ((q_add, nq), (k_add, nk)):
        assert t is None or (t.dtype == torch.float32 and t.shape == (rows, w) and t.stride(1) == 1)
    if out is None:
        out = torch.empty((n_seq * nq, w), dtype=torch.bfloat16, device=q.device)
    assert out.dtype == torch.bfloat16 and out.shape == (n_seq * nq, w) and out.stride(1) == 1
    ws_bytes = _attention_tokens_workspace_bytes(n_seq, nq, nk, n_heads, head_dim)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if ws_bytes > 0 else None
    _cost(4.0 * n_seq * nq * nk * w, 2.0 * n_seq * (2 * nq + 2 * nk) * w)
    _call(f"attention_tokens.s{n_seq}.q{nq}.k{nk}.d{head_dim}" if _PROF is not None else "attention_tokens",
          "la_attention_tokens", q, q.stride(0), k, k.stride(0), v, v.stride(0), q_add,
        q_add.stride(0) if q_add is not None else 0, k_add, k_add.stride(0) if k_add is not None else 0,
        out, out.stride(0), n_seq, nq, nk, n_heads, head_dim, head_dim ** -0.5, ws)
    return out


def mask_downscale(masks: torch.Tensor, host_weights: dict) -> torch.Tensor:
    """masks fp32 [S, H, W] -> fp32 [S, H/4, W/4, 16].  host_weights: CPU fp32 tensors w0,b0,g1,be1,w3,b3,g2,be2 + eps."""
    _require_cuda(masks)
    assert masks.dtype == torch.float32 and masks.is_contiguous() and masks.dim() == 3
    S, H, W = masks.shape
    out = torch.empty((S, H // 4, W // 4, 16), dtype=torch.float32, device=masks.device)
    hw = host_weights
    for k in ("w0", "b0", "g1", "be1", "w3", "b3", "g2", "be2"):
        assert hw[k].device.type == "cpu" and hw[k].dtype == torch.float32 and hw[k].is_contiguous()
    assert hw["w0"].numel() == 16 and hw["w3"].numel() == 256, "mask_downscaling is built for mask_in_chans = 16"
    _cost(0.0, float(S) * H * W * 4 * 2)
    _call("mask_downscale", "la_mask_downscale", masks, out, S, H, W, hw["w0"], hw["b0"],
        hw["g1"], hw["be1"], float(hw["eps1"]), hw["w3"], hw["b3"],
        hw["g2"], hw["be2"], float(hw["eps2"]))
    return out


def resize_bilinear(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """token-major fp32 [n, h, w, c] -> [n, out_h, out_w, c] (F.interpolate bilinear, align_corners=False)."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, h, w, c = x.shape
    out = torch.empty((n, out_h, out_w, c), dtype=torch.float32, device=x.device)
    _call("resize_bilinear", "la_resize_bilinear", x, out, n, h, w, out_h, out_w, c)
    return out


def build_src(feat: torch.Tensor, m16: torch.Tensor | None, mask_flags: torch.Tensor | None, w6, b6, not_a_mask,
              no_mask, code: torch.Tensor | None, n_seq: int, tokens: int, d: int, n_classes: int, examples: int,
              feat_lead: int = 0, seq_offset: int = 0) -> torch.Tensor:
    """src rows (bf16 [n_seq*tokens, d]) of sequences [seq_offset, seq_offset + n_seq); seq_offset must be a multiple
    of examples * n_classes (whole episodes), m16 / mask_flags cover exactly these sequences."""
    _require_cuda(feat, m16, mask_flags, w6, b6, not_a_mask, no_mask, code)
    assert feat.dtype == torch.float32 and feat.is_contiguous()
    assert seq_offset % (examples * n_classes) == 0 and n_seq % n_classes == 0
    ep0 = seq_offset // (examples * n_classes)
    n_img_needed = (ep0 + -(-n_seq // (examples * n_classes))) * (examples + feat_lead)
    assert feat.numel() >= min(n_img_needed, feat.numel() // (tokens * d)) * tokens * d
    feat = feat.view(-1, d)[ep0 * (examples + feat_lead) * tokens:]
    assert m16 is None or (m16.dtype == torch.float32 and m16.is_contiguous() and m16.numel() == n_seq * tokens * 16)
    assert mask_flags is None or (mask_flags.dtype == torch.uint8 and mask_flags.is_contiguous() and mask_flags.numel() == n_seq)
    for t in (w6, b6, not_a_mask, no_mask, code):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    out = torch.empty((n_seq * tokens, d), dtype=torch.bfloat16, device=feat.device)
    _cost(0.0, float(n_seq) * tokens * (d * 2 + (64 if m16 is not None else 0)) + float(n_seq // n_classes) * tokens * d * 4)
    _call("build_src", "la_build_src", feat, m16, mask_flags, w6, b6,
                                    not_a_mask, no_mask, code, out, n_seq, tokens, d,
                                    n_classes, examples, feat_lead)
    return out


def embed_sparse(points, point_labels, boxes, box_flags, gauss, not_a_point, pe_table, n_seq: int, d: int,
                 image_w: int, image_h: int) -> torch.Tensor:
    """-> fp32 [n_seq, n, d];  points [S,P,2], point_labels [S,P], boxes [S,Bx,4], box_flags [S,Bx] (fp32) or None."""
    _require_cuda(points, point_labels, boxes, box_flags, gauss, not_a_point, pe_table)
    for t in (points, point_labels, boxes, box_flags, gauss, not_a_point, pe_table):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    P = points.shape[1] if points is not None else 0
    Bx = boxes.shape[1] if boxes is not None else 0
    n = (P + (0 if boxes is not None else 1) if points is not None else 0) + 2 * Bx
    out = torch.empty((n_seq, n, d), dtype=torch.float32, device=gauss.device)
    _call("embed_sparse", "la_embed_sparse", points, point_labels, P, boxes,
                                       box_flags, Bx, gauss, not_a_point,
                                       pe_table, out, n_seq, d, image_w, image_h)
    return out


def masked_mean(emb: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """emb fp32 [B,M,C,D], flags uint8 [B,M,C] -> fp32 [B,C,D]."""
    _require_cuda(emb, flags)
    assert emb.dtype == torch.float32 and emb.is_contiguous() and flags.dtype == torch.uint8 and flags.is_contiguous()
    B, M, C, D = emb.shape
    assert flags.shape == (B, M, C)
    out = torch.empty((B, C, D), dtype=torch.float32, device=emb.device)
    _call("masked_mean", "la_masked_mean", emb, flags, out, B, M, C, D)
    return out


def classify(x: torch.Tensor, cls: torch.Tensor, batch: int, pixels: int) -> torch.Tensor:
    """x bf16 [batch*pixels, dk], cls fp32 [batch, C, dk] -> fp32 [batch, C, pixels]."""
    _require_cuda(x, cls)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and cls.dtype == torch.float32 and cls.is_contiguous()
    C, dk = cls.shape[1], cls.shape[2]
    assert x.shape == (batch * pixels, dk) and cls.shape[0] == batch
    out = torch.empty((batch, C, pixels), dtype=torch.float32, device=x.device)
    _cost(2.0 * batch * pixels * C * dk, float(batch) * pixels * (dk * 2 + C * 4))
    _call("classify", "la_classify", x, cls, out, batch, pixels, C, dk)
    return out


def postprocess_masks(logits: torch.Tensor, sizes: torch.Tensor, flag_gts: torch.Tensor | None, image_size: int,
                      out_h: int, out_w: int) -> torch.Tensor:
    """logits fp32 [B,C,lh,lw]; sizes int32 [B,4] = (oh, ow, ih, iw) on device -> fp32 [B,C,out_h,out_w]."""
    _require_cuda(logits, sizes, flag_gts)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 4
    assert sizes.dtype == torch.int32 and sizes.is_contiguous() and sizes.shape == (logits.shape[0], 4)
    assert flag_gts is None or (flag_gts.dtype == torch.uint8 and flag_gts.is_contiguous())
    B, C, lh, lw = logits.shape
    out = torch.empty((B, C, out_h, out_w), dtype=torch.float32, device=logits.device)
    _cost(0.0, float(B) * C * (out_h * out_w + lh * lw) * 4)
    _call("postprocess_masks", "la_postprocess_masks", logits, out, sizes,
                                            flag_gts, B, C, lh, lw, image_size, out_h, out_w)
    return out


def nchw_to_tokens(x: torch.Tensor, want_f32: bool = True, want_bf16: bool = False):
    """[n, C, h, w] fp32 -> token-major [n*h*w, C] fp32 and/or bf16."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    n, C, h, w = x.shape
    o32 = torch.empty((n * h * w, C), dtype=torch.float32, device=x.device) if want_f32 else None
    o16 = torch.empty((n * h * w, C), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    for s in range(0, n, 32768):
        m = min(32768, n - s)
        _call("nchw_to_tokens", "la_nchw_to_tokens", x[s:],
                                             o32[s * h * w:] if want_f32 else None,
                                             o16[s * h * w:] if want_bf16 else None, m, C, h * w)
    return o32, o16


def tokens_to_nchw(x: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    """token-major fp32 [n*h*w, C] -> [n, C, h, w] fp32."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[0] == n * h * w
    C = x.shape[1]
    out = torch.empty((n, C, h, w), dtype=torch.float32, device=x.device)
    for s in range(0, n, 32768):
        m = min(32768, n - s)
        _call("tokens_to_nchw", "la_tokens_to_nchw", x[s * h * w:], out[s:], m, C, h * w)
    return out


def copy_slabs(x: torch.Tensor, n_slabs: int, slab_rows: int, stride_rows: int, offset_rows: int = 0,
               want_f32: bool = True, want_bf16: bool = False):
    """out[s, r] = x[s*stride_rows + offset_rows + r] for r < slab_rows (fp32 in; fp32 and/or bf16 out)."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    d = x.shape[1]
    assert (n_slabs - 1) * stride_rows + offset_rows + slab_rows <= x.shape[0]
    o32 = torch.empty((n_slabs * slab_rows, d), dtype=torch.float32, device=x.device) if want_f32 else None
    o16 = torch.empty((n_slabs * slab_rows, d), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    _call("copy_slabs", "la_copy_slabs", x, stride_rows, offset_rows, o32, o16,
                                     n_slabs, slab_rows, d)
    return o32, o16


def add_bcast(a: torch.Tensor, b: torch.Tensor, row_div: int, b_mod: int) -> torch.Tensor:
    """out[r] = a[r] + b[(r // row_div) % b_mod]; fp32 [rows, d]."""
    _require_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1] and b.shape[0] >= b_mod
    out = torch.empty_like(a)
    _call("add_bcast", "la_add_bcast", a, b, out, a.shape[0], a.shape[1],
                                    row_div, b_mod)
    return out


def permute_rows(x: torch.Tensor, outer: int, na: int, nb: int) -> torch.Tensor:
    """fp32 [outer*na*nb, d] -> rows reordered (o, i, j) -> (o, j, i)."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[0] == outer * na * nb
    out = torch.empty_like(x)
    _call("permute_rows", "la_permute_rows", x, out, outer, na, nb, x.shape[1])
    return out


def label_confusion(logits: torch.Tensor | None, preds: torch.Tensor | None, gt: torch.Tensor | None,
                    label_map: torch.Tensor | None, confmat: torch.Tensor | None = None,
                    invalid: torch.Tensor | None = None, ignore_index: int = -100, want_preds: bool = True,
                    want_gt: bool = True):
    """argmax over classes (or int64 `preds`), label_map[b][.] on predictions and gt, confmat[target, pred] += 1
    (int64 [G, G], accumulated in place).  logits [B, C, *spatial] fp32, preds / gt [B, *spatial] int64,
    label_map [B, map_len] int64.  Returns (mapped preds or None, mapped gt or None)."""
    _require_cuda(logits, preds, gt, label_map, confmat, invalid)
    ref = logits if logits is not None else (preds if preds is not None else gt)
    B = ref.shape[0]
    spatial = tuple(ref.shape[2:]) if logits is not None else tuple(ref.shape[1:])
    P = 1
    for s in spatial:
        P *= s
    C = logits.shape[1] if logits is not None else 0
    if logits is not None:
        assert logits.dtype == torch.float32 and logits.is_contiguous()
    for t in (preds, gt):
        assert t is None or (t.dtype == torch.int64 and t.is_contiguous() and tuple(t.shape) == (B,) + spatial), \
            "labels must be contiguous int64 [B, *spatial]"
    assert label_map is None or (label_map.dtype == torch.int64 and label_map.is_contiguous() and label_map.shape[0] == B)
    G = 0
    if confmat is not None:
        assert confmat.dtype == torch.int64 and confmat.is_contiguous() and confmat.dim() == 2 and confmat.shape[0] == confmat.shape[1]
        assert invalid is not None and invalid.dtype == torch.int64 and invalid.numel() == 1
        G = confmat.shape[0]
    have_pred = logits is not None or preds is not None
    preds_out = torch.empty((B,) + spatial, dtype=torch.int64, device=ref.device) if (want_preds and have_pred) else None
    gt_out = torch.empty((B,) + spatial, dtype=torch.int64, device=ref.device) if (want_gt and gt is not None) else None
    _cost(0.0, float(B) * P * (4 * C + (8 if preds is not None else 0) + (8 if gt is not None else 0)
                               + (8 if preds_out is not None else 0) + (8 if gt_out is not None else 0)))
    _call("label_confusion", "la_label_confusion", logits, preds, gt, label_map,
          preds_out, gt_out, confmat, invalid, B, C, P,
          label_map.shape[1] if label_map is not None else 0, G, int(ignore_index))
    return preds_out, gt_out


def _loss_workspace(device: torch.device) -> torch.Tensor:
    """Scratch of la_focal_loss (per-CTA partial sums + a counter the entry point zeroes itself): a fresh stream-ordered
    allocation per call -- nothing cached per stream, nothing to leak."""
    return torch.empty(_focal_loss_workspace_bytes() // 8 + 1, dtype=torch.float64, device=device)


def label_class_weights(labels: torch.Tensor, classes: int, ignore_index: int = -100):
    """-> (class_w fp32 [classes], hist int64 [classes + 2]) of loss/utils.py:17-42 (see the C header)."""
    _require_cuda(labels)
    assert labels.dtype == torch.int64 and labels.is_contiguous()
    hist = torch.empty(classes + 2, dtype=torch.int64, device=labels.device)
    w = torch.empty(classes, dtype=torch.float32, device=labels.device)
    _cost(0.0, 8.0 * labels.numel())
    _call("label_class_weights", "la_label_class_weights", labels, labels.numel(), classes,
          int(ignore_index), hist, w)
    return w, hist


def focal_loss(logits: torch.Tensor, target: torch.Tensor, class_w: torch.Tensor | None, gamma: float,
               ignore_index: int = -100, mean: bool = True, want_loss: bool = True, want_grad: bool = False,
               grad_scale: torch.Tensor | None = None, want_wtarget: bool = False):
    """-> (loss fp32 [] or None, grad like logits or None, wtarget [B, *spatial] or None)."""
    _require_cuda(logits, target, class_w, grad_scale)
    if logits is None:      # weight map only
        assert class_w is not None and not want_loss and not want_grad
        B, C, P = target.shape[0], class_w.numel(), target.numel() // target.shape[0]
    else:
        assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() >= 2
        B, C = logits.shape[:2]
        P = logits.numel() // (B * C)
    assert target.dtype == torch.int64 and target.is_contiguous() and target.numel() == B * P
    assert class_w is None or (class_w.dtype == torch.float32 and class_w.is_contiguous() and class_w.numel() == C)
    assert grad_scale is None or (grad_scale.dtype == torch.float32 and grad_scale.numel() == 1)
    loss = torch.empty((), dtype=torch.float32, device=target.device) if want_loss else None
    grad = torch.empty_like(logits) if want_grad else None
    wt = torch.empty(target.shape, dtype=torch.float32, device=target.device) if want_wtarget else None
    ws = _loss_workspace(target.device) if want_loss else None
    _cost(0.0, float(B) * P * (4 * C + 8 + (4 * C if want_grad else 0) + (4 if want_wtarget else 0)))
    _call("focal_loss.grad" if want_grad else "focal_loss", "la_focal_loss", logits,
          target, class_w, grad_scale, loss, grad, wt, ws, B, C, P,
          float(gamma), int(ignore_index), 1 if mean else 0)
    return loss, grad, wt


# ---------------------------------------------------------------------------------------------- input preprocessing
def preprocess_image_u8(src: torch.Tensor, new_h: int, new_w: int, size: int, bounds_x, kk_x, ksize_x: int, bounds_y,
                        kk_y, ksize_y: int, tmp, mean, std, out: torch.Tensor) -> torch.Tensor:
    """uint8 HWC [H, W, 3] -> fp32 CHW [3, size, size] (Pillow-exact bilinear resize, /255, (x - mean) / std, zero pad)."""
    _require_cuda(src, bounds_x, kk_x, bounds_y, kk_y, tmp, out)
    assert src.dtype == torch.uint8 and src.is_contiguous() and src.dim() == 3 and src.shape[2] == 3
    H, W = int(src.shape[0]), int(src.shape[1])
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (3, size, size)
    for b, k, ks, n in ((bounds_x, kk_x, ksize_x, new_w), (bounds_y, kk_y, ksize_y, new_h)):
        assert (b is None) == (k is None)
        if b is not None:
            assert b.dtype == torch.int32 and k.dtype == torch.int32 and b.is_contiguous() and k.is_contiguous()
            assert tuple(b.shape) == (n, 2) and tuple(k.shape) == (n, ks)
    assert tmp is None or (tmp.dtype == torch.uint8 and tmp.is_contiguous() and tmp.numel() >= H * new_w * 3)
    _cost(0.0, float(H) * W * 3 + 12.0 * size * size)
    _call("preprocess_image", "la_preprocess_image_u8", src, H, W, new_h, new_w, size, bounds_x, kk_x, ksize_x, bounds_y,
          kk_y, ksize_y, tmp, float(mean[0]), float(mean[1]), float(mean[2]), float(std[0]), float(std[1]), float(std[2]),
          out)
    return out


def rasterize_masks_u8(masks, n: int, H: int, W: int, new_h: int, new_w: int, long_side: int, out_side: int,
                       out: torch.Tensor, flag=None) -> torch.Tensor:
    """OR of n uint8 masks [n, H, W] -> nearest / pad / nearest -> fp32 {0, 1} [out_side, out_side] (+ presence flag)."""
    _require_cuda(masks, out, flag)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == out_side * out_side
    assert masks is None or (masks.dtype == torch.uint8 and masks.is_contiguous() and masks.numel() == n * H * W)
    assert flag is None or (flag.dtype == torch.uint8 and flag.numel() == 1)
    _call("rasterize_masks", "la_rasterize_masks_u8", masks, n, H, W, new_h, new_w, long_side, out_side, out, flag)
    return out


def scale_coords_f64(coords: torch.Tensor, sx: float, sy: float, out: torch.Tensor) -> torch.Tensor:
    """(x, y) float64 pairs * (sx, sy) in double -> fp32."""
    _require_cuda(coords, out)
    assert coords.dtype == torch.float64 and coords.is_contiguous() and coords.shape[-1] == 2
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == coords.shape
    _call("scale_coords", "la_scale_coords_f64", coords, coords.numel() // 2, float(sx), float(sy), out)
    return out


@torch.compiler.assume_constant_result
def _error_points_workspace_bytes(batch: int, classes: int, height: int) -> int:
    return int(_native.lib().la_error_points_workspace_bytes(batch, classes, height))


def error_points(logits: torch.Tensor, gt: torch.Tensor, rand: torch.Tensor, sx: torch.Tensor, sy: torch.Tensor,
                 ignore_index: int = -100):
    """-> (points fp32 [B, C, n, 2] (x, y) scaled by (sx[b], sy[b]), labels fp32 [B, C, n]); see la_error_points."""
    _require_cuda(logits, gt, rand, sx, sy)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 4
    B, C, H, W = logits.shape
    assert gt.dtype == torch.int64 and gt.is_contiguous() and tuple(gt.shape) == (B, H, W)
    assert rand.dtype == torch.int64 and rand.is_contiguous() and rand.shape[:2] == (B, C)
    n = rand.shape[2]
    for t in (sx, sy):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == B
    ws = torch.empty(_error_points_workspace_bytes(B, C, H), dtype=torch.uint8, device=logits.device)
    points = torch.empty((B, C, n, 2), dtype=torch.float32, device=logits.device)
    labels = torch.empty((B, C, n), dtype=torch.float32, device=logits.device)
    _cost(0.0, 2.0 * B * H * W * (4 * C + 8))
    _call("error_points", "la_error_points", logits, gt, B, C, H, W, int(ignore_index), rand, n, sx, sy, ws, points, labels)
    return points, labels


# ---------------------------------------------------------------------------------------------- pooled attention
def attention_pooled_supported(n_query: int, n_heads: int, d: int, tokens: int) -> bool:
    """Shapes la_attention_pooled_bf16 is built for (everything else keeps the projected k / v path)."""
    return n_query * n_heads <= 8 and d in (64, 128, 256, 512) and tokens >= 1 and not _NO_POOLED_ATTENTION


_NO_POOLED_ATTENTION = bool(__import__("os").environ.get("LA_NO_POOLED_ATTENTION"))   # experiment switch


def attention_pooled(x: torch.Tensor, u: torch.Tensor, e: torch.Tensor | None, scale: float, n_seq: int, tokens: int,
                     rows: int) -> torch.Tensor:
    """y[s, r] = sum_t softmax_t(scale (u[s, r] . x[s, t] + e[s, r, t])) x[s, t]; x bf16 [n_seq*tokens, d] read once."""
    _require_cuda(x, u, e)
    d = x.shape[1]
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[0] == n_seq * tokens
    assert u.dtype == torch.bfloat16 and u.is_contiguous() and tuple(u.shape) == (n_seq * rows, d)
    assert e is None or (e.dtype == torch.float32 and e.dim() == 2 and e.stride(1) == 1 and e.shape[0] == n_seq * rows
                         and e.shape[1] >= tokens)
    y = torch.empty((n_seq * rows, d), dtype=torch.bfloat16, device=x.device)
    _cost(4.0 * n_seq * rows * tokens * d, 2.0 * n_seq * tokens * d + (4.0 * n_seq * rows * tokens if e is not None else 0))
    _call("attention_pooled", "la_attention_pooled_bf16", x, x.stride(0), u, e, e.stride(0) if e is not None else 0,
          float(scale), y, n_seq, tokens, rows, d)
    return y


def head_rows(t: torch.Tensor, n_seq: int, heads: int, head_dim: int, expand: bool) -> torch.Tensor:
    """expand: bf16 [n_seq, H*dh] -> [n_seq*H, H*dh] with row (s, h) keeping only head h's columns;
    gather: [n_seq*H, H*dh] -> [n_seq, H*dh] taking head h's columns from row (s, h)."""
    _require_cuda(t)
    w = heads * head_dim
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == ((n_seq if expand else n_seq * heads), w)
    out = torch.empty(((n_seq * heads if expand else n_seq), w), dtype=torch.bfloat16, device=t.device)
    _call("head_rows", "la_head_rows_bf16", t, out, n_seq, heads, head_dim, 0 if expand else 1)
    return out
