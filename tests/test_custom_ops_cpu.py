"""The C ABI is exposed to PyTorch as `torch.library` custom ops derived from include/labelanything_b200.h (one op per
`int la_xxx(void* stream, ...)` entry point, fake implementation included).  No GPU here: the schemas, the fake
implementations (FakeTensor tracing of CUDA-device tensors needs no device) and the loud failure on CPU tensors."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from labelanything_b200 import _native, ops


def test_every_launch_entry_point_is_a_custom_op_with_the_header_schema():
    decl = _native.declared_functions()
    launches = [n for n, (ret, types) in decl.items()
                if ret == "int" and types and types[0] == "void*" and n != "la_attention_set_trace"]
    assert len(launches) >= 25
    for name in launches:
        op = getattr(torch.ops.labelanything_b200, name).default
        schema = op._schema
        types = decl[name][1][1:]                                   # stream parameter dropped
        assert len(schema.arguments) == len(types), name
        for arg, ctype in zip(schema.arguments, types):
            if "*" in ctype:
                assert str(arg.type) == "Optional[Tensor]", (name, arg.name, str(arg.type))
                mutated = arg.alias_info is not None and arg.alias_info.is_write
                assert mutated == (not ctype.startswith("const")), (name, arg.name, ctype)
            else:
                assert str(arg.type) == ("float" if ctype in ("float", "double") else "int"), (name, arg.name)
        assert len(schema.returns) == 0                             # outputs are caller-allocated, mutated in place


def test_fake_tensor_tracing_runs_the_wrappers_without_a_device():
    """Under FakeTensorMode the registered fake implementation (a no-op: outputs are pre-allocated) is what runs;
    shapes / dtypes of the results come from the wrappers' own allocations."""
    with FakeTensorMode():
        a = torch.empty(300, 128, dtype=torch.bfloat16, device="cuda")
        w = torch.empty(256, 128, dtype=torch.bfloat16, device="cuda")
        b = torch.empty(256, dtype=torch.float32, device="cuda")
        out = torch.empty(300, 256, dtype=torch.float32, device="cuda")
        torch.ops.labelanything_b200.la_gemm_bf16(a, a.stride(0), w, w.stride(0), b, out, out.stride(0), ops.DT_F32,
                                                   300, 256, 128, ops.ACT_NONE)
        x = torch.empty(64, 256, dtype=torch.float32, device="cuda")
        y = torch.empty(64, 256, dtype=torch.bfloat16, device="cuda")
        g = torch.empty(256, dtype=torch.float32, device="cuda")
        torch.ops.labelanything_b200.la_add_layernorm(x, 0, None, None, None, 0, None, g, g, 1e-6, 0, y, ops.DT_BF16, None,
                                                       None, 0, None, 64, 256, 0, 0, 0, 0, 0)
        assert out.shape == (300, 256) and y.dtype == torch.bfloat16


def test_cpu_tensors_are_refused_loudly():
    a = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.gemm(a, a)
    with pytest.raises(RuntimeError, match="no CUDA tensor|no CPU fallback"):
        torch.ops.labelanything_b200.la_gemm_bf16(a, 64, a, 64, None, torch.zeros(4, 4), 4, ops.DT_F32, 4, 4, 64, 0)
