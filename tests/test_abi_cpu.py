"""The C-ABI library loads without a GPU, exports every symbol include/labelanything_b200.h declares, and its
entry points reject bad arguments before touching CUDA (no compute calls here)."""
import ctypes

import pytest

from labelanything_b200 import _native


def test_library_exports_every_declared_symbol():
    decl = _native.declared_functions()
    assert len(decl) >= 25 and "la_gemm_bf16" in decl and "la_attention_bf16" in decl
    lib = _native.lib()
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but missing from the library"
    assert lib.la_version() == 100


def test_header_prototypes_parse_to_ctypes():
    for name, (ret, args) in _native.declared_functions().items():
        assert ret in ("int", "long long", "const char*"), (name, ret)
        for a in args:
            assert a in _native._CTYPE, f"{name}: argument type {a!r} has no ctypes mapping"


def test_bad_arguments_are_rejected_with_a_message():
    lib = _native.lib()
    rc = lib.la_gemm_bf16(None, None, 0, None, 0, None, None, 0, 0, 1, 8, 8, 0)
    assert rc == -1
    assert b"null pointer" in lib.la_last_error()
    with pytest.raises(RuntimeError, match="la_gemm_bf16: null pointer"):
        _native.check(rc, "gemm")
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    p += (-p) % 16
    assert lib.la_gemm_bf16(None, p, 8, p, 8, None, p, 8, 0, 4, 12, 8, 0) == -1   # N not a multiple of 8
    assert b"multiple of 8" in lib.la_last_error()
    assert lib.la_attention_tokens(None, p, 8, p, 8, p, 8, None, 0, None, 0, p, 8, 1, 1, 1, 1, 12, 1.0, None) == -1
    assert b"head_dim" in lib.la_last_error()
    assert lib.la_attention_tokens_splits(4, 1, 4096) >= 1 and lib.la_attention_tokens_splits(4, 4096, 9) == 0
    assert lib.la_attention_tokens_workspace_bytes(1200, 1, 4096, 8, 32) >= 0
    # the entry points added for the encoder hot path state what they are built for
    assert lib.la_conv3x3_bf16(None, p, 1, 30, 30, 256, p, 2304, None, p, 256, 0, 256, 0) == -1
    assert b"64-wide feature maps" in lib.la_last_error()
    assert lib.la_gemm_bf16_accumulate(None, p, 64, p, 64, None, p, 256, 128, 256, 64) == -1
    assert b"CTA-pair kernel" in lib.la_last_error()
    assert lib.la_attention_window_bf16(None, p, 2304, 0, p, 2304, 768, 1536, 196, 1, 12, 0.125, p, 16, p, 768, 0, 0,
                                        0, 0) == -1
    assert b"rel_pad 32" in lib.la_last_error()
    assert lib.la_attention_window_bf16(None, p, 2304, 0, p, 2304, 768, 1536, 196, 1, 12, 0.125, None, 32, p, 768, 0,
                                        0, 0, 0) == -1
    assert b"rel_table is required" in lib.la_last_error()
    assert lib.la_gemm_bf16_to_grid(None, p, 64, p, 64, None, p, 64, 4096, 64, 64, 48, 70) == -1   # grid % 32 != 0
    assert b"multiple of 32" in lib.la_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):
        _native.lib()
