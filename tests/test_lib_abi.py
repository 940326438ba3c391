"""The C-ABI shared library loads (no GPU needed) and exports every symbol include/ganlab_b200.h declares."""
import ctypes

import pytest

from gan_lab_b200._lib import LIB, LIB_PATH, parse_header


def test_header_parses():
    sigs = parse_header()
    assert len(sigs) >= 40
    for must in ("glb_conv2d_fprop", "glb_conv2d_dgrad", "glb_conv2d_wgrad", "glb_style_epilogue_fwd",
                 "glb_mbstd_bwdbwd", "glb_adam_ewma_multi", "glb_last_error"):
        assert must in sigs


def test_library_exports_every_declared_symbol():
    if not LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    dll = ctypes.CDLL(str(LIB_PATH))
    for name in parse_header():
        assert hasattr(dll, name), f"{name} declared in the header but not exported by {LIB_PATH.name}"
    assert LIB.fn("glb_version")() >= 100


def test_no_cpu_fallback():
    """Product ops must refuse CPU tensors loudly (no CPU path behind the C-ABI)."""
    import torch
    from gan_lab_b200 import ops, GlbError
    with pytest.raises(GlbError):
        ops.blur3x3(torch.zeros(1, 4, 4, 4))
