"""CPU-side checks of the C-ABI boundary: the library loads (no GPU needed) and exports every
symbol include/fs2k.h declares; error paths that need no device work return the documented codes."""
import ctypes

from fastspeech2_lightning_b200 import _lib


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    lib = _lib.lib()
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.fs2k_version() >= 100


def test_strerror_and_argument_validation_without_a_device():
    lib = _lib.lib()
    assert lib.fs2k_strerror(0) == b"ok"
    assert b"unsupported" in lib.fs2k_strerror(-2)
    # negative dims / null pointers are rejected before any CUDA call
    assert lib.fs2k_mas_fwd(None, 0, None, None, -1, 4, 4, None, None, None, None, 0, None) == -1
    assert lib.fs2k_mas_fwd(None, 0, None, None, 1, 4, 5000, None, None, None, None, 0, None) == -2
    assert lib.fs2k_mas_fwd(None, 0, None, None, 1, 4, 4, None, None, None, None, 0, None) == -3
    assert lib.fs2k_lr_gather(None, None, None, 1, 4, 6, 8, None, None, None, None, None, None) == -2  # D % 4
    assert lib.fs2k_lr_gather(None, None, None, 1, 4, 8, 8, None, None, None, None, None, None) == -4
    assert lib.fs2k_mas_workspace_bytes(2, 100, 80) == 2560 + 2 * 100 * 80 * 4  # direction words (256-aligned) + log buffer
    # empty inputs are a no-op
    assert lib.fs2k_lr_scan(None, 0, 10, None, None, None) == 0


def test_product_path_has_no_cpu_fallback():
    import pytest
    import torch

    from fastspeech2_lightning_b200 import ops

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ValueError):
        ops.layernorm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8))
    with pytest.raises(_lib.Fs2kError):
        _lib.require_device()
