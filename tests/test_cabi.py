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
    # direction words (one per thread and block of 8 frames: 13 blocks x 32 threads, 256-aligned) + log buffer
    assert lib.fs2k_mas_workspace_bytes(2, 100, 80) == 2 * 13 * 32 * 4 + 2 * 100 * 80 * 4
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


def test_no_kernel_touches_global_memory_before_its_programmatic_launch_wait():
    """Every kernel of the library is launched with programmatic stream serialization: it may start while its predecessors
    are still running and must not read or write global memory before griddepcontrol.wait (SASS: ACQBULK).  ptxas hoists
    `const __restrict__` loads (LDG.E.CONSTANT) above the wait when a kernel does set-up work first — that made the tensor-core
    attention kernels read `order` / `lens` before their producer had run (first CUDA-graph replay only).  Scan the SASS."""
    import re
    import shutil
    import subprocess

    import pytest

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    access = re.compile(r"\b(LDG|LD\.E|STG|ST\.E|ATOMG|REDG|RED\.E|UTMALDG|UTMASTG|UBLKCP)\b")
    bad, fn, waited, n_fn = [], None, False, 0
    for line in sass.splitlines():
        if "Function :" in line:
            fn, waited = line.split("Function :")[1].strip(), False
            n_fn += 1
        elif "ACQBULK" in line:
            waited = True
        elif fn and not waited and access.search(line):
            bad.append((fn, line.strip()[:90]))
            waited = True  # one report per kernel
    assert n_fn > 100
    assert not bad, bad
