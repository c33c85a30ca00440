"""ctypes binding of libfs2k.so — the only way the package reaches the GPU.

There is no CPU fallback and no alternative backend: if the shared library cannot be
loaded (or built, when nvcc is present), importing any op raises.  Prototypes are parsed
from `include/fs2k.h`, so the header is the single source of truth for the C ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
import shutil
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "fs2k.h"
LIB_PATH = PKG / "libfs2k.so"

_SCALARS = {
    "int": ctypes.c_int,
    "long": ctypes.c_long,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "fs2k_stream_t": ctypes.c_void_p,
}


class Fs2kError(RuntimeError):
    pass


def parse_header(path: Path = HEADER) -> dict[str, tuple[object, list[object]]]:
    """{name: (restype, [argtypes])} for every `fs2k_*` prototype in the header."""
    text = re.sub(r"/\*.*?\*/", " ", path.read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(fs2k_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue
        if ret == "const char*" or ret == "const char *":
            restype = ctypes.c_char_p
        elif ret == "size_t":
            restype = ctypes.c_size_t
        else:
            restype = ctypes.c_int
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    continue
                ty = " ".join(a.replace("const", " ").split()[:-1])
                argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_profile_records = None  # list of (entry point, args, start event, end event) while profiling


class _ProfilingProxy:
    """Wraps every C-ABI call in a pair of CUDA events on the current stream (bench.py's roofline pass)."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if (not name.startswith("fs2k_") or name in ("fs2k_strerror", "fs2k_version", "fs2k_check_device")
                or name.endswith(("_bytes", "_supported"))):
            return fn

        def timed(*args):
            import torch

            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            _profile_records.append((name, args, s, e))
            return rc

        return timed


def start_profile() -> None:
    global _profile_records
    _profile_records = []


def stop_profile():
    """Returns [(entry point, args, milliseconds)] of every call since start_profile()."""
    global _profile_records
    import torch

    torch.cuda.synchronize()
    out = [(n, a, s.elapsed_time(e)) for n, a, s, e in _profile_records]
    _profile_records = None
    return out


def lib():
    global _lib
    if _lib is not None:
        return _ProfilingProxy(_lib) if _profile_records is not None else _lib
    from . import build as _build

    if shutil.which(_build.NVCC) or Path(_build.NVCC).exists():
        if _build.needs_build():
            _build.build()
    if not LIB_PATH.exists():
        raise Fs2kError(
            f"{LIB_PATH} is missing and nvcc is not available to build it — "
            "fastspeech2_lightning_b200 has no CPU or PyTorch fallback path"
        )
    cdll = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(cdll, name)  # AttributeError if the header declares something the library lacks
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = cdll
    if os.environ.get("FS2K_PDL", "1") == "0":  # plain launches (no programmatic dependent launch), for A/B timing
        cdll.fs2k_set_pdl(0)
    return cdll


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().fs2k_strerror(rc).decode()
        raise Fs2kError(f"{what or 'fs2k call'} failed: {msg} (code {rc})")


def require_device() -> None:
    """Fail loudly unless the current CUDA device can run the sm_100a build."""
    import torch

    if not torch.cuda.is_available():
        raise Fs2kError("no CUDA device: fastspeech2_lightning_b200 runs on B200 (sm_100a) only, there is no CPU path")
    check(lib().fs2k_check_device(), "fs2k_check_device")
