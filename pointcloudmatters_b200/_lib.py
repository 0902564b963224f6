"""ctypes binding of libpcm_b200.so -- the C ABI declared in include/pcm_b200.h.

The product path has NO fallback: if the shared object is missing or a symbol declared in the
header is not exported, importing this module raises.  Prototypes are parsed from the header so
that the binding can never drift from the ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
HEADER = _HERE.parent / "include" / "pcm_b200.h"
LIB_PATH = Path(os.environ.get("PCM_B200_LIB", _HERE / "libpcm_b200.so"))


class PcmError(RuntimeError):
    pass


_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "long": ctypes.c_long,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
    "size_t": ctypes.c_size_t,
    "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64,
    "int64_t": ctypes.c_int64,
    "pcm_stream_t": ctypes.c_void_p,
}


def _parse_header(text: str):
    """Return {name: (restype, [argtypes])} for every `int|const char * pcm_*(...)` prototype."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(long long|int|const char \*|void)\s*(pcm_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
        restype = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "const char *": ctypes.c_char_p, "void": None}[ret]
        protos[name] = (restype, argtypes)
    return protos


PROTOTYPES = _parse_header(HEADER.read_text())


def _load():
    if not LIB_PATH.exists():
        raise PcmError(
            f"{LIB_PATH} not found: build it with `python -m pointcloudmatters_b200.build` "
            "(the product path has no CPU / eager fallback)."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # declared in the header but not exported
            raise PcmError(f"{LIB_PATH} does not export {name} declared in {HEADER.name}") from e
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


class _CountingLib:
    """Thin proxy over the ctypes handle that counts launches of OUR kernels (bench `gpu_launches`)."""

    def __init__(self, handle):
        object.__setattr__(self, "_h", handle)
        object.__setattr__(self, "launches", 0)
        object.__setattr__(self, "calls", {})  # entry point -> number of calls (tests assert which path ran)
        object.__setattr__(self, "_no_count", {"pcm_abi_version", "pcm_build_info", "pcm_tune_fps_threads", "pcm_launch_count"})

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if name in self._no_count:
            return fn

        calls = self.calls
        calls.setdefault(name, 0)

        def call(*args):
            object.__setattr__(self, "launches", self.launches + 1)
            calls[name] += 1
            return fn(*args)

        object.__setattr__(self, name, call)
        return call


lib = _CountingLib(_load())


def launch_count() -> int:
    """Kernel launches issued by libpcm_b200.so so far in this process."""
    return int(lib.pcm_launch_count())


def check(status: int, what: str) -> None:
    if status != 0:
        if status > 0:
            raise PcmError(f"{what}: CUDA error {status}")
        raise PcmError(f"{what}: invalid argument / unsupported size (code {status})")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PcmError("pointcloudmatters_b200 kernels run on CUDA tensors only (no CPU fallback)")
