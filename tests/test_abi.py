"""The C-ABI shared library loads and exports every symbol include/pcm_b200.h declares (no GPU)."""
import ctypes
import re
import subprocess

from pointcloudmatters_b200 import _lib


def test_header_parses_all_prototypes():
    text = _lib.HEADER.read_text()
    declared = set(re.findall(r"\b(pcm_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert len(declared) >= 19


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _lib.PROTOTYPES:
        assert hasattr(lib, name), name
    assert lib.pcm_abi_version() >= 1


def test_no_torch_types_in_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("pcm_") for s in exported), exported
    undefined = subprocess.run(["nm", "-D", "-u", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "at::" not in undefined and "c10" not in undefined and "torch" not in undefined


def test_product_package_never_imports_oracle():
    import pathlib

    root = pathlib.Path(_lib.__file__).resolve().parent
    for p in root.rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
