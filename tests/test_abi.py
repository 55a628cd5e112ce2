"""The C-ABI library loads and exports every symbol include/vittf.h declares (no compute, no GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "vittf.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vittf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from vittf_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from vittf_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()
    lib = _lib.load()
    assert lib.vittf_version() >= 100
    assert isinstance(lib.vittf_last_error(), bytes)


def test_library_is_native_sm100a():
    """The shipped binary must contain tcgen05 / TMA machine code for sm_100a and nothing else."""
    import subprocess
    from vittf_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.sm_(\d+a?)\.", out))
    assert archs == {"100a"}, (archs, out[:300])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vittf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(_lib.VittfError):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vittf_b200/ may import it."""
    for f in (ROOT / "vittf_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
