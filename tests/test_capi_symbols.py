"""The C-ABI library builds, loads on a CPU-only box and exports every symbol include/monorec_b200.h declares."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "monorec_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from monorec_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/monorec_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in monorec_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_string_without_gpu():
    from monorec_b200 import _lib
    lib = _lib.load()
    assert lib.mr_version() >= 0x100
    assert isinstance(lib.mr_last_error(), bytes)
    # argument validation happens before any CUDA call, so it can be exercised on a CPU-only box
    rc = lib.mr_cost_volume_fwd(None, None, None, None, None, None, 1, 1, 32, 64, 64, 10.0, None, None)
    assert rc == -1 and b"null pointer" in lib.mr_last_error()
    assert lib.mr_cost_volume_host_workspace(8, 4, 32, 256, 512) > 8 * 5 * 32 * 256 * 512 * 4
    assert lib.mr_cost_volume_host_workspace(0, 4, 32, 256, 512) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under monorec_b200/ may import it (tier rule 3)."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for p in (ROOT / "monorec_b200").rglob("*.py"):
        assert not pat.search(p.read_text()), f"{p} imports the oracle"
