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


def test_header_is_plain_c_and_a_c_consumer_links(tmp_path):
    """include/monorec_b200.h compiles as C99 (-pedantic) and a C program links against the library and reaches the
    argument checks without a GPU (no compute call)."""
    import shutil
    import subprocess
    from pathlib import Path
    import pytest
    from monorec_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    _lib.load()
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "consumer.c"
    src.write_text('#include "monorec_b200.h"\n#include <stdio.h>\n#include <string.h>\n'
                   'int main(void) {\n'
                   '    mr_conv_desc d; memset(&d, 0, sizeof d);\n'
                   '    if (mr_sizeof_conv_desc() != (int)sizeof d) return 2;\n'
                   '    if (mr_conv2d_nhwc_tc(0, 16, 32, 0, 0) == MR_OK) return 3;\n'
                   '    if (strstr(mr_last_error(), "null descriptor") == 0) return 4;\n'
                   '    if (mr_cost_volume_host_workspace(8, 4, 32, 256, 512) <= 0) return 5;\n'
                   '    {   /* pack a 3x3 layer with two concatenated sources for the half tensor-core path, on the host */\n'
                   '        static float w[24 * 96 * 9]; static unsigned short packed[9 * 32 * 128];\n'
                   '        int src_c[2] = {32, 64}, n_pad = 0, k_pad = 0, i;\n'
                   '        for (i = 0; i < 24 * 96 * 9; ++i) w[i] = (float)(i % 7) - 3.0f;\n'
                   '        if (mr_pack_conv_weights_bytes(24, 2, src_c, 3, 3, MR_DT_F16, &n_pad, &k_pad) != (long long)sizeof packed) return 6;\n'
                   '        if (n_pad != 32 || k_pad != 128) return 7;\n'
                   '        if (mr_pack_conv_weights(w, 24, 2, src_c, 3, 3, MR_DT_F16, packed) != MR_OK) return 8;\n'
                   '        /* tap (1,1), output channel 5, input channel 40 (second source) lives at [4][5][64 + 8]; w = -3 .. 3 */\n'
                   '        if (packed[(4 * 32 + 5) * 128 + 72] != 0xC000 /* half -2.0 = w[(5*96+40)*9+4] = (4684 % 7) - 3 */) return 9;\n'
                   '        if (packed[(4 * 32 + 5) * 128 + 40] != 0) return 10;   /* padding of the first source */\n'
                   '        if (mr_conv_workspace_bytes(&d) != 0) return 11;\n'
                   '    }\n'
                   '    printf("%d\\n", mr_version());\n    return 0;\n}\n')
    exe = tmp_path / "consumer"
    libdir = _lib.LIB_PATH.parent
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{root / 'include'}", str(src), "-o", str(exe),
                    f"-L{libdir}", f"-l:{_lib.LIB_PATH.name}", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert int(out.stdout.strip()) == _lib.load().mr_version()
