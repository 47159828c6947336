"""The default kernels of libmonorec_b200.so must be the ones that last passed the GPU parity suite.

Round 1 ended with GPU minutes exhausted while opt-in variants were still being added (new template instantiations next to the
default ones).  This CPU test pins the SASS of every GPU-verified kernel (tests/golden/verified_sass_digest.txt, produced by
tools/sass_digest.py on the verified build): a refactor that changes one of them has to be re-verified on the GPU and the
file regenerated in the same commit."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None, reason="needs the CUDA toolkit")
def test_default_kernels_match_gpu_verified_build():
    from monorec_b200 import build
    build.build()
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_digest.py")], capture_output=True, text=True, check=True).stdout
    current = {}
    for line in out.splitlines():
        digest, rest = line.split(None, 1)
        current[rest.split("lines", 1)[1].strip()] = digest
    want = [line.rstrip("\n").split("\t") for line in (ROOT / "tests" / "golden" / "verified_sass_digest.txt").read_text().splitlines()
            if line and not line.startswith("#")]
    assert len(want) >= 19
    changed = [name for name, digest in want if current.get(name) != digest]
    assert not changed, f"kernels differ from the GPU-verified build: {changed}"
