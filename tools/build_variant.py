"""Builds a variant of the library next to the default one: python tools/build_variant.py TAG -DMR_CV_SKIP=2 ...
-> monorec_b200/variants/libmonorec_b200_TAG.so (load it with MONOREC_B200_LIB=<path>).  Experiments only."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from monorec_b200 import build as B  # noqa: E402

tag, defs = sys.argv[1], sys.argv[2:]
out = ROOT / "monorec_b200" / "variants"
out.mkdir(exist_ok=True)
lib = out / f"libmonorec_b200_{tag}.so"
cmd = [B.NVCC, *B.FLAGS, *defs, "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xptxas", "-v", "-o", str(lib)] + [str(s) for s in B.sources()]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    sys.exit(1)
lines = (res.stdout + res.stderr).splitlines()
for i, l in enumerate(lines):
    if "cost_volume_kernel" in l and "Compiling" in l:
        print(tag, " ".join(x.strip() for x in lines[i + 2:i + 4]))
print(lib)
