"""Prints an address-insensitive digest of every kernel's SASS in libmonorec_b200.so (or the library given as argv[1]).

Used to show that a refactor (a new template parameter, an opt-in variant) left the GPU-verified default kernels bit-for-bit
unchanged when there are no GPU minutes to re-run the parity suite:  python tools/sass_digest.py > before; ...; diff."""
import hashlib
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "monorec_b200" / "libmonorec_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name, body, digests = None, [], {}


def flush():
    if name is not None:
        digests[name] = (hashlib.md5("\n".join(body).encode()).hexdigest(), len(body))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, body = m.group(1), []
        continue
    if name is None or "identifier =" in line:
        continue
    line = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)            # instruction offsets
    line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line)            # encodings
    line = re.sub(r"0x[0-9a-f]{6,}", "ADDR", line)             # absolute branch targets
    line = re.sub(r"\s+", " ", line).strip()
    if line.endswith(";"):                                     # instructions only (headers mention the mangled name)
        body.append(line)
flush()
for k in sorted(digests):
    short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0]
    print(f"{digests[k][0]}  {digests[k][1]:6d} lines  {short}")
