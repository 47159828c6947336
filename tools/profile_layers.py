#!/usr/bin/env python
"""Per-layer times of the tensor-core convolutions inside one MonoRecModel forward (CUDA events around every launch):
which layers are far from their HBM time, and which kernel (halo / tap-refetch) each one takes.

    MONOREC_B200_CONV=f16 python tools/profile_layers.py [B] [F]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200 import conv as C  # noqa: E402
from monorec_b200.model import MonoRecModel  # noqa: E402
from monorec_b200.synthetic import make_inputs, to_device  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
model = MonoRecModel().cuda().eval()
data = to_device(make_inputs(B, F, 256, 512, seed=0), "cuda:0")
records = []
orig = C.conv2d_tc


def halo_kind(L, half, n_pad):
    """The eligibility rule of conv2d_nhwc_tc_impl (csrc/conv_tc.cu), restated for the label only."""
    if L.stride != (1, 1) or L.kw > 9 or L.kh > 7:
        return "refetch"
    if half:
        k64 = sum((c + 63) // 64 * 64 for c in L.src_c)
        k32 = sum((c + 31) // 32 * 32 for c in L.src_c)
        kc = 32 if all(c <= 32 for c in L.src_c) and k32 != k64 else 64
        row = kc * 2
    else:
        kc, row = 32, 128
    chunks = sum((c + kc - 1) // kc for c in L.src_c)
    bres = (L.kh * L.kw * chunks * n_pad * row + 1023) // 1024 * 1024
    a = ((16 + L.kh - 1) * (8 + L.kw - 1) * row + 1023) // 1024 * 1024
    budget = 224 * 1024 // 2 - 8 * 1024
    st = (budget - 2048 - bres) // a if bres + 2048 < budget else 0
    if st >= 2:
        return f"halo x{min(st, 4)}"
    cols = 32
    while cols < 2 * n_pad:
        cols *= 2
    bud = 228 * 1024 // 2 - 11 * 1024 - 512 - 1024
    if L.kh * L.kw > 1 and 2 * cols <= 512 and bud > 2 * a + 3 * n_pad * row:
        return f"halo stream x{min(8, (bud - 2 * a) // (n_pad * row))}"
    return "refetch"


def wrapped(srcs, L, out=None, out_hw=None, round_out=True, half=False, out_f32=False, out_coff=0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(srcs, L, out=out, out_hw=out_hw, round_out=round_out, half=half, out_f32=out_f32, out_coff=out_coff)
    e1.record()
    Bn, Hs, Ws, _ = srcs[0].shape
    ho, wo = (out_hw if out_hw is not None else (-(-Hs // L.stride[0]), -(-Ws // L.stride[1])))
    es = 2 if half else 4
    eo = r.element_size()
    byts = Bn * Hs * Ws * sum(L.src_c) * es + Bn * ho * wo * L.cout * eo
    flops = 2.0 * Bn * ho * wo * L.cout * sum(L.src_c) * L.kh * L.kw
    n_pad = (L.cout + 15) // 16 * 16
    records.append((e0, e1, f"{'+'.join(map(str, L.src_c))}->{L.cout} {L.kh}x{L.kw} s{L.stride[0]}{L.stride[1]} B{Bn} {Hs}x{Ws}",
                    byts, flops, halo_kind(L, half, n_pad)))
    return r


orig_ph = C.conv2d_tc_phases


def wrapped_phases(srcs, subs, out, out_hw, round_out=True, half=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig_ph(srcs, subs, out, out_hw, round_out=round_out, half=half)
    e1.record()
    Bn, Hs, Ws, _ = srcs[0].shape
    L = subs[0]
    es = 2 if half else 4
    byts = Bn * Hs * Ws * sum(L.src_c) * es + r.numel() * r.element_size()
    flops = sum(2.0 * Bn * Hs * Ws * S.cout * sum(S.src_c) * S.kh * S.kw for S in subs)
    records.append((e0, e1, f"{'+'.join(map(str, L.src_c))}->{L.cout} {len(subs)} sub-pixel phases B{Bn} {Hs}x{Ws}", byts, flops,
                    "refetch, one launch"))
    return r


C.conv2d_tc = wrapped
C.conv2d_tc_phases = wrapped_phases
with torch.no_grad():
    for _ in range(3):
        records.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        model(dict(data))
        t1.record()
        torch.cuda.synchronize()
total = t0.elapsed_time(t1)
rows = {}
for e0, e1, name, byts, flops, kind in records:
    t = e0.elapsed_time(e1) * 1e3
    r = rows.setdefault((name, kind), [0, 0.0, byts, flops])
    r[0] += 1
    r[1] += t
conv_total = sum(r[1] for r in rows.values())
print(f"mode {C.MODE}: forward {total:.2f} ms (eager, events around every conv), tensor-core convs {conv_total / 1e3:.2f} ms in {len(records)} launches")
print(f"{'layer':44s} {'n':>3s} {'us/launch':>9s} {'sum us':>8s} {'GB/s':>7s} {'TFLOP/s':>8s}  kernel")
for (name, kind), (n, t, byts, flops) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:44s} {n:3d} {t / n:9.1f} {t:8.1f} {byts / (t / n) / 1e3:7.0f} {flops / (t / n) / 1e6:8.1f}  {kind}")
