#!/usr/bin/env python
"""Per-layer micro-benchmark through the C ABI (CUDA events, B images): tuning aid, not the product bench.
    python tools/bench_ops.py [--batch 256] [--only dw|gemm|se|stem|tail]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import devops as D  # noqa: E402
from mtg_card_image_segmentation_b200 import arch  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    B = a.batch
    dev = "cuda"
    H, W = 160, 120
    print(f"B={B}  variant env: DW={os.environ.get('MTGSEG_DW_VARIANT')} GEMM={os.environ.get('MTGSEG_GEMM_VARIANT')}")
    tot = {}
    for i, b in enumerate(arch.BLOCKS, start=1):
        stride = 1 if b.dilation > 1 else b.stride
        pad = (b.kernel - 1) // 2 * b.dilation
        Ho = (H + 2 * pad - b.dilation * (b.kernel - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - b.dilation * (b.kernel - 1) - 1) // stride + 1
        if a.only in ("", "gemm") and b.cexp != b.cin:
            M = B * H * W
            x = torch.randn(M, b.cin, device=dev).bfloat16()
            w = torch.randn(b.cexp, b.cin, device=dev).bfloat16()
            sc = torch.ones(b.cexp, device=dev); sh = torch.zeros(b.cexp, device=dev)
            ms = timeit(lambda: D.conv1x1(x, w, sc, sh, 2))
            by = 2.0 * M * (b.cin + b.cexp)
            print(f"b{i}.expand {H}x{W} {b.cin}->{b.cexp}: {ms*1e3:8.1f} us {by/ms/1e6:7.0f} GB/s {2.0*M*b.cin*b.cexp/ms/1e9:6.1f} TF/s")
            tot["gemm"] = tot.get("gemm", 0) + ms
            del x
        if a.only in ("", "dw"):
            x = torch.randn(B, H, W, b.cexp, device=dev).bfloat16()
            w = torch.randn(b.kernel ** 2, b.cexp, device=dev).bfloat16()
            sc = torch.ones(b.cexp, device=dev); sh = torch.zeros(b.cexp, device=dev)
            ms = timeit(lambda: D.dwconv(x, w, sc, sh, 2, b.kernel, stride, b.dilation, b.use_se))
            by = 2.0 * B * b.cexp * (H * W + Ho * Wo)
            print(f"b{i}.dw k{b.kernel} s{stride} d{b.dilation} {H}x{W} C{b.cexp} gap={int(b.use_se)}: {ms*1e3:8.1f} us {by/ms/1e6:7.0f} GB/s")
            tot["dw"] = tot.get("dw", 0) + ms
            del x
        if a.only in ("", "gemm"):
            M = B * Ho * Wo
            x = torch.randn(M, b.cexp, device=dev).bfloat16()
            w = torch.randn(b.cout, b.cexp, device=dev).bfloat16()
            sc = torch.ones(b.cout, device=dev); sh = torch.zeros(b.cout, device=dev)
            res = torch.randn(M, b.cout, device=dev).bfloat16() if (b.stride == 1 and b.cin == b.cout) else None
            ase = torch.rand(B, b.cexp, device=dev) if b.use_se else None
            ms = timeit(lambda: D.conv1x1(x, w, sc, sh, 0, res, ase, Ho * Wo))
            by = 2.0 * M * (b.cexp + b.cout * (2 if res is not None else 1))
            print(f"b{i}.project {Ho}x{Wo} {b.cexp}->{b.cout} se={int(b.use_se)} res={int(res is not None)}: {ms*1e3:8.1f} us {by/ms/1e6:7.0f} GB/s {2.0*M*b.cout*b.cexp/ms/1e9:6.1f} TF/s")
            tot["gemm"] = tot.get("gemm", 0) + ms
            del x
        if a.only in ("", "se") and b.use_se:
            sq = arch.make_divisible(b.cexp // 4, 8)
            sums = torch.randn(B, 4, b.cexp, device=dev)
            w1 = torch.randn(sq, b.cexp, device=dev).bfloat16(); b1 = torch.zeros(sq, device=dev)
            w2 = torch.randn(b.cexp, sq, device=dev).bfloat16(); b2 = torch.zeros(b.cexp, device=dev)
            ms = timeit(lambda: D.se_mlp(sums, Ho * Wo, w1, b1, 1, w2, b2, 3))
            print(f"b{i}.se C{b.cexp} sq{sq}: {ms*1e3:8.1f} us")
            tot["se"] = tot.get("se", 0) + ms
        H, W = Ho, Wo
    if a.only in ("", "gemm"):
        M = B * H * W
        x = torch.randn(M, 160, device=dev).bfloat16(); w = torch.randn(960, 160, device=dev).bfloat16()
        sc = torch.ones(960, device=dev); sh = torch.zeros(960, device=dev)
        ms = timeit(lambda: D.conv1x1(x, w, sc, sh, 2))
        print(f"b16.conv 160->960: {ms*1e3:8.1f} us {2.0*M*1120/ms/1e6:7.0f} GB/s")
        tot["gemm"] = tot.get("gemm", 0) + ms
        x = torch.randn(B, H, W, 960, device=dev).bfloat16(); w = torch.randn(128, 9, 960, device=dev).bfloat16() * 0.01
        sc = torch.ones(128, device=dev); sh = torch.zeros(128, device=dev)
        ms = timeit(lambda: D.conv3x3(x, w, sc, sh, 1))
        print(f"head.cbr 3x3 960->128: {ms*1e3:8.1f} us {2.0*M*128*8640/ms/1e9:6.1f} TF/s")
        tot["cbr"] = ms
    if a.only in ("", "stem"):
        x = torch.randn(B, 3, 320, 240, device=dev)
        w = torch.randn(27, 16, device=dev); sc = torch.ones(16, device=dev); sh = torch.zeros(16, device=dev)
        ms = timeit(lambda: D.stem(x, w, sc, sh))
        print(f"stem: {ms*1e3:8.1f} us {B*(3*320*240*4+160*120*16*2)/ms/1e6:7.0f} GB/s")
        tot["stem"] = ms
    if a.only in ("", "tail"):
        lowres = torch.randn(B, 40, 30, 2, device=dev)
        for dt, nb in ((torch.float32, 4), (torch.bfloat16, 2)):
            ms = timeit(lambda: D.upsample_out(lowres, 320, 240, dt))
            print(f"upsample_out {dt}: {ms*1e3:8.1f} us {B*320*240*2*nb/ms/1e6:7.0f} GB/s")
        ms = timeit(lambda: D.upsample_out(lowres, 320, 240, torch.float32, False, True))
        print(f"upsample_out mask only: {ms*1e3:8.1f} us")
    print("totals ms:", {k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))


if __name__ == "__main__":
    main()
