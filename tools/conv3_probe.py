#!/usr/bin/env python
"""Times and checks the 3x3 tcgen05 convolutions (head 3x3: 960 -> 128 at 20x15, its dgrad shape, small and ragged maps)
through the C ABI.  `--all` runs the A/B variants in child processes (the switches are read once per process):
MTGSEG_CONV3=0 nine-shifted-boxes kernel (gemm_tc.cu), MTGSEG_CONV3=1 haloed-tile kernel (conv3_tc.cu)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = [(2, 4, 3, 64, 32), (1, 9, 7, 72, 32), (1, 20, 15, 64, 128), (3, 20, 15, 960, 128), (5, 20, 15, 64, 128),
          (2, 40, 30, 128, 128), (2, 30, 40, 136, 200), (256, 20, 15, 960, 128), (32, 20, 15, 128, 960), (256, 20, 15, 128, 960)]


def main():
    import torch
    import torch.nn.functional as F
    import devops as D
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [SHAPES[int(a.split('=')[1])] for a in sys.argv if a.startswith('--shape=')] or SHAPES
    for (B, H, W, K, N) in shapes:
        g = torch.Generator().manual_seed(B + K)
        a = torch.randn(B, H, W, K, generator=g).to(torch.bfloat16).cuda()
        w = (torch.randn(N, 9, K, generator=g) / (9 * K) ** 0.5).to(torch.bfloat16).cuda()
        scale = (torch.rand(N, generator=g) + 0.5).cuda(); shift = torch.randn(N, generator=g).cuda()
        out = D.conv3x3(a, w, scale, shift, act=1)
        ref = F.conv2d(a.float().permute(0, 3, 1, 2), w.float().view(N, 3, 3, K).permute(0, 3, 1, 2), padding=1)
        ref = F.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
        err = float((out.float() - ref).abs().max() / ref.abs().max())
        for _ in range(5):
            D.conv3x3(a, w, scale, shift, act=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            D.conv3x3(a, w, scale, shift, act=1)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"B={B} {H}x{W} {K}->{N}: max err {err:.2e}  {us:8.1f} us  {2.0 * B * H * W * N * K * 9 / us / 1e6:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if "--all" in sys.argv:
        for env in ({"MTGSEG_CONV3": "0"}, {"MTGSEG_CONV3": "1"}):
            print("==", env, flush=True)
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, **env), timeout=300)
            print("rc", r.returncode, flush=True)
    else:
        main()
