import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import devops as D
B, hw = 256, 300
dev = "cuda"; g = torch.Generator(device=dev).manual_seed(0)
for (K, N) in [(960, 160), (672, 112), (480, 112), (120, 40)]:
    hwl = 300 if K >= 480 else 1200
    M = B * hwl
    xs = [torch.randn(M, K, device=dev, generator=g).bfloat16() for _ in range(3)]
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    sc = torch.rand(N, device=dev, generator=g) + 0.5; sh = torch.randn(N, device=dev, generator=g) * 0.1
    r = torch.randn(M, N, device=dev, generator=g).bfloat16()
    gate = torch.rand(B, K, device=dev, generator=g)
    for label, res, se in (("plain", None, None), ("res", r, None), ("se", None, gate), ("se+res", r, gate)):
        for i in range(3): D.conv1x1(xs[i], w, sc, sh, 0, res, se, hwl if se is not None else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(6): D.conv1x1(xs[i % 3], w, sc, sh, 0, res, se, hwl if se is not None else 0)
        e1.record(); torch.cuda.synchronize()
        print(f"K={K} N={N} hw={hwl} {label}: {e0.elapsed_time(e1)*1e3/6:.1f} us")
    del xs, r
    torch.cuda.empty_cache()
