"""Per-operator parity of the CUDA kernels (through the C ABI) against plain PyTorch fp32 maths on the
same bf16-rounded inputs.  Tolerances: outputs are bf16 (8 mantissa bits) -> max error <= 1e-2 of the
output range and rel-L2 <= 4e-3 (half an ulp RMS)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import devops as D  # noqa: E402
from mtg_card_image_segmentation_b200 import _native as N_  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False  # the references below must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False
ACTS = {0: lambda x: x, 1: F.relu, 2: F.hardswish, 3: F.hardsigmoid, 4: torch.sigmoid}


def _rand(*shape, gen, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale)


def _check(name, got, ref, tol_max=1e-2, tol_l2=4e-3):
    emax, el2 = D.report(name, got, ref)
    assert not torch.isnan(got.float()).any(), name
    assert emax <= tol_max and el2 <= tol_l2, f"{name}: max-rel {emax:.3g} l2 {el2:.3g}"


@pytest.mark.parametrize("M,N,K,act,res,se", [
    (128, 64, 64, 0, False, False),      # one full tile, one k-block
    (300, 960, 160, 2, False, False),    # block-16 conv: 4 N tiles, partial M tile, K=160 (2.5 k-blocks)
    (4800, 160, 960, 0, True, True),     # block 14/15 project: K=960 (15 k-blocks), SE scale, residual
    (2 * 160 * 120, 64, 16, 1, False, False),  # block 2 expand: K=16 (one 16-wide k-step)
    (2 * 80 * 60, 24, 64, 0, False, False),    # N=24 -> padded MMA N=32
    (1200 * 3, 40, 72, 0, False, True),  # block 4 project with SE, K=72
    (777, 200, 80, 2, False, False),     # ragged M, N=200 -> 208
    (600, 184, 80, 2, False, False),
    (1, 16, 16, 0, False, False),        # smallest
    (148 * 128 * 9 + 5, 72, 24, 1, False, False),  # many tiles per persistent CTA
    # >= 8 tiles per SM: the decoupled per-warp epilogue (alternate tiles per warp set, per-warp TMA stores) ...
    (148 * 128 * 10, 16, 16, 0, True, False),   # ... with the residual tile double buffered and fetched by the producer (b1.project)
    (300 * 600, 160, 960, 0, True, True),       # ... with ONE residual buffer shared by both warp sets (15 k-blocks), SE gate prefetch
    (1200 * 160, 40, 120, 0, True, True),       # ... SE project at 40x30 (b5/b6)
    (148 * 128 * 9, 240, 40, 2, False, False),  # ... four 64-column slabs per tile (b7.expand)
    (300 * 512, 960, 160, 2, False, False),     # ... resident weights with five N tiles (b14.expand)
])
def test_conv1x1(M, N, K, act, res, se):
    g = torch.Generator().manual_seed(M + N + K)
    a = _rand(M, K, gen=g).bfloat16().to(DEV)
    w = _rand(N, K, gen=g, scale=K ** -0.5).bfloat16().to(DEV)
    scale = (torch.rand(N, generator=g) + 0.5).to(DEV)
    shift = _rand(N, gen=g, scale=0.2).to(DEV)
    residual = _rand(M, N, gen=g).bfloat16().to(DEV) if res else None
    hw = 300 if M % 300 == 0 else (1200 if M % 1200 == 0 else M)
    a_scale = torch.rand(M // hw, K, generator=g).to(DEV) if se else None
    got = D.conv1x1(a, w, scale, shift, act, residual, a_scale, hw)
    af = a.float()
    if se:
        af = (af.view(M // hw, hw, K) * a_scale[:, None, :]).bfloat16().float().view(M, K)
    ref = ACTS[act]((af @ w.float().t()) * scale + shift)
    if res:
        ref = ref + residual.float()
    _check(f"conv1x1 M{M} N{N} K{K}", got, ref)


# plain 3x3 on maps up to 63 pixels wide = the haloed-tile kernel (csrc/conv3_tc.cu): an odd number of M tiles (the last pair
# holds one tile), several N tiles with a ragged last one, the dgrad shape of the head conv (128 -> 960), a 40-wide map (three
# rows per tile), K with a ragged last 64-channel slab; wider maps (W = 70) keep the nine-shifted-boxes path of gemm_tc.cu
@pytest.mark.parametrize("B,H,W,N,K", [(2, 20, 15, 128, 960), (3, 4, 3, 128, 960), (5, 20, 15, 128, 64), (1, 9, 7, 32, 72),
                                       (1, 20, 15, 128, 64), (2, 30, 40, 200, 136), (3, 20, 15, 960, 128), (2, 40, 30, 48, 40),
                                       (2, 6, 70, 32, 64)])
def test_conv3x3(B, H, W, N, K):
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = _rand(B, H, W, K, gen=g).bfloat16().to(DEV)
    w = _rand(N, K, 3, 3, gen=g, scale=(9 * K) ** -0.5).bfloat16()
    wp = w.permute(0, 2, 3, 1).reshape(N, 9, K).contiguous().to(DEV)
    scale = (torch.rand(N, generator=g) + 0.5).to(DEV)
    shift = _rand(N, gen=g, scale=0.2).to(DEV)
    got = D.conv3x3(x, wp, scale, shift, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().to(DEV), None, 1, 1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1)
    _check(f"conv3x3 B{B} {H}x{W} N{N} K{K}", got, ref)


@pytest.mark.parametrize("B,H,W,C,k,stride,dil,act,gap", [
    (2, 160, 120, 16, 3, 1, 1, 1, False), (2, 160, 120, 64, 3, 2, 1, 1, False), (2, 80, 60, 72, 5, 2, 1, 1, True),
    (3, 40, 30, 120, 5, 1, 1, 1, True), (2, 40, 30, 240, 3, 2, 1, 2, False), (2, 20, 15, 184, 3, 1, 1, 2, False),
    (2, 20, 15, 672, 5, 1, 2, 2, True), (3, 20, 15, 960, 5, 1, 2, 2, True), (1, 7, 5, 200, 3, 1, 1, 2, False),
])
def test_dwconv(B, H, W, C, k, stride, dil, act, gap):
    g = torch.Generator().manual_seed(C + k)
    x = _rand(B, H, W, C, gen=g).bfloat16().to(DEV)
    w = _rand(C, 1, k, k, gen=g, scale=1.0 / k).bfloat16()
    wp = w.reshape(C, k * k).t().contiguous().to(DEV)
    scale = (torch.rand(C, generator=g) + 0.5).to(DEV)
    shift = _rand(C, gen=g, scale=0.2).to(DEV)
    got, gp = D.dwconv(x, wp, scale, shift, act, k, stride, dil, gap)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().to(DEV), None, stride, (k - 1) // 2 * dil, dil, C)
    ref = ACTS[act](ref * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1)
    _check(f"dwconv C{C} k{k} s{stride} d{dil}", got, ref)
    if gap:
        sums = gp.sum(1)
        ref_s = got.float().sum((1, 2))
        _check(f"dwconv gap C{C}", sums, ref_s, tol_max=1e-4, tol_l2=1e-5)


def test_stem():
    g = torch.Generator().manual_seed(1)
    x = _rand(3, 3, 64, 48, gen=g).to(DEV)
    w = _rand(16, 3, 3, 3, gen=g, scale=0.3)
    wp = w.reshape(16, 27).t().contiguous().to(DEV)
    scale = (torch.rand(16, generator=g) + 0.5).to(DEV)
    shift = _rand(16, gen=g, scale=0.2).to(DEV)
    got = D.stem(x, wp, scale, shift)
    ref = F.conv2d(x, w.to(DEV), None, 2, 1)
    ref = F.hardswish(ref * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1)
    _check("stem", got, ref)


@pytest.mark.parametrize("B,C,SQ,two", [(5, 960, 240, True), (3, 72, 24, True), (4, 960, 128, False), (2, 120, 32, True)])
def test_se_mlp(B, C, SQ, two):
    g = torch.Generator().manual_seed(C)
    hw = 300
    sums = (_rand(B, 3, C, gen=g) * 50).to(DEV)
    w1 = _rand(SQ, C, gen=g, scale=C ** -0.5).bfloat16().to(DEV)
    b1 = _rand(SQ, gen=g, scale=0.1).to(DEV) if two else None
    w2 = _rand(C, SQ, gen=g, scale=SQ ** -0.5).bfloat16().to(DEV) if two else None
    b2 = _rand(C, gen=g, scale=0.1).to(DEV) if two else None
    got = D.se_mlp(sums, hw, w1, b1, 1 if two else 4, w2, b2, 3)
    mean = sums.sum(1) / hw
    h = mean @ w1.float().t()
    if two:
        ref = F.hardsigmoid(F.relu(h + b1) @ w2.float().t() + b2)
    else:
        ref = torch.sigmoid(h)
    _check(f"se_mlp C{C}", got, ref, tol_max=1e-4, tol_l2=1e-5)


def test_gap():
    g = torch.Generator().manual_seed(3)
    x = _rand(3, 300, 960, gen=g).bfloat16().to(DEV)
    _check("gap", D.gap(x), x.float().sum(1), tol_max=1e-5, tol_l2=1e-6)


@pytest.mark.parametrize("NC,H,W", [(2, 320, 240), (3, 64, 48), (2, 70, 50)])
def test_head_tail(NC, H, W):
    g = torch.Generator().manual_seed(NC)
    B, IC, LC = 3, 128, 40
    Hl, Wl = (H + 7) // 8, (W + 7) // 8
    Hh, Wh = (Hl + 1) // 2, (Wl + 1) // 2
    cbr = F.relu(_rand(B, Hh, Wh, IC, gen=g)).bfloat16().to(DEV)
    s = torch.rand(B, IC, generator=g).to(DEV)
    low = _rand(B, Hl, Wl, LC, gen=g).bfloat16().to(DEV)
    wh, bh = _rand(NC, IC, gen=g, scale=0.1).to(DEV), _rand(NC, gen=g, scale=0.1).to(DEV)
    wl, bl = _rand(NC, LC, gen=g, scale=0.1).to(DEV), _rand(NC, gen=g, scale=0.1).to(DEV)
    lowres = D.head_mix(cbr, s, low, wh, bh, wl, bl)
    x = (cbr.float() * s[:, None, None, :]).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=(Hl, Wl), mode="bilinear", align_corners=False)
    ref_low = F.conv2d(low.float().permute(0, 3, 1, 2), wl[:, :, None, None], bl) + F.conv2d(x, wh[:, :, None, None], bh)
    _check("head_mix", lowres.permute(0, 3, 1, 2), ref_low, tol_max=1e-4, tol_l2=1e-5)
    t = torch.randint(0, 2, (B, H, W), generator=g).to(DEV)
    logits, mask, counts = D.upsample_out(lowres, H, W, torch.float32, True, True, t if NC == 2 else None)
    ref = F.interpolate(lowres.permute(0, 3, 1, 2).contiguous(), size=(H, W), mode="bilinear", align_corners=False)
    _check("upsample_out f32", logits, ref, tol_max=1e-5, tol_l2=1e-6)
    assert torch.equal(mask.long(), torch.argmax(logits, 1))
    if NC == 2:
        p = mask.long()
        exp = torch.bincount((t * 2 + p).reshape(-1), minlength=4)
        assert torch.equal(counts, exp)
    lb, _, _ = D.upsample_out(lowres, H, W, torch.bfloat16)
    _check("upsample_out bf16", lb, ref, tol_max=1e-2, tol_l2=4e-3)
    lh, _, _ = D.upsample_out(lowres, H, W, torch.float16)
    _check("upsample_out f16", lh, ref, tol_max=2e-3, tol_l2=1e-3)


def _guarded(shape, dtype, guard=4096):
    """A tensor of `shape` carved out of a larger buffer whose margins hold a sentinel: (view, check())."""
    n = 1
    for s in shape:
        n *= s
    sentinel = 0x5A if dtype == torch.uint8 else 12345.0
    buf = torch.full((n + 2 * guard,), sentinel, dtype=dtype, device=DEV)
    view = buf[guard:guard + n].view(*shape)

    def check():
        assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + n:] == sentinel).all()), "write outside the output tensor"
    return view, check


@pytest.mark.parametrize("M,N,K,res", [
    (148 * 128 * 9 + 77, 16, 16, True),    # per-warp epilogue, ragged last M tile, N below the slab width
    (148 * 128 * 8 + 1, 200, 80, False),   # per-warp epilogue, N = 200 padded to 208 columns, one valid row in the last tile
    (777, 24, 72, True),                   # CTA-wide epilogue (few tiles), N = 24 padded to 32
])
def test_conv1x1_writes_stay_inside_the_output(M, N, K, res):
    """compute-sanitizer is not available on the GPU pool: guard bands around the output instead (the TMA stores clip
    against the tensor map, ragged tiles and padded N columns must never touch memory past [M][N])."""
    g = torch.Generator().manual_seed(M)
    a = _rand(M, K, gen=g).bfloat16().to(DEV)
    w = _rand(N, K, gen=g, scale=K ** -0.5).bfloat16().to(DEV)
    residual = _rand(M, N, gen=g).bfloat16().to(DEV) if res else None
    out, check = _guarded((M, N), torch.bfloat16)
    lib = N_.load()
    N_.check(lib.mtgseg_conv1x1(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, None, None, 0, N_.ptr(residual), None, 0,
                                N_.stream_ptr()), "conv1x1")
    torch.cuda.synchronize()
    check()
    ref = a.float() @ w.float().t() + (residual.float() if res else 0)
    _check(f"guarded conv1x1 M{M} N{N} K{K}", out, ref)


@pytest.mark.parametrize("B,H,W,C,k,s,d", [(3, 20, 15, 184, 3, 1, 1), (2, 40, 30, 120, 5, 1, 1), (2, 21, 17, 72, 5, 2, 1), (2, 20, 15, 672, 5, 1, 2)])
def test_dwconv_writes_stay_inside_the_output(B, H, W, C, k, s, d):
    g = torch.Generator().manual_seed(C + k)
    x = _rand(B, H, W, C, gen=g).bfloat16().to(DEV)
    w = _rand(k * k, C, gen=g, scale=0.2).bfloat16().to(DEV)
    sc = (torch.rand(C, generator=g) + 0.5).to(DEV); sh = _rand(C, gen=g, scale=0.1).to(DEV)
    pad = (k - 1) // 2 * d
    Ho = (H + 2 * pad - d * (k - 1) - 1) // s + 1; Wo = (W + 2 * pad - d * (k - 1) - 1) // s + 1
    out, check = _guarded((B, Ho, Wo, C), torch.bfloat16)
    lib = N_.load()
    chunks = lib.mtgseg_dwconv_chunks(H, W, C, k, s, d, 1)
    gap, check_gap = _guarded((B, chunks, C), torch.float32)
    N_.check(lib.mtgseg_dwconv(x.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, C, k, s, d, sc.data_ptr(), sh.data_ptr(), 1,
                               gap.data_ptr(), chunks, N_.stream_ptr()), "dwconv")
    torch.cuda.synchronize()
    check(); check_gap()
    want, want_gap = D.dwconv(x, w, sc, sh, 1, k, s, d, True)
    assert torch.equal(out, want) and torch.equal(gap, want_gap)


@pytest.mark.parametrize("B,H,W", [(2, 46, 46), (1, 33, 21), (2, 320, 240)])
def test_stem_writes_stay_inside_the_output(B, H, W):
    """Two output pixels per thread: an odd output width leaves the second pixel of the last pair without a home."""
    g = torch.Generator().manual_seed(H * W)
    x = _rand(B, 3, H, W, gen=g).to(DEV)
    w = _rand(27, 16, gen=g, scale=0.3).to(DEV)
    sc = (torch.rand(16, generator=g) + 0.5).to(DEV); sh = _rand(16, gen=g, scale=0.1).to(DEV)
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out, check = _guarded((B, Ho, Wo, 16), torch.bfloat16)
    N_.check(N_.load().mtgseg_stem(x.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(), B, H, W, N_.stream_ptr()), "stem")
    torch.cuda.synchronize()
    check()
    assert torch.equal(out, D.stem(x, w, sc, sh))
