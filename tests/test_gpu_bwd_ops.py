"""Per-kernel pins of the small backward kernels of the training step (csrc/train_misc.cu: dot_pool, se_bwd, outer_sum,
head_bwd) against torch autograd on identical inputs, through the C ABI.  Tolerances: fp32 outputs <= 1e-4 of the tensor's
range (they are fp32 sums of exactly representable bf16 x fp32 products, only the summation order differs); bf16 outputs
<= 1e-2 (one rounding to bf16 of an fp32 value)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(__file__))
pytestmark = pytest.mark.gpu

from mtg_card_image_segmentation_b200 import _native as N  # noqa: E402
import devops as D  # noqa: E402


def _close(name, got, ref, tol):
    emax, _ = D.report(name, got, ref)
    assert emax <= tol, (name, emax)


@pytest.mark.parametrize("B,HW,C,SQ,chunks", [(3, 300, 72, 24, 1), (2, 1200, 120, 32, 5), (4, 300, 960, 240, 2), (33, 77, 672, 168, 16)])
def test_squeeze_excite_backward_vs_autograd(B, HW, C, SQ, chunks):
    """tv:ops/misc.py:252-261: s = hardsigmoid(fc2(relu(fc1(mean(y))))); out = s * y.  The gradient THROUGH the gate."""
    g = torch.Generator().manual_seed(B * 1000 + C)
    dev = "cuda"
    y = (torch.randn(B, HW, C, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    da = torch.randn(B, HW, C, generator=g).to(torch.bfloat16).to(dev)
    w1 = (torch.randn(SQ, C, generator=g) * (3.0 / C ** 0.5)).to(dev).requires_grad_()
    b1 = (torch.randn(SQ, generator=g) * 0.5).to(dev).requires_grad_()
    w2 = (torch.randn(C, SQ, generator=g) * (6.0 / SQ ** 0.5)).to(dev).requires_grad_()   # wide: saturates part of the hardsigmoid
    b2 = (torch.randn(C, generator=g)).to(dev).requires_grad_()
    yf = y.float()
    mean = yf.mean(1).detach().requires_grad_()
    hid = F.relu(mean @ w1.t() + b1)
    s = F.hardsigmoid(hid @ w2.t() + b2)
    loss = (s[:, None, :] * yf * da.float()).sum()
    dmean_ref, dw1_ref, db1_ref, dw2_ref, db2_ref = torch.autograd.grad(loss, [mean, w1, b1, w2, b2])
    sat = ((s.detach() <= 0) | (s.detach() >= 1)).float().mean().item()
    dead = (hid.detach() <= 0).float().mean().item()
    print(f"hardsigmoid saturated {sat:.2f}, relu dead {dead:.2f}")
    assert 0.02 < sat < 0.98 and 0.02 < dead < 0.98  # both branches of both derivatives are exercised
    # per-chunk channel sums of y, any partition of the rows (the forward's pool partials)
    bounds = [(k * HW) // chunks for k in range(chunks + 1)]
    gap = torch.stack([yf[:, bounds[k]:bounds[k + 1]].sum(1) for k in range(chunks)], 1).contiguous()
    lib = N.load()
    scratch = torch.empty(lib.mtgseg_se_bwd_scratch_floats(B, C, SQ), dtype=torch.float32, device=dev)
    dmean = torch.empty(B, C, device=dev); dw1 = torch.empty(SQ, C, device=dev); db1 = torch.empty(SQ, device=dev)
    dw2 = torch.empty(C, SQ, device=dev); db2 = torch.empty(C, device=dev)
    N.check(lib.mtgseg_se_block_bwd(da.data_ptr(), y.data_ptr(), s.detach().contiguous().data_ptr(), hid.detach().contiguous().data_ptr(),
                                    gap.data_ptr(), chunks, w1.data_ptr(), w2.data_ptr(), dmean.data_ptr(), dw1.data_ptr(),
                                    db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(), scratch.data_ptr(), B, HW, C, SQ, N.stream_ptr()),
            "se_block_bwd")
    for name, got, ref in (("dmean", dmean, dmean_ref), ("dw1", dw1, dw1_ref), ("db1", db1, db1_ref), ("dw2", dw2, dw2_ref),
                           ("db2", db2, db2_ref)):
        _close(f"SE backward {name} C={C}", got, ref, 1e-4)


@pytest.mark.parametrize("B,Hh,Wh,NC", [(2, 20, 15, 2), (5, 4, 3, 2), (3, 8, 6, 3)])
def test_head_tail_backward_vs_autograd(B, Hh, Wh, NC):
    """train/model.py:137-142: out = low_classifier(low) + high_classifier(up2(cbr * s)) at the low-level resolution."""
    IC, LC, Hl, Wl = 128, 40, 2 * Hh, 2 * Wh
    g = torch.Generator().manual_seed(B + Hh)
    dev = "cuda"
    cbr = torch.randn(B, Hh, Wh, IC, generator=g).relu().to(torch.bfloat16).to(dev)
    low = torch.randn(B, Hl, Wl, LC, generator=g).to(torch.bfloat16).to(dev)
    s = torch.rand(B, IC, generator=g).to(dev).requires_grad_()
    w_high = (torch.randn(NC, IC, generator=g) * 0.1).to(dev).requires_grad_()
    w_low = (torch.randn(NC, LC, generator=g) * 0.2).to(dev).requires_grad_()
    b_high = torch.randn(NC, generator=g).to(dev).requires_grad_()
    b_low = torch.randn(NC, generator=g).to(dev).requires_grad_()
    d_lowres = torch.randn(B, Hl, Wl, NC, generator=g).to(dev)
    cbr_f = cbr.float().requires_grad_()
    low_f = low.float().requires_grad_()
    h2 = torch.einsum("bhwi,ci->bchw", cbr_f * s[:, None, None, :], w_high)
    h2.retain_grad()
    up = F.interpolate(h2, size=(Hl, Wl), mode="bilinear", align_corners=False)
    lowres = up.permute(0, 2, 3, 1) + b_high + torch.einsum("bhwk,ck->bhwc", low_f, w_low) + b_low
    (lowres * d_lowres).sum().backward()
    d_h2 = h2.grad.permute(0, 2, 3, 1).contiguous()  # what mtgseg_upsample_bwd produces from d_lowres
    lib = N.load()
    up_t = torch.empty(B, Hh, Wh, NC, device=dev)
    N.check(lib.mtgseg_upsample_bwd(d_lowres.permute(0, 3, 1, 2).contiguous().data_ptr(), N.LOGITS_F32, up_t.data_ptr(), B, NC, Hh, Wh,
                                    Hl, Wl, N.stream_ptr()), "upsample_bwd")
    _close("x2 bilinear transpose", up_t, d_h2, 1e-5)
    dcbr = torch.empty(B, Hh, Wh, IC, dtype=torch.bfloat16, device=dev)
    dlow = torch.empty(B, Hl, Wl, LC, dtype=torch.bfloat16, device=dev)
    ds = torch.full((B, lib.mtgseg_head_bwd_segments(B), IC), float("nan"), device=dev); dwh = torch.zeros(NC, IC, device=dev); dwl = torch.zeros(NC, LC, device=dev)
    dbh = torch.zeros(NC, device=dev); dbl = torch.zeros(NC, device=dev)
    N.check(lib.mtgseg_head_bwd(d_lowres.data_ptr(), d_h2.data_ptr(), cbr.data_ptr(), s.detach().data_ptr(), low.data_ptr(),
                                w_high.data_ptr(), w_low.data_ptr(), dcbr.data_ptr(), ds.data_ptr(), dlow.data_ptr(), dwh.data_ptr(),
                                dwl.data_ptr(), dbh.data_ptr(), dbl.data_ptr(), B, Hh, Wh, Hl, Wl, IC, LC, NC, N.stream_ptr()), "head_bwd")
    _close("head bwd dcbr (bf16)", dcbr, cbr_f.grad, 1e-2)
    _close("head bwd dlow (bf16)", dlow, low_f.grad, 1e-2)
    _close("head bwd ds (segment partials summed)", ds.sum(1), s.grad, 1e-4)
    _close("head bwd dw_high", dwh, w_high.grad, 1e-4)
    _close("head bwd dw_low", dwl, w_low.grad, 1e-4)
    _close("head bwd db_high", dbh, b_high.grad, 1e-4)
    _close("head bwd db_low", dbl, b_low.grad, 1e-4)


@pytest.mark.parametrize("B,Hl,Wl,NC", [(2, 40, 30, 2), (3, 8, 6, 2), (1, 5, 7, 3)])
def test_lowres_loss_and_gradient_vs_reference_formula(B, Hl, Wl, NC):
    """CombinedLoss taken from the head's low-resolution logits (csrc/loss.cu lowres_loss_kernel, what the captured training step
    uses): equals train/utils.py:58-92 applied to F.interpolate(lowres, x8, bilinear) -- value <= 1e-5, gradient w.r.t. the
    low-resolution logits <= 1e-4 of its range; bit-identical between two runs (it feeds the backward chain)."""
    from oracle import lraspp_oracle as O
    H, W = 8 * Hl, 8 * Wl
    g = torch.Generator().manual_seed(Hl * 100 + NC)
    lowres = (torch.randn(B, Hl, Wl, NC, generator=g) * 2).cuda().requires_grad_()
    t = torch.randint(0, NC, (B, H, W), generator=g).cuda()
    logits = F.interpolate(lowres.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=False)
    loss_ref = O.combined_loss(logits, t)
    (d_ref,) = torch.autograd.grad(loss_ref, lowres)
    lib = N.load()
    outs = []
    for _ in range(2):
        d = torch.full_like(lowres, float("nan")).detach()
        scratch = torch.empty(lib.mtgseg_loss_lowres_scratch_floats(B, Hl, Wl), dtype=torch.float32, device="cuda")
        loss3 = torch.empty(3, dtype=torch.float32, device="cuda")
        N.check(lib.mtgseg_loss_lowres(lowres.detach().data_ptr(), t.data_ptr(), d.data_ptr(), scratch.data_ptr(), loss3.data_ptr(), B, Hl, Wl,
                                       H, W, NC, 0.5, 0.5, 1e-6, N.stream_ptr()), "loss_lowres")
        outs.append((loss3.clone(), d))
    assert abs(outs[0][0][0].item() - loss_ref.item()) <= 1e-5 * max(1.0, abs(loss_ref.item()))
    _close("lowres loss gradient", outs[0][1], d_ref, 1e-4)
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
