"""Test-side helpers that call the C ABI per operator with torch tensors (tests only)."""
import ctypes as C

import torch

from mtg_card_image_segmentation_b200 import _native as N


def _s():
    return N.stream_ptr()


def conv1x1(a, w, scale=None, shift=None, act=0, residual=None, a_scale=None, hw=0):
    M, K = a.shape
    Nn = w.shape[0]
    out = torch.empty(M, Nn, dtype=torch.bfloat16, device=a.device)
    N.check(N.load().mtgseg_conv1x1(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, Nn, K, N.ptr(scale), N.ptr(shift), act,
                                    N.ptr(residual), N.ptr(a_scale), hw, _s()), "conv1x1")
    return out


def conv3x3(a, w, scale=None, shift=None, act=0):
    B, H, W, K = a.shape
    Nn = w.shape[0]
    out = torch.empty(B, H, W, Nn, dtype=torch.bfloat16, device=a.device)
    N.check(N.load().mtgseg_conv3x3(a.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, Nn, K, N.ptr(scale), N.ptr(shift),
                                    act, _s()), "conv3x3")
    return out


def dwconv(x, w, scale, shift, act, k, stride, dil, need_gap=False):
    B, H, W, Cc = x.shape
    pad = (k - 1) // 2 * dil
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out = torch.empty(B, Ho, Wo, Cc, dtype=torch.bfloat16, device=x.device)
    lib = N.load()
    chunks = lib.mtgseg_dwconv_chunks(H, W, Cc, k, stride, dil, int(need_gap))
    gap = torch.zeros(B, chunks, Cc, dtype=torch.float32, device=x.device) if need_gap else None
    N.check(lib.mtgseg_dwconv(x.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, Cc, k, stride, dil, scale.data_ptr(),
                              shift.data_ptr(), act, N.ptr(gap), chunks, _s()), "dwconv")
    return out, gap


def stem(x, w, scale, shift):
    B, _, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = torch.empty(B, Ho, Wo, 16, dtype=torch.bfloat16, device=x.device)
    N.check(N.load().mtgseg_stem(x.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), B, H, W, _s()),
            "stem")
    return out


def se_mlp(sums, hw, w1, b1, act1, w2=None, b2=None, act2=0):
    B, chunks, Cc = sums.shape
    SQ = w1.shape[0]
    out = torch.empty(B, Cc if w2 is not None else SQ, dtype=torch.float32, device=sums.device)
    hidden = torch.empty(B, SQ, dtype=torch.float32, device=sums.device) if w2 is not None else None
    N.check(N.load().mtgseg_se_mlp(sums.data_ptr(), chunks, B, Cc, SQ, hw, w1.data_ptr(), N.ptr(b1), act1, N.ptr(w2), N.ptr(b2),
                                   act2, out.data_ptr(), N.ptr(hidden), _s()), "se_mlp")
    return out


def gap(x):
    B, HW, Cc = x.shape
    out = torch.empty(B, Cc, dtype=torch.float32, device=x.device)
    N.check(N.load().mtgseg_gap(x.data_ptr(), out.data_ptr(), B, HW, Cc, _s()), "gap")
    return out


def head_mix(cbr, s, low, w_high, b_high, w_low, b_low):
    B, Hh, Wh, IC = cbr.shape
    _, Hl, Wl, LC = low.shape
    NC = w_high.shape[0]
    out = torch.empty(B, Hl, Wl, NC, dtype=torch.float32, device=cbr.device)
    N.check(N.load().mtgseg_head_mix(cbr.data_ptr(), s.data_ptr(), low.data_ptr(), w_high.data_ptr(), b_high.data_ptr(),
                                     w_low.data_ptr(), b_low.data_ptr(), out.data_ptr(), B, Hh, Wh, Hl, Wl, IC, LC, NC, _s()),
            "head_mix")
    return out


def upsample_out(lowres, H, W, dtype=torch.float32, want_logits=True, want_mask=False, targets=None):
    B, Hl, Wl, NC = lowres.shape
    dt = {torch.float32: 1, torch.bfloat16: 2, torch.float16: 3}[dtype]
    logits = torch.empty(B, NC, H, W, dtype=dtype, device=lowres.device) if want_logits else None
    mask = torch.empty(B, H, W, dtype=torch.uint8, device=lowres.device) if want_mask else None
    counts = torch.zeros(4, dtype=torch.int64, device=lowres.device) if targets is not None else None
    N.check(N.load().mtgseg_upsample_out(lowres.data_ptr(), N.ptr(logits), dt, N.ptr(mask), N.ptr(targets), N.ptr(counts),
                                         B, Hl, Wl, H, W, NC, _s()), "upsample_out")
    return logits, mask, counts


def report(name, got, ref):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().clamp_min(1e-12)
    rel_l2 = (got - ref).norm() / ref.norm().clamp_min(1e-12)
    idx = int(err.argmax())
    msg = (f"[{name}] shape={tuple(ref.shape)} max_abs_err={err.max().item():.4g} at flat {idx} "
           f"(got {got.reshape(-1)[idx].item():.5g} ref {ref.reshape(-1)[idx].item():.5g}) "
           f"max_ref={denom.item():.4g} rel_l2={rel_l2.item():.4g} nan={int(torch.isnan(got).sum())}")
    print(msg)
    return err.max().item() / denom.item(), rel_l2.item()
