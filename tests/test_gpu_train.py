"""Training-step parity on the B200: per-kernel backward checks against torch autograd (fp32 maths on the same
bf16-rounded inputs) and the whole step (forward with batch statistics, loss, backward, AdamW) against the oracle.

Gradient reference: at random-init weights the gradients of this network are chaotic in the activation precision --
on the CPU alone, the fp32 oracle and the same oracle with bf16-rounded activations (straight-through rounding,
``O.forward(q=ste_bf16, wq=ste_bf16)``) agree only to a median cosine of 0.89 (64x48, B=4) / 0.95 (320x240, B=2).
The whole CUDA step is therefore checked against the oracle differentiated AT the bf16-rounded activations, i.e. the
function the kernels actually evaluate, with thresholds that reflect that residual chaos (the two forwards still
differ by accumulation order: ~1.6 % rel-L2 in the logits): at 320x240 per parameter tensor cosine >= 0.88 and norm
ratio within 30 %, median cosine >= 0.97 (0.75 / 0.94 for the tiny 64x48 fixture whose BatchNorms see 48 values) (tensors whose true gradient is structurally zero -- a BatchNorm shift followed by
another train-mode BatchNorm -- are held to an absolute bound instead); the fp32-oracle agreement is printed for
the record.  The exact correctness of every backward kernel is pinned separately, per kernel, against torch autograd
on identical inputs (second half of this file, tolerances 1e-2 .. 1e-5)."""
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
    m.inference_precision = "bf16"  # eval-mode checks in this file compare against the bf16-emulated oracle
    return m.cuda().train()


def _oracle_step(sd, x, m, q=None):
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    upd = {}
    y = O.forward(sdg, x, training=True, bn_updates=upd, q=q, wq=q)
    loss = O.combined_loss(y, m)
    loss.backward()
    return y.detach(), loss.detach(), {k: v.grad for k, v in sdg.items() if v.dtype.is_floating_point and "running" not in k}, upd


@pytest.mark.parametrize("B,H,W,seed", [(4, 64, 48, 7), (2, 320, 240, 11), (32, 320, 240, 13)])
def test_train_step_vs_oracle(B, H, W, seed):
    """(32, 320, 240) is BASELINE.json configs[2]: the batch and resolution the training bench leg runs."""
    x, m = O.synthetic_cards(B, seed=seed, height=H, width=W)
    sd = O.calibrate_running_stats(O.make_weights(seed + 1), x)
    y_ref, loss_ref, g_ref, upd = _oracle_step(sd, x, m)
    y_emu, loss_emu, g_emu, _ = _oracle_step(sd, x, m, q=O.ste_bf16)
    model = _train_model(sd)
    crit = M.CombinedLoss(0.5, 0.5)
    logits = model(x.cuda())
    loss = crit(logits, m.cuda())
    loss.backward()
    emax, el2 = D.report("train-mode logits vs bf16-emulated oracle", logits.detach().cpu(), y_emu)
    # rel-L2 is the stable number (1.4-1.7e-2 on every shape); the MAX over 4.9 M logits at B=32 sits at 1.8-2.3e-2 depending on
    # the summation order of the statistics alone (measured with MTGSEG_BN_EPI=0/1), so it gets a wider bound there
    assert emax <= (3e-2 if B >= 32 else 2e-2) and el2 <= 2e-2
    emax, el2 = D.report("train-mode logits vs fp32 oracle", logits.detach().cpu(), y_ref)
    assert emax <= 8e-2 and el2 <= 8e-2
    assert abs(loss.item() - loss_emu.item()) <= 5e-3 * abs(loss_emu.item())
    # running statistics (momentum 0.01 backbone / 0.1 head, unbiased variance) and the step counter
    got = model.state_dict()
    for k in ("model.backbone.0.1.running_mean", "model.backbone.0.1.running_var", "model.classifier.cbr.1.running_mean",
              "model.classifier.cbr.1.running_var", "model.backbone.15.block.1.1.running_var", "model.backbone.4.block.3.1.running_mean"):
        torch.testing.assert_close(got[k].cpu(), upd[k], rtol=2e-2, atol=2e-3)
    assert int(got["model.backbone.0.1.num_batches_tracked"]) == int(sd["model.backbone.0.1.num_batches_tracked"]) + 1
    norms = torch.tensor([float(v.norm()) for v in g_emu.values()])
    typical = float(norms.median())
    bad, cos_emu, cos_f32 = [], [], []
    for name, p in model.named_parameters():
        g, r, r32 = p.grad.detach().cpu().double().flatten(), g_emu[name].double().flatten(), g_ref[name].double().flatten()
        cos = float((g @ r) / (g.norm() * r.norm()).clamp_min(1e-30))
        cos_f32.append(float((g @ r32) / (g.norm() * r32.norm()).clamp_min(1e-30)))
        ratio = float(g.norm() / r.norm().clamp_min(1e-30))
        if float(r.norm()) < 1e-3 * typical:  # structurally zero gradient: only bound the noise
            ok = float(g.norm()) <= 0.2 * typical
        else:
            cos_emu.append(cos)
            ok = cos >= (0.88 if H >= 320 else 0.75) and 0.7 <= ratio <= 1.3
        if not ok:
            bad.append((name, round(cos, 4), round(ratio, 4), float(r.norm())))
    cos_emu.sort(); cos_f32.sort()
    print(f"cosine vs bf16-emulated oracle: min {cos_emu[0]:.4f} median {cos_emu[len(cos_emu)//2]:.4f}; "
          f"vs fp32 oracle: min {cos_f32[0]:.4f} median {cos_f32[len(cos_f32)//2]:.4f}")
    print(f"{len(bad)} of {len(g_emu)} parameter gradients outside tolerance: {bad[:16]}")
    assert not bad
    assert cos_emu[len(cos_emu) // 2] >= (0.97 if H >= 320 else 0.94)


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
    model.inference_precision = "bf16"
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


def test_grad_scaler_loop_matches_plain_step():
    """train/train.py:96-105: scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update().  The scaled gradient flows back
    through the fused loss and the native backward; after unscaling, one AdamW step must land where the unscaled step lands
    (the scale is a power of two: exact in fp32, up to the bf16 rounding of the scaled activations' gradients)."""
    x, m = O.synthetic_cards(4, seed=5, height=64, width=48)
    sd = O.make_weights(41)
    xc, mc = x.cuda(), m.cuda()

    def one_step(scale):
        model = _train_model(sd)
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
        crit = M.CombinedLoss()
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xc), mc)
        if scale:
            scaler = torch.amp.GradScaler("cuda", init_scale=scale)
            scaler.scale(loss).backward()
            scaler.unscale_(opt)  # as for gradient clipping in the reference's loop
            scaler.step(opt)
            scaler.update()
            assert scaler.get_scale() == scale  # no inf / nan was found
        else:
            loss.backward()
            opt.step()
        grads.append([p.grad.detach().float().cpu().clone() * (1.0 if scale else 1.0) for p in model.parameters()])
        return float(loss.detach()), [p.detach().float().cpu().clone() for p in model.parameters()]

    grads = []
    l0, p0 = one_step(0)
    l1, p1 = one_step(1024.0)
    assert l0 == l1
    worst = max(float((a - b).abs().max()) for a, b in zip(p0, p1))
    names = [n for n, _ in _train_model(sd).named_parameters()]
    for n, a, b, ga, gb in zip(names, p0, p1, grads[0], grads[1]):
        d = (a - b).abs()
        if float(d.max()) > 1e-4:
            i = int(d.argmax())
            print(f"  differs: {n} elem {i}: dparam {float(d.max()):.3e} g_plain {float(ga.flatten()[i]):.4e} g_scaled(after unscale) "
                  f"{float(gb.flatten()[i]):.4e}; elems {int((d > 1e-4).sum())}/{d.numel()}")
    moved = max(float((a - sd[k].float()).abs().max()) for a, (k, _) in zip(p0, [kv for kv in _train_model(sd).named_parameters()]))
    print(f"GradScaler step vs plain step: max param difference {worst:.3e} (a step moves parameters by up to {moved:.3e})")
    # AdamW normalises the gradient: a 1e-3 step; the two runs may differ by bf16 rounding of scaled vs unscaled gradients
    assert moved > 5e-4 and worst <= 0.05 * moved  # measured 4.5e-6 vs 1e-3


def test_graphed_train_step_matches_eager_loop():
    """engine.GraphedTrainStep = train/train.py:88-110's loop body replayed as one CUDA graph.  Building it must leave the model,
    the BatchNorm statistics and the optimizer untouched; the replays must follow the eager loop on the same batches (same loss at
    step 1 bit for bit: the forward is deterministic; later steps up to the atomics' summation order in the weight gradients),
    honour learning-rate changes made between steps (the captured scalars are frozen, the device block is not), advance
    num_batches_tracked / the optimizer step counts, and leave the model ready for eval."""
    from mtg_card_image_segmentation_b200.engine import GraphedTrainStep
    sd = O.make_weights(47)
    data = [O.synthetic_cards(4, seed=s, height=64, width=48) for s in (21, 22, 23)]
    data = [(a.cuda(), b.cuda()) for a, b in data]
    lrs = [3e-4, 1e-3, 5e-4]

    def final_state(model):
        return {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}

    def eager():
        model = _train_model(sd)
        opt, crit, losses = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4), M.CombinedLoss(), []
        for (xb, mb), lr in zip(data, lrs):
            opt.param_groups[0]["lr"] = lr
            opt.zero_grad(set_to_none=True)
            loss = crit(model(xb), mb)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return losses, final_state(model), model

    def graphed():
        model = _train_model(sd)
        opt, crit, losses = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4), M.CombinedLoss(), []
        g = GraphedTrainStep(model, crit, opt, data[2][0], data[2][1])  # example batch: any data of the right shape
        untouched = final_state(model)
        for k, v in sd.items():
            assert torch.equal(untouched[k], v.float()), f"building the graph changed {k}"
        assert all(float(st["step"]) == 0 and not st["exp_avg"].any() for st in opt.state.values())
        first = None
        for (xb, mb), lr in zip(data, lrs):
            opt.param_groups[0]["lr"] = lr
            losses.append(float(g.step(xb, mb)))
            if first is None:
                first = final_state(model)
        assert all(float(st["step"]) == 3 for st in opt.state.values())
        return losses, final_state(model), model, first, g.launches_per_replay

    le, se_, me = eager()
    lg, sg, mg, first, launches = graphed()
    print(f"losses eager {le} graphed {lg}; {launches} launches per replay")
    assert launches > 300
    assert abs(lg[0] - le[0]) <= 2e-6 * abs(le[0])  # same forward bits; the captured step sums the loss over the 40x30 grid's owners
    # step 2 sees weights that differ only by the atomics' summation order in the weight gradients; by step 3 AdamW's sign-like first
    # steps (a gradient at noise level moves its weight by +-lr either way) have amplified that: measured 4e-6 at step 2 and up to
    # 3.8e-3 at step 3 when the whole GPU suite ran before it in the same process (below 2e-3 when the file runs alone).  The
    # trajectory is chaotic from there on, so the step-3 bounds below are statistical (3x the largest value seen), not tight.
    assert abs(lg[1] - le[1]) <= 2e-3 * abs(le[1]) and abs(lg[2] - le[2]) <= 3e-2 * abs(le[2])
    # AdamW's first step moves every weight by ~lr*sign(g): the step taken must be the 3e-4 set after construction, not the 1e-3 the
    # graph was captured under
    pk = [k for k, _ in me.named_parameters()]
    step1 = max(float((first[k] - sd[k].float()).abs().max()) for k in pk)
    assert 2.4e-4 <= step1 <= 3.6e-4, step1
    moved = max(float((se_[k] - sd[k].float()).abs().max()) for k in pk)
    worst = max(float((se_[k] - sg[k]).abs().max()) for k in pk)
    n = sum(se_[k].numel() for k in pk)
    mean_moved = sum(float((se_[k] - sd[k].float()).abs().sum()) for k in pk) / n
    mean_diff = sum(float((se_[k] - sg[k]).abs().sum()) for k in pk) / n
    print(f"after 3 steps: graphed vs eager parameter difference mean {mean_diff:.3e} max {worst:.3e}; the steps moved parameters by "
          f"mean {mean_moved:.3e} max {moved:.3e}")
    # AdamW's early steps are sign-like: a weight whose gradient is at noise level (atomics' summation order, amplified by the bf16
    # activations of a random-init net) may move by up to lr in either direction, so the bound is on the mean, not on the worst
    assert mean_diff <= 0.3 * mean_moved and worst <= 1.5 * moved, (mean_diff, mean_moved, worst, moved)
    assert int(sg["model.backbone.0.1.num_batches_tracked"]) == int(se_["model.backbone.0.1.num_batches_tracked"]) == 3
    rs = [k for k in sd if k.endswith("running_var")]
    rv_diff = max(float((se_[k] - sg[k]).abs().max() / se_[k].abs().max()) for k in rs)
    assert rv_diff <= 5e-2, rv_diff
    with torch.no_grad():
        ze, zg = me.eval()(data[0][0]).float(), mg.eval()(data[0][0]).float()
    assert float((ze - zg).abs().max()) <= 0.2 * float(ze.abs().max()), (float((ze - zg).abs().max()), float(ze.abs().max()))


def test_graphed_train_step_under_autocast_and_rebuild():
    """train/train.py:96-99 runs the forward under autocast: a GraphedTrainStep built inside the same context produces reduced-
    precision logits like the eager loop (same first loss, bit for bit), a second graph built on the same model / optimizer after
    some steps continues from the current state, and an eager step afterwards still works (the graph does not leave the optimizer
    or the engine in a captured-only mode)."""
    from mtg_card_image_segmentation_b200.engine import GraphedTrainStep
    sd = O.make_weights(53)
    x, m = O.synthetic_cards(4, seed=31, height=64, width=48)
    xc, mc = x.cuda(), m.cuda()
    crit = M.CombinedLoss()

    def eager_losses(n):
        model = _train_model(sd)
        opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
        out = []
        for _ in range(n):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = model(xc)
                assert logits.dtype == torch.bfloat16
                loss = crit(logits, mc)
            loss.backward()
            opt.step()
            out.append(float(loss.detach()))
        return out

    ref = eager_losses(4)
    model = _train_model(sd)
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        g1 = GraphedTrainStep(model, crit, opt, xc, mc)
    got = [float(g1.step(xc, mc)) for _ in range(2)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        g2 = GraphedTrainStep(model, crit, opt, xc, mc, lowres_loss=True)  # rebuilt mid-training: must continue from step 2, not restart (and: loss from the 40x30 logits)
    assert all(float(st["step"]) == 2 for st in opt.state.values())
    got.append(float(g2.step(xc, mc)))
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = crit(model(xc), mc)
    loss.backward()
    opt.step()
    got.append(float(loss.detach()))
    print(f"autocast losses eager {ref} graphed/rebuilt/eager {got}")
    # (the captured step takes the loss from the fp32 low-resolution logits; the eager loop from the bf16-rounded full-resolution ones)
    assert abs(got[0] - ref[0]) <= 1e-4 * abs(ref[0])
    assert all(abs(a - b) <= 5e-3 * abs(a) for a, b in zip(ref, got))
    assert all(float(st["step"]) == 4 for st in opt.state.values())
    assert int(model.state_dict()["model.backbone.0.1.num_batches_tracked"]) == 4


def test_structured_pruning_flow():
    """train/prune.py:76-93 (`PRUNING_STRUCTURED = True`): `prune.ln_structured(module, 'weight', amount=int(out_channels * 0.3),
    n=2, dim=0)` on every Conv2d.  torch keeps the tensor shapes and masks whole output channels, so on the drop-in it is the
    mask path again: evaluation with the masks active equals the oracle on the masked weights, whole filters are zero in the
    effective weights, a training step leaves them without gradient, and prune.remove restores the 319-key layout."""
    import torch.nn.utils.prune as prune
    x, m = O.synthetic_cards(4, seed=10, height=64, width=48)
    sd = O.calibrate_running_stats(O.make_weights(44), x)
    xc, mc = x.cuda(), m.cuda()
    model = _train_model(sd).eval()
    pruned_filters = 0
    for mod in model.model.modules():  # the loop of apply_structured_pruning
        if isinstance(mod, torch.nn.Conv2d) and mod.out_channels > 1:
            n = int(mod.out_channels * 0.3)
            if n > 0:
                prune.ln_structured(mod, name="weight", amount=n, n=2, dim=0)
                pruned_filters += n
    assert pruned_filters > 1000
    masked_sd = {k: v.detach().cpu() for k, v in zip(model._ref_keys, model._state_tensors())}
    dead = sum(int((masked_sd[k].flatten(1).abs().sum(1) == 0).sum()) for k in sd if sd[k].dim() == 4)
    assert dead >= pruned_filters
    with torch.no_grad():
        got = model(xc).float().cpu()
    ref = O.forward_bf16_emulated(masked_sd, x)
    assert D.report("structured pruning (masks active) vs emulated oracle on masked weights", got, ref)[0] <= 2e-2
    model.train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0)
    crit = M.CombinedLoss()
    opt.zero_grad(set_to_none=True)
    loss = crit(model(xc), mc)
    loss.backward()
    opt.step()
    assert float(loss.detach()) == float(loss.detach())
    for mod in model.model.modules():
        if hasattr(mod, "weight_mask"):
            assert bool((mod.weight_orig.grad[mod.weight_mask == 0] == 0).all())
            prune.remove(mod, "weight")
    assert sorted(model.state_dict().keys()) == sorted(sd.keys())


def test_pruning_flow_masks_active_then_removed():
    """train/prune.py:52-113,144-175 on the drop-in model: global magnitude pruning of the conv children, evaluation and fine-tuning
    with the masks ACTIVE (weight = weight_orig * weight_mask, maintained by torch's pruning hook), then prune.remove.  No call
    into the package is needed: the state-tensor cache notices the module surgery by itself."""
    import torch.nn.utils.prune as prune
    x, m = O.synthetic_cards(4, seed=9, height=64, width=48)
    sd = O.calibrate_running_stats(O.make_weights(43), x)
    xc, mc = x.cuda(), m.cuda()
    model = _train_model(sd).eval()
    with torch.no_grad():
        dense = model(xc).float().cpu()
    convs = [(mod, "weight") for mod in model.model.modules() if isinstance(mod, torch.nn.Conv2d)]
    prune.global_unstructured(convs, pruning_method=prune.L1Unstructured, amount=0.3)
    assert len(list(model.state_dict().keys())) > 319  # weight_orig + weight_mask: the state_dict no longer has the reference layout
    # (1) evaluation with masks active == the oracle on the masked weights
    masked_sd = {k: v.detach().cpu() for k, v in zip(model._ref_keys, model._state_tensors())}
    assert list(masked_sd) == list(sd) and all(masked_sd[k].shape == sd[k].shape for k in sd)
    zeros = sum(int((masked_sd[k] == 0).sum()) for k in sd if sd[k].dim() == 4)
    total = sum(sd[k].numel() for k in sd if sd[k].dim() == 4)
    assert abs(zeros / total - 0.3) < 0.01
    with torch.no_grad():
        pruned = model(xc).float().cpu()
    ref = O.forward_bf16_emulated(masked_sd, x)
    assert D.report("pruned (masks active) vs emulated oracle on masked weights", pruned, ref)[0] <= 2e-2
    assert float((pruned - dense).abs().max()) > 1e-3  # pruning did change the function
    # (2) fine-tuning with masks active: masked positions get no gradient and stay zero in the effective weight
    model.train()
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.0)
    crit = M.CombinedLoss()
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xc), mc)
        loss.backward()
        losses.append(float(loss.detach()))
        opt.step()
    assert all(l == l for l in losses)
    for mod, _ in convs:
        g = mod.weight_orig.grad
        assert g is not None and bool((g[mod.weight_mask == 0] == 0).all())
    moved = max(float((mod.weight_orig.detach().cpu() - sd[k]).abs().max()) for k, (mod, a) in zip(model._ref_keys, model._state_slots())
                if a == "weight" and hasattr(mod, "weight_orig"))
    assert moved > 1e-4  # the surviving weights were trained
    model.eval()
    with torch.no_grad():
        tuned = model(xc).float()
    eff = model._state_tensors()
    assert all(bool((t[mod.weight_mask == 0] == 0).all()) for t, (mod, a) in zip(eff, model._state_slots()) if a == "weight" and hasattr(mod, "weight_mask"))
    # (3) prune.remove makes the masks permanent: same function, reference state_dict layout again
    for mod, _ in convs:
        prune.remove(mod, "weight")
    assert sorted(model.state_dict().keys()) == sorted(sd.keys())
    with torch.no_grad():
        final = model(xc).float()
    assert torch.equal(final, tuned)


# ------------------------------------------------------------------------------------------------------------
# per-kernel backward checks against torch autograd on identical (bf16-rounded) inputs
# ------------------------------------------------------------------------------------------------------------
from mtg_card_image_segmentation_b200 import _native as N  # noqa: E402

ACTS = {0: lambda t: t, 1: F.relu, 2: F.hardswish}


def _chk(name, got, ref, tol_max=1e-2, tol_l2=5e-3):
    emax, el2 = D.report(name, got, ref)
    assert emax <= tol_max and el2 <= tol_l2, name


@pytest.mark.parametrize("B,HW,C,act,res,se", [(3, 300, 960, 2, False, True), (2, 4800, 72, 1, False, False),
                                               (4, 1200, 40, 0, True, False), (2, 19200, 16, 2, False, False),
                                               # B * chunks > 512: the reducing kernels walk two images per CTA (last CTA: one)
                                               (21, 19200, 64, 1, False, False), (19, 19200, 64, 2, True, True)])
def test_bn_train_fwd_bwd(B, HW, C, act, res, se):
    g = torch.Generator().manual_seed(C)
    lib = N.load()
    z = (torch.randn(B, HW, C, generator=g) * 1.5 + 0.3).bfloat16().cuda()
    residual = torch.randn(B, HW, C, generator=g).bfloat16().cuda() if res else None
    gamma = (torch.rand(C, generator=g) + 0.5).cuda(); beta = (torch.randn(C, generator=g) * 0.2).cuda()
    rm = torch.zeros(C).cuda(); rv = torch.ones(C).cuda(); nbt = torch.zeros((), dtype=torch.int64).cuda()
    scale, shift, mean, rstd = (torch.empty(C).cuda() for _ in range(4))
    nscr = lib.mtgseg_bn_scratch_floats(B, HW, C)
    scratch = torch.empty(nscr + 2 * C).cuda()
    y = torch.empty_like(z)
    gap = torch.zeros(B, 4, C).cuda() if se else None
    N.check(lib.mtgseg_bn_train_fwd(z.data_ptr(), y.data_ptr(), N.ptr(residual), gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.01,
                                    rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), scratch.data_ptr(), N.ptr(gap), 4, act, B, HW, C, N.stream_ptr()), "bn_fwd")
    zf = z.float().requires_grad_(True)
    gam = gamma.clone().requires_grad_(True); bet = beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C).cuda(), torch.ones(C).cuda()
    yr = ACTS[act](F.batch_norm(zf.permute(0, 2, 1), rm2, rv2, gam, bet, True, 0.01, 1e-3).permute(0, 2, 1))
    yfull = yr + residual.float() if res else yr
    _chk("bn fwd y", y, yfull.detach())
    torch.testing.assert_close(rm, rm2, rtol=1e-4, atol=1e-6); torch.testing.assert_close(rv, rv2, rtol=1e-4, atol=1e-6)
    assert int(nbt) == 1
    if se:
        _chk("bn fwd gap", gap.sum(1), y.float().sum(1), 1e-4, 1e-5)
    dy = torch.randn(B, HW, C, generator=g).bfloat16().cuda()
    se_s = torch.rand(B, C, generator=g).cuda() if se else None
    se_dm = torch.randn(B, C, generator=g).cuda() if se else None
    dz = torch.empty_like(z); dgam = torch.empty(C).cuda(); dbet = torch.empty(C).cuda()
    N.check(lib.mtgseg_bn_train_bwd(z.data_ptr(), dy.data_ptr(), dz.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), N.ptr(se_s), N.ptr(se_dm), scratch.data_ptr(), dgam.data_ptr(), dbet.data_ptr(),
                                    act, B, HW, C, N.stream_ptr()), "bn_bwd")
    dyf = dy.float()
    if se:
        dyf = dyf * se_s[:, None, :] + se_dm[:, None, :] / HW
    yr.backward(dyf)
    _chk("bn bwd dz", dz, zf.grad)
    _chk("bn bwd dgamma", dgam, gam.grad, 1e-3, 1e-3)
    _chk("bn bwd dbeta", dbet, bet.grad, 1e-3, 1e-3)


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("M,N_,K,taps,H,W,se", [(9600, 160, 960, 1, 0, 0, True), (2 * 19200, 64, 16, 1, 0, 0, False),
                                                (4 * 300, 128, 960, 9, 20, 15, False), (777 * 3, 24, 72, 1, 0, 0, False),
                                                (32 * 300, 960, 160, 1, 0, 0, False), (6 * 1200, 40, 120, 1, 0, 0, True)])
def test_wgrad(M, N_, K, taps, H, W, se, tc):
    g = torch.Generator().manual_seed(M)
    dz = torch.randn(M, N_, generator=g).bfloat16().cuda(); x = torch.randn(M, K, generator=g).bfloat16().cuda()
    hw = 300 if M % 300 == 0 else (1200 if M % 1200 == 0 else 777)
    if M % 19200 == 0:
        hw = 19200
    a_scale = torch.rand(M // hw, K, generator=g).cuda() if se else None
    dw = torch.zeros(N_, K, taps).cuda()
    if tc:  # tensor-core path: the SE gate multiplies the fp32 accumulator (no bf16 rounding of x*s)
        N.check(N.load().mtgseg_wgrad_tc(dz.data_ptr(), x.data_ptr(), dw.data_ptr(), N.ptr(a_scale), M // hw, hw, N_, K, taps, H, W,
                                         N.stream_ptr()), "wgrad_tc")
    else:
        N.check(N.load().mtgseg_wgrad(dz.data_ptr(), x.data_ptr(), dw.data_ptr(), N.ptr(a_scale), hw, M, N_, K, taps, H, W, N.stream_ptr()), "wgrad")
    xf = x.float()
    if se:
        xf = xf.view(-1, hw, K) * a_scale[:, None, :]
        xf = (xf if tc else xf.bfloat16().float()).reshape(M, K)
    if taps == 1:
        ref = (dz.float().t() @ xf)[:, :, None]
    else:
        B = M // (H * W)
        xw = xf.view(B, H, W, K).permute(0, 3, 1, 2).requires_grad_(False)
        wt = torch.zeros(N_, K, 3, 3, device="cuda", requires_grad=True)
        F.conv2d(xw, wt, None, 1, 1).backward(dz.float().view(B, H, W, N_).permute(0, 3, 1, 2))
        ref = wt.grad.reshape(N_, K, 9)
    _chk(f"wgrad M{M} N{N_} K{K} taps{taps}", dw, ref, 2e-3, 1e-3)


@pytest.mark.parametrize("B,H,W,C,k,stride,dil", [(2, 20, 15, 960, 5, 1, 2), (2, 80, 60, 72, 5, 2, 1), (2, 40, 30, 240, 3, 2, 1),
                                                  (3, 33, 21, 16, 3, 1, 1), (5, 40, 30, 120, 5, 1, 1), (37, 20, 15, 672, 5, 1, 1),
                                                  (9, 160, 120, 64, 3, 2, 1)])
def test_dw_bwd(B, H, W, C, k, stride, dil):
    g = torch.Generator().manual_seed(C + k)
    x = torch.randn(B, H, W, C, generator=g).bfloat16().cuda()
    w = (torch.randn(C, 1, k, k, generator=g) / k).bfloat16().cuda()
    wp = w.reshape(C, k * k).t().contiguous()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    out = F.conv2d(xr, wr, None, stride, (k - 1) // 2 * dil, dil, C)
    dz = torch.randn(out.shape, generator=g).bfloat16().cuda()
    out.backward(dz.float())
    dzp = dz.permute(0, 2, 3, 1).contiguous()
    dx = torch.empty_like(x); dw = torch.zeros(C, k * k).cuda()
    N.check(N.load().mtgseg_dw_bwd(dzp.data_ptr(), x.data_ptr(), wp.data_ptr(), dx.data_ptr(), dw.data_ptr(), B, H, W, C, k, stride, dil,
                                   N.stream_ptr()), "dw_bwd")
    _chk("dw dgrad", dx, xr.grad.permute(0, 2, 3, 1))
    _chk("dw wgrad", dw, wr.grad.reshape(C, k * k), 2e-3, 1e-3)


def test_stem_wgrad_and_upsample_bwd():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 3, 64, 48, generator=g).cuda()
    wr = torch.zeros(16, 3, 3, 3, device="cuda", requires_grad=True)
    out = F.conv2d(x, wr, None, 2, 1)
    dz = torch.randn(out.shape, generator=g).bfloat16().cuda()
    out.backward(dz.float())
    dw = torch.zeros(16, 27).cuda()
    dzp = dz.permute(0, 2, 3, 1).contiguous()
    N.check(N.load().mtgseg_stem_wgrad(x.data_ptr(), dzp.data_ptr(), dw.data_ptr(), 3, 64, 48, N.stream_ptr()), "stem_wgrad")
    _chk("stem wgrad", dw, wr.grad.reshape(16, 27), 2e-3, 1e-3)
    for (Hc, Wc, Hf, Wf) in [(40, 30, 320, 240), (5, 4, 9, 7), (20, 15, 40, 30)]:
        lo = torch.randn(2, 2, Hc, Wc, generator=g).cuda().requires_grad_(True)
        up = F.interpolate(lo, size=(Hf, Wf), mode="bilinear", align_corners=False)
        gup = torch.randn(up.shape, generator=g).cuda()
        up.backward(gup)
        outb = torch.empty(2, Hc, Wc, 2).cuda()
        N.check(N.load().mtgseg_upsample_bwd(gup.data_ptr(), 1, outb.data_ptr(), 2, 2, Hc, Wc, Hf, Wf, N.stream_ptr()), "up_bwd")
        _chk(f"upsample bwd {Hc}x{Wc}", outb.permute(0, 3, 1, 2), lo.grad, 1e-5, 1e-5)


def test_parity_on_short_trained_weights():
    """SURVEY.md §8c fixture C ("short-trained"): random-init weights are the worst case for bf16 storage noise (chaotic,
    unstructured).  Train the network for a few hundred steps on synthetic cards WITH THE CUDA STEP ITSELF, then compare
    CUDA inference against the fp32 oracle on those weights at config.py resolution: this is the regime the north-star
    tolerances (2e-2 relative logits, >= 99.9 % identical masks) are stated for."""
    torch.manual_seed(0)
    model = M.create_model(2, pretrained=False).cuda().train()
    model.inference_precision = "bf16"
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = M.CombinedLoss()
    first = last = None
    for step in range(240):
        x, m = O.synthetic_cards(16, seed=1000 + step % 24)
        xc, mc = x.cuda(), m.cuda()
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xc), mc)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    print(f"short training: loss {first:.4f} -> {last:.4f}")
    assert last < 0.5 * first
    # 240 steps at BatchNorm momentum 0.01 leave the running statistics far from converged (eval mode would predict one
    # class): re-estimate them on a calibration batch like a longer training run would have
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    xcal, _ = O.synthetic_cards(16, seed=77)
    sd = O.calibrate_running_stats(sd, xcal)
    model.load_state_dict(sd, strict=True)
    model.eval()
    x, m = O.synthetic_cards(4, seed=4242)
    with torch.no_grad():
        z = model(x.cuda()).cpu()
        ref = O.forward(sd, x)
        emu = O.forward_bf16_emulated(sd, x)
    fmax, fl2 = D.report("trained weights: CUDA bf16 vs fp32 oracle", z, ref)
    smax, sl2 = D.report("trained weights: bf16-emulated oracle vs fp32 oracle", emu, ref)
    agree = ((z[:, 1] > z[:, 0]) == (ref[:, 1] > ref[:, 0])).float().mean().item()
    iou = O.metrics_from_counts(O.confusion_counts(z, m))["iou"]
    print(f"trained weights: mask agreement (all pixels) {agree:.5f}; IoU vs ground truth {iou}")
    assert min(iou) > 0.5, "degenerate prediction: the fixture must actually segment the cards"
    assert fmax <= 2e-2 and fl2 <= 2e-2
    assert agree >= 0.999


def test_fp16_autocast_grad_scaler_step_matches_bf16_and_fp32_steps():
    """The reference's AMP dtype (train/train.py:96: torch.autocast('cuda') = fp16, GradScaler at 65536).  The unscaled loss
    gradient at B*H*W = 2.4 M pixels is ~2e-7 per logit -- an fp16 subnormal -- so the loss kernel hands fp16 logits an fp32
    gradient and the scale is applied before the narrowing cast.  The unscaled parameter gradients of the fp16 step must equal
    those of the bf16-autocast and no-autocast steps (same kernels, same bf16 arithmetic inside) up to the logits' rounding."""
    x, m = O.synthetic_cards(32, seed=21)
    sd = O.calibrate_running_stats(O.make_weights(22), x[:4])
    xc, mc = x.cuda(), m.cuda()
    crit = M.CombinedLoss(0.5, 0.5)
    grads = {}
    for mode in ("fp32", "bf16", "fp16"):
        model = _train_model(sd)
        scaler = torch.amp.GradScaler("cuda", enabled=mode == "fp16")
        ctx = torch.autocast("cuda", dtype=torch.bfloat16 if mode == "bf16" else torch.float16, enabled=mode != "fp32")
        with ctx:
            out = model(xc)
            loss = crit(out, mc)
        assert out.dtype == {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[mode]
        scaler.scale(loss).backward()
        inv = 1.0 / scaler.get_scale() if mode == "fp16" else 1.0
        grads[mode] = torch.cat([p.grad.flatten() for p in model.parameters()]).double().cpu() * inv
        assert torch.isfinite(grads[mode]).all()
    ref = grads["fp32"]
    for mode in ("bf16", "fp16"):
        g = grads[mode]
        cos = float((g @ ref) / (g.norm() * ref.norm()))
        ratio = float(g.norm() / ref.norm())
        print(f"{mode} autocast step vs no-autocast step: cosine {cos:.6f}, norm ratio {ratio:.4f}")
        # before the fix the fp16 step kept 2-3 bits of most logit gradients: cosine ~0.9, norm ratio ~0.8
        assert cos >= 0.999 and 0.99 <= ratio <= 1.01


def test_adamw_table_survives_load_state_dict():
    """Optimizer.load_state_dict replaces the moment tensors; a chunk table cached before the load must not be
    reused (the kernel would update freed memory and the loaded moments would never move)."""
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(40000, device="cuda"))
    q = torch.nn.Parameter(p.detach().clone())
    opt, ref = FusedAdamW([p], lr=1e-2, weight_decay=1e-4), torch.optim.AdamW([q], lr=1e-2, weight_decay=1e-4)
    gbuf, gref = torch.empty_like(p), torch.empty_like(q)
    p.grad, q.grad = gbuf, gref
    gen = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        gbuf.normal_(generator=gen); gref.copy_(gbuf)
        opt.step(); ref.step()
        if step == 2:  # same gradient address before and after: only the moment pointers change
            opt.load_state_dict(opt.state_dict())
            ref.load_state_dict(ref.state_dict())
    torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-7)
    # (the moments differ from torch's in the last bits: b2*v + (1-b2)*g*g here, mul_ / addcmul_ there)
    torch.testing.assert_close(opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(opt.state[p]["exp_avg"], ref.state[q]["exp_avg"], rtol=1e-4, atol=1e-7)
    assert float(opt.state[p]["step"]) == 6.0


def test_backward_of_a_stale_forward_raises():
    """ONE workspace holds the saved activations; backward of an older forward must fail loudly, not mix tensors."""
    x, m = O.synthetic_cards(2, seed=5, height=64, width=48)
    model = _train_model(O.make_weights(3))
    crit = M.CombinedLoss()
    xc, mc = x.cuda(), m.cuda()
    first = crit(model(xc), mc)
    second = crit(model(xc), mc)
    with pytest.raises(RuntimeError, match="stale forward"):
        first.backward()
    second.backward()  # the most recent forward is fine
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    third = crit(model(xc), mc)
    opt.step()         # weights (and the packed arena, at the next forward) change ...
    model.eval()
    with torch.no_grad():
        model(xc)      # ... an eval forward re-packs
    with pytest.raises(RuntimeError, match="stale forward"):
        third.backward()
