"""Pins the oracle (oracle/lraspp_oracle.py) against vectors recorded from the unmodified
reference + torchvision (oracle/make_golden.py), and against the live reference when present."""
import functools
import sys
import types

import pytest
import torch

from conftest import has_reference, load_golden
from oracle import lraspp_oracle as O


def _weights(g):
    return O.make_weights(g["weights_seed"], running_stats=g["running_stats"])


def test_state_dict_layout():
    spec = O.state_dict_spec()
    assert len(spec) == 319
    sd = O.make_weights(0)
    n_params = sum(v.numel() for (k, _, kind), v in zip(spec, sd.values()) if kind in ("conv", "gamma", "beta", "bias"))
    assert n_params == 4_201_348  # SURVEY.md finding 2
    assert spec[0][0] == "model.backbone.0.0.weight" and spec[-1][0] == "model.classifier.high_classifier.bias"


def test_eval_forward_small():
    g = load_golden("seg_small.pt")
    x, m = O.synthetic_cards(g["batch"], seed=g["input_seed"], height=g["height"], width=g["width"])
    taps = {}
    with torch.no_grad():
        y = O.forward(_weights(g), x, taps=taps)
    torch.testing.assert_close(y, g["eval_logits"], rtol=1e-4, atol=1e-5)
    for k, ref in g["eval_acts"].items():
        torch.testing.assert_close(taps[k], ref, rtol=1e-4, atol=1e-5)


def test_train_forward_backward_small():
    g = load_golden("seg_small.pt")
    x, m = O.synthetic_cards(g["batch"], seed=g["input_seed"], height=g["height"], width=g["width"])
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
          for k, v in _weights(g).items()}
    upd = {}
    y = O.forward(sd, x, training=True, bn_updates=upd)
    torch.testing.assert_close(y.detach(), g["train_logits"], rtol=1e-4, atol=1e-5)
    loss = O.combined_loss(y, m)
    torch.testing.assert_close(loss.detach(), g["train_loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.combined_loss_closed_form(y.detach(), m).float(), g["train_loss"], rtol=1e-5, atol=1e-6)
    for k, ref in g["bn_after"].items():
        torch.testing.assert_close(upd[k], ref, rtol=1e-4, atol=1e-6)
    loss.backward()
    for k, ref in g["grads"].items():
        got = sd[k].grad
        got = got if got.numel() <= 8192 else O.sample(got, 4096)
        scale = float(g["grad_norms"][k]) / max(1.0, sd[k].numel() ** 0.5)
        torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3 * scale + 1e-9)
    for k, ref in g["grad_norms"].items():
        torch.testing.assert_close(sd[k].grad.norm(), ref, rtol=2e-3, atol=1e-9)


def test_eval_forward_full_resolution():
    g = load_golden("seg_full.pt")
    x, m = O.synthetic_cards(g["calib_batch"], seed=g["input_seed"])
    taps = {}
    with torch.no_grad():
        y = O.forward(_weights(g), x[:1], taps=taps)
    torch.testing.assert_close(taps["lowres_logits"], g["lowres_logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(O.sample(y, 4096), g["logits_sample"], rtol=1e-4, atol=1e-5)
    assert abs(float(y.double().sum() - g["logits_sum"])) < 1e-4 * float(g["logits_abs_sum"])
    pred = torch.argmax(y, 1).to(torch.uint8)
    assert (pred == g["mask_u8"]).float().mean().item() >= 0.9999
    torch.testing.assert_close(O.combined_loss(y, m[:1]), g["loss"], rtol=1e-5, atol=1e-6)
    cm = O.confusion_counts(y, m[:1]).reshape(2, 2)
    assert int((cm - g["confusion"]).abs().sum()) <= 4  # argmax flips only on ~1e-7 logit margins


def test_metrics_and_counts():
    g = load_golden("metrics.pt")
    gen = torch.Generator().manual_seed(5)
    for case in g["cases"]:
        b, hh, ww = case["seed_shape"]
        z = torch.randn(b, 2, hh, ww, generator=gen)
        z[:, 1, ::3, ::2] = z[:, 0, ::3, ::2]
        t = torch.randint(0, 2, (b, hh, ww), generator=gen)
        if case["logits"] is not None:
            assert torch.equal(z, case["logits"]) and torch.equal(t, case["targets"])
        c = O.confusion_counts(z, t)
        assert torch.equal(c, case["counts"])
        assert int(c.sum()) == b * hh * ww
        mm = O.metrics_from_counts(c)
        torch.testing.assert_close(torch.tensor(mm["iou"]), case["iou"], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(torch.tensor(mm["dice"]), case["dice"], rtol=1e-6, atol=1e-7)
        assert abs(mm["acc"] - float(case["acc"])) < 1e-6
        torch.testing.assert_close(O.combined_loss(z, t), case["loss"], rtol=1e-5, atol=1e-6)
    pcm = O.per_class_metrics(g["cm"])
    for name, ref in g["per_class"].items():
        for k, v in ref.items():
            assert abs(pcm[name][k] - v) <= 1e-12 * max(1.0, abs(v))


def test_adamw_trajectory():
    g = load_golden("adamw.pt")
    p = g["p0"].clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step, (grad, ref) in enumerate(zip(g["g"], g["p"]), start=1):
        p, m, v = O.adamw_step(p, grad, m, v, step)
        torch.testing.assert_close(p, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not has_reference(), reason="/root/reference only exists in the build container")
def test_live_reference_matches_oracle():
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn", "albumentations", "albumentations.pytorch"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["albumentations.pytorch"].ToTensorV2 = object
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_model", "/root/reference/train/model.py")
    ref_model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_model)
    ref_model.lraspp_mobilenet_v3_large = functools.partial(ref_model.lraspp_mobilenet_v3_large, weights_backbone=None)
    torch.manual_seed(1)
    model = ref_model.create_model(num_classes=2, pretrained=False).eval()
    sd = model.state_dict()  # the reference's own random init (incl. ctor side effects)
    assert [k for k, _, _ in O.state_dict_spec()] == list(sd.keys())
    for (k, shape, _), v in zip(O.state_dict_spec(), sd.values()):
        assert tuple(v.shape) == tuple(shape), k
    x, _ = O.synthetic_cards(1, seed=99, height=96, width=64)
    with torch.no_grad():
        torch.testing.assert_close(O.forward(sd, x), model(x), rtol=1e-4, atol=1e-5)
