"""fp32-exact inference path (csrc/f32net.cu) against the fp32 oracle: what train/evaluate.py:66 runs (float32, no autocast)
and what BASELINE.json's north_star holds to "1e-4 in fp32".  Tolerances written here: logits max |err| <= 1e-4 of the logit
range AND rel-L2 <= 1e-4; thresholded masks >= 99.9 % identical over ALL pixels (no margin band); counts bit-exact against the
oracle's argmax of the CUDA logits."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from conftest import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu

from oracle import lraspp_oracle as O  # noqa: E402
import mtg_card_image_segmentation_b200 as M  # noqa: E402
import devops as D  # noqa: E402


def _model(sd):
    m = M.create_model(2, pretrained=False)
    m.load_state_dict(sd, strict=True)
    assert m.inference_precision == "auto"
    return m.cuda().eval()


def _check(name, z, ref):
    emax, el2 = D.report(name, z, ref)
    agree = ((z[:, 1] > z[:, 0]) == (ref[:, 1] > ref[:, 0])).float().mean().item()
    print(f"[{name}] mask agreement over all pixels {agree:.6f}")
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)
    assert agree >= 0.999
    return emax


@pytest.mark.parametrize("golden,batch", [("seg_small.pt", 2), ("seg_full.pt", 1)])
def test_fp32_forward_vs_reference_golden(golden, batch):
    """Vectors recorded from the UNMODIFIED reference (oracle/make_golden.py) on its own fp32 CPU path; seg_full is 320x240.
    The 64x48 fixture stores the full eval logits; the 320x240 one a 4096-element sample, its checksum and the argmax mask."""
    g = load_golden(golden)
    sd = O.make_weights(g["weights_seed"], running_stats=g["running_stats"])
    x, m = O.synthetic_cards(g.get("calib_batch", g["batch"]), seed=g["input_seed"], height=g["height"], width=g["width"])
    x = x[:batch]
    model = _model(sd)
    with torch.no_grad():
        z = model(x.cuda())  # float32 batch, no autocast -> "auto" picks the fp32-exact path
        ref = O.forward(sd, x)
    assert z.dtype == torch.float32
    z = z.cpu()
    _check(f"fp32 path vs fp32 oracle on {golden}", z, ref)
    rng = ref.abs().max().item()
    if "eval_logits" in g:
        assert ((z - g["eval_logits"]).abs().max().item()) <= 1e-4 * rng
    else:
        assert (O.sample(z, 4096) - g["logits_sample"]).abs().max().item() <= 1e-4 * rng
        assert abs(z.double().sum().item() - float(g["logits_sum"])) <= 1e-4 * float(g["logits_abs_sum"])
        agree = ((z[:, 1] > z[:, 0]).to(torch.uint8) == g["mask_u8"]).float().mean().item()
        print(f"mask agreement vs the reference's recorded mask: {agree:.6f}")
        assert agree >= 0.999


def test_fp32_forward_320x240_batch_and_outputs():
    """Config resolution, a ragged batch (5), BN-calibrated deterministic weights: logits, mask and confusion counts of ONE call."""
    x, m = O.synthetic_cards(5, seed=11)
    sd = O.calibrate_running_stats(O.make_weights(7), x)
    model = _model(sd)
    with torch.no_grad():
        out = model.predict(x.cuda(), targets=m.cuda(), want_logits=True)
        ref = O.forward(sd, x)
    z = out["logits"].cpu()
    _check("fp32 path 320x240 B=5 vs fp32 oracle", z, ref)
    assert torch.equal(out["mask"].cpu().bool(), z[:, 1] > z[:, 0])
    assert torch.equal(out["counts"].cpu(), O.confusion_counts(z, m))
    # batch invariance: every image alone gives the same bits (fixed summation order per output element)
    with torch.no_grad():
        single = torch.cat([model(x[i:i + 1].cuda()) for i in range(5)]).cpu()
    assert torch.equal(single, z)


def test_precision_rule_follows_autocast():
    """evaluate.py:66 (no autocast) -> fp32-exact; train.py:142 (autocast) -> tensor-core bf16 path; explicit override wins."""
    x, _ = O.synthetic_cards(2, seed=5, height=96, width=64)
    sd = O.calibrate_running_stats(O.make_weights(3), x)
    model = _model(sd)
    xc = x.cuda()
    with torch.no_grad():
        ref = O.forward(sd, x)
        z32 = model(xc).cpu()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            zb = model(xc)
        model.inference_precision = "bf16"
        zforced = model(xc).cpu()
        model.inference_precision = "fp32"
        with torch.autocast("cuda", dtype=torch.bfloat16):
            z32b = model(xc)
    assert zb.dtype == torch.bfloat16 and z32b.dtype == torch.bfloat16
    rng = ref.abs().max()
    e32 = ((z32 - ref).abs().max() / rng).item()
    eb = ((zforced - ref).abs().max() / rng).item()
    print(f"auto/no-autocast err {e32:.2e}; forced bf16 err {eb:.2e}")
    assert e32 <= 1e-4 < eb
    # bf16-typed output of the fp32 path = the fp32 logits rounded once
    assert torch.equal(z32b.cpu(), z32.to(torch.bfloat16))
    with pytest.raises(RuntimeError):
        model.predict(torch.zeros(1, 96, 64, 3, dtype=torch.uint8, device="cuda"), precision="fp32")
