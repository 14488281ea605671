"""Pose head (BASELINE.json configs[4]): oracle pinned to the reference's HRNetPoseHead /
decode_heatmaps (CPU), CUDA head vs oracle (GPU)."""
import pytest
import torch

from conftest import load_golden
from oracle import pose_oracle as PO


def _inputs(g):
    sd = PO.make_weights(g["weights_seed"], g["cin"])
    x = torch.randn((2 if g["cin"] == 64 else 1), g["cin"], *g["feat"], generator=torch.Generator().manual_seed(g["input_seed"]))
    return sd, x


def test_pose_oracle_vs_reference_golden():
    g = load_golden("pose.pt")
    sd, x = _inputs(g["small"])
    with torch.no_grad():
        hm = PO.forward(sd, x, g["small"]["target_hw"])
    torch.testing.assert_close(hm, g["small"]["heatmaps"], rtol=1e-4, atol=1e-5)
    assert torch.equal(PO.decode_heatmaps(hm), g["small"]["coords"])
    sd, x = _inputs(g["ref_geometry"])
    with torch.no_grad():
        hm = PO.forward(sd, x, g["ref_geometry"]["target_hw"])
    assert tuple(hm.shape) == g["ref_geometry"]["shape"] == (1, 4, 120, 160)
    torch.testing.assert_close(hm.reshape(-1)[::37], g["ref_geometry"]["heatmaps_sample"], rtol=1e-4, atol=1e-5)
    assert torch.equal(PO.decode_heatmaps(hm), g["ref_geometry"]["coords"])


def test_pose_module_state_dict_layout():
    from mtg_card_image_segmentation_b200.pose import HRNetPoseHead
    m = HRNetPoseHead(64)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in PO.state_dict_spec(64)]
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(torch.zeros(1, 64, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,batch", [("small", 2), ("ref_geometry", 1)])
def test_pose_head_cuda_vs_oracle(fixture, batch):
    import devops as D
    from mtg_card_image_segmentation_b200.pose import HRNetPoseHead, decode_heatmaps
    g = load_golden("pose.pt")[fixture]
    sd, x = _inputs(g)
    head = HRNetPoseHead(g["cin"], 4, (g["target_hw"][1], g["target_hw"][0]))
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    with torch.no_grad():
        hm, coords = head(x.cuda(), return_coords=True)
        emu = PO.forward_bf16_emulated(sd, x, g["target_hw"])
        ref = PO.forward(sd, x, g["target_hw"])
    emax, el2 = D.report(f"pose {fixture} vs bf16-emulated oracle", hm.cpu(), emu)
    assert emax <= 1.5e-2 and el2 <= 1.5e-2
    fmax, fl2 = D.report(f"pose {fixture} vs fp32 oracle", hm.cpu(), ref)
    assert fmax <= 3e-2 and fl2 <= 3e-2
    # decode: integer argmax is exact on identical heatmaps
    assert torch.equal(coords.cpu(), PO.decode_heatmaps(hm.cpu()))
    assert torch.equal(decode_heatmaps(ref.cuda()).cpu(), PO.decode_heatmaps(ref))
    t = torch.zeros(2, 4, 12, 16); t[0, 1, 3, 5] = 1; t[0, 1, 7, 2] = 1  # tie -> first (lowest index) maximum
    assert torch.equal(decode_heatmaps(t.cuda()).cpu(), PO.decode_heatmaps(t))


@pytest.mark.gpu
def test_pose_head_full_size():
    """configs[4]: 512-channel 40x30 feature (a 640x480 input) -> (B,4,120,160) heatmaps + (B,8) coords."""
    import devops as D
    from mtg_card_image_segmentation_b200.pose import HRNetPoseHead
    sd = PO.make_weights(3, 512)
    x = torch.randn(2, 512, 40, 30, generator=torch.Generator().manual_seed(4))
    head = HRNetPoseHead(512)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().eval()
    with torch.no_grad():
        hm, coords = head(x.cuda(), return_coords=True)
        emu = PO.forward_bf16_emulated(sd, x)
    assert tuple(hm.shape) == (2, 4, 120, 160) and tuple(coords.shape) == (2, 8)
    emax, el2 = D.report("pose full vs bf16-emulated oracle", hm.cpu(), emu)
    assert emax <= 1.5e-2 and el2 <= 1.5e-2


def test_corner_metrics_and_loss_oracle_vs_reference_golden():
    """CornerMetrics.update/compute and CornerLoss of the unmodified reference (recorded by oracle/make_golden_pose.py) vs the
    oracle restatement: every distance bit-identical (float32 arithmetic in the reference's order), compute() identical."""
    g = load_golden("pose.pt")
    p, t = g["metrics"]["pred"].float(), g["metrics"]["target"].float()
    for case in g["metrics"]["cases"]:
        d = PO.corner_distances(p[:3], t[:3], case["image_size"]) + PO.corner_distances(p[3:], t[3:], case["image_size"])
        assert [float(v) for v in d] == case["distances"]
        got = PO.corner_compute(d)
        assert {k: float(v) for k, v in got.items()} == case["compute"]
    assert PO.corner_compute([]) == g["metrics"]["empty"]
    assert float(PO.corner_loss(p, t)) == pytest.approx(g["loss"]["value"], rel=1e-6)


@pytest.mark.gpu
def test_corner_metrics_and_loss_cuda_vs_golden():
    """The CUDA CornerMetrics / CornerLoss against the reference's recorded results: threshold counts exact (integer work),
    mean distance and loss to 1e-6 relative, gradient of the loss to 1e-6."""
    from mtg_card_image_segmentation_b200.pose import CornerLoss, CornerMetrics
    g = load_golden("pose.pt")
    p, t = g["metrics"]["pred"].float().cuda(), g["metrics"]["target"].float().cuda()
    for case in g["metrics"]["cases"]:
        m = CornerMetrics(case["image_size"])
        assert m.compute() == g["metrics"]["empty"]
        m.update(p[:3], t[:3])
        m.update(p[3:], t[3:])
        got, want = m.compute(), case["compute"]
        assert got["corner_acc_3px"] == want["corner_acc_3px"] and got["corner_acc_6px"] == want["corner_acc_6px"]
        assert got["mean_corner_distance"] == pytest.approx(want["mean_corner_distance"], rel=1e-6)
        m.reset()
        m.update(p[:1], t[:1])
        d = PO.corner_distances(p[:1].cpu(), t[:1].cpu(), case["image_size"])
        assert m.compute()["mean_corner_distance"] == pytest.approx(float(sum(float(v) for v in d) / len(d)), rel=1e-6)
    pg = p.clone().requires_grad_(True)
    loss = CornerLoss()(pg, t)
    assert float(loss.detach()) == pytest.approx(g["loss"]["value"], rel=1e-6)
    (3.0 * loss).backward()
    want = g["loss"]["grad_sample"].cuda() * 3.0
    assert torch.allclose(pg.grad.reshape(-1)[::53], want, rtol=1e-6, atol=1e-12)
    # run-to-run determinism of the two-stage reduction
    assert float(CornerLoss()(p, t)) == float(CornerLoss()(p, t))
