"""TEST INFRASTRUCTURE ONLY — records tests/golden/pose.pt from the unmodified reference pose head.
    python oracle/make_golden_pose.py   (build container only: needs /root/reference)"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pose_oracle as PO  # noqa: E402

sys.modules.setdefault("timm", types.ModuleType("timm"))  # model.py imports timm at the top; the backbone is never built
spec = importlib.util.spec_from_file_location("_ref_pose_model", "/root/reference/train-pose-estimation_custom/model.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.manual_seed(0)
cin, hf, wf = 64, 10, 8
head = ref.HRNetPoseHead(cin, 4, (32, 24)).eval()  # target (width, height) = (32, 24): pool 40x32 -> 24x32 (H shrinks, W stays)
sd = PO.make_weights(5, cin)
head.load_state_dict(sd, strict=True)
assert list(head.state_dict().keys()) == [k for k, _ in PO.state_dict_spec(cin)]
x = torch.randn(2, cin, hf, wf, generator=torch.Generator().manual_seed(6))
with torch.no_grad():
    hm = head(x)
coords = ref.LiteHRNet.decode_heatmaps(None, hm)
# second fixture: the reference geometry (40x30 feature -> 160x120 -> pooled to 120x160, aspect swapped), tiny channel count
head2 = ref.HRNetPoseHead(16, 4).eval()
sd2 = PO.make_weights(7, 16)
head2.load_state_dict(sd2, strict=True)
x2 = torch.randn(1, 16, 40, 30, generator=torch.Generator().manual_seed(8))
with torch.no_grad():
    hm2 = head2(x2)
# CornerMetrics / CornerLoss of the unmodified reference (metrics.py does `from model import LiteHRNet` inside update)
sys.path.insert(0, "/root/reference/train-pose-estimation_custom")
sys.modules["model"] = ref
mspec = importlib.util.spec_from_file_location("_ref_pose_metrics", "/root/reference/train-pose-estimation_custom/metrics.py")
refm = importlib.util.module_from_spec(mspec)
mspec.loader.exec_module(refm)
gm = torch.Generator().manual_seed(9)
yy, xx = torch.meshgrid(torch.arange(30.0), torch.arange(40.0), indexing="ij")


def blobs(cx, cy, sigma=1.5):  # (B,K) centres -> gaussian heatmaps (B,K,30,40)
    return torch.exp(-((xx[None, None] - cx[..., None, None]) ** 2 + (yy[None, None] - cy[..., None, None]) ** 2) / (2 * sigma * sigma))


tcx, tcy = torch.rand(6, 4, generator=gm) * 39, torch.rand(6, 4, generator=gm) * 29
target_hm = blobs(tcx.round(), tcy.round())
# predictions: targets displaced by 0..3 heatmap pixels (12 image pixels per heatmap pixel: below / between / above 3 and 6 px
# never happens by luck, so some keypoints are exact hits) + noise + exact ties on a plateau
dx = torch.randint(-2, 3, (6, 4), generator=gm).float(); dy = torch.randint(-2, 3, (6, 4), generator=gm).float()
pred_hm = blobs((tcx.round() + dx).clamp(0, 39), (tcy.round() + dy).clamp(0, 29)) + 0.01 * torch.randn(6, 4, 30, 40, generator=gm)
pred_hm[0, 0] = 0.0; pred_hm[0, 0, 7, 5] = 1.0; pred_hm[0, 0, 20, 30] = 1.0  # tie: the first maximum wins
# the fixture stores fp16 to stay small: the reference runs on exactly the stored values
pred_hm = pred_hm.half().float(); target_hm = target_hm.half().float()
metrics_cases = []
for image_size in ((480, 640), (39, 29), (100, 75)):  # (39, 29): one image pixel per heatmap pixel -> distances around the thresholds
    m = refm.CornerMetrics(image_size)
    m.update(pred_hm[:3], target_hm[:3])
    m.update(pred_hm[3:], target_hm[3:])
    metrics_cases.append({"image_size": image_size, "distances": [float(d) for d in m.all_distances], "compute": {k: float(v) for k, v in m.compute().items()}})
loss_val = refm.CornerLoss()(pred_hm, target_hm)
pg = pred_hm.clone().requires_grad_(True)
refm.CornerLoss()(pg, target_hm).backward()
torch.save({"small": {"weights_seed": 5, "cin": cin, "feat": (hf, wf), "target_hw": (24, 32), "input_seed": 6, "heatmaps": hm, "coords": coords},
            "ref_geometry": {"weights_seed": 7, "cin": 16, "feat": (40, 30), "target_hw": (120, 160), "input_seed": 8,
                             "heatmaps_sample": hm2.reshape(-1)[::37].clone(), "shape": tuple(hm2.shape),
                             "coords": ref.LiteHRNet.decode_heatmaps(None, hm2)},
            "metrics": {"pred": pred_hm.half(), "target": target_hm.half(), "cases": metrics_cases, "empty": refm.CornerMetrics().compute()},
            "loss": {"value": float(loss_val), "grad_sample": pg.grad.reshape(-1)[::53].clone()}},
           os.path.join(ROOT, "tests", "golden", "pose.pt"))
print("pose.pt", os.path.getsize(os.path.join(ROOT, "tests", "golden", "pose.pt")))
