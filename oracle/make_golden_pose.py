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
torch.save({"small": {"weights_seed": 5, "cin": cin, "feat": (hf, wf), "target_hw": (24, 32), "input_seed": 6, "heatmaps": hm, "coords": coords},
            "ref_geometry": {"weights_seed": 7, "cin": 16, "feat": (40, 30), "target_hw": (120, 160), "input_seed": 8,
                             "heatmaps_sample": hm2.reshape(-1)[::37].clone(), "shape": tuple(hm2.shape),
                             "coords": ref.LiteHRNet.decode_heatmaps(None, hm2)}},
           os.path.join(ROOT, "tests", "golden", "pose.pt"))
print("pose.pt", os.path.getsize(os.path.join(ROOT, "tests", "golden", "pose.pt")))
