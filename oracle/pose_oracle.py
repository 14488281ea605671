"""TEST INFRASTRUCTURE ONLY — CPU oracle of the pose head (train-pose-estimation_custom/model.py:10-77,133-164).

Plain torch.nn.functional fp32 restatement driven by the 28-entry ``HRNetPoseHead`` state_dict; pinned by
tests/golden/pose.pt, which ``oracle/make_golden_pose.py`` records from the unmodified reference class (imported
with a ``timm`` stub: the backbone is never built).  ``forward_bf16_emulated`` applies the roundings of the CUDA path
(bf16 weights and layer outputs, fp32 accumulation)."""
import torch
import torch.nn.functional as F


def state_dict_spec(in_channels, num_keypoints=4):
    spec = []

    def bn(p):
        spec.extend([(p + ".weight", (256,)), (p + ".bias", (256,)), (p + ".running_mean", (256,)), (p + ".running_var", (256,)),
                     (p + ".num_batches_tracked", ())])
    spec.append(("deconv_layers.0.0.weight", (in_channels, 256, 4, 4))); bn("deconv_layers.0.1")
    spec.append(("deconv_layers.1.0.weight", (256, 256, 4, 4))); bn("deconv_layers.1.1")
    spec.extend([("conv_layers.0.weight", (256, 256, 3, 3)), ("conv_layers.0.bias", (256,))]); bn("conv_layers.1")
    spec.extend([("conv_layers.3.weight", (256, 256, 3, 3)), ("conv_layers.3.bias", (256,))]); bn("conv_layers.4")
    spec.extend([("final_layer.weight", (num_keypoints, 256, 1, 1)), ("final_layer.bias", (num_keypoints,))])
    return spec


def make_weights(seed, in_channels, num_keypoints=4):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in state_dict_spec(in_channels, num_keypoints):
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("running_mean") or k.endswith(".bias"):
            sd[k] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1:
            sd[k] = torch.rand(shape, generator=g) + 0.5
        else:
            fan = shape[1] * shape[2] * shape[3] if "conv_layers" in k or "final" in k else shape[0] * 4  # deconv: 4 taps hit a pixel
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan) ** 0.5
    return sd


def forward(sd, x, target_hw=(120, 160), q=None, wq=None):
    q = q or (lambda t: t)
    wq = wq or (lambda t: t)

    def bn(t, p):
        return F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)

    t = q(x)
    for i in range(2):
        t = q(F.relu(bn(F.conv_transpose2d(t, wq(sd[f"deconv_layers.{i}.0.weight"]), None, 2, 1), f"deconv_layers.{i}.1")))
    for c, b in ((0, 1), (3, 4)):
        t = q(F.relu(bn(F.conv2d(t, wq(sd[f"conv_layers.{c}.weight"]), sd[f"conv_layers.{c}.bias"], 1, 1), f"conv_layers.{b}")))
    t = q(F.conv2d(t, wq(sd["final_layer.weight"]), sd["final_layer.bias"]))
    return F.adaptive_avg_pool2d(t, target_hw)


def forward_bf16_emulated(sd, x, target_hw=(120, 160)):
    r = lambda t: t.bfloat16().float()
    return forward(sd, x, target_hw, q=r, wq=r)


def decode_heatmaps(hm):
    """model.py:133-164 (first maximum on ties, as torch.max on the CPU)."""
    B, K, H, W = hm.shape
    idx = hm.reshape(B, K, -1).max(dim=2).indices
    coords = torch.zeros(B, K * 2)
    coords[:, 0::2] = (idx % W).float() / (W - 1)
    coords[:, 1::2] = (idx // W).float() / (H - 1)
    return coords


def corner_loss(pred, target):
    """CornerLoss.forward (metrics.py:125-136): nn.MSELoss() on the heatmaps."""
    return ((pred.float() - target.float()) ** 2).mean()


def corner_distances(pred, target, image_size=(480, 640)):
    """CornerMetrics.update (metrics.py:29-73): per (image, keypoint) pixel distance between the argmax of the predicted and of the
    target heatmap; float32 arithmetic in the reference's order (torch float32 coordinates, numpy float32 distance)."""
    import numpy as np
    B, K, H, W = pred.shape
    pi = pred.reshape(B, K, -1).max(dim=2).indices
    ti = target.reshape(B, K, -1).max(dim=2).indices
    px = ((pi % W).float() * image_size[0] / (W - 1)).numpy(); py = ((pi // W).float() * image_size[1] / (H - 1)).numpy()
    tx = ((ti % W).float() * image_size[0] / (W - 1)).numpy(); ty = ((ti // W).float() * image_size[1] / (H - 1)).numpy()
    out = []
    for i in range(B):
        for j in range(K):
            out.append(np.sqrt((px[i, j] - tx[i, j]) ** 2 + (py[i, j] - ty[i, j]) ** 2))
    return out


def corner_compute(distances):
    """CornerMetrics.compute (metrics.py:75-100)."""
    import numpy as np
    if not distances:
        return {"corner_acc_3px": 0.0, "corner_acc_6px": 0.0, "mean_corner_distance": 0.0}
    d = np.array(distances)
    return {"corner_acc_3px": np.mean(d <= 3.0) * 100, "corner_acc_6px": np.mean(d <= 6.0) * 100, "mean_corner_distance": np.mean(d)}
