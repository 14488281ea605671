"""Host-side drop-in surface (no GPU): module tree / state_dict layout, checkpoint round trip, loud failure
without CUDA, the exporter-only composite forward, repo layout rules."""
import os
import re

import pytest
import torch

from conftest import ROOT, has_reference
from oracle import lraspp_oracle as O

import mtg_card_image_segmentation_b200 as M


def test_state_dict_layout_matches_reference_contract():
    m = M.create_model(num_classes=2, pretrained=False)
    sd = m.state_dict()
    spec = O.state_dict_spec()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for (k, shape, _), v in zip(spec, sd.values()):
        assert tuple(v.shape) == tuple(shape), k
    assert M.count_parameters(m) == (4_201_348, 4_201_348)
    assert abs(M.get_model_size(m) - 16.12) < 0.01
    assert len(list(m.parameters())) == 178  # AdamW param order contract
    # module tree the reference's tools walk (train/prune.py:52-58 iterates nn.Conv2d children)
    assert isinstance(m.model.backbone["0"][0], torch.nn.Conv2d)
    assert isinstance(m.model.classifier.cbr[0], torch.nn.Conv2d) and m.model.classifier.cbr[0].kernel_size == (3, 3)
    assert m.model.backbone["4"].block[2].fc1.bias is not None
    assert m.model.classifier.cbr[1].eps == 1e-5 and m.model.backbone["3"].block[0][1].eps == 1e-3
    assert m.model.backbone["3"].block[0][1].momentum == 0.01 and m.model.classifier.cbr[1].momentum == 0.1


def test_constructor_contract():
    with pytest.raises(RuntimeError, match="pretrained"):
        M.create_model(2, pretrained=True)
    m = M.CardSegmentationModel(num_classes=3, pretrained=False)
    assert m.model.classifier.low_classifier.out_channels == 3


def test_cpu_forward_fails_loudly():
    m = M.create_model(2, False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(torch.zeros(1, 3, 64, 48))
    with pytest.raises(RuntimeError, match="CUDA"):
        M.CombinedLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA"):
        M.calculate_iou(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))


def test_checkpoint_round_trip(tmp_path):
    m = M.create_model(2, False)
    m.load_state_dict(O.make_weights(4), strict=True)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=100, eta_min=1e-5)
    M.save_checkpoint(m, opt, sched, 7, 0.875, str(tmp_path), "ck.pth")
    ck = torch.load(tmp_path / "ck.pth", map_location="cpu", weights_only=True)  # tensors/primitives only
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_metric"}
    assert list(ck["model_state_dict"].keys()) == [k for k, _, _ in O.state_dict_spec()]
    m2 = M.create_model(2, False)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    epoch, best = M.load_checkpoint(m2, opt2, None, str(tmp_path / "ck.pth"))
    assert (epoch, best) == (7, 0.875)
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)


@pytest.mark.skipif(not has_reference(), reason="/root/reference only exists in the build container")
def test_checkpoint_interchange_with_reference(tmp_path):
    """A checkpoint written by our save_checkpoint loads into the reference model and vice versa."""
    import functools
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_model2", "/root/reference/train/model.py")
    ref_model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_model)
    ref_model.lraspp_mobilenet_v3_large = functools.partial(ref_model.lraspp_mobilenet_v3_large, weights_backbone=None)
    ref = ref_model.create_model(2, pretrained=False)
    ours = M.create_model(2, False)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    o1 = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    o2 = torch.optim.AdamW(ours.parameters(), lr=1e-3)
    o2.load_state_dict(o1.state_dict())


def test_export_composite_forward_matches_oracle():
    """torch.jit.trace (what train/export.py:177-182 does) goes through the ATen composite and reproduces the
    oracle; the trace contains the 66 convolutions of the reference's exported graph."""
    m = M.create_model(2, False)
    x, _ = O.synthetic_cards(1, seed=2, height=64, width=48)
    sd = O.calibrate_running_stats(O.make_weights(8), x)
    m.load_state_dict(sd, strict=True)
    m.eval()
    traced = torch.jit.trace(m, x, check_trace=False)
    with torch.no_grad():
        torch.testing.assert_close(traced(x), O.forward(sd, x), rtol=1e-4, atol=1e-5)
    assert str(traced.inlined_graph).count("aten::_convolution") == 66


def _onnx_graph(model, x):
    """The opset-11 ONNX graph torch.onnx.export(model, x, opset_version=11, do_constant_folding=True, input_names=['input'],
    output_names=['output']) builds (train/export.py:68-79), taken from the TorchScript exporter right before serialisation
    (writing the file needs the `onnx` wheel, absent here; building the graph does not)."""
    from torch.onnx._internal.torchscript_exporter import utils as TU
    from torch.onnx._internal.torchscript_exporter._globals import GLOBALS
    GLOBALS.export_onnx_opset_version = 11
    with TU.exporter_context(model, torch.onnx.TrainingMode.EVAL, False):
        graph, params, _ = TU._model_to_graph(model, (x,), do_constant_folding=True, input_names=["input"], output_names=["output"])
    hist = {}
    for n in graph.nodes():
        hist[n.kind()] = hist.get(n.kind(), 0) + 1
    return graph, params, hist


def test_onnx_export_graph_histogram():
    """a15 / f2: the exported graph is the reference's -- 66 Conv (BatchNorm folded), 28 HardSigmoid, 29 Mul, 20 Relu, 11 Add,
    9 GlobalAveragePool, 2 Resize, 1 Sigmoid, 131 initialisers, tensors named 'input' / 'output' (SURVEY.md §3E)."""
    import warnings
    m = M.create_model(2, False).eval()
    x = torch.randn(1, 3, 320, 240)  # export.py:62-66: dummy input at Config resolution
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        graph, params, hist = _onnx_graph(m, x)
    want = {"onnx::Conv": 66, "onnx::HardSigmoid": 28, "onnx::Mul": 29, "onnx::Relu": 20, "onnx::Add": 11,
            "onnx::GlobalAveragePool": 9, "onnx::Resize": 2, "onnx::Sigmoid": 1}
    assert {k: hist.get(k, 0) for k in want} == want, hist
    assert "onnx::BatchNormalization" not in hist
    assert len(params) == 131
    assert [i.debugName() for i in graph.inputs()][0] == "input" and [o.debugName() for o in graph.outputs()] == ["output"]
    # initialisers carry the reference's parameter names (what the demo's ORT session and onnx_fp16_converter.py see)
    assert "model.classifier.cbr.0.weight" not in params  # folded with its BatchNorm into an anonymous Conv weight
    assert "model.classifier.low_classifier.weight" in params and "model.backbone.4.block.2.fc1.bias" in params
    resize = [n for n in graph.nodes() if n.kind() == "onnx::Resize"]
    for n in resize:  # opset-11 Resize: 'half_pixel' is the default coordinate_transformation_mode and is omitted when it applies
        assert n.s("mode") == "linear"
        assert "coordinate_transformation_mode" not in n.attributeNames() or n.s("coordinate_transformation_mode") == "half_pixel"
    if has_reference():  # node for node the same op sequence as the unmodified reference's export
        from oracle import ref_loader as R
        ref = R.load_reference(("model",))["model"].create_model(2, pretrained=False).eval()
        R.unload()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rgraph, rparams, rhist = _onnx_graph(ref, x)
        assert rhist == hist and len(rparams) == len(params)
        assert [n.kind() for n in rgraph.nodes()] == [n.kind() for n in graph.nodes()]


def test_repo_layout_rules():
    """Only tests/, bench.py and __graft_entry__.py may touch oracle/; the product never does."""
    pkg = os.path.join(ROOT, "mtg_card_image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src, f
    for f in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(ROOT, f)).read()


def test_arch_table_in_sync_with_cuda_plan():
    from mtg_card_image_segmentation_b200 import arch
    src = open(os.path.join(ROOT, "mtg_card_image_segmentation_b200", "csrc", "net.cu")).read()
    rows = re.findall(r"\{(\d+), (\d+), (\d+), (\d+), (true|false), ACT_(RELU|HSWISH),\s*(\d+), (\d+)\}", src)
    assert len(rows) == 15
    for r, b in zip(rows, arch.BLOCKS):
        assert (int(r[0]), int(r[1]), int(r[2]), int(r[3]), r[4] == "true", {"RELU": "RE", "HSWISH": "HS"}[r[5]],
                int(r[6]), int(r[7])) == tuple(b)
    assert [tuple(b) for b in arch.BLOCKS] == [tuple(b) for b in O.BLOCKS]


def test_graphed_train_step_rejects_what_it_cannot_capture():
    """engine.GraphedTrainStep fails loudly (before touching CUDA) for setups whose step it would not reproduce: a foreign
    optimizer (its scalar arguments would be frozen into the graph), a foreign criterion, eval mode, CPU tensors."""
    from mtg_card_image_segmentation_b200.engine import GraphedTrainStep
    from mtg_card_image_segmentation_b200.optim import FusedAdamW
    model = M.create_model(2, pretrained=False).train()
    x, y = torch.zeros(2, 3, 64, 48), torch.zeros(2, 64, 48, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="FusedAdamW"):
        GraphedTrainStep(model, M.CombinedLoss(), torch.optim.AdamW(model.parameters()), x, y)
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    with pytest.raises(RuntimeError, match="CUDA"):
        GraphedTrainStep(model, M.CombinedLoss(), opt, x, y)
    with pytest.raises(RuntimeError, match="model.train"):
        GraphedTrainStep(model.eval(), M.CombinedLoss(), opt, x, y)
    two = FusedAdamW([{"params": list(model.parameters())[:10]}, {"params": list(model.parameters())[10:]}], lr=1e-3)
    with pytest.raises(RuntimeError, match="one parameter group"):
        GraphedTrainStep(model.train(), M.CombinedLoss(), two, x, y)


def test_state_slots_of_a_model_pruned_before_its_first_forward():
    """train/prune.py prunes a freshly loaded model.  torch's pruning moves `weight` out of `module._parameters` (it becomes
    `weight_orig` + `weight_mask` and a plain attribute recomputed by a hook): the 319 reference slots and the 178 parameter slots
    must be recognised all the same, whether the slot table is built before or after the surgery, for unstructured and for
    structured (`ln_structured`, train/prune.py:76-93) pruning, and `prune.remove` must restore the reference layout."""
    import torch.nn.utils.prune as prune
    for structured in (False, True):
        model = M.create_model(2, pretrained=False)
        ref_keys = list(model.state_dict().keys())
        convs = [m for m in model.model.modules() if isinstance(m, torch.nn.Conv2d)]
        if structured:
            for m in convs:
                n = int(m.out_channels * 0.3)
                if m.out_channels > 1 and n > 0:
                    prune.ln_structured(m, name="weight", amount=n, n=2, dim=0)
        else:
            prune.global_unstructured([(m, "weight") for m in convs], pruning_method=prune.L1Unstructured, amount=0.3)
        tensors = model._state_tensors()  # first call AFTER the surgery
        assert len(tensors) == 319 and len(model._param_slots) == 178
        assert sum(t.requires_grad for t in tensors) == 178
        masked = [m for m in convs if hasattr(m, "weight_mask")]
        assert masked and all(bool((m.weight[m.weight_mask == 0] == 0).all()) for m in masked)
        for m in masked:
            prune.remove(m, "weight")
        assert sorted(model.state_dict().keys()) == sorted(ref_keys)  # (prune.remove re-registers `weight` after `bias`: same keys, new order)
        assert sum(t.requires_grad for t in model._state_tensors()) == 178
