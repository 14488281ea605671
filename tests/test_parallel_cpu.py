"""world_size-2 gloo tests (CPU) of the multi-process host logic: batch sharding, the flat-gradient average that the
data-parallel training step performs, global confusion-count merge, and bench.py's rank discipline."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtg_card_image_segmentation_b200 import parallel as P
    g = torch.Generator().manual_seed(rank)
    flat = torch.randn(4_201_348 // 64, generator=g)
    mine = flat.clone()
    views = [flat[:1000].view(10, 100), flat[1000:]]           # per-parameter views alias the bucket
    P.average_gradients(flat)
    others = [torch.randn(4_201_348 // 64, generator=torch.Generator().manual_seed(r)) for r in range(world)]
    want = sum(others) / world
    ok = torch.allclose(flat, want, atol=1e-6) and torch.allclose(views[0].reshape(-1), want[:1000], atol=1e-6)
    counts = torch.tensor([10 + rank, 1, 2, 3 * rank], dtype=torch.int64)
    P.merge_counts(counts)
    ok = ok and counts.tolist() == [10 * world + sum(range(world)), world, 2 * world, 3 * sum(range(world))]
    sl = P.shard_batch(257, rank, world)
    # gradients that do NOT alias the flat buffer (zero_grad(set_to_none=False), accumulation, cloning hooks) are still averaged
    class _Holder:
        def __init__(self, r):
            g = torch.Generator().manual_seed(100 + r)
            self.last_flat_grad = torch.randn(64, generator=g)
            self.p = [torch.nn.Parameter(torch.zeros(40)), torch.nn.Parameter(torch.zeros(24))]
            self.p[0].grad = self.last_flat_grad[:40].clone()    # a clone: not a view of the flat buffer
            self.p[1].grad = self.last_flat_grad[40:]            # a view

        def parameters(self):
            return self.p
    h = _Holder(rank)
    P.average_gradients(h)
    want_h = sum(_Holder(r).last_flat_grad for r in range(world)) / world
    ok = ok and torch.allclose(h.p[0].grad, want_h[:40], atol=1e-6) and torch.allclose(h.p[1].grad, want_h[40:], atol=1e-6)
    # a captured training step in a multi-process job must contain the gradient exchange: without it, it refuses
    import mtg_card_image_segmentation_b200 as M
    from mtg_card_image_segmentation_b200.engine import GraphedTrainStep
    from mtg_card_image_segmentation_b200.optim import FusedAdamW
    model = M.create_model(2, pretrained=False).train()
    try:
        GraphedTrainStep(model, M.CombinedLoss(), FusedAdamW(model.parameters()), torch.zeros(1, 3, 64, 48),
                         torch.zeros(1, 64, 48, dtype=torch.int64))
        ok = False
    except RuntimeError as e:
        ok = ok and "enable_gradient_exchange" in str(e)
    out[rank] = (bool(ok), sl.start, sl.stop, float((mine - flat).abs().max()) > 0)
    dist.destroy_process_group()


def test_gradient_average_and_sharding_world2():
    world = 2
    with mp.Manager() as man:
        out = man.dict()
        mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
        res = dict(out)
    assert all(res[r][0] for r in range(world))
    assert (res[0][1], res[0][2]) == (0, 129) and (res[1][1], res[1][2]) == (129, 257)  # contiguous, covers the batch once


def test_reference_arm_only_rank0_prints():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints one JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29612", OMP_NUM_THREADS="4")
    outs = []
    for rank in (1, 0):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "1"], env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == ""
    line = json.loads(outs[1])
    # "reference": the staged unmodified reference (oracle/_ref, made by build()); "port": the oracle, when it is not staged
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] in ("reference", "port") and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "images/s"
