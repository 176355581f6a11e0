"""Multi-GPU gradient step against a record made from the REAL reference
methods (tools/make_golden_parallel.py execs ``run_gradient_descent`` / ``_get_parallel_grad`` /
``_sum_parallel_grad`` of sup3r/models/abstract.py with a stand-in ``get_single_grad``).  Here
THIS repo's ``run_gradient_descent`` runs in world_size 2 and 4 process groups (gloo, CPU), one
rank per reference "GPU", every rank fed the same batch: each rank must hand ``get_single_grad``
exactly the shard the reference gives ``/gpu:<rank>``, every rank must apply the SUM of the shard
gradients once, and every rank must return the last shard's loss details."""
import importlib.util
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "make_golden_parallel", os.path.join(ROOT, "tools", "make_golden_parallel.py"))
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)
G = json.load(open(os.path.join(ROOT, "tests", "golden", "parallel_grad.json")))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_case(name, rank):
    """One golden case through this repo's run_gradient_descent on this rank."""
    from types import SimpleNamespace
    from sup3r_b200.models.abstract import AbstractSingleModel
    n_gpus, multi_gpu, with_mask = T.CASES[name]
    calls = []

    class Scripted(AbstractSingleModel):
        def __init__(self):
            self._optimizer = T.Optimizer()
            self._graphed_steps = SimpleNamespace(run=lambda *a, **k: None)   # eager path
            self.default_device = "/cpu:0"

        def get_single_grad(self, low_res, hi_res_true, training_weights, device_name=None,
                            **kw):
            calls.append([[int(v) for v in low_res.shape],
                          sorted(k for k in kw if k != "mask")])
            mask = kw.get("mask")
            grads, det = T.shard_gradient(np.asarray(low_res), np.asarray(hi_res_true),
                                          None if mask is None else np.asarray(mask))
            return [torch.tensor(g, dtype=torch.float32) for g in grads], det
    obj = Scripted()
    lr, hr, mask = T.batch()
    kw = {"weight_gen_advers": 0.01, "train_gen": True}
    if with_mask:
        kw["mask"] = mask
    weights = [f"w{i}" for i in range(len(T.WEIGHT_SHAPES))]
    details = obj.run_gradient_descent(lr, hr, weights, multi_gpu=multi_gpu, **kw)
    want = G[name]
    sharded = multi_gpu and n_gpus > 1
    # this rank saw the shard the reference hands to /gpu:<rank> (or the whole batch)
    assert calls == [want["calls"][rank if sharded else 0][1:]], (name, calls)
    assert len(obj.optimizer.applied) == 1
    applied = obj.optimizer.applied[0]
    assert [v for _, v in applied] == want["applied_to"]
    for (g, _), w in zip(applied, want["applied"]):
        np.testing.assert_allclose(np.array(g), np.array(w), rtol=2e-6, err_msg=name)
    assert sorted(details) == sorted(want["details"])
    for k, v in want["details"].items():
        assert float(details[k]) == pytest.approx(v, rel=2e-6), (name, k)
    if sharded:
        # the reference's private entry point (abstract.py:807-841) gives the same sums
        total, det = obj._get_parallel_grad(lr, hr, weights, **kw)
        for g, w in zip(total, want["applied"]):
            np.testing.assert_allclose(g.numpy(), np.array(w), rtol=2e-6, err_msg=name)
        assert float(det["first"]) == pytest.approx(want["details"]["first"], rel=2e-6)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name, (n_gpus, multi_gpu, _) in T.CASES.items():
            # the flag-off case runs unsharded in any world; the others need one rank per GPU
            if n_gpus == world or (n_gpus == 2 and not multi_gpu):
                _run_case(name, rank)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 4])
def test_gradient_step_over_ranks_matches_reference(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=200) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_single_process_step_matches_reference():
    _run_case("one_gpu_flag_on", 0)


def test_golden_is_reproducible_from_the_reference_when_present():
    if not os.path.isdir(T.REF):
        pytest.skip("reference source not present")
    assert json.loads(json.dumps(T.scenario(T.load_reference()))) == G


def test_sum_parallel_grad_matches_reference():
    """``_sum_parallel_grad`` on per-shard results held in one process (abstract.py:785-805):
    SUM of the gradients, loss details of the last shard."""
    import time
    from types import SimpleNamespace
    from sup3r_b200.models.abstract import AbstractSingleModel
    for name, n in (("two_gpus_mask", 2), ("four_gpus", 4)):
        lr, hr, mask = T.batch()
        with_mask = T.CASES[name][2]
        futures = []
        for a, b, m in zip(np.split(lr, n), np.split(hr, n), np.split(mask, n)):
            grads, det = T.shard_gradient(a, b, m if with_mask else None)
            futures.append(SimpleNamespace(result=lambda g=grads, d=det: (
                [torch.tensor(x) for x in g], d)))
        obj = AbstractSingleModel.__new__(AbstractSingleModel)
        total, det = obj._sum_parallel_grad(futures, start_time=time.time())
        for g, w in zip(total, G[name]["applied"]):
            np.testing.assert_allclose(g.numpy(), np.array(w), rtol=1e-12, err_msg=name)
        assert {k: float(v) for k, v in det.items()} == pytest.approx(G[name]["details"])
