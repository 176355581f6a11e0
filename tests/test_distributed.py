"""N > 1 host logic on CPU with the gloo backend (world_size 2): gradient SUM all-reduce
(reference semantics sup3r/models/abstract.py:785-805), weight broadcast, and the chunk
partition over ranks (sup3r/pipeline/strategy.py:363-372)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sup3r_b200 import configs as C


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sup3r_b200 import parallel
        from sup3r_b200.network import CustomNetwork
        from sup3r_b200.models import Sup3rGan
        from sup3r_b200.pipeline import ArrayInputHandler, ForwardPassStrategy
        assert parallel.world_size() == world and parallel.rank() == rank
        # gradient all-reduce is a SUM over shards, not a mean
        grads = [torch.full((3, 2), float(rank + 1)), torch.arange(4.0) * (rank + 1)]
        red = parallel.allreduce_sum_grads(grads)
        assert torch.equal(red[0], torch.full((3, 2), 3.0))
        assert torch.equal(red[1], torch.arange(4.0) * 3)
        # the reduced gradients are views of ONE persistent flat arena (reused across steps)
        red2 = parallel.allreduce_sum_grads([torch.ones(3, 2), torch.ones(4)])
        assert red2[0].data_ptr() == red[0].data_ptr()
        assert red2[1].data_ptr() == red[0].data_ptr() + 6 * 4
        assert torch.equal(red2[1], torch.full((4,), 2.0))
        # tf.split semantics of the batch shards (abstract.py:819-825)
        batch = torch.arange(8.0).reshape(4, 2)
        assert torch.equal(parallel.shard_batch(batch), batch[2 * rank:2 * rank + 2])
        try:
            parallel.shard_batch(torch.zeros(3, 2))
            raise AssertionError("uneven batch must raise")
        except ValueError:
            pass
        # loss details of the LAST shard reach every rank (same GAN schedule everywhere)
        det = parallel.broadcast_loss_details({"loss_gen": torch.tensor(1.0 + rank),
                                               "loss_disc": 10.0 * (rank + 1)})
        assert float(det["loss_gen"]) == 2.0 and float(det["loss_disc"]) == 20.0
        # weights differ per rank before the broadcast, are identical to rank 0's afterwards
        CustomNetwork.seed(100 + rank)
        net = CustomNetwork(C.spatial_generator(2, (2,), n_blocks=1), name="g", device="cpu")
        net.build((1, 6, 6, 2))
        before = [w.copy() for w in net.get_weights()]
        parallel.broadcast_weights([net], src=0)
        gathered = [None] * world
        dist.all_gather_object(gathered, [w.tobytes() for w in net.get_weights()])
        assert gathered[0] == gathered[1]
        if rank == 1:
            assert any(not np.array_equal(a, b) for a, b in zip(before, net.get_weights()))
        # chunk partition: every chunk on exactly one rank, np.array_split sizes
        model = Sup3rGan(C.spatiotemporal_generator(2, 2, (2,), n_blocks=1),
                         C.discriminator(3, "same", (8,)), default_device="/cpu:0",
                         meta={"lr_features": ["u", "v"], "hr_out_features": ["u", "v"],
                               "s_enhance": 2, "t_enhance": 2})
        data = np.zeros((20, 12, 30, 2), np.float32)
        strat = ForwardPassStrategy(model=model, input_handler=ArrayInputHandler(data, ["u", "v"]),
                                    fwp_chunk_shape=(8, 8, 10), spatial_pad=2, temporal_pad=2,
                                    max_nodes=world)
        mine = [int(c) for c in strat.node_chunks[rank]]
        allc = [None] * world
        dist.all_gather_object(allc, mine)
        flat = sorted(c for part in allc for c in part)
        assert flat == list(range(strat.n_chunks)) and strat.n_chunks == 18
        assert abs(len(allc[0]) - len(allc[1])) <= 1
        assert parallel.split_chunks(list(range(7)), 2) == [[0, 1, 2, 3], [4, 5, 6]]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
