"""Host-side logic of the training fast paths, on CPU: the piece table of the fused sum + Adam
kernel (``parallel.GradArena``) and the bookkeeping of the CUDA-graph step cache
(``train_graph.GraphedSteps``).  The kernels themselves are covered by the ``-m gpu`` tests."""
import numpy as np
import torch

from sup3r_b200 import parallel, train_graph
from sup3r_b200.network import Variable
from sup3r_b200.optimizers import Adam


def test_adam_piece_table_covers_every_weight_once():
    shapes = [(3, 3, 3, 6, 64), (64,), (5000,), (4096,), (1,)]
    weights = [Variable(f"w{i}:0", torch.zeros(s)) for i, s in enumerate(shapes)]
    grads = [torch.zeros(s) for s in shapes]
    arena = parallel.LocalArena(grads)
    assert arena.arena.numel() == sum(int(np.prod(s)) for s in shapes)
    assert [tuple(v.shape) for v in arena.in_views] == shapes
    opt = Adam(1e-3)
    slots = [opt.slots_for(w) for w in weights]
    ptrs = tuple((w.value.data_ptr(), m.data_ptr(), v.data_ptr())
                 for w, (m, v) in zip(weights, slots))
    rec = arena.segment_table(ptrs, weights)
    assert rec.dtype == np.uint64 and rec.shape[1] == 5
    assert rec[:, 4].max() <= arena.PIECE and rec[:, 4].min() >= 1
    off = 0
    for (wp, mp, vp), s in zip(ptrs, shapes):
        n = int(np.prod(s))
        mine = rec[(rec[:, 3] >= off) & (rec[:, 3] < off + n)]
        assert mine[:, 4].sum() == n                       # every element exactly once
        assert np.array_equal(mine[:, 3], off + np.arange(len(mine)) * arena.PIECE)
        k = (mine[:, 3] - off).astype(np.int64)
        assert np.array_equal(mine[:, 0], wp + 4 * k) and np.array_equal(mine[:, 1], mp + 4 * k)
        assert np.array_equal(mine[:, 2], vp + 4 * k)
        off += n
    assert len(rec) == sum(-(-int(np.prod(s)) // arena.PIECE) for s in shapes)
    # staging copies the gradients into the arena views
    g2 = [torch.full(s, float(i + 1)) for i, s in enumerate(shapes)]
    arena.stage(g2)
    assert float(arena.arena.sum()) == sum((i + 1) * int(np.prod(s)) for i, s in enumerate(shapes))


def test_step_arena_only_for_this_librarys_adam_on_cuda():
    g = [torch.zeros(4)]
    assert parallel.step_arena(g, Adam(1e-3)) is None           # CPU tensors: optimiser's own step
    assert parallel.step_arena(g, object()) is None
    assert parallel.step_arena([], Adam(1e-3)) is None


class _Net:
    def __init__(self, n):
        self.weights = [Variable(f"k{i}:0", torch.zeros(2)) for i in range(n)]
        self.layers = []


class _Model:
    _graph_safe = True
    precision = "fp16c"

    def __init__(self):
        self.generator, self.discriminator = _Net(2), _Net(1)

    def torch_device(self):
        return torch.device("cpu")


def test_graph_cache_bookkeeping(monkeypatch):
    gs = train_graph.GraphedSteps(_Model())
    # never eligible off the GPU: run() hands the step back to the eager path
    assert gs.run(np.zeros((1, 2)), np.zeros((1, 4)), [], None, False, {}) is None
    assert gs.stats == {"captures": 0, "replays": 0, "recaptures": 0}
    monkeypatch.setenv("SUP3R_B200_TRAIN_GRAPH", "0")
    assert not gs.enabled()
    monkeypatch.delenv("SUP3R_B200_TRAIN_GRAPH")
    assert gs.enabled()
    # subclasses have to opt in themselves
    class Child(_Model):
        pass
    assert type(gs.model).__dict__.get("_graph_safe") and not Child.__dict__.get("_graph_safe")
    # stochastic layers keep a model eager
    class GaussianNoiseAxis:
        pass
    assert gs._deterministic()
    gs.model.generator.layers.append(GaussianNoiseAxis())
    assert not gs._deterministic()
    # version bump reaches every weight of both networks
    gs._bump_versions()
    assert [w.version for n in gs._nets() for w in n.weights] == [1, 1, 1]
    # at most MAX_GRAPHS live graphs, failed keys (False) do not count, oldest goes first
    for k in range(6):
        gs._steps[k] = object() if k != 1 else False
        gs._seen[k] = 2
        gs._evict(k)
    live = [k for k, v in gs._steps.items() if v is not False]
    assert live == [2, 3, 4, 5][-train_graph.MAX_GRAPHS:] and gs._steps[1] is False
    assert train_graph._shape([[1, 2, 3]]) == (1, 3) and train_graph._shape(torch.zeros(2, 5)) == (2, 5)
