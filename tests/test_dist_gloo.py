"""CPU, world_size 2, gloo: the rank-to-rank plumbing of the replicas-only multi-GPU path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200 import dist as mdist
    from mintime_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = mdist.shard_range(total, rank, world)
    # each rank "computes" logits for its shard of the seeded batch (the value encodes the global clip index)
    meta = synth.make_batch_meta(total, 8, [1, 2], seed=5)
    local = torch.arange(a, b, dtype=torch.float32).unsqueeze(1) + meta["mask"][a:b].float().sum(1, keepdim=True) * 100
    full = mdist.gather_rows(local, total)
    ms = mdist.max_over_ranks(10.0 + rank)
    mdist.barrier()
    q.put((rank, a, b, full.flatten().tolist(), ms))
    dist.destroy_process_group()


def test_shard_gather_and_max_over_ranks():
    world, total = 2, 7            # ragged: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200 import synth
    meta = synth.make_batch_meta(total, 8, [1, 2], seed=5)
    expect = (torch.arange(total, dtype=torch.float32) + meta["mask"].float().sum(1) * 100).tolist()
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 7)]
    for r in res:
        assert r[3] == expect               # every rank sees the whole batch, in clip order
        assert r[4] == 11.0                 # max over ranks


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200.dist import shard_range
    for total in (0, 1, 5, 32, 64, 100):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _grad_sync_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200 import training
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sync = training.GradSync()
    # three "layer buckets" launched one after the other as the backward would, finished together
    buckets = [torch.full((1000 + 7 * i,), float(rank + 1) * (i + 1)) for i in range(3)]
    views = [b[10:20] for b in buckets]                 # parameter gradients are views of the flat buckets
    for b in buckets:
        sync.launch(b)
    sync.finish()
    # deferred mode (GraphedTrainStep): the backward only records its buckets, exchange() reduces them afterwards
    sync.deferred = True
    late = [torch.full((64,), float(rank + 1) * 10.0), torch.full((8,), float(rank))]
    for b in late:
        sync.launch(b)
    sync.finish()
    untouched = [float(b[0]) for b in late]
    sync.exchange()
    q.put((rank, [float(b[0]) for b in buckets], [float(v.sum()) for v in views], len(sync.pending), untouched,
           [float(b[0]) for b in late]))
    dist.destroy_process_group()


def test_grad_sync_averages_layer_buckets_over_ranks():
    """training.GradSync (the gradient exchange of train.py under data parallelism): per-layer flat buckets are
    all-reduced asynchronously and averaged; views carved out of a bucket see the averaged values."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] == [1.5 * (i + 1) for i in range(3)]          # mean of (1, 2) * (i + 1)
        assert r[2] == [15.0 * (i + 1) for i in range(3)]
        assert r[3] == 0
        assert r[4] == [10.0 * (r[0] + 1), float(r[0])]             # nothing moved before exchange()
        assert r[5] == [15.0, 0.5]                                   # mean over the two ranks afterwards


def _attach_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200 import SizeInvariantTimeSformer, training
    from mintime_b200.spec import default_tsf_config
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = default_tsf_config(num_frames=8)
    cfg["model"]["depth"] = 1
    torch.manual_seed(100 + rank)                       # every rank initialises differently ...
    model = SizeInvariantTimeSformer(config=cfg)
    before = float(model.to_out[1].weight.double().sum())
    training.attach_grad_sync(model)                    # ... and starts from rank 0's parameters, like DDP
    after = float(model.to_out[1].weight.double().sum())
    bad = None
    try:
        model._grad_sync.launch(torch.zeros(8, 8)[:, :4])
    except ValueError as e:
        bad = str(e)
    q.put((rank, before, after, bad))
    dist.destroy_process_group()


def test_attach_grad_sync_broadcasts_rank0_parameters_and_rejects_copies():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_attach_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] != res[1][1]                       # different seeds
    assert res[0][2] == res[1][2] == res[0][1]          # both hold rank 0's values afterwards
    assert all("contiguous" in r[3] for r in res)


def _extractor_sync_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import mintime_b200  # noqa: F401
    from mintime_b200 import training
    from mintime_b200.efficientnet_train import EffnetTrainFunction
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7 + rank)
    ext = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))     # stands in for the extractor's containers
    ext[1].running_mean.fill_(float(rank + 1))
    training.attach_grad_sync(ext)                       # parameters AND buffers of rank 0 everywhere
    ext[1].weight.requires_grad_(False)                  # a frozen tensor gets no gradient and is not exchanged

    class Ctx:
        pass

    ctx = Ctx()
    ctx.ext, ctx.names = ext, [k for k, _ in ext.named_parameters()]
    G = {"0.weight": torch.full((2, 3), float(rank + 1)), "0.bias": torch.full((2,), 10.0 * (rank + 1)),
         "1.weight": torch.full((2,), 99.0), "1.bias": torch.full((2,), float(rank))}
    out = EffnetTrainFunction._pack(ctx, G)
    q.put((rank, float(ext[1].running_mean[0]), [None if o is None else float(o.flatten()[0]) for o in out],
           float(ext[0].weight.double().sum())))
    dist.destroy_process_group()


def test_extractor_gradients_are_averaged_over_ranks():
    """an unfrozen extractor under data parallelism (train.py:153-170 + :294-296): attach_grad_sync(extractor) broadcasts
    rank 0's parameters and BatchNorm buffers, and the backward's gradients leave as one averaged bucket"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_extractor_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 1.0                 # rank 0's running_mean
    assert res[0][3] == res[1][3]                        # rank 0's weights
    for r in res:
        assert r[2] == [None, None, None, 1.5, 15.0, None, 0.5]
