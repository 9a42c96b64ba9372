"""world_size-2 gloo tests (CPU) for the multi-process host logic: batch sharding, bucketed and
overlapped gradient averaging, loss reduction (mirror of Miscellaneous/distributed.py)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "3d-fm-gan_b200"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from Miscellaneous import distributed as D
    r, w, _ = D.init_distributed(backend="gloo")
    assert (r, w) == (rank, world) and D.get_rank() == rank and D.get_world_size() == world
    D.synchronize()
    # reduce_sum leaves its input alone
    t = torch.full((3,), float(rank + 1))
    s = D.reduce_sum(t)
    assert torch.equal(s, torch.full((3,), 3.0)) and torch.equal(t, torch.full((3,), float(rank + 1)))
    # batch sharding covers the batch exactly once
    spans = D.all_gather(D.shard_batch(7))
    assert spans == [(0, 4), (4, 7)]
    # bucketed gather_grad == mean of per-rank grads
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(10 + rank))
    model(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    D.gather_grad(model.parameters(), bucket_bytes=256)
    others = D.all_gather(local)
    for p, a, b in zip(model.parameters(), others[0], others[1]):
        assert torch.allclose(p.grad, (a + b) / 2, atol=1e-6)
    # overlapped reducer gives the same averages (hooks fire during backward), twice in a row
    red = D.GradBucketReducer(model.parameters(), bucket_mb=256 / (1 << 20))
    assert len(red.buckets) >= 2
    for it in range(2):
        model.zero_grad(set_to_none=True)
        xi = torch.randn(5, 8, generator=torch.Generator().manual_seed(20 + 2 * it + rank))
        model(xi).pow(2).sum().backward()
        red.finish()
        mine = [p.grad.clone() for p in model.parameters()]
        # recompute the expected mean without communication hooks
        ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        ref.load_state_dict(model.state_dict())
        exp = []
        for rr in range(world):
            ref.zero_grad(set_to_none=True)
            xr = torch.randn(5, 8, generator=torch.Generator().manual_seed(20 + 2 * it + rr))
            ref(xr).pow(2).sum().backward()
            exp.append([p.grad.clone() for p in ref.parameters()])
        for g, a, b in zip(mine, exp[0], exp[1]):
            assert torch.allclose(g, (a + b) / 2, atol=1e-6)
    # contract: a second parameter-gradient backward before finish() raises instead of silently dropping gradients
    model.zero_grad(set_to_none=True)
    model(x).pow(2).sum().backward()
    try:
        model(x).pow(2).sum().backward()
        raise AssertionError("second backward before finish() did not raise")
    except RuntimeError as e:
        assert "finish()" in str(e)
    red.finish()
    # a parameter without a gradient keeps its zero-filled slot: equal message sizes on every rank, no hang
    model.zero_grad(set_to_none=True)
    model[0](x).pow(2).sum().backward()               # only the first Linear receives gradients
    red.finish()
    assert model[2].weight.grad is None and model[0].weight.grad is not None
    g0 = D.all_gather(model[0].weight.grad.clone())
    assert torch.allclose(g0[0], g0[1])
    red.remove()
    # loss dict: rank 0 holds the mean
    out = D.reduce_loss_dict({"b": torch.tensor(float(rank)), "a": torch.tensor(2.0 * rank)})
    if rank == 0:
        assert abs(float(out["a"]) - 1.0) < 1e-6 and abs(float(out["b"]) - 0.5) < 1e-6
    torch.distributed.destroy_process_group()
    ret[rank] = True


def test_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_single_process_noops():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "3d-fm-gan_b200"))
    from Miscellaneous import distributed as D
    assert D.get_rank() == 0 and D.get_world_size() == 1
    D.synchronize()
    t = torch.ones(2)
    assert D.reduce_sum(t) is t
    d = {"x": torch.tensor(1.0)}
    assert D.reduce_loss_dict(d) is d
    assert D.all_gather(5) == [5]
    assert D.shard_batch(10, 1, 4) == (3, 6)
    D.gather_grad([torch.nn.Parameter(torch.ones(1))])
