"""CPU, world_size 2, gloo: the data-parallel adapter-gradient exchange (flat arena, one
all-reduce, averaging) — host logic of engine/ddp.py; NCCL takes the same code path on GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lycoris_b200.engine.ddp import FlatGradAllReduce

        torch.manual_seed(0)  # identical adapter init on every rank
        params = [nn.Parameter(torch.randn(5, 3)), nn.Parameter(torch.randn(7)), nn.Parameter(torch.randn(2, 2, 2))]
        dp = FlatGradAllReduce(params, overlap=False)
        assert dp.num_elements == 15 + 7 + 8
        # grads are views into one arena
        base = params[0].grad.untyped_storage().data_ptr()
        assert all(p.grad.untyped_storage().data_ptr() == base for p in params)
        # rank-dependent "shard" of the batch
        torch.manual_seed(100 + rank)
        xs = [torch.randn_like(p) for p in params]
        loss = sum((p * x).sum() for p, x in zip(params, xs))
        loss.backward()
        assert params[0].grad.untyped_storage().data_ptr() == base, "autograd must accumulate in place"
        local = [x.clone() for x in xs]
        dp.allreduce()
        dp.wait()
        # expected: mean over ranks of the local grads
        expect = []
        for r in range(world):
            torch.manual_seed(100 + r)
            expect.append([torch.randn_like(p) for p in params])
        for i, p in enumerate(params):
            mean = sum(e[i] for e in expect) / world
            assert torch.allclose(p.grad, mean, atol=1e-6), (rank, i)
            assert not torch.allclose(p.grad, local[i]) or world == 1
        dp.zero_grad()
        assert all(float(p.grad.abs().sum()) == 0 for p in params)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_a_noop():
    from lycoris_b200.engine.ddp import FlatGradAllReduce

    p = nn.Parameter(torch.ones(4))
    dp = FlatGradAllReduce([p], overlap=False)
    (p * 2).sum().backward()
    dp.allreduce()
    assert torch.equal(p.grad, torch.full((4,), 2.0))
