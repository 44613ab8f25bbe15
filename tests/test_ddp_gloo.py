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


def _worker_bucketed(rank, world, port, out):
    """overlap="backward": buckets are reduced from inside backward by the post-accumulate hooks, in reverse
    parameter order; plus the two failure modes the round-1 review found: set_to_none detaches the aliases,
    and ranks start from different adapter initialisations."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lycoris_b200.engine.ddp import FlatGradAllReduce

        torch.manual_seed(rank)  # DIFFERENT init per rank: the constructor must broadcast rank 0's
        params = [nn.Parameter(torch.randn(64, 8)), nn.Parameter(torch.randn(100)), nn.Parameter(torch.randn(4, 4, 4)),
                  nn.Parameter(torch.randn(33))]
        dp = FlatGradAllReduce(params, overlap="backward", bucket_bytes=4 * 60)
        torch.manual_seed(0)
        ref0 = [torch.randn(64, 8), torch.randn(100), torch.randn(4, 4, 4), torch.randn(33)]
        assert all(torch.equal(p.detach(), r) for p, r in zip(params, ref0)), "parameters not broadcast from rank 0"
        assert len(dp.buckets) >= 3
        # arena is laid out in backward (reverse) order: the LAST parameter sits at offset 0
        assert params[-1].grad.data_ptr() == dp.buckets[0].arena.data_ptr()

        def run_step(clear):
            clear()
            torch.manual_seed(100 + rank)
            xs = [torch.randn_like(p) for p in params]
            sum((p * x).sum() for p, x in zip(params, xs)).backward()
            dp.allreduce()
            dp.wait()
            expect = []
            for r in range(world):
                torch.manual_seed(100 + r)
                expect.append([torch.randn_like(p) for p in params])
            for i, p in enumerate(params):
                mean = sum(e[i] for e in expect) / world
                assert torch.allclose(p.grad, mean, atol=1e-6), (rank, i)

        run_step(dp.zero_grad)
        assert dp.buckets_overlapped == len(dp.buckets), (dp.buckets_overlapped, len(dp.buckets))
        assert dp.realiased == 0
        run_step(dp.zero_grad)  # counters re-arm
        assert dp.buckets_overlapped == len(dp.buckets)

        def set_to_none():
            for p in params:
                p.grad = None

        run_step(set_to_none)  # what optimizer.zero_grad() does by default: must still reduce the REAL grads
        assert dp.realiased == len(params)
        run_step(dp.zero_grad)  # and recover the overlapped path afterwards
        assert dp.buckets_overlapped == len(dp.buckets)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback

        out.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_bucketed_overlap_and_alias_recovery_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_bucketed, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


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
