"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: gradient all-reduce over the arena,
gather / pad_across_processes, parameter broadcast, and the contiguous split of conditionings over ranks."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videometamaterials_b200.accel import Accelerator, broadcast_object_list
    from videometamaterials_b200.blocks_bwd import GradArena
    acc = Accelerator(cpu=True)
    assert acc.num_processes == world and acc.process_index == rank
    torch.manual_seed(rank)                       # different initial weights per rank: prepare() must broadcast rank 0's
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    net = acc.prepare(net)
    w0 = net[0].weight.detach().clone()
    arena = GradArena(net)
    net._vmm_arena = arena
    arena.flat_grad.fill_(float(rank + 1))        # rank-dependent gradients
    acc.all_reduce_gradients()
    mean = sum(range(1, world + 1)) / world
    ok_grad = bool(torch.allclose(arena.flat_grad, torch.full_like(arena.flat_grad, mean)))
    g = acc.gather(torch.tensor(float(rank)))
    v = torch.ones(rank + 1, 2) * rank
    padded = acc.pad_across_processes(v, dim=0)
    gathered = acc.gather(padded)
    obj = broadcast_object_list([{"cond": [1, 2, 3]} if rank == 0 else None])[0]

    class T:      # minimal stand-in carrying what cond_to_gpu reads
        accelerator, test_batch_size = acc, 2
    from videometamaterials_b200.trainer import Trainer
    chunks = Trainer.cond_to_gpu(T, torch.arange(10).reshape(5, 2))
    acc.wait_for_everyone()
    q.put((rank, w0, ok_grad, g.tolist(), tuple(padded.shape), tuple(gathered.shape), obj, [c.tolist() for c in chunks]))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, ok0, g0, ps0, gs0, o0, c0), (r1, w1, ok1, g1, ps1, gs1, o1, c1) = res
    assert torch.equal(w0, w1)                                    # parameters broadcast from rank 0
    assert ok0 and ok1                                            # averaged gradient on both ranks
    assert g0 == g1 == [0.0, 1.0]
    assert ps0 == ps1 == (2, 2) and gs0 == gs1 == (4, 2)
    assert o0 == o1 == {"cond": [1, 2, 3]}
    # VDDP:1506-1532: rank 0 gets rows [0, 2), the last rank the remainder [2, 5), chunked by test_batch_size
    assert c0 == [[[0, 1], [2, 3]]] and c1 == [[[4, 5], [6, 7]], [[8, 9]]]
