"""Host-side data-parallel logic on CPU: world_size 2, gloo backend (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from starcop_b200 import parallel
    r, lr, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # gradient sync: sum across ranks, scale = 1/world (DDP mean)
    flat = torch.full((1000,), float(rank + 1))
    scale = parallel.GradSync()(flat)
    assert scale == 0.5 and torch.all(flat == 3.0)
    # bucketed exchange: the tail bucket (decoder + head) early, the rest at the end -- same sums as one all-reduce
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    gs = parallel.GradSync(bucketed=True)
    gs.early(flat, 340)
    assert torch.equal(flat[340:], torch.arange(340, 1000, dtype=torch.float32) * 3) and \
        torch.equal(flat[:340], torch.arange(340, dtype=torch.float32) * (rank + 1))
    assert gs(flat) == 0.5 and torch.equal(flat, torch.arange(1000, dtype=torch.float32) * 3)
    flat = torch.full((10,), float(rank + 1))
    assert gs(flat) == 0.5 and torch.all(flat == 3.0)          # the next step without an early bucket: one all-reduce
    # parameter broadcast from rank 0
    p = torch.full((10,), float(rank))
    parallel.broadcast_parameters(p, [torch.full((3,), float(rank))])
    assert torch.all(p == 0)
    # confusion-matrix reduction is exact integer arithmetic
    cm = torch.tensor([[10 + rank, 1], [2, 3 * rank]], dtype=torch.long)
    parallel.reduce_confusion(cm)
    assert cm.tolist() == [[21, 2], [4, 3]]
    # tile sharding covers every tile exactly once
    mine = parallel.shard_tiles(9, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(9))
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]
