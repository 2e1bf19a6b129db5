"""The N > 1 host logic on CPU: two gloo ranks shard the tokens, run the oracle's backward on their shard,
all-reduce the parameter gradients and must reproduce the single-process gradients (what DDP relies on)."""
import os
import socket
import sys

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


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from hvs_b200 import dist as hd
    from oracle import mhc_ref
    g = torch.Generator().manual_seed(0)
    t = 37
    x = torch.randn(t, 4, 512, generator=g).to(torch.bfloat16)
    dy = torch.randn(t, 4, 512, generator=g).to(torch.bfloat16)
    phi = torch.randn(2048, 24, generator=g) * 0.02
    bias, alpha, scale = torch.zeros(24), torch.full((3,), 0.2), torch.ones(2048)
    lo, hi = hd.shard_range(t, world, rank)
    local = mhc_ref.stream_mhc_backward(x[lo:hi], dy[lo:hi], phi, bias, alpha, scale)
    summed = hd.allreduce_param_grads(local)
    tmax = hd.max_over_ranks(float(rank + 1), "cpu")
    dets = hd.gather_image_shards([f"img{lo + i}" for i in range(hi - lo)], world)
    if rank == 0:
        full = mhc_ref.stream_mhc_backward(x, dy, phi, bias, alpha, scale)
        ok = all(((summed[k] - full[k]).norm() <= 1e-4 * full[k].norm() + 1e-7) for k in hd.PARAM_GRAD_KEYS)
        ok = ok and tmax == float(world) and dets == [f"img{i}" for i in range(t)]
        out.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_shard_range_covers_everything():
    from hvs_b200 import dist as hd
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [hd.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
