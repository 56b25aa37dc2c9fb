"""world_size-2 gloo test of the N>1 host logic: image partition + detection all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leanyolo_b200.dist import ShardedDetector, gather_detections, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_dets(n_images, k=7):
    g = torch.Generator().manual_seed(5)
    return torch.rand(n_images, k, 6, generator=g)


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _fake_dets(n_images)
    lo, hi = shard_range(n_images, rank, world)
    got = gather_detections(full[lo:hi].clone(), n_images)      # partition known
    got2 = gather_detections(full[lo:hi].clone())               # sizes exchanged
    # the overlapped form (ring of pre-allocated gather buffers, tickets): even shards
    even = _fake_dets(2 * world)
    sd = ShardedDetector(model=None, group=None, depth=2)
    t0 = sd.submit_detections(even[2 * rank:2 * rank + 2].clone())
    t1 = sd.submit_detections(even[2 * rank:2 * rank + 2].clone() + 1.0)
    ok3 = bool(torch.equal(sd.collect(t0), even)) and bool(torch.equal(sd.collect(t1), even + 1.0))
    t2 = sd.submit_detections(even[2 * rank:2 * rank + 2].clone() + 2.0)
    try:
        sd.collect(t0)
        ok3 = False            # ticket 0 was overwritten by ticket 2 (depth 2): must raise
    except RuntimeError:
        pass
    ok3 = ok3 and bool(torch.equal(sd.collect(t2), even + 2.0))
    q.put((rank, bool(torch.equal(got, full)), bool(torch.equal(got2, full)) and ok3, (lo, hi)))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (1, 2, 7, 256, 2048):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_gather_detections_world2_uneven():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
    assert sorted(r[3] for r in res) == [(0, 3), (3, 5)]
