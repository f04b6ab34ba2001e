"""World-size-2 gloo test of the multi-GPU host logic (shard -> local rows -> all-gather -> same table on every rank)."""
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


def _worker(rank, world, port, n_clips, genre, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from audio_key_estimation_b200 import distributed as akd
    r, _, w = akd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = akd.shard_range(n_clips, rank, world)
    ids = torch.arange(lo, hi, dtype=torch.float32)
    # stand-in "predictions" that encode the clip id so ordering errors are visible
    key = ids[:, None] + torch.arange(12)[None] / 100
    tonic = -ids[:, None] + torch.arange(12)[None] / 100
    gen = ids[:, None] * 2 + torch.arange(11)[None] / 100 if genre else None
    table = akd.gather_rows(akd.pack_rows(key, tonic, gen), n_clips)
    # asynchronous form (bench.py overlaps step i's gather with step i+1's kernels): two gathers in flight, waited in order
    rows = akd.pack_rows(key, tonic, gen)
    g1, g2 = akd.RowGather(rows, n_clips), akd.RowGather(rows * 2, n_clips)
    assert torch.equal(g1.wait(), table) and torch.equal(g2.wait(), table * 2)
    counters = akd.reduce_counters(torch.tensor([hi - lo, 1], dtype=torch.int64))
    out = akd.unpack_rows(table, genre)
    q.put((rank, table.clone(), counters.clone(), len(out)))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(target, args):
    """Spawn two ranks; a rendezvous that loses the race for its TCP port (another process grabbed it between _free_port()
    and the store's bind) is retried on a fresh port."""
    import queue as _queue
    ctx = mp.get_context("spawn")
    last = None
    for _ in range(3):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, 2, port) + tuple(args) + (q,)) for r in range(2)]
        for p in procs:
            p.start()
        results = []
        try:
            results = [q.get(timeout=120) for _ in procs]
        except _queue.Empty as e:
            last = e
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
        if len(results) == 2 and all(p.exitcode == 0 for p in procs):
            return results
        last = last or RuntimeError(f"exit codes {[p.exitcode for p in procs]}")
    raise AssertionError(f"world-size-2 run failed three times: {last}")


@pytest.mark.parametrize("n_clips,genre", [(8, True), (7, False), (1, True)])
def test_shard_gather_world2(n_clips, genre):
    results = _run_world2(_worker, (n_clips, genre))
    ids = torch.arange(n_clips, dtype=torch.float32)
    for rank, table, counters, n_out in results:
        assert table.shape == (n_clips, 35)
        assert torch.allclose(table[:, 0], ids) and torch.allclose(table[:, 12], -ids)
        assert torch.allclose(table[:, 24], ids * 2 if genre else torch.zeros(n_clips))
        assert counters.tolist() == [n_clips, 2]
        assert n_out == (3 if genre else 2)


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from audio_key_estimation_b200 import distributed as akd
    akd.init_from_env("gloo")
    flat = torch.full((167031,), float(rank + 1))  # the flat gradient bucket of the genre architecture
    flat[rank] = 10.0
    akd.allreduce_gradients(flat)
    q.put((rank, flat[:3].clone(), float(flat[100])))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    """Config 5's only collective: ONE averaged all-reduce of the flat gradient buffer."""
    results = _run_world2(_grad_worker, ())
    for rank, head, mid in results:
        assert head.tolist() == [6.0, 5.5, 1.5] and mid == 1.5
