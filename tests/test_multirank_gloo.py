"""The N > 1 path on CPU: two gloo ranks each take their slice of the shard plan, analyse it (oracle standing in for
the kernel -- this test is about the host logic), and the ordered gather equals the unsharded result; the benchmark's
max-over-ranks timing reduction is exercised the same way."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import aus_oracle as orc
    import aus_rust_b200 as aus
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, hop, sr = 512, 128, 44100
    x = orc.synth_signal(60 * hop + 77, seed=21)                     # every rank can regenerate the input
    vo, vl, vr, vd = aus.plan_shards([0], [x.size], N, hop, world)
    mine = [i for i in range(len(vo)) if vd[i] == rank]
    parts = [orc.stft_features(x[vo[i]:vo[i] + vl[i]], N, hop, orc.HANNING, sr) for i in mine]
    local = np.concatenate(parts) if parts else np.zeros((0, 14))
    rows = torch.tensor([local.shape[0]], dtype=torch.int64)
    all_rows = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    maxr = int(max(r.item() for r in all_rows))
    pad = torch.zeros((maxr, 14), dtype=torch.float64)
    pad[:local.shape[0]] = torch.from_numpy(local)
    gathered = [torch.zeros((maxr, 14), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, pad)
    full = np.concatenate([g[:int(r.item())].numpy() for g, r in zip(gathered, all_rows)])
    t = torch.tensor([10.0 + rank], dtype=torch.float64)             # bench.py: elapsed ms, MAX over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((full, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_equals_unsharded():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aus_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = orc.synth_signal(60 * 128 + 77, seed=21)
    ref = orc.stft_features(x, 512, 128, orc.HANNING, 44100)
    assert np.array_equal(full, ref)
    assert tmax == 11.0
