"""N_train-sharded DSA / LSA over NCCL (needs >= 2 GPUs; skipped otherwise).  Every rank must
reproduce the single-GPU (= oracle) result bit for bit for DSA, and within the LSA tolerance."""
import os

import numpy as np
import pytest

from oracle import np_oracle

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from simple_tip_b200 import engine as E
        from simple_tip_b200.core.surprise import DSA

        comm = E.TrainShardComm()
        for n_train, n_test, d, classes, dt, seed in [(6000, 700, 128, 10, np.float32, 2), (2500, 300, 200, 4, np.float32, 3),
                                                      (1500, 200, 24, 3, np.float64, 4)]:
            xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(n_train, n_test, d, classes, seed=seed, dtype=dt)
            xtr[100:110] = xtr[0:10]            # duplicates that land on different shards
            ytr[100:110] = ytr[0:10]
            xte[:10], pte[:10] = xtr[0:10], ytr[0:10]
            want = np_oracle.dsa_oracle(xtr, ytr, xte, pte)
            sa = DSA(xtr, ytr, comm=comm)
            assert sa._engine.n < n_train
            got = sa(xte, pte)
            assert np.array_equal(got, want["dsa"], equal_nan=True), (rank, "dsa")
            assert np.array_equal(sa.last_winner_index, want["idx_a"]), (rank, "winner")
            assert np.array_equal(sa.last_dist_b, want["dist_b"]), (rank, "dist_b")
        q.put((rank, "ok"))
    except Exception:
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_dsa_nccl():
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results
