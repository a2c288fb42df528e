"""CPU, world_size 2 over gloo: the host-side sharding contract of the multi-GPU path.
Each rank takes mlp_shard_pairs(rank, world), computes its pairs (the oracle stands in for the device), and the
fixed-layout "sum == union" merge that mlp_exchange performs with ncclAllReduce is replayed with gloo all_reduce.
The merged distances / per-pair cell counts must equal the unsharded result."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, ret):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import mlprobs_b200 as M
    from mlprobs_b200 import synth
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = synth.family(7, 40, seed=21)
    n = len(seqs)
    lens = [len(s) for s in seqs]
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    mine = M.shard_pairs(lens, rank, world)
    d = torch.zeros(n, n, dtype=torch.float32)
    cnt = torch.zeros(n, n, dtype=torch.int32)
    for a, b in mine.tolist():
        post, dd, _ = O.pair_posterior(O.QP, 3, ht, pt, seqs[a], seqs[b])
        d[a, b] = d[b, a] = dd
        cnt[a, b] = cnt[b, a] = int((post >= np.float32(0.01)).sum())
    dist.all_reduce(d); dist.all_reduce(cnt)          # foreign slots are zero: sum == union
    sizes = [None] * world
    dist.all_gather_object(sizes, len(mine))
    if rank == 0:
        full, S, _ = O.posterior_stage(O.QP, 3, ht, pt, seqs, threads=2)
        ok = np.array_equal(d.numpy(), full) and sum(sizes) == n * (n - 1) // 2
        ok = ok and all(int(cnt[a, b]) == len(S.get(a, b)[1]) for a in range(n) for b in range(a + 1, n))
        ret.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_shard_and_merge():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
