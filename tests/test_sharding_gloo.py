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


def _quant(v):
    """QuickProbs' uint16 fixed point (SparseEntry.h:31-32) in float32 arithmetic: trunc(v * 65535) / 65535."""
    v = np.asarray(v, np.float32)
    code = np.trunc(v * np.float32(65535.0)).astype(np.uint32) & 0xffff
    return (code.astype(np.float32) / np.float32(65535.0)).astype(np.float32)


def _worker_streamed(rank, world, port, ret):
    """The streamed two-pass flow of BASELINE config #5 at world size 2, the oracle standing in for the device:
    (1) the ranks' restricted shards (mlp_shard_pairs_within = what mlp_restrict_pairs keeps) are disjoint and their union is exactly
    the set of pairs inside a <= selectivity-leaf subtree; (2) for every OTHER pair the reference's consistency repetition is the
    posterior-stage matrix with one more quantisation -- what the streamed stage emits; (3) the repetition of the pairs inside
    gives the same matrices when every matrix outside has been dropped from the set -- what the second pass keeps resident."""
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import mlprobs_b200 as M
    from mlprobs_b200 import synth
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = synth.family_clustered(3, 4, 40, seed=5)
    n = len(seqs)
    lens = [len(s) for s in seqs]
    ht, pt = O.hmm_tables(), O.part_tables(O.QP)
    SEL = 4.0                                                     # small selectivity: pairs in different sub-families are outside
    mine = M.shard_pairs(lens, rank, world)
    d = torch.zeros(n, n, dtype=torch.float32)
    for a, b in mine.tolist():
        _, dd, _ = O.pair_posterior(O.QP, 3, ht, pt, seqs[a], seqs[b])
        d[a, b] = d[b, a] = dd
    dist.all_reduce(d)                                            # mlp_exchange_distances
    w, sd, _, _ = M.qp_guide_tree(d.numpy())                      # every rank, same tree
    w = np.maximum(w, np.float32(1e-6))
    sd = np.asarray(sd, np.float32).reshape(n, n)
    within = M.shard_pairs_within(lens, rank, world, sd, SEL)
    gathered = [None] * world
    dist.all_gather_object(gathered, [tuple(p) for p in within.tolist()])
    ok = True
    if rank == 0:
        union = [p for g in gathered for p in g]
        expect = {(a, b) for a in range(n) for b in range(a + 1, n) if sd[a, b] <= SEL}
        ok = ok and len(union) == len(set(union)) and set(union) == expect and 0 < len(expect) < n * (n - 1) // 2
        full_d, S, _ = O.posterior_stage(O.QP, 3, ht, pt, seqs, threads=2)
        ok = ok and np.array_equal(full_d, d.numpy())
        cutoff = float(np.float32(1e-5))
        R = O.relax_qp(S, w, sd, cutoff, SEL, 3.0, threads=2)
        # (2) pairs outside: posterior-stage matrix, quantised once more
        for a in range(n):
            for b in range(n):
                if a != b and sd[a, b] > SEL:
                    rp0, c0, v0 = S.get(a, b); rp1, c1, v1 = R.get(a, b)
                    ok = ok and np.array_equal(rp0, rp1) and np.array_equal(c0, c1) and np.array_equal(_quant(v0), v1)
        # (3) pairs inside: unchanged when every matrix outside is emptied
        P = O.CsrSet(S.lens, S.nz_cap)
        P.rp_off[:] = S.rp_off; P.nz_off[:] = S.nz_off; P.rowptr[:] = S.rowptr; P.col[:] = S.col; P.val[:] = S.val
        P.c.rp_used = S.c.rp_used; P.c.nz_used = S.c.nz_used
        for a in range(n):
            for b in range(n):
                if a != b and sd[a, b] > SEL:
                    ro = int(P.rp_off[a * n + b]); P.rowptr[ro:ro + lens[a] + 2] = 0
        R2 = O.relax_qp(P, w, sd, cutoff, SEL, 3.0, threads=2)
        for a in range(n):
            for b in range(n):
                if a != b and sd[a, b] <= SEL:
                    ok = ok and all(np.array_equal(x, y) for x, y in zip(R.get(a, b), R2.get(a, b)))
        ret.put(bool(ok))
    dist.destroy_process_group()


def test_two_rank_streamed_flow_contract():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_streamed, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
