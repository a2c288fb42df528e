#!/usr/bin/env python3
"""torchrun --nproc-per-node W tools/multigpu_check.py : sharded posterior + exchange + relax + exchange must give
every rank the same bytes a single GPU produces (developer/CI check for the NCCL path)."""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mlprobs_b200 as M
from mlprobs_b200 import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
seqs = synth.family_fast(int(sys.argv[1]) if len(sys.argv) > 1 else 60, 200, seed=5) + [synth.family(1, 700, seed=9)[0]]
n = len(seqs)


def run(sharded):
    eng = M.Engine(lr)
    h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(seqs)
    if sharded:
        uid = [M.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    if sharded: eng.exchange_begin()       # split exchange: distances() overlaps the cell broadcasts, the next stage call ends it
    d = eng.distances()
    if sharded: eng.exchange_end()
    w, sd, _, _ = M.qp_guide_tree(d)
    eng.relax(M.QP, np.maximum(w, np.float32(1e-6)), sd, 200.0, 3.0, float(np.float32(1e-5)))
    if sharded: eng.exchange()
    t = M.qp_guide_tree_ex(d)      # the tail runs on every rank over the exchanged (complete) set
    rows = eng.qp_finish_alignment(np.maximum(t["weights"], np.float32(1e-6)), t["left"], t["right"], 6)
    nnz, rp, col, val = eng.csr_bulk()
    tr = [eng.csr(b, a) for a, b in [(0, 1), (n - 2, n - 1), (3, n // 2)]]
    sig = (zlib.crc32(d.tobytes()), zlib.crc32(nnz.tobytes()), zlib.crc32(rp.tobytes()), zlib.crc32(col.tobytes()), zlib.crc32(val.tobytes()),
           tuple(zlib.crc32(x[1].tobytes()) ^ zlib.crc32(x[2].tobytes()) for x in tr), zlib.crc32(b"".join(rows)))
    eng.close()
    return sig



def run_selective(sharded, fam):
    """QuickProbs flow with the selective exchange: distances all-reduced, tree on every rank, only the matrices the consistency
    can read are imported, the relaxed set stays sharded.  Returns (crc of distances, crc of the rank-summed per-matrix digests,
    crc of distances() after a following full exchange, final alignment after gathering the set)."""
    eng = M.Engine(lr)
    h, p = M.default_tables(M.QP); eng.set_tables(h, p); eng.set_sequences(fam)
    if sharded:
        uid = [M.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
    eng.posterior_all_pairs(M.QP, 3, 0.01)
    if sharded: eng.exchange_distances()
    d = eng.distances()
    t = M.qp_guide_tree_ex(d)
    w = np.maximum(t["weights"], np.float32(1e-6))
    if sharded: eng.exchange_needed(t["seldist"], 200.0)
    eng.relax(M.QP, w, t["seldist"], 200.0, 3.0, float(np.float32(1e-5)))
    # sharded packed read-back (split call): the owned matrices' row sizes are shipped compact and scattered into the fixed layout
    lay = eng.csr_layout()
    out = M.PinnedPackedBuffers(len(fam), lay[1], lay[2])
    eng.csr_packed_begin(out); eng.csr_packed_end()
    own = M.shard_pairs(eng.lens, rank if sharded else 0, world if sharded else 1)
    for a, b in [tuple(x) for x in own[:: max(1, len(own) // 7)]]:
        for x, y in ((a, b), (b, a)):
            rp, col, val = eng.csr(x, y)
            got_rp, got_col, got_val = out.matrix(x, y, eng.lens)
            assert np.array_equal(rp, got_rp) and np.array_equal(col, got_col) and np.array_equal(val, got_val), (x, y)
    dg = torch.from_numpy(eng.set_digest().view(np.int64)).cuda()
    if sharded: dist.all_reduce(dg)                   # int64 wrap-around sum == sum mod 2^64
    dg = dg.cpu().numpy()
    if sharded:
        eng.exchange()                                # gather the relaxed set (a tail needs all of it) ...
        eng.exchange()                                # ... and a second call on the complete set must change nothing
    d2 = eng.distances()
    rows = eng.qp_finish_alignment(w, t["left"], t["right"], 4)
    nacc = int(((t["seldist"].reshape(len(fam), len(fam)) <= 200).sum() - len(fam)) // 2)
    eng.close()
    return (zlib.crc32(d.tobytes()), zlib.crc32(dg.tobytes()), zlib.crc32(d2.tobytes()), zlib.crc32(b"".join(rows))), nacc

single = run(False)
multi = run(True)
ok = single == multi
fam = synth.family_clustered(6, 45, 90, seed=11)      # 270 sequences in six sub-families: the selectivity filter rejects most third sequences
s_sel, nacc = run_selective(False, fam)
m_sel, _ = run_selective(True, fam)
ok = ok and (s_sel == m_sel) and s_sel[0] == s_sel[2]
if rank == 0:
    print("selective exchange (world=%d, n=%d, %d of %d matrices importable):" % (world, len(fam), nacc, len(fam) * (len(fam) - 1) // 2),
          "OK" if s_sel == m_sel else "MISMATCH", s_sel, m_sel)
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print("multi-GPU exchange parity (world=%d, n=%d):" % (world, n), "OK" if all(flags) else "MISMATCH %s" % flags, single[:2], multi[:2])
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
