"""Worker of tests/test_gpu_sharded_build.py: one process per rank. Builds the DIndex of a seeded case with
lnr_index_build_sharded (the library's own NCCL exchange; the unique id travels through a file) and compares the assembled
dir / hs with the CPU oracle, then maps the case's reads with it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, idfile, case_name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    import linear_b200 as lb
    from cases import make_case
    from cpu_checkers import Oracle
    g, reads, bases, offs, T, preset = make_case(case_name)
    import torch
    dev = rank % max(torch.cuda.device_count(), 1)
    ctx = lb.Context(dev)

    def exchange(idb):
        if idb is not None:
            with open(idfile + ".tmp", "wb") as f:
                f.write(idb)
            os.replace(idfile + ".tmp", idfile)
            return idb
        for _ in range(600):
            if os.path.exists(idfile):
                return open(idfile, "rb").read()
            time.sleep(0.1)
        raise TimeoutError("no NCCL unique id from rank 0")

    comm = lb.Comm(ctx, rank, world, exchange)
    gen = lb.Genome(ctx, g)
    index = lb.create_index_sharded(ctx, gen, comm, 1, T)
    d1, h1 = index.export_dindex()
    O = Oracle(g, threads=T, preset=preset)
    d0, h0 = O.dindex()
    assert np.array_equal(d0, d1), "dir differs on rank %d" % rank
    assert np.array_equal(h0, h1), "hs differs on rank %d" % rank
    feats = lb.create_features(ctx, gen, 2, T)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    oc, oo = O.map_batch(bases, offs, map_threads=2)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords), "cords differ on rank %d" % rank
    comm.close()
    print("rank %d ok: n_hs=%d cords=%d" % (rank, len(h1), len(cords)))


if __name__ == "__main__":
    main()
