"""Worker of tests/test_gpu_sharded_build.py: one process per rank. Builds the DIndex (or, argv[5] = 2, the HIndex) of a seeded
case with lnr_index_build_sharded (the library's own NCCL exchange; the unique id travels through a file) and compares the
assembled dir / hs (ysa / directory) with the CPU oracle, then maps the case's reads with it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, idfile, case_name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    index_type = int(sys.argv[5]) if len(sys.argv) > 5 else 1
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
    index = lb.create_index_sharded(ctx, gen, comm, index_type, T)
    O = Oracle(g, threads=T, preset=preset, index_type=index_type)
    if index_type == 1:
        d1, h1 = index.export_dindex()
        d0, h0 = O.dindex()
        assert np.array_equal(d0, d1), "dir differs on rank %d" % rank
        assert np.array_equal(h0, h1), "hs differs on rank %d" % rank
        n_rec = len(h1)
    else:
        y1, e1, kv1, tl1 = index.export_hindex()
        y0, e0, kv0, tl0 = O.hindex()
        assert len(y0) == len(y1) and e0 == e1 and tl0 == tl1, "HIndex sizes differ on rank %d" % rank
        assert np.array_equal(y0, y1), "ysa differs on rank %d" % rank
        assert np.array_equal(kv0, kv1), "directory differs on rank %d" % rank
        single = lb.create_index(ctx, gen, 2, T)          # and the single-GPU build of the same library
        y2, e2, kv2, tl2 = single.export_hindex()
        assert e2 == e1 and tl2 == tl1 and np.array_equal(y2, y1) and np.array_equal(kv2, kv1), "sharded != single-GPU build on rank %d" % rank
        n_rec = len(y1)
    feats = lb.create_features(ctx, gen, 2, T)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    oc, oo = O.map_batch(bases, offs, map_threads=2)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords), "cords differ on rank %d" % rank
    comm.close()
    print("rank %d ok: n_hs=%d cords=%d" % (rank, n_rec, len(cords)))


if __name__ == "__main__":
    main()
