"""Developer probe: where does the -i 2 (HIndex) path spend its time at a given genome size?
usage: python tools/hindex_probe.py GENOME_BASES N_READS [index_type]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import linear_b200 as lb  # noqa: E402
from linear_b200 import datagen  # noqa: E402

G, n_reads = int(float(sys.argv[1])), int(sys.argv[2])
itype = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lens = datagen.contig_lengths(G, 24 if G >= 1e9 else 4, seed=31)
dev = torch.device("cuda", 0)
genome = bench.gen_genome(torch, dev, lens)
torch.cuda.synchronize()
print("genome", G, flush=True)
ctx = lb.Context(0)
ctx.set_profiling(True)
gen = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=[int(x) for x in lens])
t0 = time.time(); feats = lb.create_features(ctx, gen, 2, 16); torch.cuda.synchronize(); print("features s", round(time.time() - t0, 3), flush=True)
ctx.reset_kernel_times()
t0 = time.time(); ix = lb.create_index(ctx, gen, itype, 16); torch.cuda.synchronize()
print("index build s", round(time.time() - t0, 3), {k: round(v[0], 2) for k, v in ctx.kernel_times().items()}, flush=True)
bases_t, offs = bench.gen_reads(torch, dev, genome, lens, n_reads, seed=1000)
bases = bases_t.cpu().numpy()
offs = np.asarray(offs, dtype=np.uint64)
for nb in (64, n_reads):
    ctx.reset_kernel_times()
    t0 = time.time()
    cords, coff = lb.apx_map_batch(ctx, ix, feats, bases[: int(offs[nb])], offs[: nb + 1], preset=1)
    dt = time.time() - t0
    print("map", nb, "reads s", round(dt, 3), "cords", len(cords), ctx.counters(), {k: round(v[0], 2) for k, v in ctx.kernel_times().items()}, flush=True)
