"""Developer tool: one small pass over every entry point, meant to run under `compute-sanitizer --tool memcheck`."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import linear_b200 as lb  # noqa: E402
from linear_b200 import datagen  # noqa: E402

lens = datagen.contig_lengths(300_000, 2, seed=3)
g = datagen.make_genome(5, lens, n_families=2, copies=40, n_tandem=5)
rs = datagen.simulate_reads(9, g, 12, mean_len=5000, sd_len=1500, err=0.08, sv_frac=0.3)
rng = np.random.default_rng(1)
junk = rng.integers(0, 4, size=4000, dtype=np.uint8)
bases = np.concatenate([rs.bases, junk])
offs = np.concatenate([rs.offsets, [rs.offsets[-1] + len(junk)]]).astype(np.uint64)
ctx = lb.Context(0)
gen = lb.Genome(ctx, g)
for ft in (2, 1):
    feats = lb.create_features(ctx, gen, ft, 4)
    for it in (1, 2):
        index = lb.create_index(ctx, gen, it, 4)
        c, o = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=1)
        c0, o0 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=1, no_chain=True)
        c1, o1 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=1, no_chain=True, gdl_state=1)
        print("ft", ft, "index", it, "cords", len(c), len(c0), len(c1), flush=True)
        if ft == 2:
            p = os.path.join(tempfile.mkdtemp(), "x.lnridx")
            index.save(p)
            back = lb.Index.load(ctx, p)
            cb, ob = lb.apx_map_batch(ctx, back, feats, bases, offs, preset=1)
            assert np.array_equal(cb, c)
            back.close()
        index.close()
    if ft == 2:
        recs = lb.cords_to_records(ctx, c, o, np.diff(offs).astype(np.uint64))
        print("records", len(recs[0]), flush=True)
    feats.close()
print("done")
