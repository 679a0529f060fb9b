"""Developer script (GPU box): where a small-block call (64 reads, the p_calRecords pattern) spends its time.
LNR_TRACE=1 prints the host-side phase times of every lnr_apxmap_batch call; kernel times come from the context."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import linear_b200 as lb
os.environ.setdefault("LNR_BENCH_GENOME", "3100000000")
dev = torch.device("cuda", 0)
lens = bench.contig_lengths()
genome = bench.gen_genome(torch, dev, lens)
ctx = lb.Context(0)
gen = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=[int(x) for x in lens])
feats = lb.create_features(ctx, gen, 2, 16)
index = lb.create_index(ctx, gen, 1, 16)
bases_t, offs = bench.gen_reads(torch, dev, genome, lens, 64 * 64, seed=1000)
bases = bases_t.cpu().numpy()
blk = 64
def block(i):
    so = (offs[i * blk:(i + 1) * blk + 1] - offs[i * blk]).astype(np.uint64)
    return bases[int(offs[i * blk]):int(offs[(i + 1) * blk])], so
for i in range(4):
    lb.apx_map_batch(ctx, index, feats, *block(i), preset=1)
ctx.set_profiling(True); ctx.reset_kernel_times()
t0 = time.time()
for i in range(4, 36):
    lb.apx_map_batch(ctx, index, feats, *block(i), preset=1)
dt = (time.time() - t0) / 32
kt = ctx.kernel_times()
print("one thread: %.3f ms per 64-read block; kernels sum %.3f ms" % (1000 * dt, sum(v[0] for v in kt.values()) / 32))
for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:14]:
    print("   %-22s %.3f ms/block (%d launches)" % (k, v[0] / 32, v[1]))
ctx.set_profiling(False)
os.environ["LNR_TRACE"] = "1"
lb.apx_map_batch(ctx, index, feats, *block(40), preset=1)
del os.environ["LNR_TRACE"]
ctxs = [ctx] + [lb.Context(0) for _ in range(3)]
for c in ctxs[1:]:
    lb.apx_map_batch(c, index, feats, *block(0), preset=1)
def work(c, t):
    for i in range(16):
        lb.apx_map_batch(c, index, feats, *block(t * 16 + i), preset=1)
for n in (1, 2, 4):
    th = [threading.Thread(target=work, args=(ctxs[t], t)) for t in range(n)]
    t0 = time.time()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.time() - t0
    print("%d threads: %.0f reads/s, %.3f ms per block per thread" % (n, n * 16 * blk / dt, 1000 * dt / 16))
