"""where the first (cold) index build spends its wall time: per call, twice"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
import linear_b200 as lb
lb.load_library()
dev = torch.device("cuda", 0)
lens = bench.contig_lengths()
genome = bench.gen_genome(torch, dev, lens)
ctx = lb.Context(0); ctx.set_profiling(True)
torch.cuda.synchronize()
for rnd in range(2):
    t0 = time.time(); gen = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=[int(x) for x in lens]); torch.cuda.synchronize(); t1 = time.time()
    feats = lb.create_features(ctx, gen, 2, 16); torch.cuda.synchronize(); t2 = time.time()
    index = lb.create_index(ctx, gen, 1, 16); torch.cuda.synchronize(); t3 = time.time()
    print(f"round {rnd}: genome {t1 - t0:.3f} s, features {t2 - t1:.3f} s, index {t3 - t2:.3f} s")
    index.close(); feats.close(); gen.close()
