"""kernel times of one device ingest call on ~160 MB of 80-column FASTA"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import linear_b200 as lb
lb.load_library()
ctx = lb.Context(0)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
parts = []
for k in range(8192):
    s = alpha[rng.integers(0, 4, size=20000)].tobytes()
    parts.append(b">read%d\n" % k)
    parts.extend(s[i:i + 80] + b"\n" for i in range(0, len(s), 80))
text = b"".join(parts)
R = lb.Reads(ctx, text); R.close()
ctx.set_profiling(True); ctx.reset_kernel_times()
t0 = time.time(); R = lb.Reads(ctx, text); t1 = time.time()
print("bytes", len(text), "wall ms", 1000 * (t1 - t0), "GB/s", len(text) / (t1 - t0) / 1e9)
print({k: round(v[0], 3) for k, v in ctx.kernel_times().items()})
