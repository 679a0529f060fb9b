#!/usr/bin/env python
"""bench.py -- reads/sec of apx-map + chain (and index-build seconds) of the B200 path of `linear filter`.

Workload (BASELINE.json configs[2], the one the metric is quoted on): a synthetic 3.1-Gbase, 24-contig genome
with planted repeats, ONT-like simulated reads (lognormal mean 20 kb, ~10 % error 4:3:3 sub:ins:del, 20 % of
reads with one planted ins/del/inv/dup SV, 50 % reverse strand). One "step" = one batch of reads through the
whole hot path: reverse complement + 2x read features + seeding + anchor filter/sort + chaining + window
extension + block chaining = cords (what Mapper::p_calRecords does per read before mapGaps).

  value  reads/s, whole job, inputs already resident in HBM           (lnr_apxmap_batch_device)
  e2e    reads/s through the C-ABI call with pinned HOST buffers, H2D of the 2-bit packed bases and D2H of the
         cords inside the timed region                                (lnr_apxmap_batch_packed)
  e2e_dna5   the same through the 1-byte-per-base call the SeqAn shim makes (lnr_apxmap_batch); at N >= 4 this one
         is bound by the host's upload ceiling (e2e_dna5.h2d_ceiling), not by the GPUs
  roofline      dominant kernel of the step: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  cpu_baseline  the reference's own CPU code (oracle/_ref, or the oracle port) on a bounded sample of the batch
  --impl reference   times that CPU implementation alone, same config / metric

N > 1 (torchrun): reads shard across ranks with no data-path collective (weak scaling, fixed reads per GPU); the
index is replicated. Timing: barrier + cuda sync both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_BASES = int(os.environ.get("LNR_BENCH_GENOME", 3_100_000_000))
N_CONTIGS = 24
INDEX_TYPE = int(os.environ.get("LNR_BENCH_INDEX", 1))     # 2: HIndex (-i 2), a side measurement (BASELINE configs[3]); N = 1 only
READ_PROFILE = os.environ.get("LNR_BENCH_PROFILE", "ont")   # "hifi": 15 kb reads at 1 % error, a side measurement
THREADS_SEM = 16          # the reference's code default -t (base.cpp:26-54); semantic for the index
SEED_COUNT_WRITE = 4      # bytes k_seed_count writes per sample: its match count (+ 4 B per match entry, below)
SEED_FILL_READ = 4        # bytes k_seed_fill reads per sample: the count (+ 4 B per match entry, below)
TRAFFIC_FILE = "r2_ncu_traffic.json"   # dram bytes per read of each kernel from the last `ncu --set full` capture
METRIC = "reads/sec apx-map+chain at 1/2/4/8 B200 (3.1-Gbase synth); index build sec"   # BASELINE.json's metric, verbatim
try:
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
except Exception:  # noqa: BLE001
    pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-reads", type=int, default=int(os.environ.get("LNR_BENCH_BATCH", 65536)))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("LNR_BENCH_CPU_SAMPLE", 2048)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("LNR_BENCH_STREAMS", 4)),
                    help="concurrent host threads, each with its own lnr_ctx (the reference calls p_calRecords from -t threads)")
    return ap.parse_args()


def contig_lengths():
    from linear_b200 import datagen
    return datagen.contig_lengths(GENOME_BASES, N_CONTIGS, seed=31)


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs, generated on the GPU (torch is plumbing here: device memory + RNG)
# ---------------------------------------------------------------------------------------------------------
def gen_genome(torch, dev, lens, seed=1234):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    total = int(sum(lens))
    genome = torch.empty(total, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for s in range(0, total, step):   # chunked: randint materialises int64 internally
        n = min(step, total - s)
        genome[s:s + n] = torch.randint(0, 4, (n,), generator=g, device=dev, dtype=torch.uint8)
    # planted interspersed repeat families: 8 families x 2000 copies x 300 bp, ~12 % divergence, and 2000 tandem arrays
    for fam in range(8):
        base = torch.randint(0, 4, (300,), generator=g, device=dev, dtype=torch.uint8)
        pos = torch.randint(0, total - 400, (2000,), generator=g, device=dev)
        idx = (pos[:, None] + torch.arange(300, device=dev)[None, :]).reshape(-1)
        copies = base.repeat(2000)
        mut = torch.rand(copies.shape, generator=g, device=dev) < 0.12
        copies = torch.where(mut, torch.randint(0, 4, copies.shape, generator=g, device=dev, dtype=torch.uint8), copies)
        genome[idx] = copies
    unit_len = 40
    pos = torch.randint(0, total - 2000, (2000,), generator=g, device=dev)
    units = torch.randint(0, 4, (2000, unit_len), generator=g, device=dev, dtype=torch.uint8)
    arr = units.repeat(1, 20)
    idx = (pos[:, None] + torch.arange(unit_len * 20, device=dev)[None, :]).reshape(-1)
    genome[idx] = arr.reshape(-1)
    return genome


def gen_reads(torch, dev, genome, lens, n_reads, seed, mean=20000, sigma=0.5, err=0.10, mix=(4, 3, 3), sv_frac=0.2):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mu = float(np.log(mean) - sigma * sigma / 2)
    L = torch.exp(torch.randn(n_reads, generator=g, device=dev) * sigma + mu).long().clamp_(2000, 150000)
    off = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(L, 0)
    total = int(off[-1].item())
    rid = torch.repeat_interleave(torch.arange(n_reads, device=dev), L)
    pos = torch.arange(total, device=dev) - off[rid]
    tot = float(sum(mix))
    p_sub, p_ins, p_del = (err * m / tot for m in mix)
    u = torch.rand(total, generator=g, device=dev)
    is_ins = u < p_ins
    is_del = (u >= p_ins) & (u < p_ins + p_del)
    is_sub = (u >= p_ins + p_del) & (u < err)
    step = torch.ones(total, dtype=torch.int64, device=dev)
    step[is_ins] = 0
    step[is_del] = 2
    # one planted SV in sv_frac of the reads
    has_sv = (torch.rand(n_reads, generator=g, device=dev) < sv_frac) & (L > 9000)
    kind = torch.randint(0, 4, (n_reads,), generator=g, device=dev)           # 0 ins, 1 del, 2 inv, 3 dup
    p = (2000 + torch.rand(n_reads, generator=g, device=dev) * (L - 7000).clamp_min(1)).long()
    r01 = torch.rand(n_reads, generator=g, device=dev)
    lo = torch.tensor([100, 100, 500, 300], device=dev)[kind]
    hi = torch.tensor([2000, 3000, 3000, 2000], device=dev)[kind]
    m = (lo + r01 * (hi - lo)).long()
    sv_r, p_r, m_r, k_r = has_sv[rid], p[rid], m[rid], kind[rid]
    in_sv = sv_r & (pos >= p_r) & (pos < p_r + m_r)
    ins_reg = in_sv & (k_r == 0)
    inv_reg = in_sv & (k_r == 2)
    is_ins = is_ins | ins_reg
    step[ins_reg] = 0
    step[inv_reg] = 1
    at_p = sv_r & (pos == p_r)
    step = torch.where(at_p & (k_r == 1), step + m_r, step)
    step = torch.where(at_p & (k_r == 3), step - m_r, step)
    step[pos == 0] = 0
    cs = torch.cumsum(step, 0)
    rel = cs - cs[off[:-1]][rid]
    # template start: uniform over the genome, kept inside one contig
    coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    coff_t = torch.tensor(coff, device=dev)
    gpos = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * float(coff[-1])).long()
    ci = (torch.searchsorted(coff_t, gpos, right=True) - 1).clamp_(0, len(lens) - 1)
    span = (L.double() * 1.35).long() + 8000
    cstart, cend = coff_t[ci], coff_t[ci + 1]
    start = torch.minimum(gpos, cend - span).clamp_min(0)
    start = torch.maximum(start, cstart + 3000)
    tpos = start[rid] + rel
    t_at_p = tpos[(off[:-1] + p).clamp_max(total - 1)]
    tpos = torch.where(inv_reg, 2 * t_at_p[rid] + m_r - 1 - tpos, tpos)
    tpos = torch.minimum(torch.maximum(tpos, cstart[rid]), cend[rid] - 1)
    b = genome[tpos]
    b = torch.where(inv_reg, 3 - b, b)
    rnd = torch.randint(0, 4, (total,), generator=g, device=dev, dtype=torch.uint8)
    b = torch.where(is_sub & ~inv_reg, (b + 1 + rnd % 3) % 4, b)
    b = torch.where(is_ins, rnd, b)
    rev = torch.rand(n_reads, generator=g, device=dev) < 0.5
    src = torch.where(rev[rid], off[rid] + (L[rid] - 1 - pos), off[rid] + pos)
    out = b[src]
    out = torch.where(rev[rid], 3 - out, out).to(torch.uint8).contiguous()
    return out, off.cpu().numpy().astype(np.uint64)


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU while the timed regions run. NVML (the library nvidia-smi reads from) when
    it is importable -- forking nvidia-smi five times a second from a process with GBs of pinned memory, on every rank,
    perturbs what it observes -- else nvidia-smi itself."""

    def __init__(self, gpu_index, pci_bus_id=None):
        super().__init__(daemon=True)
        self.gpu, self.pci, self.stop_flag, self.rows, self.source = gpu_index, pci_bus_id, False, [], "nvidia-smi"

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByPciBusId(self.pci.encode() if isinstance(self.pci, str) else self.pci) if self.pci else \
            nv.nvmlDeviceGetHandleByIndex(self.gpu)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (0x8, 0x40, 0x20, 0x4)   # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        self.source = "nvml"
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:  # noqa: BLE001
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max(int(r[1]) for r in self.rows if r[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows), "source": self.source}


def bind_to_gpu_numa(pci_bus_id):
    """Run this rank (and so first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU hangs off: with N
    ranks on one box, uploads from a remote socket's memory cross the inter-socket link and stop scaling at 4-8 GPUs.
    Returns the node, or None when sysfs does not tell (numa_node -1: single-node box or a container that hides it)."""
    try:
        dom_bus = pci_bus_id.lower()
        node = int(open(f"/sys/bus/pci/devices/{dom_bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001
        return None


def cpu_reference(contigs_host, bases, offs, n_sample, cores, reps=1):
    """The reference's own CPU path on a bounded sample: oracle/_ref (kind 'reference') if it travelled, else the
    oracle port. Returns (dict(value reads/s, index_build_s, kind, cores, sample), checker, cords, cords_off): the
    checker and its cords are what the GPU output is compared with (the `parity` object of the line)."""
    from cpu_checkers import Oracle, RefImpl, have_ref
    kind = "reference" if have_ref() else "port"
    cls = RefImpl if have_ref() else Oracle
    t0 = time.time()
    chk = cls(contigs_host, threads=THREADS_SEM, preset=1, index_type=INDEX_TYPE)
    t_index = time.time() - t0
    n = min(n_sample, len(offs) - 1)
    sb = bases[: int(offs[n])]
    so = offs[: n + 1].copy()
    best = None
    for _ in range(max(reps, 1)):
        t0 = time.time()
        cords, coff = chk.map_batch(sb, so, map_threads=cores)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.time()
    chk.map_batch(sb, so, map_threads=min(4, cores))
    dt4 = time.time() - t0
    info = {"value": n / best, "unit": "reads/s", "cores": cores, "kind": kind, "index_build_s": round(t_index, 2),
            "index_threads": THREADS_SEM, "value_t4": n / dt4,
            "sample": f"first {n} reads of the rank-0 batch ({int(so[-1])} bases) through apxMap with {cores} OpenMP threads "
                      f"(value) and with 4 (value_t4, BASELINE configs[0] says -t 4); genome features + DIndex built by the same "
                      f"code at -t {THREADS_SEM} in {t_index:.1f} s"}
    return info, chk, cords, coff


def check_parity(torch, dev, chk, index, ref_cords, ref_coff, gpu_cords, gpu_coff):
    """Bit-exact comparison of the GPU results with the reference on the bench's own 3.1-Gbase inputs: the cords of the
    sampled reads, and the whole DIndex (dir + hs) against the arrays the reference built on the host."""
    n = len(ref_coff) - 1
    g_off = np.asarray(gpu_coff[: n + 1], dtype=np.uint64)
    cords_equal = bool(np.array_equal(g_off, np.asarray(ref_coff, dtype=np.uint64)) and
                       np.array_equal(np.asarray(gpu_cords[: int(g_off[-1])], dtype=np.uint64), np.asarray(ref_cords, dtype=np.uint64)))
    n_diff = 0
    if not cords_equal:
        for r in range(n):
            a = gpu_cords[int(g_off[r]):int(g_off[r + 1])]
            b = ref_cords[int(ref_coff[r]):int(ref_coff[r + 1])]
            if len(a) != len(b) or not np.array_equal(a, b):
                n_diff += 1
    out = {"reads": n, "cords": int(ref_coff[-1]), "cords_equal": cords_equal, "reads_differing": n_diff}
    try:
        if INDEX_TYPE != 1:
            raise RuntimeError("HIndex run: only the cords are compared")
        rdir, rhs = chk.dindex_views()
        d_dev, hs_dev = index.export_device(torch, dev)
        ok = len(rhs) == hs_dev.numel() and len(rdir) == d_dev.numel()
        if ok:
            ok = bool(torch.equal(torch.from_numpy(rdir).to(dev), d_dev))
        step = 1 << 26
        hs_i64 = rhs.view(np.int64)
        for a in range(0, len(rhs), step):
            if not ok:
                break
            ok = bool(torch.equal(torch.from_numpy(hs_i64[a:a + step]).to(dev), hs_dev[a:a + step]))
        out.update({"dindex_equal": ok, "n_hs": int(len(rhs))})
        del d_dev, hs_dev
    except Exception as e:  # noqa: BLE001
        out.update({"dindex_equal": None, "dindex_error": repr(e)})
    return out


def measure_ingest(lb, ctx, torch, dev, bases_np, offs, n=8192, wrap=80):
    """SURVEY 8(f) row 3, read ingest: FASTA text (80-column lines) of the first n reads of the batch -> Dna5 ordinals +
    offsets on the device. GB/s of text, with the text resident in HBM and from host memory (upload inside), next to a
    one-core numpy statement of the same parse (a port; the reference's own seqan reader is quoted at 8358 reads/thread/s
    in its README, about 0.17 GB/s)."""
    import ctypes as C
    n = min(n, len(offs) - 1)
    alpha = np.frombuffer(b"ACGTN", np.uint8)
    parts = []
    for k in range(n):
        s = alpha[bases_np[int(offs[k]):int(offs[k + 1])]].tobytes()
        parts.append(b">read%d\n" % k)
        parts.extend(s[i:i + wrap] + b"\n" for i in range(0, len(s), wrap))
    text = b"".join(parts)
    vp = C.c_void_p

    t_pin = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    t_dev = t_pin.to(dev)

    def parse_host():   # text in pinned host memory: the upload is inside
        h = vp()
        ctx.check(ctx.lib.lnr_reads_parse(ctx.h, C.cast(t_pin.data_ptr(), C.c_char_p), len(text), 0, C.byref(h)))
        ctx.lib.lnr_reads_destroy(h)

    def parse_dev():
        h = vp()
        ctx.check(ctx.lib.lnr_reads_parse_device(ctx.h, vp(t_dev.data_ptr()), len(text), text[0], 0, C.byref(h)))
        ctx.lib.lnr_reads_destroy(h)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.time() - t0) / reps

    th, td = timed(parse_host), timed(parse_dev)
    # one-core numpy port: strip header lines and newlines, table lookup
    t0 = time.time()
    a = np.frombuffer(text, np.uint8)
    nl = np.flatnonzero(a == 10)
    starts = np.concatenate([[0], nl[:-1] + 1])
    is_hdr = a[starts] == ord(">")
    keep = np.ones(len(a), bool)
    keep[nl] = False
    hb, he = starts[is_hdr], nl[is_hdr]
    for b_, e_ in zip(hb, he):
        keep[b_:e_] = False
    lut = np.full(256, 4, np.uint8)
    for ch, v in (("Aa", 0), ("Cc", 1), ("Gg", 2), ("TtUu", 3)):
        for c in ch:
            lut[ord(c)] = v
    out = lut[a[keep]]
    t_cpu = time.time() - t0
    assert len(out) == int(offs[n] - offs[0])
    gb = len(text) / 1e9
    return {"text_bytes": len(text), "reads": n, "line_width": wrap, "device_resident_GBps": gb / td, "from_host_GBps": gb / th,
            "cpu_port_GBps": gb / t_cpu, "cpu_port": "numpy, 1 core", "ms_device_resident": 1000 * td}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    import torch
    lens = contig_lengths()
    cores = os.cpu_count() or 1
    config = {"workload": f"BASELINE configs[2] (apx map + chaining; index build = configs[1]): {GENOME_BASES / 1e9:.2f}-Gbase synthetic genome "
                          f"({N_CONTIGS} contigs, planted repeats) + simulated ONT-like reads (lognormal mean 20 kb, 10% error, 20% planted "
                          f"ins/del/inv/dup SVs), one step = one batch of {args.batch_reads} reads per GPU",
              "index": "DIndex (-i 1)" if INDEX_TYPE == 1 else "HIndex (-i 2), side measurement", "features": "2-mer/48 (-f 2)", "threads_sem": THREADS_SEM, "preset": 1,
              "batch_reads_per_gpu": args.batch_reads, "parallelism": f"reads sharded x{world} (no collective); index built by minimizer range x{world} + one NCCL all-gather",
              "l2_policy": "inputs larger than L2 (batch bases + index >> 126 MB)",
              "host_threads": args.streams}
    if READ_PROFILE == "hifi":
        config["workload"] = (f"side measurement, not the headline: {GENOME_BASES / 1e9:.2f}-Gbase synthetic genome + HiFi-like reads (mean 15 kb, "
                              f"1% error, no SVs), one step = one batch of {args.batch_reads} reads per GPU")

    if args.impl == "reference":
        if rank != 0:
            return
        if not torch.cuda.is_available():
            # no GPU needed for this arm, but the synthetic genome recipe is the GPU one; fall back to numpy
            rng = np.random.default_rng(1234)
            contigs = [rng.integers(0, 4, size=l, dtype=np.uint8) for l in lens]
            from linear_b200 import datagen
            rs = datagen.simulate_reads(77, contigs, args.cpu_sample, mean_len=20000, err=0.10, mix=(4, 3, 3), sv_frac=0.2, lognormal=True)
            bases, offs = rs.bases, rs.offsets
        else:
            dev = torch.device("cuda", local_rank)
            genome = gen_genome(torch, dev, lens)
            # the B200 arm's rank-0 batch, generated the same way (same seed, same size), so that the sample is a prefix of it
            bases_t, offs = gen_reads(torch, dev, genome, lens, args.batch_reads, seed=1000)
            keep = int(offs[min(args.cpu_sample, len(offs) - 1)])
            bases = bases_t[:keep].cpu().numpy()
            gh = genome.cpu().numpy()
            coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            contigs = [gh[coff[i]:coff[i + 1]] for i in range(len(lens))]
            del genome, bases_t
            torch.cuda.empty_cache()
        from cpu_checkers import Oracle, RefImpl, have_ref
        cls = RefImpl if have_ref() else Oracle
        t0 = time.time()
        chk = cls(contigs, threads=THREADS_SEM, preset=1)
        t_index = time.time() - t0
        # one step = the whole sample (the first --cpu-sample reads of the very batch the B200 arm maps), every step the
        # same reads -- exactly as the B200 arm re-maps one batch per step
        n = min(args.cpu_sample, len(offs) - 1)
        so = offs[: n + 1].copy()
        sb = bases[: int(offs[n])]
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.time()
            chk.map_batch(sb, so, map_threads=cores)
            if s >= args.warmup:
                times.append(time.time() - t0)
        tot = sum(times)
        val = n * len(times) / tot
        t0 = time.time()
        chk.map_batch(sb, so, map_threads=min(4, cores))
        val_t4 = n / (time.time() - t0)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000 * tot / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "index_build_s": round(t_index, 2), "reads_per_step": n, "value_t4": val_t4,
                "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": "reference" if have_ref() else "port",
                                 "value_t4": val_t4,
                                 "sample": f"the first {n} reads of the B200 arm's rank-0 batch ({int(so[-1])} bases), all of them every step, through "
                                           f"the reference's apxMap with {cores} OpenMP threads (value_t4: 4 threads; index + features built "
                                           f"at -t {THREADS_SEM} in {t_index:.1f} s)"},
                "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------------------------------------- B200 arm
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its
    # version banner and NCCL_DEBUG output to stdout from C) and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import linear_b200 as lb
    lb.load_library()
    pci = None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        pci = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:  # noqa: BLE001
        pass
    numa_node = bind_to_gpu_numa(pci) if (pci and world > 1 and not os.environ.get("LNR_BENCH_NO_NUMA")) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    genome = gen_genome(torch, dev, lens)
    if world > 1:
        # one genome for the whole job: the planted repeats are written with a scatter whose overlapping writes land in an
        # order that is not reproducible from GPU to GPU
        dist.broadcast(genome, 0)
    ctx = lb.Context(local_rank)
    lens64 = [int(x) for x in lens]
    comm = None
    if world > 1:
        def exchange_id(idb):   # rank 0's NCCL unique id to everybody (the library's communicator is its own)
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if idb is not None:
                t.copy_(torch.frombuffer(bytearray(idb), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())
        comm = lb.Comm(ctx, rank, world, exchange_id)
    # ---- index build, inputs resident in HBM (genome features + DIndex = createFeatures + createIndexDynamic)
    def build_device_resident():
        g_ = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=lens64)
        f_ = lb.create_features(ctx, g_, 2, THREADS_SEM)
        if world == 1:
            i_ = lb.create_index(ctx, g_, INDEX_TYPE, THREADS_SEM)
        else:
            # hash-range sharded build inside the C ABI: every rank builds 2^26/N buckets straight into its slice of the final
            # arrays, one grouped NCCL exchange over NVLink completes them (lnr_index_build_sharded)
            i_ = lb.create_index_sharded(ctx, g_, comm, 1, THREADS_SEM)
        torch.cuda.synchronize()
        return g_, f_, i_

    # the first call also pays the process's one-time costs (kernel images, first multi-GB cudaMalloc); it is reported
    # as seconds_first_call and the build is timed on the second call
    barrier()
    t0 = time.time()
    gen, feats, index = build_device_resident()
    t_index_first = max_over_ranks(time.time() - t0)
    index.close(); feats.close(); gen.close()
    ctx.set_profiling(True)
    barrier()
    t0 = time.time()
    gen, feats, index = build_device_resident()
    t_index = max_over_ranks(time.time() - t0)
    idx_kernels = ctx.kernel_times()
    n_hs = index.n_hs
    ctx.reset_kernel_times()
    sharded_equal = None
    if world > 1:
        # the index every rank assembled over NCCL == the one a single GPU builds from the same genome
        ix1 = lb.create_index(ctx, gen, 1, THREADS_SEM)
        d_a, h_a = index.export_device(torch, dev)
        d_b, h_b = ix1.export_device(torch, dev)
        ok = torch.tensor([int(d_a.shape == d_b.shape and h_a.shape == h_b.shape and bool(torch.equal(d_a, d_b)) and bool(torch.equal(h_a, h_b)))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_equal = bool(ok.item())
        del d_a, h_a, d_b, h_b
        ix1.close()
        ctx.reset_kernel_times()
    # ---- index build end to end from host memory (rank 0, N = 1 only): upload + features + index
    t_index_e2e = None
    contigs_host = None
    if world == 1:
        gh = torch.empty(genome.shape, dtype=torch.uint8, pin_memory=True)
        gh.copy_(genome)
        ghn = gh.numpy()
        coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        contigs_host = [ghn[coff[i]:coff[i + 1]] for i in range(len(lens))]
        index.close(); feats.close(); gen.close()
        torch.cuda.synchronize()
        t0 = time.time()
        gen = lb.Genome(ctx, contigs_host)
        feats = lb.create_features(ctx, gen, 2, THREADS_SEM)
        index = lb.create_index(ctx, gen, INDEX_TYPE, THREADS_SEM)
        torch.cuda.synchronize()
        t_index_e2e = time.time() - t0
        ctx.reset_kernel_times()
    # ---- reads of this rank
    if READ_PROFILE == "hifi":   # side measurement (BASELINE configs[0]/[3] flavour), not the headline workload
        bases_t, offs = gen_reads(torch, dev, genome, lens, args.batch_reads, seed=1000 + rank, mean=15000, sigma=0.2, err=0.01, mix=(1, 1, 1), sv_frac=0.0)
    else:
        bases_t, offs = gen_reads(torch, dev, genome, lens, args.batch_reads, seed=1000 + rank)
    del genome
    torch.cuda.empty_cache()
    n_reads = len(offs) - 1
    total_bases = int(offs[-1])
    cap = total_bases // 16 + 64 * n_reads + 1024
    import ctypes as C
    from linear_b200.api import Params, u64p
    prm = Params(preset=1, feature_type=2)
    offs_c = offs.ctypes.data_as(u64p)
    bases_pin = torch.empty(total_bases, dtype=torch.uint8, pin_memory=True)
    bases_pin.copy_(bases_t)
    bases_np = bases_pin.numpy()
    # the same batch 2-bit packed (lnr_pack_dna5, the converter a shim calls once per read block), in pinned memory
    packed_pin = torch.empty((total_bases + 3) // 4, dtype=torch.uint8, pin_memory=True)
    packed_np = packed_pin.numpy()
    has_n = C.c_int()
    t0 = time.time()
    lb.load_library().lnr_pack_dna5(C.c_void_p(bases_np.ctypes.data), total_bases, C.c_void_p(packed_np.ctypes.data), None, C.byref(has_n))
    t_pack = time.time() - t0
    assert not has_n.value
    n_str = max(1, args.streams)
    ctxs = [ctx] + [lb.Context(local_rank) for _ in range(n_str - 1)]

    class Stream:
        """one host thread's resources: its own lnr_ctx (CUDA stream + workspace) and output buffers"""
        def __init__(self, c):
            self.ctx = c
            self.cords_dev = torch.empty(cap, dtype=torch.int64, device=dev)
            self.coff_dev = torch.empty(n_reads + 4200, dtype=torch.int64, device=dev)
            self.ntot = C.c_uint64()
            self.cords_pin = torch.empty(cap, dtype=torch.int64, pin_memory=True)
            self.cords_np = self.cords_pin.numpy().view(np.uint64)
            self.coff_np = np.zeros(n_reads + 1, dtype=np.uint64)
            self.last = None

        def step_device(self):
            c = self.ctx
            c.check(c.lib.lnr_apxmap_batch_device(c.h, index.h, feats.h, C.byref(prm), n_reads, C.c_void_p(bases_t.data_ptr()), offs_c,
                                                  C.c_void_p(self.cords_dev.data_ptr()), C.c_void_p(self.coff_dev.data_ptr()), cap,
                                                  C.byref(self.ntot)))

        def step_host(self):
            self.last = lb.apx_map_batch(self.ctx, index, feats, bases_np, offs, preset=1, cords_out=self.cords_np, cords_off_out=self.coff_np)

        def step_packed(self):
            self.last_packed = lb.apx_map_batch_packed(self.ctx, index, feats, packed_np, None, offs, preset=1, cords_out=self.cords_np,
                                                       cords_off_out=self.coff_np)

    streams = [Stream(c) for c in ctxs]

    def run_steps(kind, k):
        """k steps in total, dealt round-robin to the host threads; returns wall seconds (barrier + sync both sides)"""
        per = [k // n_str + (1 if i < k % n_str else 0) for i in range(n_str)]

        errors = []

        def work(st, n):
            try:
                for _ in range(n):
                    getattr(st, kind)()
            except BaseException as e:  # noqa: BLE001  -- a failed host thread must fail the run, not shorten it
                errors.append(e)
        barrier()
        t0 = time.time()
        if n_str == 1:
            work(streams[0], per[0])
        else:
            th = [threading.Thread(target=work, args=(streams[i], per[i])) for i in range(n_str) if per[i]]
            for t in th:
                t.start()
            for t in th:
                t.join()
        torch.cuda.synchronize()
        if errors:
            raise errors[0]
        barrier()
        return time.time() - t0

    # ---- timed region: device-resident inputs
    for c in ctxs:
        c.set_profiling(False)
    sampler = ClockSampler(local_rank, ("0000" + pci) if pci else None)
    if rank == 0:          # one sampler per job: the line reports rank 0's GPU
        sampler.start()
    run_steps("step_device", args.warmup * n_str)
    for c in ctxs:
        c.set_profiling(True)
        c.reset_kernel_times()
    dt = max_over_ranks(run_steps("step_device", args.steps))
    kt = {}
    for c in ctxs:
        for k, v in c.kernel_times().items():
            a = kt.get(k, (0.0, 0))
            kt[k] = (a[0] + v[0], a[1] + v[1])
    counters = ctx.counters()
    diag = ctx.diag()
    stage_cycles = ctx.stage_cycles()
    n_cords = int(streams[0].ntot.value)
    value = world * n_reads * args.steps / dt
    # ---- end to end through the host-buffer call
    for c in ctxs:
        c.set_profiling(False)
    run_steps("step_host", n_str)
    dt_e2e = max_over_ranks(run_steps("step_host", args.steps))
    e2e_value = world * n_reads * args.steps / dt_e2e
    c_host, coff_host = streams[0].last
    c_host = c_host.copy()
    coff_host = coff_host.copy()
    d2h = int(len(c_host) * 8 + (n_reads + 1) * 8)
    # ---- the same through the 2-bit packed call (a quarter of the upload)
    run_steps("step_packed", n_str)
    dt_packed = max_over_ranks(run_steps("step_packed", args.steps))
    packed_value = world * n_reads * args.steps / dt_packed
    packed_equal = bool(np.array_equal(streams[0].last_packed[1], coff_host) and np.array_equal(streams[0].last_packed[0], c_host))
    # ---- the reference's own call pattern (Mapper::p_calRecords, mapper.cpp:404-473): every host thread maps blocks of 64
    # reads, one lnr_apxmap_batch per block, concurrently on its own context -- no large batches
    small = None
    if world == 1 and not os.environ.get("LNR_BENCH_NO_SMALL"):
        blk, n_blk = 64, 24
        def small_work(st, t_id, out):
            for b in range(n_blk):
                i0 = (t_id * n_blk + b) * blk
                so = (offs[i0:i0 + blk + 1] - offs[i0]).astype(np.uint64)
                out.append(lb.apx_map_batch(st.ctx, index, feats, bases_np[int(offs[i0]):int(offs[i0 + blk])], so, preset=1)[1][-1])
        for rep in range(2):                      # first round warms the contexts' buffers for this shape
            outs = [[] for _ in range(n_str)]
            th = [threading.Thread(target=small_work, args=(streams[i], i, outs[i])) for i in range(n_str)]
            torch.cuda.synchronize()
            t0 = time.time()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt_small = time.time() - t0
        small = {"reads_per_s": n_blk * n_str * blk / dt_small, "block_reads": blk, "blocks": n_blk * n_str, "host_threads": n_str,
                 "ms_per_block": 1000 * dt_small / n_blk,
                 "what": "each host thread calls lnr_apxmap_batch on blocks of 64 reads (host buffers in and out), the way "
                         "Mapper::p_calRecords is driven with -b 1"}
        # the same pattern from as many calling threads as the reference would run with -t 16: a block keeps 64 warps busy, so
        # the device has room for many blocks in flight and the throughput follows the number of callers
        try:
            n_thr = int(os.environ.get("LNR_BENCH_SMALL_THREADS", 16))
            class _C:
                pass
            extra = []
            for _ in range(n_thr):
                o = _C(); o.ctx = lb.Context(local_rank); extra.append(o)
            for rep in range(2):
                outs = [[] for _ in range(n_thr)]
                th = [threading.Thread(target=small_work, args=(extra[i], i, outs[i])) for i in range(n_thr)]
                torch.cuda.synchronize()
                t0 = time.time()
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                dt_small = time.time() - t0
            small["more_threads"] = {"host_threads": n_thr, "reads_per_s": n_blk * n_thr * blk / dt_small, "blocks": n_blk * n_thr,
                                     "ms_per_block": 1000 * dt_small / n_blk}
            for o in extra:
                o.ctx.close()
        except Exception as e:  # noqa: BLE001
            small["more_threads"] = {"error": repr(e)}
    # ---- host-side ceiling: what this box delivers when every rank only uploads its batch (pinned host -> HBM), all ranks at once
    up = torch.empty(total_bases, dtype=torch.uint8, device=dev)
    up.copy_(bases_pin, non_blocking=True)
    barrier()
    t0 = time.time()
    for _ in range(4):
        up.copy_(bases_pin, non_blocking=True)
    torch.cuda.synchronize()
    dt_up = max_over_ranks((time.time() - t0) / 4)
    barrier()
    del up
    h2d_ceiling = {"GBps_per_rank": total_bases / dt_up / 1e9, "GBps_all_ranks": world * total_bases / dt_up / 1e9,
                   "reads_per_s_if_upload_only": world * n_reads / dt_up, "numa_node_bound": numa_node,
                   "what": "every rank uploads its batch (1 B/base, pinned) at the same time, nothing else running"}
    # ---- per-kernel times for the roofline: ONE host thread, so that an event pair brackets one kernel and not the other
    # threads' kernels interleaved with it (the 4-thread numbers above stay `value` / `e2e`)
    ctx.set_profiling(True)
    ctx.reset_kernel_times()
    single_steps = max(2, min(4, args.steps))
    streams[0].step_device()
    ctx.reset_kernel_times()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(single_steps):
        streams[0].step_device()
    torch.cuda.synchronize()
    dt_single = (time.time() - t0) / single_steps
    kt1 = ctx.kernel_times()
    ctx.set_profiling(False)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    # every rank leaves the process group at the same point; rank 0 alone goes on to the report
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline. Algorithmic bytes per launch, one launch = one batch; DESIGN.md section 4 states every term.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    S, H, A, Hits, W, Cc = (counters[k] for k in ("S_seeds", "H_records_scanned", "A_raw_anchors", "Hits", "W_windows", "C_cords"))
    nf_bytes = 2 * 12 * (total_bases // 16)
    alg = {
        "k_feat_reads": total_bases + nf_bytes,            # bases in, both strands' int96 features out
        "k_seed_count": total_bases + 32 * S + H + SEED_COUNT_WRITE * S + 4 * A,   # bases, one 32-B lookup sector per seed, 1-byte Y keys, counts + match entries out
        "k_seed_fill": SEED_FILL_READ * S + 4 * A + 8 * A + 8 * A,   # counts in, match entries in, matched hs records in, anchors out
        "k_hits_sort": 8 * A,
        "k_hits_chain": 24 * Hits,
        "k_hits_blocks": 24 * Hits,
        "k_map_extend": 144 * W + 8 * Cc,
        "k_map_finish": 16 * Cc,
    }
    # the whole step: SURVEY 8(d)'s per-read sum, every term once
    step_alg = total_bases + 8 * S + 8 * H + 8 * A + nf_bytes + 48 * Hits + 144 * W + 16 * Cc
    per_kernel = {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1]} for k, v in kt.items()}
    per_kernel_single = {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1]} for k, v in kt1.items()}
    step_ms_kernels = sum(v[0] for v in kt1.values()) / single_steps
    cand = {k: v for k, v in kt1.items() if k in alg}
    dom = max(cand.items(), key=lambda kv: kv[1][0])[0] if cand else None
    roof = None
    if dom:
        ms = kt1[dom][0] / max(kt1[dom][1], 1)
        ab = alg[dom]
        ach = ab / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        traffic = None
        try:   # dram__bytes_read+write of one `ncu --set full` capture (profiles/), scaled from its 32768-read batch
            tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
            if dom in tj:
                traffic = tj[dom]["dram_bytes_per_read"] * n_reads
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": ab, "ms_per_launch": ms,
                "timing": f"CUDA events on the launching stream, {single_steps} steps with ONE host thread after the timed region",
                "ms_per_launch_4_threads": kt[dom][0] / max(kt[dom][1], 1) if dom in kt else None,
                "share_of_step": (kt1[dom][0] / single_steps) / step_ms_kernels if step_ms_kernels else None,
                "whole_step": {"algorithmic_bytes": step_alg, "ms_per_step": 1000 * dt / args.steps,
                               "achieved_GBps": step_alg / (dt / args.steps) / 1e9, "frac": step_alg / (dt / args.steps) / 1e9 / peak,
                               "ms_per_step_one_thread": 1000 * dt_single},
                "per_kernel": {k: {"ms_per_launch": round(kt1[k][0] / max(kt1[k][1], 1), 4), "algorithmic_bytes": int(alg[k]),
                                   "algorithmic_GBps": round(alg[k] / (kt1[k][0] / max(kt1[k][1], 1) * 1e-3) / 1e9, 1),
                                   "frac": round(alg[k] / (kt1[k][0] / max(kt1[k][1], 1) * 1e-3) / 1e9 / peak, 4)}
                               for k in alg if k in kt1 and kt1[k][0] > 0}}
    idx_alg = GENOME_BASES + 8 * n_hs + 4 * ((1 << 26) + 1) + GENOME_BASES + 12 * (GENOME_BASES // 16)
    index_info = {"seconds": round(t_index, 4), "seconds_first_call": round(t_index_first, 4), "seconds_e2e_from_host": None if t_index_e2e is None else round(t_index_e2e, 4),
                  "n_hs": n_hs, "algorithmic_bytes": idx_alg, "achieved_GBps": idx_alg / t_index / 1e9,
                  "frac_of_hbm_peak": idx_alg / t_index / 1e9 / peak,
                  "kernels_ms": {k: round(v[0], 3) for k, v in idx_kernels.items()}}
    if roof is not None:
        roof["index"] = index_info     # index-build seconds live inside a key the driver keeps
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline and contigs_host is not None:
        try:
            cpu, chk, r_cords, r_coff = cpu_reference(contigs_host, bases_np, offs, args.cpu_sample, cores)
            parity = check_parity(torch, dev, chk, index, r_cords, r_coff, c_host, coff_host)
            chk.close()
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "reads/s", "cores": cores, "kind": "unavailable", "sample": repr(e)}
    ingest = None
    if world == 1:
        try:
            ingest = measure_ingest(lb, ctx, torch, dev, bases_np, offs)
        except Exception as e:  # noqa: BLE001
            ingest = {"error": repr(e)}
    launches = sum(v[1] for v in kt.values())
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": config,
            # e2e = the packed call: north_star's 2-bit input format, a quarter of the upload, so the number is the device's and
            # not the host's PCIe / memory ceiling at N >= 4 ("next round" item 4); e2e_dna5 = the 1-byte call the
            # SeqAn shim makes with the reference's own String<Dna5> buffers. Both return the same cords (checked every run).
            "e2e": {"value": packed_value, "unit": "reads/s", "h2d_bytes_per_step": (total_bases + 3) // 4 + (n_reads + 1) * 8,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1000 * dt_packed / args.steps,
                    "input": "2-bit packed bases + read offsets in pinned host memory (lnr_apxmap_batch_packed); cords + offsets copied back",
                    "cords_equal_to_dna5_call": packed_equal, "host_pack_seconds_one_core": round(t_pack, 3),
                    "h2d_GBps_per_rank": ((total_bases + 3) // 4) / (dt_packed / args.steps) / 1e9},
            "e2e_dna5": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": total_bases + (n_reads + 1) * 8, "d2h_bytes_per_step": d2h,
                         "ms_per_step": 1000 * dt_e2e / args.steps, "input": "Dna5, 1 byte/base (lnr_apxmap_batch)",
                         "h2d_GBps_per_rank": total_bases / (dt_e2e / args.steps) / 1e9, "h2d_ceiling": h2d_ceiling},
            "e2e_small_blocks": small,
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "parity": parity if parity is not None else ({"sharded_dindex_equals_single_gpu_build": sharded_equal} if sharded_equal is not None else None), "index_build": index_info,
            "clocks": sampler.summary(), "ingest": ingest, "kernels": per_kernel, "kernels_one_thread": per_kernel_single,
            "counters": counters, "fallback_paths_last_batch": diag, "stage_cycles_last_batch": stage_cycles, "cords_per_step": n_cords,
            "bases_per_step": total_bases}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if sharded_equal is False:
        sys.stderr.write("PARITY FAILURE: the sharded index differs from the single-GPU build\n")
        sys.exit(3)
    if parity is not None and (parity.get("cords_equal") is False or parity.get("dindex_equal") is False):
        sys.stderr.write("PARITY FAILURE vs the reference: %r\n" % (parity,))
        sys.exit(3)


if __name__ == "__main__":
    main()
