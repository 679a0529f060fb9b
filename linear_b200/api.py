"""Host-side Python binding of the C ABI in include/lnr_b200.h (ctypes; used by tests/ and bench.py).

The names follow the reference interface each call replaces:
  Genome            <- Mapper::loadGenomes / StringSet<String<Dna5>>                    (mapper.cpp:389)
  create_features   <- createFeatures(genomes, f2, feature_type, threads)               (pmpfinder.cpp:775)
  create_index      <- createIndexDynamic(genomes, index, 0, n, threads, false)         (index_util.cpp:2478)
  apx_map_batch     <- per read: _compltRvseStr + 2x createFeatures + apxMap            (mapper.cpp:438-447)

There is no CPU fallback: if the CUDA library is missing or no device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblnr_b200.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)

LNR_OK, LNR_E_CUDA, LNR_E_ARG, LNR_E_CAPACITY, LNR_E_UNSUPPORTED, LNR_E_LIMIT = 0, -1, -2, -3, -4, -5

EXPORTS = [
    "lnr_ctx_create", "lnr_ctx_destroy", "lnr_last_error", "lnr_ctx_set_profiling", "lnr_ctx_kernel_times",
    "lnr_ctx_reset_kernel_times", "lnr_genome_upload", "lnr_genome_from_device", "lnr_genome_destroy",
    "lnr_features_build", "lnr_features_count", "lnr_features_download", "lnr_features_destroy",
    "lnr_index_build", "lnr_index_build_shard", "lnr_index_export_dindex", "lnr_index_export_hindex", "lnr_index_export_dindex_device",
    "lnr_index_from_device", "lnr_index_save", "lnr_index_load", "lnr_index_destroy", "lnr_nccl_unique_id", "lnr_comm_create", "lnr_comm_from_nccl", "lnr_comm_destroy",
    "lnr_index_build_sharded", "lnr_hindex_shard_cuts", "lnr_apxmap_batch", "lnr_apxmap_batch_packed", "lnr_pack_dna5",
    "lnr_apxmap_batch_device", "lnr_cords_to_records", "lnr_last_batch_counters", "lnr_last_batch_diag", "lnr_last_batch_stage_cycles", "lnr_read_features", "lnr_selftest_sort",
    "lnr_reads_parse", "lnr_reads_parse_device", "lnr_reads_info", "lnr_reads_download", "lnr_reads_device", "lnr_reads_destroy", "lnr_apxmap_reads",
]


class LnrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lnr_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("preset", C.c_int), ("feature_type", C.c_int), ("no_chain", C.c_int), ("gdl_state", C.c_int), ("reserved", C.c_int * 4)]


class BamParms(C.Structure):
    _fields_ = [("window", C.c_uint32), ("reserved", C.c_uint32), ("thd_large_x", C.c_uint64), ("thd_di", C.c_int64), ("thd_x", C.c_int64)]


BAM_REC_DTYPE = np.dtype([("rid", np.int32), ("begin_pos", np.int32), ("flag", np.uint32), ("s1", np.int32), ("s2", np.int32), ("s3", np.int32),
                          ("cigar_begin", np.uint32), ("cigar_end", np.uint32)])


class DebugOut(C.Structure):
    _fields_ = [("raw_anchors", u64p), ("raw_anchors_cap", C.c_uint64), ("raw_anchors_off", u64p),
                ("hits", u64p), ("hits_cap", C.c_uint64), ("hits_off", u64p),
                ("cords1", u64p), ("cords1_cap", C.c_uint64), ("cords1_off", u64p)]


_lib = None


def load_library() -> C.CDLL:
    """Load liblnr_b200.so (built in-tree by __graft_entry__.build()). Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.lnr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.lnr_ctx_destroy.argtypes = [vp]
    lib.lnr_ctx_destroy.restype = None
    lib.lnr_last_error.argtypes = [vp]
    lib.lnr_last_error.restype = C.c_char_p
    lib.lnr_ctx_set_profiling.argtypes = [vp, C.c_int]
    lib.lnr_ctx_kernel_times.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), u64p]
    lib.lnr_ctx_reset_kernel_times.argtypes = [vp]
    lib.lnr_genome_upload.argtypes = [vp, C.c_uint32, C.POINTER(u8p), u64p, C.POINTER(vp)]
    lib.lnr_genome_from_device.argtypes = [vp, C.c_uint32, vp, u64p, C.POINTER(vp)]
    lib.lnr_genome_destroy.argtypes = [vp]
    lib.lnr_genome_destroy.restype = None
    lib.lnr_features_build.argtypes = [vp, vp, C.c_int, C.c_uint, C.POINTER(vp)]
    lib.lnr_features_count.argtypes = [vp, C.c_uint32, u64p]
    lib.lnr_features_download.argtypes = [vp, C.c_uint32, vp, C.c_uint64, u64p]
    lib.lnr_features_destroy.argtypes = [vp]
    lib.lnr_features_destroy.restype = None
    lib.lnr_index_build.argtypes = [vp, vp, C.c_int, C.c_uint, C.POINTER(vp)]
    lib.lnr_index_export_dindex.argtypes = [vp, i32p, u64p, C.c_uint64, u64p]
    lib.lnr_index_export_hindex.argtypes = [vp, u64p, C.c_uint64, u64p, u64p, C.c_uint64, u64p, u64p, u64p]
    lib.lnr_index_build_shard.argtypes = [vp, vp, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.POINTER(vp)]
    lib.lnr_index_export_dindex_device.argtypes = [vp, vp, vp, C.c_uint64]
    lib.lnr_index_from_device.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(vp)]
    lib.lnr_index_destroy.argtypes = [vp]
    lib.lnr_index_destroy.restype = None
    lib.lnr_index_save.argtypes = [vp, C.c_char_p]
    lib.lnr_index_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.lnr_nccl_unique_id.argtypes = [vp]
    lib.lnr_comm_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lnr_comm_from_nccl.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lnr_comm_destroy.argtypes = [vp]
    lib.lnr_comm_destroy.restype = None
    lib.lnr_index_build_sharded.argtypes = [vp, vp, C.c_int, C.c_uint, vp, C.POINTER(vp)]
    lib.lnr_hindex_shard_cuts.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
    lib.lnr_apxmap_batch.argtypes = [vp, vp, vp, C.POINTER(Params), C.c_uint32, vp, u64p, vp, u64p, C.c_uint64,
                                     C.POINTER(DebugOut)]
    lib.lnr_apxmap_batch_packed.argtypes = [vp, vp, vp, C.POINTER(Params), C.c_uint32, vp, vp, u64p, vp, u64p, C.c_uint64,
                                            C.POINTER(DebugOut)]
    lib.lnr_pack_dna5.argtypes = [vp, C.c_uint64, vp, vp, C.POINTER(C.c_int)]
    lib.lnr_apxmap_batch_device.argtypes = [vp, vp, vp, C.POINTER(Params), C.c_uint32, vp, u64p, vp, vp, C.c_uint64, u64p]
    lib.lnr_cords_to_records.argtypes = [vp, C.c_uint32, vp, u64p, u64p, C.POINTER(BamParms), vp, C.c_uint64, u64p, vp, C.c_uint64, u64p]
    lib.lnr_last_batch_counters.argtypes = [vp, u64p]
    lib.lnr_last_batch_stage_cycles.argtypes = [vp, u64p]
    lib.lnr_last_batch_diag.argtypes = [vp, u64p]
    lib.lnr_read_features.argtypes = [vp, u8p, C.c_uint64, C.c_int, vp, vp, C.c_uint64, u64p]
    lib.lnr_selftest_sort.argtypes = [vp, u64p, C.c_uint32]
    lib.lnr_reads_parse.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(vp)]
    lib.lnr_reads_parse_device.argtypes = [vp, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lnr_reads_info.argtypes = [vp, u64p, u64p]
    lib.lnr_reads_download.argtypes = [vp, u8p, u64p, u64p, C.POINTER(C.c_uint32)]
    lib.lnr_reads_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.lnr_reads_destroy.argtypes = [vp]
    lib.lnr_reads_destroy.restype = None
    lib.lnr_apxmap_reads.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint64, vp]
    _lib = lib
    return lib


STAGE_NAMES = ("binning", "sort_asc", "run_filter", "sort_x", "chain_dp", "traceback", "hit_blocks", "hit_window_filter",
               "window_extension", "clean_gaps", "cord_block_chaining", "n_tie_fallback", "n_reads", "max_read_cycles", "sum_read_cycles", "r15")
COUNTER_NAMES = ("S_seeds", "H_records_scanned", "A_raw_anchors", "Hits", "W_windows", "C_cords", "bases", "remap_tasks")


class Context:
    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.lnr_ctx_create(device, C.byref(h))
        if rc != 0:
            raise LnrError(rc, "lnr_ctx_create failed (no CUDA device?)")
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise LnrError(rc, (self.lib.lnr_last_error(self.h) or b"").decode())

    def set_profiling(self, on: bool):
        self.check(self.lib.lnr_ctx_set_profiling(self.h, int(on)))

    def reset_kernel_times(self):
        self.check(self.lib.lnr_ctx_reset_kernel_times(self.h))

    def kernel_times(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = (C.c_uint64 * cap)()
        k = self.lib.lnr_ctx_kernel_times(self.h, cap, names, ms, n)
        return {names[i].decode(): (float(ms[i]), int(n[i])) for i in range(min(k, cap))}

    def counters(self):
        c = (C.c_uint64 * 8)()
        self.check(self.lib.lnr_last_batch_counters(self.h, c))
        return dict(zip(COUNTER_NAMES, [int(v) for v in c]))

    def diag(self):
        c = (C.c_uint64 * 8)()
        self.check(self.lib.lnr_last_batch_diag(self.h, c))
        return {"hits_big_tasks": int(c[0]), "finish_big_reads": int(c[1]), "seed_rescans": int(c[2]), "heavy_lane_tasks": int(c[3])}

    def stage_cycles(self):
        c = (C.c_uint64 * 16)()
        self.check(self.lib.lnr_last_batch_stage_cycles(self.h, c))
        return dict(zip(STAGE_NAMES, [int(v) for v in c]))

    def close(self):
        if getattr(self, "h", None):
            self.lib.lnr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Genome:
    """Device-resident genome (StringSet<String<Dna5>> of the reference)."""

    def __init__(self, ctx: Context, contigs: Optional[Sequence[np.ndarray]] = None, device_ptr: int = 0,
                 lens: Optional[Sequence[int]] = None):
        self.ctx = ctx
        h = C.c_void_p()
        if contigs is not None:
            self._keep = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
            n = len(self._keep)
            ptrs = (u8p * n)(*[c.ctypes.data_as(u8p) for c in self._keep])
            ln = np.array([len(c) for c in self._keep], dtype=np.uint64)
            ctx.check(ctx.lib.lnr_genome_upload(ctx.h, n, ptrs, ln.ctypes.data_as(u64p), C.byref(h)))
            self.lens = [int(v) for v in ln]
        else:
            ln = np.array(list(lens), dtype=np.uint64)
            ctx.check(ctx.lib.lnr_genome_from_device(ctx.h, len(ln), C.c_void_p(device_ptr), ln.ctypes.data_as(u64p), C.byref(h)))
            self.lens = [int(v) for v in ln]
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.lnr_genome_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Features:
    def __init__(self, ctx: Context, genome: Genome, feature_type: int = 2, threads: int = 4):
        self.ctx, self.genome, self.feature_type = ctx, genome, feature_type
        h = C.c_void_p()
        ctx.check(ctx.lib.lnr_features_build(ctx.h, genome.h, feature_type, threads, C.byref(h)))
        self.h = h

    def download(self, contig: int) -> np.ndarray:
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.lnr_features_count(self.h, contig, C.byref(n)))
        out = np.zeros((n.value, 1), dtype=np.int16) if self.feature_type == 1 else np.zeros((n.value, 3), dtype=np.int32)
        self.ctx.check(self.ctx.lib.lnr_features_download(self.h, contig, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.lnr_features_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    def __init__(self, ctx: Context, genome: Optional[Genome] = None, index_type: int = 1, threads: int = 4, shard: int = 0,
                 n_shards: int = 1, handle=None):
        self.ctx, self.genome, self.index_type = ctx, genome, index_type
        if handle is not None:
            self.h = handle
            return
        h = C.c_void_p()
        if n_shards == 1:
            ctx.check(ctx.lib.lnr_index_build(ctx.h, genome.h, index_type, threads, C.byref(h)))
        else:
            ctx.check(ctx.lib.lnr_index_build_shard(ctx.h, genome.h, index_type, threads, shard, n_shards, C.byref(h)))
        self.h = h

    def save(self, path: str):
        """lnr_index_save: the index arrays behind a 64-byte header (the reference has no on-disk index)"""
        self.ctx.check(self.ctx.lib.lnr_index_save(self.h, os.fsencode(path)))

    @staticmethod
    def load(ctx: "Context", path: str) -> "Index":
        h = C.c_void_p()
        ctx.check(ctx.lib.lnr_index_load(ctx.h, os.fsencode(path), C.byref(h)))
        t = 1 if ctx.lib.lnr_index_export_dindex(h, None, None, 0, C.c_uint64()) == 0 else 2
        return Index(ctx, None, index_type=t, handle=h)

    def export_hindex(self):
        """(ysa uint64[], emptyDir, sorted (val1, val2) directory entries, table length) of an HIndex (-i 2)"""
        n, nk, e, tl = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.lnr_index_export_hindex(self.h, None, 0, C.byref(n), None, 0, C.byref(nk), C.byref(e), C.byref(tl)))
        ysa = np.zeros(n.value, np.uint64)
        kv = np.zeros(2 * nk.value, np.uint64)
        self.ctx.check(self.ctx.lib.lnr_index_export_hindex(self.h, ysa.ctypes.data_as(u64p), n.value, C.byref(n), kv.ctypes.data_as(u64p),
                                                            nk.value, C.byref(nk), C.byref(e), C.byref(tl)))
        return ysa, int(e.value), kv.reshape(-1, 2), int(tl.value)

    def export_device(self, torch, device):
        """(dir int32[2^26+1], hs int64[n_hs]) as torch device tensors (device-to-device copy)"""
        n = self.n_hs
        d = torch.empty((1 << 26) + 1, dtype=torch.int32, device=device)
        hs = torch.empty(max(n, 1), dtype=torch.int64, device=device)
        self.ctx.check(self.ctx.lib.lnr_index_export_dindex_device(self.h, C.c_void_p(d.data_ptr()), C.c_void_p(hs.data_ptr()), max(n, 1)))
        return d, hs[:n]

    @staticmethod
    def from_device(ctx: Context, dir_t, hs_t) -> "Index":
        h = C.c_void_p()
        ctx.check(ctx.lib.lnr_index_from_device(ctx.h, C.c_void_p(dir_t.data_ptr()), C.c_void_p(hs_t.data_ptr()), int(hs_t.numel()), C.byref(h)))
        return Index(ctx, None, handle=h)

    @property
    def n_hs(self) -> int:
        if self.index_type == 2:
            return 0
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.lnr_index_export_dindex(self.h, None, None, 0, C.byref(n)))
        return int(n.value)

    def export_dindex(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.n_hs
        dir_ = np.zeros((1 << 26) + 1, dtype=np.int32)
        hs = np.zeros(n, dtype=np.uint64)
        nn = C.c_uint64()
        self.ctx.check(self.ctx.lib.lnr_index_export_dindex(self.h, dir_.ctypes.data_as(i32p), hs.ctypes.data_as(u64p), n, C.byref(nn)))
        return dir_, hs

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.lnr_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """One rank's NCCL communicator for lnr_index_build_sharded. `exchange_id(id_bytes or None) -> id_bytes` distributes
    rank 0's 128-byte unique id to the other ranks (e.g. a torch.distributed broadcast); the library itself only needs NCCL."""

    def __init__(self, ctx: Context, rank: int, n_ranks: int, exchange_id):
        self.ctx = ctx
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            rc = ctx.lib.lnr_nccl_unique_id(buf)
            if rc != 0:
                raise LnrError(rc, "lnr_nccl_unique_id (libnccl.so.2 missing?)")
        raw = exchange_id(bytes(buf) if rank == 0 else None)
        idb = (C.c_uint8 * 128).from_buffer_copy(raw)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.lnr_comm_create(ctx.h, idb, rank, n_ranks, C.byref(self.h)))
        self.rank, self.n_ranks = rank, n_ranks

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.lnr_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def hindex_shard_cuts(pairs_per_x: np.ndarray, n_ranks: int) -> np.ndarray:
    """X ranges of the sharded HIndex build (lnr_hindex_shard_cuts; host arithmetic, no device needed): cuts[0..n_ranks]"""
    lib = load_library()
    h = np.ascontiguousarray(pairs_per_x, dtype=np.uint32)
    cuts = np.zeros(n_ranks + 1, dtype=np.uint32)
    rc = lib.lnr_hindex_shard_cuts(h.ctypes.data_as(C.POINTER(C.c_uint32)), len(h), n_ranks, cuts.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc != 0:
        raise LnrError(rc, "lnr_hindex_shard_cuts")
    return cuts


def create_index_sharded(ctx, genome, comm: "Comm", index_type=1, threads=4) -> Index:
    """createIndexDynamic across the ranks of `comm`: minimizer-range shards + one NCCL exchange (lnr_index_build_sharded)"""
    h = C.c_void_p()
    ctx.check(ctx.lib.lnr_index_build_sharded(ctx.h, genome.h, index_type, threads, comm.h, C.byref(h)))
    ix = Index(ctx, genome, index_type, handle=h)
    return ix


def create_features(ctx, genome, feature_type=2, threads=4) -> Features:
    return Features(ctx, genome, feature_type, threads)


def create_index(ctx, genome, index_type=1, threads=4) -> Index:
    return Index(ctx, genome, index_type, threads)


def read_features(ctx: Context, read: np.ndarray, feature_type: int = 2):
    read = np.ascontiguousarray(read, dtype=np.uint8)
    n = C.c_uint64()
    ctx.check(ctx.lib.lnr_read_features(ctx.h, read.ctypes.data_as(u8p), len(read), feature_type, None, None, 0, C.byref(n)))
    shape, dt = ((n.value, 1), np.int16) if feature_type == 1 else ((n.value, 3), np.int32)
    f = np.zeros(shape, dt)
    r = np.zeros(shape, dt)
    if n.value:
        ctx.check(ctx.lib.lnr_read_features(ctx.h, read.ctypes.data_as(u8p), len(read), feature_type,
                                            f.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return f, r


class Reads:
    """FASTA / FASTQ text parsed on the device (lnr_reads_parse): Dna5 ordinals, read offsets and ids.
    Mirrors readRecords + the Dna5 conversion in front of p_calRecords (loadRecords base.cpp:154)."""

    def __init__(self, ctx: Context, text: bytes, cut_id_at_space: bool = False):
        self.ctx, self.text = ctx, bytes(text)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.lnr_reads_parse(ctx.h, self.text, len(self.text), int(cut_id_at_space), C.byref(self.h)))
        n, tb = C.c_uint64(), C.c_uint64()
        ctx.check(ctx.lib.lnr_reads_info(self.h, C.byref(n), C.byref(tb)))
        self.n_reads, self.total_bases = int(n.value), int(tb.value)

    def download(self):
        """(bases uint8[total], offsets uint64[n+1], ids list[str])"""
        bases = np.empty(self.total_bases, np.uint8)
        off = np.zeros(self.n_reads + 1, np.uint64)
        io = np.zeros(max(self.n_reads, 1), np.uint64)
        il = np.zeros(max(self.n_reads, 1), np.uint32)
        self.ctx.check(self.ctx.lib.lnr_reads_download(self.h, bases.ctypes.data_as(u8p), off.ctypes.data_as(u64p), io.ctypes.data_as(u64p),
                                                       il.ctypes.data_as(C.POINTER(C.c_uint32))))
        ids = [self.text[int(io[i]):int(io[i]) + int(il[i])].decode("latin1") for i in range(self.n_reads)]
        return bases, off, ids

    def apx_map(self, index: "Index", feats: "Features", first: int = 0, n: Optional[int] = None, preset: int = 1, cap: Optional[int] = None):
        """lnr_apxmap_reads: map reads [first, first + n) straight from the parsed device buffers"""
        n = self.n_reads - first if n is None else n
        cap = int(self.total_bases // 16 + 64 * n + 1024) if cap is None else cap
        cords = np.empty(cap, dtype=np.uint64)
        coff = np.zeros(n + 1, dtype=np.uint64)
        prm = Params(preset=preset, feature_type=feats.feature_type)
        self.ctx.check(self.ctx.lib.lnr_apxmap_reads(self.ctx.h, index.h, feats.h, C.byref(prm), self.h, first, n, cords.ctypes.data_as(u64p),
                                                     coff.ctypes.data_as(u64p), cap, None))
        return cords[:int(coff[n])].copy(), coff

    def close(self):
        if self.h:
            self.ctx.lib.lnr_reads_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def selftest_sort(ctx: Context, records: np.ndarray) -> np.ndarray:
    """std::sort-order sort of uint64 records by their high 32 bits with one warp (lnr_selftest_sort)"""
    a = np.ascontiguousarray(records, dtype=np.uint64).copy()
    ctx.check(ctx.lib.lnr_selftest_sort(ctx.h, a.ctypes.data_as(u64p), len(a)))
    return a


def apx_map_batch(ctx: Context, index: Index, feats: Features, bases, offsets, preset: int = 1, debug: bool = False,
                  cords_out: Optional[np.ndarray] = None, cords_off_out: Optional[np.ndarray] = None, no_chain: bool = False,
                  gdl_state: int = 0):
    """bases: uint8 host buffer (numpy array; pass the numpy view of a pinned torch tensor for full PCIe speed),
    offsets: uint64[n+1]. Returns (cords uint64[], cords_off uint64[n+1][, debug dict]).
    no_chain = the reference's -c 0 (apxMap with f_chain = 0); gdl_state: see lnr_params in include/lnr_b200.h."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    cap = int(len(bases) // 16 + 64 * n + 1024) if cords_out is None else len(cords_out)
    cords = np.empty(cap, dtype=np.uint64) if cords_out is None else cords_out
    coff = np.zeros(n + 1, dtype=np.uint64) if cords_off_out is None else cords_off_out
    prm = Params(preset=preset, feature_type=feats.feature_type, no_chain=int(no_chain), gdl_state=gdl_state)
    dbg = None
    keep = {}
    if debug:
        dbg = DebugOut()
        ra_cap = int(len(bases) * 4 + 1024)
        keep["ra"] = np.zeros(ra_cap, np.uint64); keep["ra_off"] = np.zeros(n + 1, np.uint64)
        keep["h"] = np.zeros(int(len(bases) // 8 + 64 * n), np.uint64); keep["h_off"] = np.zeros(n + 1, np.uint64)
        keep["c1"] = np.zeros(int(len(bases) // 4 + 16 * n + 8), np.uint64); keep["c1_off"] = np.zeros(n + 1, np.uint64)
        dbg.raw_anchors = keep["ra"].ctypes.data_as(u64p); dbg.raw_anchors_cap = ra_cap; dbg.raw_anchors_off = keep["ra_off"].ctypes.data_as(u64p)
        dbg.hits = keep["h"].ctypes.data_as(u64p); dbg.hits_cap = len(keep["h"]); dbg.hits_off = keep["h_off"].ctypes.data_as(u64p)
        dbg.cords1 = keep["c1"].ctypes.data_as(u64p); dbg.cords1_cap = len(keep["c1"]); dbg.cords1_off = keep["c1_off"].ctypes.data_as(u64p)
    rc = ctx.lib.lnr_apxmap_batch(ctx.h, index.h, feats.h, C.byref(prm), n, bases.ctypes.data_as(C.c_void_p),
                                  offsets.ctypes.data_as(u64p), cords.ctypes.data_as(C.c_void_p), coff.ctypes.data_as(u64p), cap,
                                  C.byref(dbg) if dbg is not None else None)
    ctx.check(rc)
    res = cords[: int(coff[-1])]
    if debug:
        return res, coff, keep
    return res, coff


def pack_dna5(bases: np.ndarray):
    """Dna5 ordinals -> (packed2 uint8[(n+3)/4], n_mask uint8[(n+7)/8] or None when the batch has no N) -- lnr_pack_dna5"""
    lib = load_library()
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    n = len(bases)
    packed = np.zeros((n + 3) // 4, np.uint8)
    nmask = np.zeros((n + 7) // 8, np.uint8)
    has_n = C.c_int()
    rc = lib.lnr_pack_dna5(C.c_void_p(bases.ctypes.data), n, C.c_void_p(packed.ctypes.data), C.c_void_p(nmask.ctypes.data), C.byref(has_n))
    if rc != 0:
        raise LnrError(rc, "lnr_pack_dna5")
    return packed, (nmask if has_n.value else None)


def apx_map_batch_packed(ctx: Context, index: Index, feats: Features, packed: np.ndarray, n_mask: Optional[np.ndarray], offsets,
                         preset: int = 1, cords_out: Optional[np.ndarray] = None, cords_off_out: Optional[np.ndarray] = None):
    """lnr_apxmap_batch_packed: 2-bit packed reads (pack_dna5) in host memory -> (cords, cords_off)"""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1])
    cap = int(total // 16 + 64 * n + 1024) if cords_out is None else len(cords_out)
    cords = np.empty(cap, dtype=np.uint64) if cords_out is None else cords_out
    coff = np.zeros(n + 1, dtype=np.uint64) if cords_off_out is None else cords_off_out
    prm = Params(preset=preset, feature_type=feats.feature_type)
    rc = ctx.lib.lnr_apxmap_batch_packed(ctx.h, index.h, feats.h, C.byref(prm), n, C.c_void_p(packed.ctypes.data),
                                         C.c_void_p(n_mask.ctypes.data) if n_mask is not None else None,
                                         offsets.ctypes.data_as(u64p), cords.ctypes.data_as(C.c_void_p), coff.ctypes.data_as(u64p), cap, None)
    ctx.check(rc)
    return cords[: int(coff[-1])], coff


def cords_to_records(ctx: Context, cords: np.ndarray, cords_off: np.ndarray, read_len: np.ndarray, window: int = 96, thd_large_x: int = 8000,
                     thd_di: int = (1 << 60) - 1, thd_x: int = (1 << 60) - 1):
    """cords2BamLink (f_io.cpp:883) of a block of reads on the GPU: (records structured array, rec_off, cigar elements, cigar_off)"""
    cords = np.ascontiguousarray(cords, dtype=np.uint64)
    cords_off = np.ascontiguousarray(cords_off, dtype=np.uint64)
    read_len = np.ascontiguousarray(read_len, dtype=np.uint64)
    n = len(cords_off) - 1
    prm = BamParms(window=window, thd_large_x=thd_large_x, thd_di=thd_di, thd_x=thd_x)
    rec_off = np.zeros(n + 1, np.uint64)
    cig_off = np.zeros(n + 1, np.uint64)
    rc = ctx.lib.lnr_cords_to_records(ctx.h, n, C.c_void_p(cords.ctypes.data), cords_off.ctypes.data_as(u64p), read_len.ctypes.data_as(u64p), C.byref(prm),
                                      None, 0, rec_off.ctypes.data_as(u64p), None, 0, cig_off.ctypes.data_as(u64p))
    if rc not in (0, LNR_E_CAPACITY):
        ctx.check(rc)
    recs = np.zeros(int(rec_off[n]), BAM_REC_DTYPE)
    cig = np.zeros(int(cig_off[n]), np.uint64)
    if len(recs):
        ctx.check(ctx.lib.lnr_cords_to_records(ctx.h, n, C.c_void_p(cords.ctypes.data), cords_off.ctypes.data_as(u64p), read_len.ctypes.data_as(u64p),
                                               C.byref(prm), C.c_void_p(recs.ctypes.data), len(recs), rec_off.ctypes.data_as(u64p),
                                               C.c_void_p(cig.ctypes.data), len(cig), cig_off.ctypes.data_as(u64p)))
    return recs, rec_off, cig, cig_off


def cords_end(cords_str: np.ndarray, window: int = 96) -> np.ndarray:
    """cords_end[i] = cords_str[i] + ((W << 20) | W) (pmpfinder.cpp:2790-2801); W = 96 for -f 2, 192 for -f 1."""
    return cords_str + np.uint64((window << 20) | window)
