"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE, never imported by linear_b200/):

* `Oracle`  -> oracle/liblnr_oracle.so       : our CPU restatement of the reference algorithm
* `RefImpl` -> oracle/_ref/libref_harness.so : the UNMODIFIED reference, compiled from /root/reference
  by oracle/build_ref.sh (present in this container and, as a prebuilt file, on the GPU box)

Both expose the same stage checkpoints (see oracle/ref_harness.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblnr_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)


def build_oracle() -> str:
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), ORACLE_SO])
    return ORACLE_SO


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class _Checker:
    prefix = ""

    def __init__(self, so: str, contigs: Sequence[np.ndarray], index_type=1, feature_type=2, threads=4, preset=1,
                 build_index=True):
        self.lib = C.CDLL(so)
        p = self.prefix
        L = self.lib
        self._create = getattr(L, p + "create")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.c_int, C.POINTER(u8p), u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        self._destroy = getattr(L, p + "destroy")
        self._destroy.argtypes = [C.c_void_p]
        for name, extra in (("dindex_dir", [C.POINTER(i32p)]), ("dindex_hs", [C.POINTER(u64p)]),
                            ("hindex_ysa", [C.POINTER(u64p), u64p]), ("hindex_dir_kv", [C.POINTER(u64p), u64p]),
                            ("genome_features", [C.c_int, C.POINTER(i32p)]),
                            ("read_features", [u8p, C.c_uint64, C.c_int, C.POINTER(i32p)]),
                            ("read_stage", [u8p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(u64p)])):
            fn = getattr(L, p + name)
            fn.restype = C.c_int64
            fn.argtypes = [C.c_void_p] + extra
            setattr(self, "_" + name, fn)
        self._cords2bam = getattr(L, p + "cords2bam", None)
        if self._cords2bam is not None:
            self._cords2bam.restype = C.c_int64
            self._cords2bam.argtypes = [C.c_void_p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_uint64, C.c_int64, C.c_int64,
                                        C.POINTER(C.POINTER(C.c_int64)), C.POINTER(u64p), u64p]
        self._map_batch = getattr(L, p + "map_batch")
        self._map_batch.restype = C.c_int
        self._map_batch.argtypes = [C.c_void_p, C.c_uint32, u8p, u64p, C.c_int, u64p, u64p, C.c_uint64]
        self._map_batch_c0 = getattr(L, p + "map_batch_c0", None)
        if self._map_batch_c0 is not None:
            self._map_batch_c0.restype = C.c_int
            self._map_batch_c0.argtypes = [C.c_void_p, C.c_uint32, u8p, u64p, C.c_int, C.c_int, u64p, u64p, C.c_uint64]
        self.contigs = [np.ascontiguousarray(c, dtype=np.uint8) for c in contigs]
        n = len(self.contigs)
        ptrs = (u8p * n)(*[c.ctypes.data_as(u8p) for c in self.contigs])
        lens = np.array([len(c) for c in self.contigs], dtype=np.uint64)
        self.h = self._create(n, ptrs, lens.ctypes.data_as(u64p), index_type, feature_type, threads, preset,
                              int(build_index))
        self.feature_type = feature_type

    def close(self):
        if self.h:
            self._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- index ------------------------------------------------------------------------------------
    def dindex(self):
        p = i32p()
        n = self._dindex_dir(self.h, C.byref(p))
        dir_ = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        q = u64p()
        m = self._dindex_hs(self.h, C.byref(q))
        hs = np.ctypeslib.as_array(q, shape=(m,)).copy() if m else np.zeros(0, np.uint64)
        return dir_, hs

    def dindex_views(self):
        """(dir, hs) as views of the checker's own arrays (no copy: hs is 2.75 GB at 3.1 Gbase); valid while it lives"""
        p = i32p()
        n = self._dindex_dir(self.h, C.byref(p))
        q = u64p()
        m = self._dindex_hs(self.h, C.byref(q))
        dir_ = np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0, np.int32)
        hs = np.ctypeslib.as_array(q, shape=(m,)) if m else np.zeros(0, np.uint64)
        return dir_, hs

    def hindex(self):
        q = u64p()
        e = C.c_uint64()
        m = self._hindex_ysa(self.h, C.byref(q), C.byref(e))
        ysa = np.ctypeslib.as_array(q, shape=(m,)).copy() if m else np.zeros(0, np.uint64)
        k = u64p()
        tl = C.c_uint64()
        nk = self._hindex_dir_kv(self.h, C.byref(k), C.byref(tl))
        kv = np.ctypeslib.as_array(k, shape=(2 * nk,)).copy().reshape(-1, 2) if nk else np.zeros((0, 2), np.uint64)
        return ysa, int(e.value), kv, int(tl.value)

    # -- features ---------------------------------------------------------------------------------
    def genome_features(self, contig: int) -> np.ndarray:
        p = i32p()
        n = self._genome_features(self.h, contig, C.byref(p))
        w = 3 if self.feature_type == 2 else 1
        return np.ctypeslib.as_array(p, shape=(n * w,)).copy().reshape(n, w) if n else np.zeros((0, w), np.int32)

    def read_features(self, read: np.ndarray, strand: int) -> np.ndarray:
        read = np.ascontiguousarray(read, dtype=np.uint8)
        p = i32p()
        n = self._read_features(self.h, read.ctypes.data_as(u8p), len(read), strand, C.byref(p))
        w = 3 if self.feature_type == 2 else 1
        return np.ctypeslib.as_array(p, shape=(n * w,)).copy().reshape(n, w) if n else np.zeros((0, w), np.int32)

    # -- per-read stages ---------------------------------------------------------------------------
    def stage(self, read: np.ndarray, stage: int, str_: int = 0, end: int = -1, toggle: int = 0) -> np.ndarray:
        read = np.ascontiguousarray(read, dtype=np.uint8)
        if end < 0:
            end = len(read)
        q = u64p()
        n = self._read_stage(self.h, read.ctypes.data_as(u8p), len(read), stage, str_, end, toggle, C.byref(q))
        return np.ctypeslib.as_array(q, shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def cords(self, read: np.ndarray) -> np.ndarray:
        return self.stage(read, 0)

    def cords2bam(self, read_len: int, cords: np.ndarray, window: int = 96, thd_large_x: int = 8000, thd_di: int = (1 << 60) - 1,
                  thd_x: int = (1 << 60) - 1):
        """cords2BamLink (f_io.cpp:883) of one read: (records int64[n, 8] = rID, beginPos, flag, s1, s2, s3, cigar begin, cigar end;
        cigar elements uint64[] = (operation << 32) | count)"""
        cords = np.ascontiguousarray(cords, dtype=np.uint64)
        pr = C.POINTER(C.c_int64)()
        pc = u64p()
        nc = C.c_uint64()
        n = self._cords2bam(self.h, read_len, cords.ctypes.data_as(u64p), len(cords), window, thd_large_x, thd_di, thd_x,
                            C.byref(pr), C.byref(pc), C.byref(nc))
        recs = np.ctypeslib.as_array(pr, shape=(n * 8,)).copy().reshape(n, 8) if n else np.zeros((0, 8), np.int64)
        cig = np.ctypeslib.as_array(pc, shape=(nc.value,)).copy() if nc.value else np.zeros(0, np.uint64)
        return recs, cig

    def map_batch(self, bases: np.ndarray, offsets: np.ndarray, map_threads: int = 1, cap_per_base: float = 0.25,
                  no_chain: bool = False, gdl_state: int = 0):
        """no_chain: the reference's -c 0 (apxMap with f_chain = 0); gdl_state: see oracle/ref_harness.cpp ref_map_batch_c0"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        cap = int(len(bases) * cap_per_base) + 64 * n + 1024
        cords = np.zeros(cap, dtype=np.uint64)
        coff = np.zeros(n + 1, dtype=np.uint64)
        if no_chain:
            rc = self._map_batch_c0(self.h, n, bases.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p), map_threads, gdl_state,
                                    cords.ctypes.data_as(u64p), coff.ctypes.data_as(u64p), cap)
        else:
            rc = self._map_batch(self.h, n, bases.ctypes.data_as(u8p), offsets.ctypes.data_as(u64p), map_threads,
                                 cords.ctypes.data_as(u64p), coff.ctypes.data_as(u64p), cap)
        if rc != 0:
            raise RuntimeError("map_batch capacity too small")
        return cords[: int(coff[-1])].copy(), coff


class Oracle(_Checker):
    prefix = "orc_"

    def __init__(self, contigs, **kw):
        super().__init__(build_oracle(), contigs, **kw)


class _quiet_stderr:
    """The reference prints a progress panel to stderr (index_util.cpp:1690); silence it."""

    def __enter__(self):
        self.saved = os.dup(2)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 2)

    def __exit__(self, *a):
        os.dup2(self.saved, 2)
        os.close(self.null)
        os.close(self.saved)


class RefImpl(_Checker):
    prefix = "ref_"

    def __init__(self, contigs, **kw):
        if not have_ref():
            raise FileNotFoundError(REF_SO + " missing: run `make -C oracle ref` where /root/reference exists")
        with _quiet_stderr():
            super().__init__(REF_SO, contigs, **kw)


EMU_SO = os.path.join(ROOT, "tests", "host_emu", "libemu.so")


def build_emu() -> str:
    src = os.path.join(ROOT, "tests", "host_emu", "emu.cpp")
    hdrs = [os.path.join(ROOT, "linear_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "linear_b200", "csrc"))
            if f.endswith(".h")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-strict-aliasing", "-Wno-unknown-pragmas", src, "-o", EMU_SO])
    return EMU_SO


class HostEmu(_Checker):
    """Product headers (linear_b200/csrc/lnr_*.h) compiled for the host with a single-lane warp."""
    prefix = "emu_"

    def __init__(self, contigs, **kw):
        super().__init__(build_emu(), contigs, **kw)
