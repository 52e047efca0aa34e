"""ctypes binding of the synthetic genome generator (bench_support/libsynth.so). Inputs only; not the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsynth.so")
PRIMARY_SEED = 20260924
_lib = None


def build():
    subprocess.check_call(["/usr/bin/g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-fopenmp", "-shared",
                           "-o", LIB, os.path.join(HERE, "synth.cpp")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.synth_layout.restype = C.c_uint64
        L.synth_layout.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.synth_generate.restype = None
        L.synth_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_double,
                                     C.c_void_p, C.c_int]
        L.synth_layout_ids.restype = C.c_uint64
        L.synth_layout_ids.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.synth_generate_ids.restype = None
        L.synth_generate_ids.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_double,
                                         C.c_void_p, C.c_int]
        _lib = L
    return _lib


def shuffled_ids(n_total, perm_seed):
    """Seeded permutation of the genome ids 0..n_total-1 (slot p of the shuffled set holds genome ids[p])."""
    return np.random.Generator(np.random.PCG64(perm_seed)).permutation(n_total).astype(np.uint64)


def layout_ids(ids, L, G=20, seed=PRIMARY_SEED):
    ids = np.ascontiguousarray(ids, np.uint64)
    n = lib().synth_layout_ids(seed, ids.ctypes.data, len(ids), L, G, None, None)
    off = np.zeros(n + 1, np.uint64); goc = np.zeros(n, np.uint32)
    lib().synth_layout_ids(seed, ids.ctypes.data, len(ids), L, G, off.ctypes.data, goc.ctypes.data)
    return off, goc


def generate_ids(ids, L, G=20, dmin=0.001, dmax=0.05, seed=PRIMARY_SEED, out=None, threads=None):
    """Like generate(), but slot p holds genome ids[p] (any order, any subset)."""
    ids = np.ascontiguousarray(ids, np.uint64)
    n = len(ids) * L
    if out is None:
        out = np.empty(n, np.uint8)
    assert out.size >= n and out.dtype == np.uint8
    threads = threads or min(os.cpu_count() or 1, 64)
    lib().synth_generate_ids(seed, ids.ctypes.data, len(ids), L, G, dmin, dmax, out.ctypes.data, threads)
    off, goc = layout_ids(ids, L, G, seed)
    return out, off, goc


def layout(g_begin, g_end, L, G=20, seed=PRIMARY_SEED):
    n = lib().synth_layout(seed, g_begin, g_end, L, G, None, None)
    off = np.zeros(n + 1, np.uint64); goc = np.zeros(n, np.uint32)
    lib().synth_layout(seed, g_begin, g_end, L, G, off.ctypes.data, goc.ctypes.data)
    return off, goc


def generate(g_begin, g_end, L, G=20, dmin=0.001, dmax=0.05, seed=PRIMARY_SEED, out=None, threads=None):
    """Returns (bases uint8[(g_end-g_begin)*L], contig_off, genome_of_contig). `out` may be a preallocated
    (e.g. pinned) uint8 numpy array."""
    n = (g_end - g_begin) * L
    if out is None:
        out = np.empty(n, np.uint8)
    assert out.size >= n and out.dtype == np.uint8
    threads = threads or min(os.cpu_count() or 1, 64)
    lib().synth_generate(seed, g_begin, g_end, L, G, dmin, dmax, out.ctypes.data, threads)
    off, goc = layout(g_begin, g_end, L, G, seed)
    return out, off, goc
