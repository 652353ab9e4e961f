"""ctypes binding of tests/emu/libglx_emu.so — the device cell logic compiled for the host (TEST INFRASTRUCTURE).
It exists so that the kernels' control flow can be diffed against the oracle on the CPU-only build box; the GPU
tests (-m gpu) run the real CUDA build of the same headers through the C ABI."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_PATH = os.path.join(EMU_DIR, "libglx_emu.so")
_lib = None


def build():
    srcs = [os.path.join(EMU_DIR, "glx_emu.cc"), os.path.join(EMU_DIR, "cuda_host_shim.h")]
    srcs += [os.path.join(ROOT, "glnexus_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "glnexus_b200", "csrc")) if f.endswith(".cuh")]
    srcs.append(os.path.join(ROOT, "include", "glx.h"))
    if os.path.exists(EMU_PATH) and all(os.path.getmtime(EMU_PATH) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-sign-compare",
                           "-o", EMU_PATH, os.path.join(EMU_DIR, "glx_emu.cc")])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(EMU_PATH)
        vp = C.c_void_p
        L.glx_emu_general_batch.argtypes = [vp, vp, vp, vp]
        L.glx_emu_general_batch.restype = C.c_int
        L.glx_emu_tiled_batch.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_uint64), C.c_int]
        L.glx_emu_tiled_batch.restype = C.c_int
        L.glx_emu_check_hints.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
        L.glx_emu_check_hints.restype = C.c_int
        L.glx_emu_bcf_records.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.glx_emu_bcf_records.restype = C.c_int
        L.glx_emu_bgzf.argtypes = [vp, vp, vp, vp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.glx_emu_bgzf.restype = C.c_int
        L.glx_emu_wire_shapes_check.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.glx_emu_wire_shapes_check.restype = C.c_longlong
        _lib = L
    return _lib


def wire_shapes_check(batch):
    """(records whose decoded shape differs from the record store, shapes in the table, records coded as exceptions); the first is -1
    when the batch's wire form carries no shape codes."""
    ptr, _ = batch.wire()
    ns, ne = C.c_uint64(), C.c_uint64()
    bad = lib().glx_emu_wire_shapes_check(ptr, batch.records_ptr, C.byref(ns), C.byref(ne))
    return bad, ns.value, ne.value


def general_batch(batch, out):
    return lib().glx_emu_general_batch(batch.config_ptr, batch.records_ptr, batch.sites_ptr, out.ptr)


def tiled_batch(batch, out, dynamic_only=False):
    """Returns (status, number of cells the fast path declined). dynamic_only skips the signature-specialised lanes."""
    n = C.c_uint64()
    rc = lib().glx_emu_tiled_batch(batch.config_ptr, batch.records_ptr, batch.sites_ptr, out.ptr, C.byref(n), 1 if dynamic_only else 0)
    return rc, n.value


def check_hints(batch, tile_sites, block):
    """Mismatches between the resolve kernels' hint-confined tile searches and the full searches (0 = identical)."""
    return lib().glx_emu_check_hints(batch.records_ptr, batch.sites_ptr, tile_sites, block)



def bcf_records(batch, out, span=0, chunk=0):
    """(bytes, rec_off) from the device BCF encoder's stages run on the host (one virtual thread per block)."""
    p, n = C.c_void_p(), C.c_uint64()
    off = (C.c_uint64 * (batch.n_sites + 1))()
    rc = lib().glx_emu_bcf_records(batch.config_ptr, batch.sites_ptr, batch.bcf_meta_ptr, out.ptr, span, chunk, C.byref(p), C.byref(n), off)
    assert rc == 0
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    try:
        return C.string_at(p.value, n.value), list(off)
    finally:
        libc.free(p)


def bgzf_blocks(batch, out, span=0):
    """(bytes, stored blocks): the BGZF blocks of the batch's record stream from the device encoder's stages run on the host."""
    p, n, ns = C.c_void_p(), C.c_uint64(), C.c_uint32()
    rc = lib().glx_emu_bgzf(batch.config_ptr, batch.sites_ptr, batch.bcf_meta_ptr, out.ptr, span, C.byref(p), C.byref(n), C.byref(ns))
    assert rc == 0
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    try:
        return C.string_at(p.value, n.value), ns.value
    finally:
        libc.free(p)
