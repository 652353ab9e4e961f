"""ctypes binding of oracle/libglx_oracle.so — the CPU restatement of the reference algorithm.
Test infrastructure: imported only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libglx_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(ORACLE_PATH)
        vp = C.c_void_p
        L.glx_oracle_genotype_batch.argtypes = [vp, vp, vp, vp, C.c_int, C.c_char_p, C.c_size_t]
        L.glx_oracle_genotype_batch.restype = C.c_int
        L.glx_oracle_genotypes.argtypes = [C.c_uint]
        L.glx_oracle_genotypes.restype = C.c_uint
        L.glx_oracle_alleles_gt.argtypes = [C.c_uint, C.c_uint]
        L.glx_oracle_alleles_gt.restype = C.c_uint
        L.glx_oracle_gt_alleles.argtypes = [C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.glx_oracle_gt_alleles.restype = None
        L.glx_oracle_pl_to_gll.argtypes = [C.c_int32]
        L.glx_oracle_pl_to_gll.restype = C.c_double
        L.glx_oracle_revise_kat.argtypes = [vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
        L.glx_oracle_revise_kat.restype = C.c_int
        L.glx_oracle_bcf_records.argtypes = [vp, vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.glx_oracle_bcf_records.restype = C.c_int
        L.glx_oracle_free.argtypes = [vp]
        L.glx_oracle_free.restype = None
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        self.code = code
        super().__init__(f"oracle status {code}: {msg}")


def genotype_batch(batch, out, threads=1):
    """Run the oracle over a glnexus_b200.Batch into a glnexus_b200.Out."""
    err = C.create_string_buffer(1024)
    rc = lib().glx_oracle_genotype_batch(batch.config_ptr, batch.records_ptr, batch.sites_ptr, out.ptr, threads, err, len(err))
    if rc != 0:
        raise OracleError(rc, err.value.decode(errors="replace"))


def revise_kat(batch):
    a1, a2, gq = C.c_int(), C.c_int(), C.c_int()
    err = C.create_string_buffer(1024)
    rc = lib().glx_oracle_revise_kat(batch.config_ptr, batch.records_ptr, batch.sites_ptr, C.byref(a1), C.byref(a2), C.byref(gq), err, len(err))
    if rc != 0:
        raise OracleError(rc, err.value.decode(errors="replace"))
    return a1.value, a2.value, gq.value


def bcf_records(batch, out):
    """(bytes, rec_off): the batch's BCF2 records, uncompressed, from the packed output columns (BCF oracle)."""
    p, n = C.c_void_p(), C.c_uint64()
    off = (C.c_uint64 * (batch.n_sites + 1))()
    rc = lib().glx_oracle_bcf_records(batch.config_ptr, batch.sites_ptr, batch.bcf_meta_ptr, out.ptr, C.byref(p), C.byref(n), off)
    if rc != 0:
        raise OracleError(rc, "glx_oracle_bcf_records")
    try:
        return C.string_at(p.value, n.value), list(off)
    finally:
        lib().glx_oracle_free(p)
