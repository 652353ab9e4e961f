"""ctypes binding of libglx.so (include/glx.h + include/glx_host.h).

The library is the product: CUDA kernels for sm_100a behind a C ABI plus the C++ host pipeline
(decode -> pack -> assemble). There is no CPU fallback: if the shared library is missing or a CUDA call
fails, the calls below raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLX_LIB_PATH") or os.path.join(_HERE, "libglx.so")  # override: profiling builds of the same sources

GLX_OK = 0
STATUS_NAMES = {0: "OK", -1: "Failure", -2: "Invalid", -3: "NotFound", -4: "Exists", -5: "IOError",
                -6: "NotImplemented", -7: "Aborted", -100: "CudaError", -101: "Unsupported"}


class GlxError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        self.status = STATUS_NAMES.get(code, str(code))
        super().__init__(f"{self.status}: {msg}" if msg else self.status)


class SynthParams(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("n_sites", C.c_uint32), ("seed", C.c_uint64),
                ("region_beg", C.c_int32), ("region_len", C.c_int32), ("max_alt", C.c_uint32),
                ("frac_multiallelic", C.c_float), ("frac_indel", C.c_float), ("af_max", C.c_float),
                ("band_mean_len", C.c_float), ("frac_uncovered", C.c_float), ("frac_two_record_cells", C.c_float),
                ("frac_lost_allele", C.c_float), ("frac_homalt", C.c_float), ("gatk_style", C.c_uint8),
                ("sites_only", C.c_uint8), ("_pad", C.c_uint8 * 2)]


_lib = None


def lib():
    """Load libglx.so; fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make` or __graft_entry__.build(); "
                          "glnexus_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    L.glxh_batch_from_text.argtypes = [cp, cp, cp, C.c_int, C.POINTER(cp), C.POINTER(cp), cp, C.c_int, C.POINTER(vp), cp, sz]
    L.glxh_batch_from_text.restype = C.c_int
    L.glxh_synth_batch.argtypes = [cp, C.POINTER(SynthParams), C.POINTER(vp), cp, sz]
    L.glxh_synth_batch.restype = C.c_int
    L.glxh_partition_sites.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32)]
    L.glxh_partition_sites.restype = C.c_int
    L.glxh_partition_cohort.argtypes = [C.POINTER(vp), C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.glxh_partition_cohort.restype = C.c_int
    L.glxh_batch_slice.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(vp), cp, sz]
    L.glxh_batch_slice.restype = C.c_int
    L.glxh_out16_alloc.argtypes = [vp, C.c_int]
    L.glxh_out16_alloc.restype = vp
    L.glxh_out16_free.argtypes = [vp]
    L.glxh_out16_free.restype = None
    L.glxh_out16_struct.argtypes = [vp]
    L.glxh_out16_struct.restype = vp
    L.glxh_out16_array.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_out16_array.restype = C.c_int
    L.glxh_out16_widen.argtypes = [vp, vp]
    L.glxh_out16_widen.restype = C.c_int
    L.glx_download_narrow_on.argtypes = [vp, vp, vp, vp, C.c_int]
    L.glx_download_narrow_on.restype = C.c_int
    L.glx_narrow_finish.argtypes = [vp, vp, vp]
    L.glx_narrow_finish.restype = C.c_int
    L.glxh_genotype_sites.argtypes = [cp, cp, cp, C.c_int, C.POINTER(cp), C.POINTER(cp), cp, C.c_int, C.c_uint32, C.c_int, vp, C.POINTER(C.c_uint64), cp, sz]
    L.glxh_genotype_sites.restype = C.c_int
    L.glxh_batch_pin.argtypes = [vp]
    L.glxh_batch_pin.restype = C.c_int
    L.glx_upload_on.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.glx_upload_on.restype = C.c_int
    L.glx_download_on.argtypes = [vp, vp, vp, vp, C.c_int]
    L.glx_download_on.restype = C.c_int
    L.glx_status.argtypes = [vp, vp]
    L.glx_status.restype = C.c_int
    L.glx_stream_create.argtypes = [vp]
    L.glx_stream_create.restype = vp
    L.glx_stream_destroy.argtypes = [vp, vp]
    L.glx_stream_destroy.restype = None
    L.glx_stream_sync.argtypes = [vp, vp]
    L.glx_stream_sync.restype = C.c_int
    L.glxh_batch_free.argtypes = [vp]
    L.glxh_batch_free.restype = None
    for fn in ("glxh_config", "glxh_records", "glxh_sites"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = vp
    L.glxh_n_fields.argtypes = [vp]
    L.glxh_n_fields.restype = C.c_uint32
    L.glxh_device_supported.argtypes = [vp]
    L.glxh_device_supported.restype = C.c_int
    L.glxh_has_fast_path.argtypes = [vp]
    L.glxh_has_fast_path.restype = C.c_int
    L.glxh_out_alloc.argtypes = [vp, C.c_int]
    L.glxh_out_alloc.restype = vp
    L.glxh_out_free.argtypes = [vp]
    L.glxh_out_free.restype = None
    L.glxh_out_struct.argtypes = [vp]
    L.glxh_out_struct.restype = vp
    L.glxh_out_array.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_out_array.restype = C.c_int
    L.glxh_assemble_vcf.argtypes = [vp, vp, C.POINTER(vp), cp, sz]
    L.glxh_assemble_vcf.restype = C.c_int
    L.glxh_free_text.argtypes = [vp]
    L.glxh_free_text.restype = None
    L.glxh_preset_text.argtypes = [cp, C.POINTER(vp)]
    L.glxh_preset_text.restype = C.c_int
    L.glxh_bcf_meta.argtypes = [vp]
    L.glxh_bcf_meta.restype = vp
    L.glxh_bcf_prologue.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_bcf_prologue.restype = C.c_int
    L.glxh_bgzf_stored.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_bgzf_stored.restype = C.c_int
    L.glxh_bgzf_eof.argtypes = [C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_bgzf_eof.restype = C.c_int
    L.glxh_batch_wire.argtypes = [vp, C.POINTER(vp)]
    L.glxh_batch_wire.restype = C.c_int
    L.glxh_debug_wire_shape_cap.argtypes = [C.c_uint32]
    L.glxh_debug_wire_shape_cap.restype = None
    L.glxh_debug_pack_threads.argtypes = [C.c_int]
    L.glxh_debug_pack_threads.restype = None
    L.glx_upload_wire_on.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.glx_upload_wire_on.restype = C.c_int
    L.glx_debug_compare_records.argtypes = [vp, vp, vp, C.POINTER(C.c_uint32)]
    L.glx_debug_compare_records.restype = C.c_int
    L.glxh_stream_use_wire.argtypes = [vp, C.c_int]
    L.glxh_stream_use_wire.restype = None
    L.glxh_stream_open.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), cp, sz]
    L.glxh_stream_open.restype = C.c_int
    L.glxh_stream_submit.argtypes = [vp, vp, cp, sz]
    L.glxh_stream_submit.restype = C.c_int
    L.glxh_stream_drain.argtypes = [vp, cp, sz]
    L.glxh_stream_drain.restype = C.c_int
    L.glxh_stream_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.glxh_stream_stats.restype = None
    L.glxh_stream_trace.argtypes = [vp, C.c_int]
    L.glxh_stream_trace.restype = None
    L.glxh_stream_trace_rows.argtypes = [vp, C.POINTER(C.c_float), C.c_uint32]
    L.glxh_stream_trace_rows.restype = C.c_uint32
    L.glxh_stream_output.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.glxh_stream_output.restype = C.c_int
    L.glxh_stream_close.argtypes = [vp]
    L.glxh_stream_close.restype = None
    L.glxh_load_config.argtypes = [cp, cp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.glxh_load_config.restype = C.c_int
    L.glxh_batch_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.glxh_batch_stats.restype = None
    # device ABI
    L.glx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.glx_create.restype = C.c_int
    L.glx_destroy.argtypes = [vp]
    L.glx_destroy.restype = None
    L.glx_last_error.argtypes = [vp]
    L.glx_last_error.restype = cp
    L.glx_genotype_batch.argtypes = [vp, vp, vp, vp, vp]
    L.glx_genotype_batch.restype = C.c_int
    L.glx_upload.argtypes = [vp, vp, vp, vp, C.POINTER(vp)]
    L.glx_upload.restype = C.c_int
    L.glx_launch.argtypes = [vp, vp, vp]
    L.glx_launch.restype = C.c_int
    L.glx_download.argtypes = [vp, vp, vp]
    L.glx_download.restype = C.c_int
    L.glx_check.argtypes = [vp, vp]
    L.glx_check.restype = C.c_int
    L.glx_out_bytes.argtypes = [vp]
    L.glx_out_bytes.restype = C.c_uint64
    L.glx_in_bytes.argtypes = [vp]
    L.glx_in_bytes.restype = C.c_uint64
    L.glx_launch_count.argtypes = [vp]
    L.glx_launch_count.restype = C.c_int
    L.glx_free_batch.argtypes = [vp, vp]
    L.glx_free_batch.restype = None
    L.glx_synchronize.argtypes = [vp]
    L.glx_synchronize.restype = C.c_int
    L.glx_set_profile.argtypes = [vp, vp, C.c_int]
    L.glx_set_profile.restype = C.c_int
    L.glx_kernel_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.glx_kernel_ms.restype = C.c_int
    L.glx_work_count.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.glx_work_count.restype = C.c_int
    L.glx_fused_count.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.glx_fused_count.restype = C.c_int
    L.glx_stream.argtypes = [vp]
    L.glx_stream.restype = vp
    L.glx_selftest_diploid.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.glx_selftest_diploid.restype = C.c_int
    L.glx_bcf_attach.argtypes = [vp, vp, vp, vp, vp, vp]
    L.glx_bcf_attach.restype = C.c_int
    L.glx_encode_bcf_on.argtypes = [vp, vp, C.c_int, vp]
    L.glx_encode_bcf_on.restype = C.c_int
    L.glx_bcf_length.argtypes = [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.glx_bcf_length.restype = C.c_int
    L.glx_download_bcf_on.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_int]
    L.glx_download_bcf_on.restype = C.c_int
    L.glx_bcf_record_offsets.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.glx_bcf_record_offsets.restype = C.c_int
    L.glx_debug_bgzf_phases.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.glx_debug_bgzf_phases.restype = C.c_int
    L.glx_bcf_bound.argtypes = [vp]
    L.glx_bcf_bound.restype = C.c_uint64
    L.glx_pinned_alloc.argtypes = [sz]
    L.glx_pinned_alloc.restype = vp
    L.glx_pinned_free.argtypes = [vp]
    L.glx_pinned_free.restype = None
    L.glx_set_option.argtypes = [vp, cp, C.c_int]
    L.glx_set_option.restype = C.c_int
    L.glx_upload_kernel_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.glx_upload_kernel_ms.restype = C.c_int
    _lib = L
    return L


def _b(s):
    return None if s is None else s.encode()


class Out:
    """Host output columns of one batch (glx_out)."""
    WHICH = {"gt": (0, np.int8), "rnc": (1, np.uint8), "allele_used": (2, np.uint32), "site_flags": (3, np.uint8)}

    def __init__(self, batch, pinned=False):
        self._batch = batch
        self.handle = lib().glxh_out_alloc(batch.handle, 1 if pinned else 0)
        if not self.handle:
            raise MemoryError("glxh_out_alloc")
        self.ptr = lib().glxh_out_struct(self.handle)

    def array(self, which):
        if isinstance(which, str):
            idx, dt = self.WHICH[which]
        else:
            idx, dt = 16 + int(which), np.int32
        p, n = C.c_void_p(), C.c_uint64()
        rc = lib().glxh_out_array(self.handle, idx, C.byref(p), C.byref(n))
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_out_array")
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_uint8 * n.value).from_address(p.value)
        return np.frombuffer(buf, dtype=dt)

    def arrays(self):
        d = {k: self.array(k) for k in self.WHICH}
        for k in range(self._batch.n_fields):
            d[f"fld{k}"] = self.array(k)
        return d

    def nbytes(self):
        return sum(a.nbytes for a in self.arrays().values())

    def close(self):
        if self.handle:
            lib().glxh_out_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Out16:
    """Host buffers of the narrow download (glx_out16): int16 columns where every value of the field fits."""

    def __init__(self, batch, pinned=True):
        self._batch = batch
        self.handle = lib().glxh_out16_alloc(batch.handle, 1 if pinned else 0)
        if not self.handle:
            raise MemoryError("glxh_out16_alloc")
        self.ptr = lib().glxh_out16_struct(self.handle)

    def widen(self, out):
        """Expand into an Out (int32 columns, 32-bit sentinels)."""
        rc = lib().glxh_out16_widen(self.handle, out.ptr)
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_out16_widen")
        return out

    def narrow_nbytes(self):
        """Bytes that crossed PCIe for GT, RNC and the 16-bit columns."""
        tot = 0
        for which in (0, 1) + tuple(16 + k for k in range(self._batch.n_fields)):
            p, n = C.c_void_p(), C.c_uint64()
            lib().glxh_out16_array(self.handle, which, C.byref(p), C.byref(n))
            tot += n.value
        return tot

    def close(self):
        if self.handle:
            lib().glxh_out16_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host bytes (cudaHostAlloc) for asynchronous downloads."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().glx_pinned_alloc(max(16, self.nbytes))
        if not self.ptr:
            raise MemoryError("glx_pinned_alloc")

    def bytes(self, n):
        return C.string_at(self.ptr, n)

    def close(self):
        if self.ptr:
            lib().glx_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """One contig-range batch: flattened config + columnar record store + unified sites."""

    def __init__(self, handle):
        self.handle = handle
        L = lib()
        self.config_ptr = L.glxh_config(handle)
        self.records_ptr = L.glxh_records(handle)
        self.sites_ptr = L.glxh_sites(handle)
        self.n_fields = L.glxh_n_fields(handle)
        st = (C.c_uint64 * 8)()
        L.glxh_batch_stats(handle, st)
        self.n_sites, self.n_samples, self.n_records, self.n_datasets = st[0], st[1], st[2], st[3]
        self.input_bytes = st[4]
        self.variant_records = st[5]

    @classmethod
    def from_text(cls, sites_text, datasets, preset=None, config_text=None, contig=None, test_ingest_rules=False):
        L = lib()
        n = len(datasets)
        names = (C.c_char_p * max(n, 1))(*[_b(d[0]) for d in datasets])
        texts = (C.c_char_p * max(n, 1))(*[_b(d[1]) for d in datasets])
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = L.glxh_batch_from_text(_b(preset), _b(config_text), _b(sites_text), n, names, texts, _b(contig),
                                    1 if test_ingest_rules else 0, C.byref(h), err, len(err))
        if rc != GLX_OK:
            raise GlxError(rc, err.value.decode(errors="replace"))
        return cls(h)

    @classmethod
    def synthetic(cls, params, preset="DeepVariant"):
        L = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = L.glxh_synth_batch(_b(preset), C.byref(params), C.byref(h), err, len(err))
        if rc != GLX_OK:
            raise GlxError(rc, err.value.decode(errors="replace"))
        return cls(h)

    def pin(self):
        """Page-lock the batch's host arrays (uploads then run at PCIe rate and asynchronously)."""
        rc = lib().glxh_batch_pin(self.handle)
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_batch_pin")
        return self

    def partition(self, n_parts):
        """Cut points (n_parts+1) of contiguous site ranges of equal output weight, one range per GPU."""
        b = (C.c_uint32 * (n_parts + 1))()
        rc = lib().glxh_partition_sites(self.handle, n_parts, b)
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_partition_sites")
        return list(b)

    def slice(self, site_lo, site_hi):
        """The shard [site_lo, site_hi): those sites plus every record they can touch."""
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().glxh_batch_slice(self.handle, site_lo, site_hi, C.byref(h), err, len(err))
        if rc != GLX_OK:
            raise GlxError(rc, err.value.decode(errors="replace"))
        return Batch(h)

    @property
    def device_supported(self):
        return bool(lib().glxh_device_supported(self.handle))

    @property
    def has_fast_path(self):
        """single-sample datasets in output order: tiled kernels and a wire form (else the multi-sample kernel, full-width upload)"""
        return bool(lib().glxh_has_fast_path(self.handle))

    def alloc_out(self, pinned=False):
        return Out(self, pinned)

    def alloc_out16(self, pinned=True):
        return Out16(self, pinned)

    def assemble_vcf(self, out):
        t = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = lib().glxh_assemble_vcf(self.handle, out.ptr, C.byref(t), err, len(err))
        if rc != GLX_OK:
            raise GlxError(rc, err.value.decode(errors="replace"))
        try:
            return C.string_at(t.value).decode()
        finally:
            lib().glxh_free_text(t)

    def wire(self):
        """(pointer, bytes) of the batch's compact wire form (glx_wire), built on first use."""
        p = C.c_void_p()
        rc = lib().glxh_batch_wire(self.handle, C.byref(p))
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_batch_wire")
        n_bytes = C.c_uint64.from_address(p.value + 8).value
        return p.value, n_bytes

    @property
    def bcf_meta_ptr(self):
        """glx_bcf_meta of the batch (per-site strings + header dictionary for the BCF record encoder)."""
        return lib().glxh_bcf_meta(self.handle)

    def bcf_prologue(self):
        """Uncompressed bytes before the first record of a BCF file (magic, l_text, header text)."""
        p, n = C.c_void_p(), C.c_uint64()
        rc = lib().glxh_bcf_prologue(self.handle, C.byref(p), C.byref(n))
        if rc != GLX_OK:
            raise GlxError(rc, "glxh_bcf_prologue")
        try:
            return C.string_at(p.value, n.value)
        finally:
            lib().glxh_free_text(p)

    def close(self):
        if self.handle:
            lib().glxh_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBatch:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def launch(self, stream=None):
        self.ctx._check(lib().glx_launch(self.ctx.handle, self.handle, C.c_void_p(stream or 0)))

    def check(self):
        self.ctx._check(lib().glx_check(self.ctx.handle, self.handle))

    def download(self, out):
        self.ctx._check(lib().glx_download(self.ctx.handle, self.handle, out.ptr))

    def download_async(self, out, stream):
        self.ctx._check(lib().glx_download_on(self.ctx.handle, self.handle, out.ptr, C.c_void_p(stream), 0))

    def download_narrow_async(self, out16, stream):
        self.ctx._check(lib().glx_download_narrow_on(self.ctx.handle, self.handle, out16.ptr, C.c_void_p(stream), 0))

    def narrow_finish(self, out16):
        """After the stream is synchronised: fix the per-field widths (fetches int32 for fields that did not fit)."""
        self.ctx._check(lib().glx_narrow_finish(self.ctx.handle, self.handle, out16.ptr))

    def status(self):
        """Error word of a batch whose asynchronous download has been synchronised."""
        self.ctx._check(lib().glx_status(self.ctx.handle, self.handle))

    # ---- N1: BCF2 records / BGZF blocks encoded on the device ----
    def bcf_attach(self, batch, stream=None):
        self.ctx._check(lib().glx_bcf_attach(self.ctx.handle, self.handle, batch.sites_ptr, batch.bcf_meta_ptr, batch.config_ptr, C.c_void_p(stream or 0)))
        self._n_sites = batch.n_sites

    def encode_bcf(self, mode=0, stream=None):
        """mode 0: uncompressed record stream; mode 1: BGZF blocks."""
        self.ctx._check(lib().glx_encode_bcf_on(self.ctx.handle, self.handle, int(mode), C.c_void_p(stream or 0)))

    def bcf_ready(self):
        """True once the encoder has finished (bcf_length will not block)."""
        rc = lib().glx_bcf_ready(self.ctx.handle, self.handle)
        if rc < 0:
            self.ctx._check(rc)
        return rc == 1

    def bcf_length(self):
        """(encoded bytes, uncompressed stream bytes, BGZF blocks); waits for the encoder's size event."""
        n, r, b = C.c_uint64(), C.c_uint64(), C.c_uint32()
        self.ctx._check(lib().glx_bcf_length(self.ctx.handle, self.handle, C.byref(n), C.byref(r), C.byref(b)))
        return n.value, r.value, b.value

    @property
    def bcf_bound(self):
        return lib().glx_bcf_bound(self.handle)

    def download_bcf(self, host_buf=None, stream=None, sync=True):
        """The encoded bytes. host_buf: a PinnedBuffer to receive them (asynchronously when sync is False); default: a fresh bytes object."""
        if host_buf is None:
            n, _, _ = self.bcf_length()
            tmp = C.create_string_buffer(max(1, n))
            self.ctx._check(lib().glx_download_bcf_on(self.ctx.handle, self.handle, tmp, n, C.c_void_p(stream or 0), 1))
            return tmp.raw[:n]
        self.ctx._check(lib().glx_download_bcf_on(self.ctx.handle, self.handle, C.c_void_p(host_buf.ptr), host_buf.nbytes, C.c_void_p(stream or 0), 1 if sync else 0))
        return None

    def bcf_record_offsets(self):
        off = (C.c_uint64 * (self._n_sites + 1))()
        self.ctx._check(lib().glx_bcf_record_offsets(self.ctx.handle, self.handle, off))
        return list(off)

    def set_profile(self, on=True):
        self.ctx._check(lib().glx_set_profile(self.ctx.handle, self.handle, 1 if on else 0))

    def kernel_ms(self):
        """(resolve, tiled, single-record worklist, general worklist) kernel ms of the last launch, from CUDA events on
        the launching stream."""
        ms = (C.c_float * 4)()
        self.ctx._check(lib().glx_kernel_ms(self.ctx.handle, self.handle, ms))
        return tuple(ms)

    def upload_kernel_ms(self):
        """(glx_prepare_inputs_kernel, glx_block_hints_kernel) ms of this batch's upload (context option profile_upload)."""
        ms = (C.c_float * 2)()
        self.ctx._check(lib().glx_upload_kernel_ms(self.ctx.handle, self.handle, ms))
        return tuple(ms)

    def work_count(self):
        n = C.c_uint64()
        self.ctx._check(lib().glx_work_count(self.ctx.handle, self.handle, C.byref(n)))
        return n.value

    def fused_count(self):
        n = C.c_uint64()
        self.ctx._check(lib().glx_fused_count(self.ctx.handle, self.handle, C.byref(n)))
        return n.value

    @property
    def out_bytes(self):
        return lib().glx_out_bytes(self.handle)

    @property
    def in_bytes(self):
        return lib().glx_in_bytes(self.handle)

    @property
    def launch_count(self):
        return lib().glx_launch_count(self.handle)

    def close(self):
        if self.handle:
            lib().glx_free_batch(self.ctx.handle, self.handle)
            self.handle = None


class Context:
    """glx_ctx: one per process and GPU."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().glx_create(C.byref(h), int(device))
        if rc != GLX_OK:
            raise GlxError(rc, "glx_create failed (no usable CUDA device?)")
        self.handle = h

    def _check(self, rc):
        if rc != GLX_OK:
            msg = lib().glx_last_error(self.handle)
            raise GlxError(rc, msg.decode(errors="replace") if msg else "")

    PATHS = {"sig": {}, "sig-unfused": {"no_fuse": 1, "no_resolve_sig": 1}, "dynamic": {"force_dynamic_tiled": 1}, "general": {"force_general": 1}}

    def set_option(self, name, value):
        """Kernel-path switch for batches uploaded afterwards (glx_set_option)."""
        self._check(lib().glx_set_option(self.handle, _b(name), int(value)))

    def select_path(self, path):
        """Parity tests: 'sig' (default kernels), 'sig-unfused', 'dynamic' (dynamic tiled kernel), 'general' (general kernel alone)."""
        for k in ("no_fuse", "no_resolve_sig", "force_dynamic_tiled", "force_general"):
            self.set_option(k, self.PATHS[path].get(k, 0))

    def bgzf_phase_cycles(self):
        out = (C.c_uint64 * 16)()
        self._check(lib().glx_debug_bgzf_phases(self.handle, out))
        return [int(x) for x in out]

    def selftest_diploid(self, n_allele):
        """[(alleles_gt(i,j), alleles_gt(j,i))] in genotype order, from the device's index math."""
        n = n_allele * (n_allele + 1) // 2
        buf = (C.c_uint32 * n)()
        self._check(lib().glx_selftest_diploid(self.handle, n_allele, buf))
        return [(v & 0xffff, v >> 16) for v in buf]

    def genotype_batch(self, batch, out):
        """Host buffers in, host buffers out (H2D + kernels + D2H)."""
        self._check(lib().glx_genotype_batch(self.handle, batch.config_ptr, batch.records_ptr, batch.sites_ptr, out.ptr))

    def upload(self, batch):
        d = C.c_void_p()
        self._check(lib().glx_upload(self.handle, batch.config_ptr, batch.records_ptr, batch.sites_ptr, C.byref(d)))
        return DeviceBatch(self, d)

    def upload_wire(self, batch, stream=None, sync=True):
        """Upload through the compact wire form; the device expands it into the same record store glx_upload builds."""
        d = C.c_void_p()
        wp, _ = batch.wire()
        self._check(lib().glx_upload_wire_on(self.handle, batch.config_ptr, C.c_void_p(wp), batch.sites_ptr, C.c_void_p(stream or 0), 1 if sync else 0, C.byref(d)))
        return DeviceBatch(self, d)

    def compare_records(self, dev, batch):
        """Names of the device's record-store arrays that differ from the batch's host arrays (test aid)."""
        names = ["beg", "end", "maxend", "qual", "n_allele", "flags", "filt", "gt", "allele_off", "col_val", "col_len", "pool", "al_len", "al_flags",
                 "al_prefix", "al_str", "al_pool"]
        diff = C.c_uint32()
        self._check(lib().glx_debug_compare_records(self.handle, dev.handle, batch.records_ptr, C.byref(diff)))
        return [n for i, n in enumerate(names) if diff.value >> i & 1] + (["size"] if diff.value == 0xffffffff else [])

    def upload_async(self, batch, stream):
        """Enqueue the batch's H2D copies on `stream`; the batch must stay alive (and should be pinned) until synchronised."""
        d = C.c_void_p()
        self._check(lib().glx_upload_on(self.handle, batch.config_ptr, batch.records_ptr, batch.sites_ptr, C.c_void_p(stream), 0, C.byref(d)))
        return DeviceBatch(self, d)

    def stream_create(self):
        s = lib().glx_stream_create(self.handle)
        if not s:
            raise GlxError(-100, "glx_stream_create")
        return s

    def mark(self, stream=None):
        """a timing / ordering mark (CUDA event) recorded on the stream"""
        lib().glx_mark.restype = C.c_void_p
        m = lib().glx_mark(self.handle, C.c_void_p(stream or 0))
        if not m:
            raise GlxError(-100, "glx_mark")
        return m

    def mark_ms(self, a, b):
        lib().glx_mark_ms.restype = C.c_float
        return float(lib().glx_mark_ms(self.handle, C.c_void_p(a), C.c_void_p(b)))

    def mark_wait(self, m):
        self._check(lib().glx_mark_wait(self.handle, C.c_void_p(m)))

    def mark_free(self, m):
        lib().glx_mark_free(self.handle, C.c_void_p(m))

    def stream_destroy(self, s):
        lib().glx_stream_destroy(self.handle, C.c_void_p(s))

    def stream_sync(self, s):
        self._check(lib().glx_stream_sync(self.handle, C.c_void_p(s)))

    def synchronize(self):
        self._check(lib().glx_synchronize(self.handle))

    @property
    def stream_ptr(self):
        return lib().glx_stream(self.handle)

    def close(self):
        if self.handle:
            lib().glx_destroy(self.handle)
            self.handle = None


class RangeStream:
    """glxh_stream: the streamed range-batch driver (depth batches in flight on one device, retired in order).
    mode: 'columns' | 'bcf' (uncompressed BCF2 records) | 'bgzf' (BGZF blocks)."""
    MODES = {"columns": 0, "bcf": 2, "bgzf": 3}

    def __init__(self, device=0, mode="bgzf", depth=3, capture=False, wire=False):
        h = C.c_void_p()
        self._err = C.create_string_buffer(1024)
        rc = lib().glxh_stream_open(int(device), self.MODES[mode], int(depth), 1 if capture else 0, C.byref(h), self._err, len(self._err))
        if rc != GLX_OK:
            raise GlxError(rc, self._err.value.decode(errors="replace"))
        self.handle = h
        self._keep = []
        if wire:
            lib().glxh_stream_use_wire(h, 1)

    def submit(self, batch):
        self._keep.append(batch)  # the batch's arrays must outlive its time in flight
        if len(self._keep) > 16:
            self._keep.pop(0)
        rc = lib().glxh_stream_submit(self.handle, batch.handle, self._err, len(self._err))
        if rc != GLX_OK:
            raise GlxError(rc, self._err.value.decode(errors="replace"))

    def drain(self):
        rc = lib().glxh_stream_drain(self.handle, self._err, len(self._err))
        if rc != GLX_OK:
            raise GlxError(rc, self._err.value.decode(errors="replace"))

    def stats(self):
        st = (C.c_uint64 * 8)()
        lib().glxh_stream_stats(self.handle, st)
        return dict(zip(("batches", "batches_retired", "sites", "cells", "records", "h2d_bytes", "d2h_bytes", "raw_bcf_bytes"), [int(x) for x in st]))

    def trace(self, on=True):
        """start (or stop) the per-batch timeline; rows come from trace_rows() once batches have retired"""
        lib().glxh_stream_trace(self.handle, 1 if on else 0)

    def trace_rows(self):
        """[(upload begins, upload done, kernels done, encoder done, download done | host: submit entered, returned, download queued)] ms"""
        n = lib().glxh_stream_trace_rows(self.handle, None, 0)
        buf = (C.c_float * (8 * max(1, n)))()
        lib().glxh_stream_trace_rows(self.handle, buf, n)
        return [tuple(round(float(buf[8 * i + k]), 3) for k in range(8)) for i in range(n)]

    def output(self):
        p, n = C.c_void_p(), C.c_uint64()
        lib().glxh_stream_output(self.handle, C.byref(p), C.byref(n))
        return C.string_at(p.value, n.value) if n.value else b""

    def close(self):
        if self.handle:
            lib().glxh_stream_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def genotype_sites(sites_text, datasets, filename, preset=None, config_text=None, device=0, batch_sites=0, test_ingest_rules=False, abort=None):
    """Service::genotype_sites on the device path (glxh_genotype_sites): text in, pVCF file out. `abort` is an optional
    ctypes c_int polled between range batches. Returns {sites, rows, batches, cells}."""
    L = lib()
    n = len(datasets)
    names = (C.c_char_p * max(n, 1))(*[_b(d[0]) for d in datasets])
    texts = (C.c_char_p * max(n, 1))(*[_b(d[1]) for d in datasets])
    err = C.create_string_buffer(1024)
    st = (C.c_uint64 * 4)()
    rc = L.glxh_genotype_sites(_b(preset), _b(config_text), _b(sites_text), n, names, texts, _b(filename), int(device), int(batch_sites),
                               1 if test_ingest_rules else 0, C.addressof(abort) if abort is not None else None, st, err, len(err))
    if rc != GLX_OK:
        raise GlxError(rc, err.value.decode(errors="replace"))
    return {"sites": st[0], "rows": st[1], "batches": st[2], "cells": st[3]}


def out_bytes_of(batch):
    """Bytes of the packed output columns of a batch (GT + RNC + lifted FORMAT values)."""
    o = batch.alloc_out()
    try:
        return int(sum(a.nbytes for k, a in o.arrays().items() if k not in ("allele_used", "site_flags")))
    finally:
        o.close()


def load_config(name=None, config_text=None, more_PL=False, squeeze=False, trim_uncalled_alleles=False):
    """glnexus_cli's load_config: preset and/or config text, then the -P/-S/-a switches OR-ed in; returns the config text."""
    t = C.c_void_p()
    rc = lib().glxh_load_config(_b(name), _b(config_text), int(more_PL), int(squeeze), int(trim_uncalled_alleles), C.byref(t))
    if rc != GLX_OK:
        raise GlxError(rc, f"load_config({name})")
    try:
        return C.string_at(t.value).decode()
    finally:
        lib().glxh_free_text(t)


def bgzf_stored(data):
    """data as BGZF blocks of stored deflate blocks (the header block of a BCF file)."""
    p, n = C.c_void_p(), C.c_uint64()
    rc = lib().glxh_bgzf_stored(data, len(data), C.byref(p), C.byref(n))
    if rc != GLX_OK:
        raise GlxError(rc, "glxh_bgzf_stored")
    try:
        return C.string_at(p.value, n.value)
    finally:
        lib().glxh_free_text(p)


def bgzf_eof():
    p, n = C.c_void_p(), C.c_uint64()
    lib().glxh_bgzf_eof(C.byref(p), C.byref(n))
    try:
        return C.string_at(p.value, n.value)
    finally:
        lib().glxh_free_text(p)


def partition_cohort(segments, n_parts):
    """Cut points [(segment, site)] * (n_parts + 1) of equal output weight over a cohort held as consecutive segment batches."""
    arr = (C.c_void_p * len(segments))(*[b.handle for b in segments])
    cs, ci = (C.c_uint32 * (n_parts + 1))(), (C.c_uint32 * (n_parts + 1))()
    rc = lib().glxh_partition_cohort(arr, len(segments), n_parts, cs, ci)
    if rc != GLX_OK:
        raise GlxError(rc, "glxh_partition_cohort")
    return list(zip(cs, ci))


def preset_text(name):
    t = C.c_void_p()
    rc = lib().glxh_preset_text(_b(name), C.byref(t))
    if rc != GLX_OK:
        raise GlxError(rc, f"unknown configuration preset {name}")
    try:
        return C.string_at(t.value).decode()
    finally:
        lib().glxh_free_text(t)


def wire_shape_cap(cap=255):
    """Test aid: rows of the shape table of wire forms built from now on (default and maximum 255)."""
    lib().glxh_debug_wire_shape_cap(cap)


def pack_threads(n=0):
    """Test aid: threads of the record packer for batches built from now on (0 = chosen from the input's size)."""
    lib().glxh_debug_pack_threads(n)
