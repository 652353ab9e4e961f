"""glnexus_b200 — B200-native joint genotyper behind the GLnexus genotype_site interface.

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, built into libglx.so) and the thin
host-side mirror of the reference's interface (api.py). See DESIGN.md.
"""
from .api import Batch, Context, GlxError, Out, Out16, PinnedBuffer, RangeStream, SynthParams, bgzf_eof, bgzf_stored, genotype_sites, lib, load_config, out_bytes_of, partition_cohort, preset_text, pack_threads, wire_shape_cap  # noqa: F401
