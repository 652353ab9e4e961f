"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(glxh?_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["glx.h", "glx_host.h"])
def test_exports(built, header):
    lib = C.CDLL(os.path.join(ROOT, "glnexus_b200", "libglx.so"))
    names = declared_functions(header)
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/{header} but not exported: {missing}"


def test_no_gpu_fails_loudly(built):
    """Without a CUDA device the product refuses to run; it never falls back to a host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from glnexus_b200 import Context, GlxError
    with pytest.raises(GlxError) as ei:
        Context(0)
    assert ei.value.code == -100


def test_presets_load(built):
    from glnexus_b200 import preset_text, GlxError
    for name in ("gatk", "gatk_unfiltered", "xAtlas", "xAtlas_unfiltered", "weCall", "weCall_unfiltered", "DeepVariant", "DeepVariantWGS",
                 "DeepVariantWES", "DeepVariantWES_MED_DP", "DeepVariant_unfiltered", "Strelka2", "GxS"):
        t = preset_text(name)
        assert "revise_genotypes" in t and "field\t" in t
    dv = preset_text("DeepVariant")
    assert "snv_prior_calibration 0.6000" in dv and "indel_prior_calibration 0.44999" in dv and "more_PL 1" in dv and "trim_uncalled_alleles 1" in dv
    g = preset_text("gatk")
    assert "required_dp 1" in g and "revise_genotypes 1" in g and "field\tSB\tSB" in g
    with pytest.raises(GlxError):
        preset_text("nonesuch")


def test_service_fails_loudly_without_gpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import glnexus_b200
    import golden_util as gu
    case = gu.load_case("deepvariant")
    ds = [(d["name"], d["vcf"]) for d in case["datasets"]]
    with pytest.raises(glnexus_b200.GlxError) as ei:
        glnexus_b200.genotype_sites(case["sites_text"], ds, str(tmp_path / "o.vcf"), preset=case["preset"], config_text=case["config_text"])
    assert ei.value.code == -1 and "no host fallback" in str(ei.value)
