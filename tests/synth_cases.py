"""Seeded synthetic cohorts shared by the parity tests, smoke() and bench.py (shapes of BASELINE.json configs[1..4])."""
from glnexus_b200 import Batch, SynthParams

# name -> (preset, overrides); small sizes are applied by the caller
SHAPES = {
    # configs[1]: DeepVariant-style exome-like cohort, biallelic-dominant, dense sites
    "c2": ("DeepVariant", dict(seed=20200210, max_alt=2, frac_multiallelic=0.005, frac_indel=0.025, af_max=0.1, band_mean_len=120.0,
                               frac_uncovered=0.01)),
    # configs[2]: WGS cohort with reference bands / MIN_DP, multi-record cells
    "c3": ("DeepVariant", dict(seed=20200211, max_alt=3, frac_multiallelic=0.02, frac_indel=0.1, af_max=0.1, band_mean_len=400.0,
                               frac_uncovered=0.01, frac_two_record_cells=0.004)),
    # configs[3]: large sparse cohort, revise_genotypes AF priors on, hom-ALT and lost alleles
    "c4": ("DeepVariant", dict(seed=20200212, max_alt=2, frac_multiallelic=0.01, frac_indel=0.05, af_max=0.01, band_mean_len=400.0,
                               frac_uncovered=0.01, frac_homalt=0.2, frac_lost_allele=0.02)),
    # configs[4]: multi-allelic stress, --config gatk, up to 6 ALT alleles (28 diploid genotypes per PL vector)
    "c5": ("gatk", dict(seed=20200213, max_alt=6, frac_multiallelic=0.85, frac_indel=0.3, af_max=0.3, band_mean_len=200.0,
                        frac_uncovered=0.01, frac_homalt=0.2, gatk_style=1)),
}


def make(shape, n_samples, n_sites, region_len, region_beg=10_000_000, preset=None, **extra):
    pre, kw = SHAPES[shape]
    p = SynthParams(n_samples=n_samples, n_sites=n_sites, region_beg=region_beg, region_len=region_len, seed=0, max_alt=1,
                    frac_multiallelic=0.0, frac_indel=0.0, af_max=0.5, band_mean_len=120.0, frac_uncovered=0.0,
                    frac_two_record_cells=0.0, frac_lost_allele=0.0, frac_homalt=0.0, gatk_style=0)
    for k, v in {**kw, **extra}.items():
        setattr(p, k, v)
    return Batch.synthetic(p, preset or pre)
