"""Pin the CPU oracle against the reference's golden vectors (CPU only).

test/data/gvcf_test_cases/*.yml -> tests/golden/*.json (make_golden.py). Each case: input gVCFs + config +
truth_unified_sites run through the product's packer, the ORACLE, and the product's pVCF assembler, then
compared with truth_output_vcf on the whitelisted fields exactly as test/gvcf_test_cases.cc:415-461 does.
"""
import pytest

import golden_util as gu
import oracle_lib
import vcfcmp


@pytest.mark.parametrize("name", gu.case_names())
def test_oracle_matches_truth_vcf(built, name):
    case = gu.load_case(name)
    texts = []
    for batch in gu.batches_of_case(case):
        out = batch.alloc_out()
        oracle_lib.genotype_batch(batch, out)
        texts.append(batch.assemble_vcf(out))
    got = gu.merge_vcf_texts(texts)
    diffs = vcfcmp.compare(got, case["truth_vcf"], case["formats"], case["infos"])
    assert not diffs, "\n".join(diffs[:20]) + "\n---- got ----\n" + "\n".join(l for l in got.split("\n") if not l.startswith("##"))
