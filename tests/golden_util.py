import glob
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.json"))
                  if not os.path.basename(p).startswith("_"))


def load_case(name):
    return json.load(open(os.path.join(GOLDEN, name + ".json")))


def contigs_of_sites(case):
    seen = []
    for line in case["sites_text"].split("\n"):
        t = line.split("\t")
        if t[0] == "site" and t[1] not in seen:
            seen.append(t[1])
    return seen


def batches_of_case(case):
    """One glnexus_b200.Batch per contig that has sites (a batch is one contig range)."""
    from glnexus_b200 import Batch
    ds = [(d["name"], d["vcf"]) for d in case["datasets"]]
    out = []
    for contig in contigs_of_sites(case):
        out.append(Batch.from_text(case["sites_text"], ds, preset=case["preset"], config_text=case["config_text"],
                                   contig=contig, test_ingest_rules=True))
    return out


def merge_vcf_texts(texts):
    """Concatenate per-contig pVCF texts (header from the first)."""
    if not texts:
        return ""
    out = texts[0].rstrip("\n").split("\n")
    for t in texts[1:]:
        out += [l for l in t.rstrip("\n").split("\n") if not l.startswith("#")]
    return "\n".join(out) + "\n"
