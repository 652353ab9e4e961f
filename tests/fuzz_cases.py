"""Randomised small cohorts as gVCF TEXT (the path a real decoder feeds) with the irregular features the seeded
synthetic generator does not produce: haploid / half-missing / phased GTs, records without PL, AD, DP or MIN_DP,
0/0 variant records, overlapping and adjacent variant records, long (> 8 bp) alleles, deletions spanning later sites,
several representations of one allele in the unification map, lost alleles, monoallelic sites, and config toggles
(squeeze, top_two_half_calls, required_dp, allow_partial_data, more_PL, revise_genotypes)."""
import random

HEADER = """##fileformat=VCFv4.2
##FILTER=<ID=PASS,Description="All filters passed">
##FILTER=<ID=RefCall,Description="">
##FILTER=<ID=LowQual,Description="">
##ALT=<ID=NON_REF,Description="">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="">
##FORMAT=<ID=DP,Number=1,Type=Integer,Description="">
##FORMAT=<ID=MIN_DP,Number=1,Type=Integer,Description="">
##FORMAT=<ID=AD,Number=R,Type=Integer,Description="">
##FORMAT=<ID=PL,Number=G,Type=Integer,Description="">
##FORMAT=<ID=SB,Number=4,Type=Integer,Description="">
##INFO=<ID=END,Number=1,Type=Integer,Description="">
##contig=<ID=chrF,length=1000000>
"""
BASES = "ACGT"


def _seq(rng, n):
    return "".join(rng.choice(BASES) for _ in range(n))


def make_case(seed, preset=None, multi=False):
    """multi: datasets of one to three samples each (multi-sample gVCFs): every record carries all of its dataset's samples, each
    with its own GT / GQ / DP / AD / PL / SB values under the record's shared alleles and FORMAT keys."""
    rng = random.Random(seed)
    n_samples = rng.randint(1, 9) if not multi else rng.randint(1, 5)
    region = (1000, 1000 + rng.randint(40, 160))
    ref_genome = _seq(rng, region[1] - region[0] + 64)

    def refseq(beg, end):  # 0-based half-open, relative to region start
        return ref_genome[beg - region[0]:end - region[0]]

    # ---- unified sites ----
    sites = []
    pos = region[0] + rng.randint(0, 5)
    while pos < region[1] - 12:
        kind = rng.random()
        rlen = 1 if kind < 0.6 else rng.randint(2, 12)
        ref = refseq(pos, pos + rlen)
        n_alt = 1 if rng.random() < 0.6 else rng.randint(2, 4)
        alts = []
        for _ in range(n_alt):
            r = rng.random()
            if r < 0.5:
                a = rng.choice([b for b in BASES if b != ref[0]]) + ref[1:]
            elif r < 0.75 and rlen > 1:
                a = ref[:rng.randint(1, rlen - 1)]
            else:
                a = ref + _seq(rng, rng.randint(1, 11))
            if a != ref and a not in alts:
                alts.append(a)
        if not alts:
            alts = [rng.choice([b for b in BASES if b != ref[0]]) + ref[1:]]
        mono = rng.random() < 0.12
        if mono:
            alts = alts[:1]
        sites.append({"beg": pos, "end": pos + rlen, "ref": ref, "alts": alts, "mono": mono,
                      "af": [round(rng.choice([1e-5, 0.001, 0.01, 0.2, 0.5]), 6) for _ in alts],
                      "lost_af": rng.choice([0.0, 0.001, 0.05]), "qual": rng.randint(0, 999)})
        pos += rlen + rng.randint(0, 14)
    lines = []
    for s in sites:
        lines.append("\t".join(["site", "chrF", str(s["beg"]), str(s["end"]), str(s["qual"]), "1" if s["mono"] else "0", repr(s["lost_af"])]))
        lines.append("\t".join(["allele", s["ref"], str(s["beg"]), str(s["end"]), s["ref"], "0", "nan"]))
        for a, af in zip(s["alts"], s["af"]):
            lines.append("\t".join(["allele", a, str(s["beg"]), str(s["end"]), a, str(rng.randint(1, 999)), repr(af)]))
        # alternative representations: the same allele anchored one base earlier / with a shorter REF
        for j, a in enumerate(s["alts"]):
            if rng.random() < 0.35 and s["beg"] - 1 >= region[0]:
                pre = refseq(s["beg"] - 1, s["beg"])
                lines.append("\t".join(["unif", str(s["beg"] - 1), str(s["end"]), pre + a, str(j + 1)]))
    sites_text = "\n".join(lines) + "\n"

    # ---- per-sample gVCFs ----
    datasets = []
    for n in range(n_samples):
        name = f"S{n:02d}"
        ns = rng.randint(1, 3) if multi else 1
        rows = []
        pos = region[0]
        gatk = rng.random() < 0.5
        sym = "<NON_REF>" if gatk else "<*>"
        while pos < region[1]:
            r = rng.random()
            site_here = [s for s in sites if s["beg"] == pos]
            if site_here and site_here[0]["end"] - site_here[0]["beg"] >= 3 and r < 0.08:
                # two short variant records inside one multi-base site: disjoint ranges => UnphasedVariants
                s = site_here[0]
                for p2 in (s["beg"], s["beg"] + rng.randint(1, s["end"] - s["beg"] - 1)):
                    rb = refseq(p2, p2 + 1)
                    alt = rng.choice([b for b in BASES if b != rb])
                    rows.append((p2, "\t".join(["chrF", str(p2 + 1), ".", rb, alt + "," + sym, "30", ".", ".", "GT:GQ:DP:AD:PL"] +
                                                [f"{'0/1' if k == 0 or rng.random() < 0.7 else rng.choice(['0/0', '1/1', './.'])}:{rng.randint(0, 99)}:{rng.randint(1, 60)}:"
                                                 f"{rng.randint(0, 30)},{rng.randint(0, 30)},0:30,0,300,99,300,990" for k in range(ns)])))
                pos = s["end"]
                continue
            if site_here and r < 0.55:
                s = site_here[0]
                variants = rng.randint(1, 2) if rng.random() < (0.4 if multi else 0.12) else 1
                for _ in range(variants):
                    rep = rng.random()
                    beg, ref = s["beg"], s["ref"]
                    pool = list(s["alts"])
                    if rng.random() < 0.15:
                        pool.append(ref + "TTTTTTTTTTTT")  # not in the unification: lost allele
                    if rng.random() < 0.1 and len(ref) > 1:
                        pool.append(ref[0])  # a deletion that may or may not be unified
                    k = rng.randint(1, min(3, len(pool)))
                    alts = rng.sample(pool, k)
                    if rep < 0.2 and beg - 1 >= region[0] and any(l.startswith("unif") for l in lines):
                        pre = refseq(beg - 1, beg)
                        beg, ref, alts = beg - 1, pre + ref, [pre + a for a in alts]
                    has_sym = rng.random() < 0.8
                    als = [ref] + alts + ([sym] if has_sym else [])
                    na = len(als)
                    nG = na * (na + 1) // 2
                    fmt = ["GT"]
                    if rng.random() < 0.995:
                        fmt.append("GQ")
                    if rng.random() < 0.9:
                        fmt.append("DP")
                    if rng.random() < 0.9:
                        fmt.append("AD")
                    if rng.random() < 0.99:
                        fmt.append("PL")
                    if gatk and rng.random() < 0.7:
                        fmt.append("SB")
                    hap_rec = ns == 1 and rng.random() < 0.08  # (a multi-sample record keeps one ploidy)
                    all00 = multi and rng.random() < 0.25          # a variant record nobody in the dataset carries
                    cols = []
                    for k in range(ns):
                        if hap_rec:
                            gt = str(rng.randint(0, na - 1 - (1 if has_sym else 0)))  # haploid
                        else:
                            a0 = rng.randint(0, len(alts))
                            a1 = rng.randint(0, len(alts))
                            if all00 or (k > 0 and rng.random() < 0.35):
                                a0 = a1 = 0  # other samples of the dataset are often hom-ref at the site
                            sep = "|" if rng.random() < 0.1 else "/"
                            gt = f"{'.' if rng.random() < 0.06 else a0}{sep}{'.' if rng.random() < 0.06 else a1}"
                        pl = [rng.choice([0, 0, 3, 10, 20, 50, 99, 300, 990]) for _ in range(nG)]
                        if rng.random() < 0.1:
                            pl[rng.randrange(nG)] = "."
                        ad = [rng.randint(0, 40) for _ in range(na)]
                        vals = {"GT": gt, "GQ": str(rng.randint(0, 99)), "DP": str(rng.randint(0, 80)), "AD": ",".join(map(str, ad)),
                                "PL": ",".join(map(str, pl)), "SB": ",".join(str(rng.randint(0, 30)) for _ in range(4))}
                        cols.append(":".join(vals[f] for f in fmt))
                    qual = rng.choice([".", "0", "12.5", "50", "50", "468.22"])
                    flt = rng.choice([".", "PASS", "PASS", "LowQual", "RefCall"])
                    rows.append((beg, "\t".join(["chrF", str(beg + 1), ".", ref, ",".join(als[1:]), qual, flt, ".", ":".join(fmt)] + cols)))
                pos = s["end"] if rng.random() < 0.8 else s["beg"] + 1
                continue
            if r < 0.93:
                ln = min(rng.randint(1, 40), region[1] - pos)
                nxt = [s["beg"] for s in sites if pos < s["beg"] < pos + ln]
                if nxt and rng.random() < 0.6:
                    ln = nxt[0] - pos
                fmt = ["GT"]
                if rng.random() < 0.95:
                    fmt.append("GQ")
                if rng.random() < 0.998:
                    fmt.append("MIN_DP")
                if rng.random() < 0.5:
                    fmt.append("DP")
                if rng.random() < 0.9:
                    fmt.append("PL")
                cols = []
                for k in range(ns):
                    gt = rng.choice(["0/0", "0/0", "0/0", "0/0", "./.", "0", "0|0"] if ns == 1 else ["0/0"] * 12 + ["./.", "0|0"])
                    vals = {"GT": gt, "GQ": str(rng.randint(0, 60)), "MIN_DP": str(rng.randint(0, 50)), "DP": str(rng.randint(0, 60)),
                            "PL": rng.choice(["0,30,300", "0,0,0", "0,9,90", "20,0,50", "0,.,."])}
                    cols.append(":".join(vals[f] for f in fmt))
                alt = sym if rng.random() < 0.95 else "."
                rows.append((pos, "\t".join(["chrF", str(pos + 1), ".", refseq(pos, pos + 1), alt, ".", ".", f"END={pos + ln}", ":".join(fmt)] + cols)))
                pos += ln
            else:
                pos += rng.randint(1, 20)  # coverage gap
        rows.sort(key=lambda x: x[0])
        names = [name] if ns == 1 else [f"{name}_{k}" for k in range(ns)]
        body = HEADER + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n" + "\n".join(r[1] for r in rows) + "\n"
        datasets.append((name, body))

    preset = preset or rng.choice(["DeepVariant", "gatk", "DeepVariant_unfiltered", "gatk_unfiltered", "weCall"])
    toggles = {}
    if rng.random() < 0.2:
        toggles["squeeze"] = 1
    if rng.random() < 0.25:
        toggles["top_two_half_calls"] = 1
    if rng.random() < 0.3:
        toggles["required_dp"] = rng.choice([0, 1, 5, 13])
    if rng.random() < 0.3:
        toggles["allow_partial_data"] = rng.choice([0, 1])
    return sites_text, datasets, preset, toggles


def build_batch(seed, preset=None, multi=False):
    """-> glnexus_b200.Batch for the seed (config = preset text with the toggles applied)."""
    from glnexus_b200 import Batch, preset_text
    sites_text, datasets, preset, toggles = make_case(seed, preset, multi)
    cfg_lines = []
    for line in preset_text(preset).split("\n"):
        key = line.split(" ")[0]
        if key in toggles:
            cfg_lines.append(f"{key} {toggles[key]}")
        else:
            cfg_lines.append(line)
    return Batch.from_text(sites_text, datasets, config_text="\n".join(cfg_lines))
