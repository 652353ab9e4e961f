"""Semantic pVCF comparison with the rules of the reference's test/testOutputVcf.py (which needs PyVCF,
absent here): rows keyed by (CHROM, POS, POS+len(REF)); REF, ALT, QUAL, FILTER always compared; INFO/FORMAT
fields compared only when whitelisted AND declared in both headers; values parsed by header Type/Number the
way PyVCF does ('.' -> None, Integer -> int, Float -> float, comma lists)."""
import re


def _conv(tok, typ):
    if tok == "." or tok == "":
        return None
    if typ == "Integer":
        try:
            return int(tok)
        except ValueError:
            return float(tok)
    if typ == "Float":
        return float(tok)
    return tok


def _parse_val(text, typ, number):
    if text is None or text == "." or text == "":
        return None
    if typ in ("String", "Character") and "," not in text:
        return text
    if number == "1" or "," not in text:
        v = _conv(text, typ)
        return v if (number == "1" or typ in ("String", "Character")) else [v]
    return [_conv(t, typ) for t in text.split(",")]


class Vcf:
    def __init__(self, text):
        self.infos, self.formats, self.samples, self.rows = {}, {}, [], []
        for line in text.split("\n"):
            if not line.strip():
                continue
            if line.startswith("##"):
                m = re.match(r"##(INFO|FORMAT)=<ID=([^,]+),Number=([^,]+),Type=([^,>]+)", line)
                if m:
                    (self.infos if m.group(1) == "INFO" else self.formats)[m.group(2)] = (m.group(4), m.group(3))
                continue
            t = line.split("\t")
            if line.startswith("#"):
                self.samples = t[9:]
                continue
            self.rows.append(t)

    def key(self, t):
        return (t[0], int(t[1]), int(t[1]) + len(t[3]))

    def parse_row(self, t):
        info = {}
        if t[7] not in (".", ""):
            for kv in t[7].split(";"):
                k, _, v = kv.partition("=")
                typ, num = self.infos.get(k, ("String", "."))
                pv = _parse_val(v, typ, num)
                if typ in ("Integer", "Float") and num != "1" and not isinstance(pv, list):
                    pv = [pv]
                info[k] = pv
        keys = t[8].split(":") if len(t) > 8 else []
        samples = []
        for s in t[9:]:
            vals = s.split(":")
            d = {}
            for i, k in enumerate(keys):
                txt = vals[i] if i < len(vals) else None
                if k == "GT":
                    d[k] = txt
                    continue
                typ, num = self.formats.get(k, ("String", "."))
                d[k] = _parse_val(txt, typ, num)
            samples.append(d)
        qual = None if t[5] == "." else (int(t[5]) if re.fullmatch(r"-?\d+", t[5]) else float(t[5]))
        flt = None if t[6] == "." else ([] if t[6] == "PASS" else t[6].split(";"))
        return {"REF": t[3], "ALT": t[4].split(","), "QUAL": qual, "FILTER": flt, "INFO": info, "FORMAT": keys, "samples": samples}


def compare(input_text, truth_text, formats, infos):
    """Returns a list of human-readable differences (empty = match)."""
    a, b = Vcf(input_text), Vcf(truth_text)
    diffs = []
    if set(a.samples) != set(b.samples):
        return [f"samples differ: {a.samples} vs {b.samples}"]
    if a.samples != b.samples:
        return [f"sample order differs: {a.samples} vs {b.samples}"]
    info_fs = set(infos) & set(a.infos) & set(b.infos)
    format_fs = set(formats) & (set(a.formats) | {"GT"}) & (set(b.formats) | {"GT"})
    truth = {b.key(t): t for t in b.rows}
    for t in a.rows:
        k = a.key(t)
        if k not in truth:
            diffs.append(f"unmatched input row {k}")
            continue
        ra, rb = a.parse_row(t), b.parse_row(truth.pop(k))
        for f in ("REF", "ALT", "QUAL", "FILTER"):
            if ra[f] != rb[f]:
                diffs.append(f"{k}: {f} {ra[f]} != {rb[f]}")
        for f in sorted(info_fs):
            pa, pb = f in ra["INFO"], f in rb["INFO"]
            if pa != pb:
                diffs.append(f"{k}: INFO {f} present {pa} vs {pb}")
            elif pa and ra["INFO"][f] != rb["INFO"][f]:
                diffs.append(f"{k}: INFO {f} {ra['INFO'][f]} != {rb['INFO'][f]}")
        for f in sorted(format_fs):
            pa, pb = f in ra["FORMAT"], f in rb["FORMAT"]
            if not pa and not pb:
                continue
            if pa != pb:
                diffs.append(f"{k}: FORMAT {f} present {pa} vs {pb}")
                continue
            for sn, sa, sb in zip(a.samples, ra["samples"], rb["samples"]):
                if sa.get(f) != sb.get(f):
                    diffs.append(f"{k}: {sn} {f} {sa.get(f)} != {sb.get(f)}")
    for k in truth:
        diffs.append(f"unmatched truth row {k}")
    return diffs
