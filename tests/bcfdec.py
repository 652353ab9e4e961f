"""Independent BCF2 reader (test infrastructure): BGZF -> BCF header + records -> VCF text rows, written from the
BCFv2.2 specification. It renders rows the way htslib's vcf_format does (and glx::vcf_rows_text restates), so a decoded
BCF can be compared string for string with the text assembler and, through vcfcmp, with the reference's truth pVCFs."""
import struct
import zlib

INT_MISSING = {1: -128, 2: -32768, 3: -2147483648}
INT_VEND = {1: -127, 2: -32767, 3: -2147483647}
FLOAT_MISSING, FLOAT_VEND = 0x7F800001, 0x7F800002
FMT = {1: ("b", 1), 2: ("h", 2), 3: ("i", 4), 5: ("I", 4), 7: ("c", 1)}


def bgzf_decompress(data):
    """Concatenated BGZF blocks -> bytes; checks the BC extra field, BSIZE, CRC32 and ISIZE of every block."""
    out, p, n_blocks = [], 0, 0
    while p < len(data):
        assert data[p:p + 4] == b"\x1f\x8b\x08\x04", f"bad gzip/BGZF magic at {p}"
        xlen = struct.unpack_from("<H", data, p + 10)[0]
        extra = data[p + 12:p + 12 + xlen]
        assert extra[:4] == b"BC\x02\x00", "BC subfield missing"
        bsize = struct.unpack_from("<H", extra, 4)[0] + 1
        cdata = data[p + 12 + xlen:p + bsize - 8]
        crc, isize = struct.unpack_from("<II", data, p + bsize - 8)
        raw = zlib.decompress(cdata, wbits=-15)
        assert len(raw) == isize and len(raw) <= 65536, "ISIZE mismatch"
        assert zlib.crc32(raw) == crc, f"CRC mismatch in block {n_blocks}"
        out.append(raw)
        p += bsize
        n_blocks += 1
    return b"".join(out), n_blocks


def fmt_float(bits):
    f = struct.unpack("<f", struct.pack("<I", bits))[0]
    return "%g" % f


class Reader:
    def __init__(self, data, p=0):
        self.d, self.p = data, p

    def typed_desc(self):
        b = self.d[self.p]
        self.p += 1
        n, t = b >> 4, b & 15
        if n == 15:
            n = self.typed_int()
        return n, t

    def typed_int(self):
        n, t = self.typed_desc()
        assert n == 1 and t in (1, 2, 3)
        return self.values(1, t)[0]

    def values(self, n, t):
        if t == 0 or n == 0:
            return []
        c, w = FMT[t]
        v = struct.unpack_from("<%d%s" % (n, c), self.d, self.p)
        self.p += n * w
        return list(v)

    def typed_vec(self):
        n, t = self.typed_desc()
        return t, self.values(n, t)

    def typed_str(self):
        t, v = self.typed_vec()
        assert t in (7, 0)
        return b"".join(v).decode()


def parse_header(prologue):
    """(text, dict ids in BCF_DT_ID order, contig names, sample names, end offset) of the uncompressed file start."""
    assert prologue[:5] == b"BCF\x02\x02"
    l_text = struct.unpack_from("<I", prologue, 5)[0]
    text = prologue[9:9 + l_text]
    assert text[-1:] == b"\x00"
    text = text[:-1].decode()
    ids, contigs, samples = [], [], []
    for line in text.split("\n"):
        if line.startswith(("##FILTER=", "##INFO=", "##FORMAT=")):
            i = line.index("ID=") + 3
            name = line[i:min(x for x in (line.find(",", i), line.find(">", i)) if x >= 0)]
            if name not in ids:
                if "IDX=" in line:  # htslib writes the index explicitly; it must agree with the order of appearance
                    assert int(line[line.rindex("IDX=") + 4:line.rindex(">")]) == len(ids), line
                ids.append(name)
        elif line.startswith("##contig="):
            i = line.index("ID=") + 3
            contigs.append(line[i:min(x for x in (line.find(",", i), line.find(">", i)) if x >= 0)])
        elif line.startswith("#CHROM"):
            samples = line.split("\t")[9:]
    return text, ids, contigs, samples, 9 + l_text


def vcf_header_from_bcf(text):
    """Header text with the IDX attributes removed (what bcftools view prints)."""
    import re
    return re.sub(r",IDX=\d+>", ">", text)


def int_vec_text(vals, t):
    out = []
    for v in vals:
        if v == INT_VEND[t]:
            break
        out.append("." if v == INT_MISSING[t] else str(v))
    return ",".join(out) if out else "."


def float_vec_text(vals):
    out = []
    for v in vals:
        if v == FLOAT_VEND:
            break
        out.append("." if v == FLOAT_MISSING else fmt_float(v))
    return ",".join(out) if out else "."


def records_to_rows(stream, ids, contigs, n_samples_expected=None):
    """Uncompressed BCF record stream -> list of VCF rows (no trailing newline)."""
    rows, p = [], 0
    while p < len(stream):
        l_shared, l_indiv = struct.unpack_from("<II", stream, p)
        r = Reader(stream, p + 8)
        chrom, pos, rlen = struct.unpack_from("<iii", stream, r.p)
        qual_bits, = struct.unpack_from("<I", stream, r.p + 12)
        nai, nfs = struct.unpack_from("<II", stream, r.p + 16)
        r.p += 24
        n_allele, n_info, n_fmt, n_sample = nai >> 16, nai & 0xffff, nfs >> 24, nfs & 0xffffff
        if n_samples_expected is not None:
            assert n_sample == n_samples_expected
        vid = r.typed_str()
        alleles = [r.typed_str() for _ in range(n_allele)]
        assert rlen == len(alleles[0])
        ft, fv = r.typed_vec()
        filt = ";".join(ids[i] for i in fv) if fv else "."
        info = []
        for _ in range(n_info):
            key = ids[r.typed_int()]
            t, v = r.typed_vec()
            info.append(key + "=" + (float_vec_text(v) if t == 5 else int_vec_text(v, t)))
        assert r.p == p + 8 + l_shared, "l_shared mismatch"
        keys, cols = [], []
        for _ in range(n_fmt):
            key = ids[r.typed_int()]
            n, t = r.typed_desc()
            vals = r.values(n * n_sample, t)
            col = []
            for s in range(n_sample):
                v = vals[s * n:(s + 1) * n]
                if key == "GT":
                    parts = []
                    for j, x in enumerate(v):
                        if x == INT_VEND[t]:
                            break
                        parts.append(("|" if (x & 1) and j else "/") if j else "")
                        parts.append(str((x >> 1) - 1) if x >> 1 else ".")
                    col.append("".join(parts))
                elif t == 7:
                    sv = b"".join(v).split(b"\x00")[0].decode()
                    col.append(sv if sv else ".")
                elif t == 5:
                    col.append(float_vec_text(v))
                else:
                    col.append(int_vec_text(v, t))
            keys.append(key)
            cols.append(col)
        assert r.p == p + 8 + l_shared + l_indiv, "l_indiv mismatch"
        row = [contigs[chrom], str(pos + 1), vid if vid else ".", alleles[0], ",".join(alleles[1:]) if n_allele > 1 else ".",
               fmt_float(qual_bits) if qual_bits != FLOAT_MISSING else ".", filt, ";".join(info) if info else ".", ":".join(keys)]
        for s in range(n_sample):
            row.append(":".join(c[s] for c in cols))
        rows.append("\t".join(row))
        p = r.p
    return rows


def bcf_file_to_vcf_text(data):
    """A complete BGZF-compressed BCF file -> VCF text (header without IDX + rows)."""
    raw, _ = bgzf_decompress(data)
    text, ids, contigs, samples, end = parse_header(raw)
    rows = records_to_rows(raw[end:], ids, contigs, len(samples))
    return vcf_header_from_bcf(text) + "".join(r + "\n" for r in rows)
