// glx_bcf_host.cc — host half of the BCF2 output path: the header Service::genotype_sites prepares
// (prepare_bcf_header, src/service.cc:190-241), its BCF_DT_ID dictionary, the per-site strings the device encoder needs
// (glx_bcf_meta, include/glx.h) and the BGZF framing of the header / EOF blocks. Records are encoded and compressed on the
// device (glx_bcf.cu); nothing here touches per-cell data.
#include <cstring>

#include "glx_host.hpp"

namespace glx {

static std::string attr_of(const std::string& line, const std::string& key) {
    size_t p = line.find(key + "=");
    if (p == std::string::npos) return "";
    p += key.size() + 1;
    return line.substr(p, line.find_first_of(",>", p) - p);
}

// header lines after "##fileformat" in the order prepare_bcf_header appends them (bcf_hdr_init("w") supplies PASS)
static std::vector<std::string> header_lines(const PackedBatch& b) {
    std::vector<std::string> h;
    h.push_back("##FILTER=<ID=PASS,Description=\"All filters passed\">");
    h.push_back("##GLnexusVersion=glnexus_b200");
    h.push_back("##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency estimate for each alternate allele\">");
    h.push_back("##INFO=<ID=AQ,Number=A,Type=Integer,Description=\"Allele Quality score reflecting evidence for each alternate allele (Phred scale)\">");
    h.push_back("##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes\">");
    h.push_back("##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles in called genotypes\">");
    h.push_back("##FILTER=<ID=MONOALLELIC,Description=\"Site represents one ALT allele in a region with multiple variants that could not be unified into non-overlapping multi-allelic sites\">");
    h.push_back("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">");
    h.push_back("##FORMAT=<ID=RNC,Number=2,Type=Character,Description=\"Reason for No Call in GT: . = n/a, M = Missing data, P = Partial data, I = gVCF input site is non-called, D = insufficient Depth of coverage, - = unrepresentable overlapping deletion, L = Lost/unrepresentable allele (other than deletion), U = multiple Unphased variants present, O = multiple Overlapping variants present, 1 = site is Monoallelic, no assertion about presence of REF or ALT allele\">");
    for (const auto& f : b.gcfg.liftover_fields) h.push_back(f.description);
    return h;
}

static bool is_dict_line(const std::string& l) { return l.compare(0, 9, "##FILTER=") == 0 || l.compare(0, 7, "##INFO=") == 0 || l.compare(0, 9, "##FORMAT=") == 0; }

const glx_bcf_meta& build_bcf_meta(PackedBatch& b) {
    BcfMetaStore& m = b.bcf;
    if (m.built) return m.meta;
    m = BcfMetaStore();
    // dictionary: IDs of FILTER/INFO/FORMAT lines in order of first appearance (htslib bcf_hdr_sync)
    for (const auto& l : header_lines(b)) {
        if (!is_dict_line(l)) continue;
        const std::string id = attr_of(l, "ID");
        bool seen = false;
        for (const auto& d : m.dict) seen = seen || d == id;
        if (!seen) m.dict.push_back(id);
    }
    auto idx = [&](const std::string& id) {
        for (size_t i = 0; i < m.dict.size(); i++)
            if (m.dict[i] == id) return (int32_t)i;
        return (int32_t)-1;
    };
    glx_bcf_meta& g = m.meta;
    memset(&g, 0, sizeof g);
    g.id_AF = idx("AF");
    g.id_AQ = idx("AQ");
    g.id_MONOALLELIC = idx("MONOALLELIC");
    g.id_GT = idx("GT");
    g.id_RNC = idx("RNC");
    for (size_t k = 0; k < b.gcfg.liftover_fields.size() && k < GLX_MAX_FIELDS; k++) {
        const auto& f = b.gcfg.liftover_fields[k];
        g.id_field[k] = idx(attr_of(f.description, "ID"));
        const std::string n = attr_of(f.description, "Number");
        g.field_vl[k] = (n == "A" || n == "R" || n == "G") ? (uint8_t)n[0] : 0;
    }
    m.dna_off.assign(1, 0);
    m.norm_off.assign(1, 0);
    for (const auto& us : b.usites) {
        m.rid.push_back(us.pos.rid);
        m.qual.push_back((float)us.qual);
        for (const auto& al : us.alleles) {
            m.dna_pool.insert(m.dna_pool.end(), al.dna.begin(), al.dna.end());
            m.dna_off.push_back((uint32_t)m.dna_pool.size());
            m.norm_beg.push_back(al.normalized.pos.beg);
            m.norm_end.push_back(al.normalized.pos.end);
            m.norm_pool.insert(m.norm_pool.end(), al.normalized.dna.begin(), al.normalized.dna.end());
            m.norm_off.push_back((uint32_t)m.norm_pool.size());
            m.aq.push_back(al.quality);
        }
    }
    m.contig_off.assign(1, 0);
    for (const auto& c : b.contigs) {
        m.contig_pool.insert(m.contig_pool.end(), c.first.begin(), c.first.end());
        m.contig_off.push_back((uint32_t)m.contig_pool.size());
    }
    m.filter_off.assign(1, 0);
    for (const auto& f : b.filter_names) {
        m.filter_pool.insert(m.filter_pool.end(), f.begin(), f.end());
        m.filter_off.push_back((uint32_t)m.filter_pool.size());
    }
    auto ptr = [](auto& v) { return v.empty() ? nullptr : v.data(); };
    static const char none[1] = {0};
    g.n_sites = (uint32_t)b.usites.size();
    g.n_contigs = (uint32_t)b.contigs.size();
    g.n_filters = (uint32_t)b.filter_names.size();
    g.rid = ptr(m.rid);
    g.qual = ptr(m.qual);
    g.dna_off = m.dna_off.data();
    g.dna_pool = m.dna_pool.empty() ? none : m.dna_pool.data();
    g.norm_beg = ptr(m.norm_beg);
    g.norm_end = ptr(m.norm_end);
    g.norm_off = m.norm_off.data();
    g.norm_pool = m.norm_pool.empty() ? none : m.norm_pool.data();
    g.aq = ptr(m.aq);
    g.contig_off = m.contig_off.data();
    g.contig_pool = m.contig_pool.empty() ? none : m.contig_pool.data();
    g.filter_off = m.filter_off.data();
    g.filter_pool = m.filter_pool.empty() ? none : m.filter_pool.data();
    g.n_dna_pool = m.dna_pool.size();
    g.n_norm_pool = m.norm_pool.size();
    g.n_contig_pool = m.contig_pool.size();
    g.n_filter_pool = m.filter_pool.size();
    m.built = true;
    return g;
}

std::string bcf_header_text(const PackedBatch& b) {
    std::vector<std::string> dict;
    std::string o = "##fileformat=VCFv4.2\n";
    for (const auto& l : header_lines(b)) {
        if (!is_dict_line(l)) {
            o += l + "\n";
            continue;
        }
        const std::string id = attr_of(l, "ID");
        size_t i = 0;
        while (i < dict.size() && dict[i] != id) i++;
        if (i == dict.size()) dict.push_back(id);
        const size_t close = l.rfind('>');
        o += l.substr(0, close) + ",IDX=" + std::to_string(i) + ">\n";
    }
    size_t ci = 0;
    for (const auto& ctg : b.contigs) {
        o += "##contig=<ID=" + ctg.first;
        if (ctg.second != 1000000042 && ctg.second != 0) o += ",length=" + std::to_string(ctg.second);
        o += ",IDX=" + std::to_string(ci++) + ">\n";
    }
    o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (const auto& sm : b.sample_names) o += "\t" + sm;
    o += "\n";
    return o;
}

std::string bcf_file_prologue(const PackedBatch& b) {
    const std::string text = bcf_header_text(b);
    std::string o("BCF\2\2", 5);
    const uint32_t l_text = (uint32_t)text.size() + 1;
    for (int i = 0; i < 4; i++) o.push_back((char)(l_text >> (8 * i) & 0xff));
    o += text;
    o.push_back('\0');
    return o;
}

uint32_t crc32_ieee(const uint8_t* p, size_t n, uint32_t crc) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = c & 1 ? (c >> 1) ^ 0xedb88320u : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

// SAM spec §4.1: gzip member with FLG.FEXTRA, XLEN 6, subfield 'B','C', SLEN 2, BSIZE = total block size - 1
static void bgzf_wrap(std::string& o, const std::string& deflate, const uint8_t* p, size_t n) {
    const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    o.append(reinterpret_cast<const char*>(hdr), 12);
    const uint32_t bsize = (uint32_t)(12 + 6 + deflate.size() + 8 - 1);
    const uint8_t sub[6] = {'B', 'C', 2, 0, (uint8_t)(bsize & 0xff), (uint8_t)(bsize >> 8)};
    o.append(reinterpret_cast<const char*>(sub), 6);
    o += deflate;
    const uint32_t crc = crc32_ieee(p, n), isize = (uint32_t)n;
    for (int i = 0; i < 4; i++) o.push_back((char)(crc >> (8 * i) & 0xff));
    for (int i = 0; i < 4; i++) o.push_back((char)(isize >> (8 * i) & 0xff));
}

std::string bgzf_block_stored(const uint8_t* p, size_t n) {
    std::string d;
    d.push_back(1);  // BFINAL = 1, BTYPE = 00 (stored)
    d.push_back((char)(n & 0xff));
    d.push_back((char)(n >> 8 & 0xff));
    d.push_back((char)(~n & 0xff));
    d.push_back((char)(~n >> 8 & 0xff));
    d.append(reinterpret_cast<const char*>(p), n);
    std::string o;
    bgzf_wrap(o, d, p, n);
    return o;
}

std::string bgzf_eof_block() {
    std::string d("\x03\x00", 2);  // empty fixed-Huffman block, as htslib's EOF marker
    std::string o;
    bgzf_wrap(o, d, nullptr, 0);
    return o;
}

}  // namespace glx
