// Host-side I/O of the cluster step, native so that the GPU path is not starved by Python:
//   * .pairs / .pairs.gz tokenizer with name -> id translation and the alignments.bed side product
//     (pairs_generator / pairs_generator_inter_ctgs, scripts/HapHiC_cluster.py:1539-1583),
//   * paired_links.clm text writer (output_clm, 376-392).
// Pure C++ (no CUDA); part of libhaphic_b200.so, declared in include/haphic_b200.h.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/haphic_b200.h"

void hh_set_error(const char* fmt, ...);

struct hh_pairs_reader {
    gzFile gz = nullptr;              // zlib reads plain files transparently as well
    FILE* bed = nullptr;
    std::unordered_map<std::string, int32_t> ids;
    std::vector<char> buf;            // unconsumed bytes
    size_t pos = 0, len = 0;
    bool eof = false;
    int inter_only = 1;
    int64_t lines = 0;
    std::vector<char> bedbuf;
    size_t bedlen = 0;
};

extern "C" int hh_pairs_open(const char* path, const char* names_blob, int32_t n_names, const char* bed_path, int inter_only,
                             hh_pairs_reader** out) {
    if (!path || !names_blob || !out || n_names < 0) {
        hh_set_error("hh_pairs_open: bad argument");
        return HH_ERR_ARG;
    }
    *out = nullptr;
    hh_pairs_reader* r = new hh_pairs_reader();
    r->gz = gzopen(path, "rb");
    if (!r->gz) {
        hh_set_error("hh_pairs_open: cannot open %s", path);
        delete r;
        return HH_ERR_ARG;
    }
    gzbuffer(r->gz, 1 << 20);
    if (bed_path && *bed_path) {
        r->bed = fopen(bed_path, "w");
        if (!r->bed) {
            hh_set_error("hh_pairs_open: cannot create %s", bed_path);
            gzclose(r->gz);
            delete r;
            return HH_ERR_ARG;
        }
        setvbuf(r->bed, nullptr, _IOFBF, 1 << 22);
    }
    const char* p = names_blob;
    r->ids.reserve((size_t)n_names * 2);
    for (int32_t i = 0; i < n_names; ++i) {
        const size_t l = strlen(p);
        r->ids.emplace(std::string(p, l), i);
        p += l + 1;
    }
    r->inter_only = inter_only;
    r->buf.resize(1 << 24);
    *out = r;
    return HH_OK;
}

static inline char* put_i64(char* p, int64_t v) {
    char tmp[24];
    int n = 0;
    uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

static inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// parses a Python-int()-compatible decimal (optional sign, surrounding blanks already stripped)
static inline bool parse_int(const char* s, const char* e, int64_t* out) {
    if (s == e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') {
        neg = *s == '-';
        ++s;
    }
    if (s == e) return false;
    int64_t v = 0;
    for (; s < e; ++s) {
        if (*s == '_') continue;                 // int('1_000') is valid Python
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (*s - '0');
    }
    *out = neg ? -v : v;
    return true;
}

extern "C" int hh_pairs_next(hh_pairs_reader* r, int32_t* rec, int64_t max_records, int64_t* n_out) {
    if (!r || !rec || !n_out || max_records <= 0) {
        hh_set_error("hh_pairs_next: bad argument");
        return HH_ERR_ARG;
    }
    int64_t n = 0;
    *n_out = 0;
    while (n < max_records) {
        // make sure a whole line is buffered
        char* base = r->buf.data();
        char* nl = (char*)memchr(base + r->pos, '\n', r->len - r->pos);
        if (!nl && !r->eof) {
            if (r->pos > 0) {
                memmove(base, base + r->pos, r->len - r->pos);
                r->len -= r->pos;
                r->pos = 0;
            }
            if (r->len == r->buf.size()) {
                r->buf.resize(r->buf.size() * 2);
                base = r->buf.data();
            }
            const int got = gzread(r->gz, base + r->len, (unsigned)(r->buf.size() - r->len));
            if (got < 0) {
                hh_set_error("hh_pairs_next: read error");
                return HH_ERR_ARG;
            }
            if (got == 0) r->eof = true;
            r->len += (size_t)got;
            continue;
        }
        if (!nl && r->pos >= r->len) break;       // EOF, nothing left
        char* ls = base + r->pos;
        char* le = nl ? nl : base + r->len;        // last line without newline
        r->pos = nl ? (size_t)(nl - base) + 1 : r->len;
        r->lines++;
        // `if not line.strip() or line.startswith('#'): continue`
        char* t = ls;
        while (t < le && is_ws(*t)) ++t;
        if (t == le || *ls == '#') continue;
        // cols = line.split(): first five whitespace-separated tokens
        const char* tok[5];
        const char* tend[5];
        int k = 0;
        char* c = t;
        while (k < 5 && c < le) {
            while (c < le && is_ws(*c)) ++c;
            if (c == le) break;
            tok[k] = c;
            while (c < le && !is_ws(*c)) ++c;
            tend[k] = c;
            ++k;
        }
        if (k < 5) {
            hh_set_error("hh_pairs_next: line %lld has fewer than 5 columns", (long long)r->lines);
            return HH_ERR_ARG;
        }
        int64_t p1, p2;
        if (!parse_int(tok[2], tend[2], &p1) || !parse_int(tok[4], tend[4], &p2)) {
            hh_set_error("hh_pairs_next: line %lld: position is not an integer", (long long)r->lines);
            return HH_ERR_ARG;
        }
        p1 -= 1;                                    // pysam / BED are 0-based
        p2 -= 1;
        if (r->bed) {
            // '{ref}\t{pos}\t{pos}\t{readID}/1\t255\t.\n{mref}\t{mpos}\t{mpos}\t{readID}/2\t255\t.\n'  (1557, 1580)
            const size_t need = 2 * (size_t)(tend[0] - tok[0]) + (size_t)(tend[1] - tok[1]) + (size_t)(tend[3] - tok[3]) + 128;
            if (r->bedbuf.size() < r->bedlen + need) r->bedbuf.resize((r->bedlen + need) * 2);
            char* q = r->bedbuf.data() + r->bedlen;
            for (int side = 0; side < 2; ++side) {
                const char* cs = side ? tok[3] : tok[1];
                const size_t cl = side ? (size_t)(tend[3] - tok[3]) : (size_t)(tend[1] - tok[1]);
                const int64_t pv = side ? p2 : p1;
                memcpy(q, cs, cl);
                q += cl;
                *q++ = '\t';
                q = put_i64(q, pv);
                *q++ = '\t';
                q = put_i64(q, pv);
                *q++ = '\t';
                memcpy(q, tok[0], (size_t)(tend[0] - tok[0]));
                q += tend[0] - tok[0];
                memcpy(q, side ? "/2\t255\t.\n" : "/1\t255\t.\n", 9);
                q += 9;
            }
            r->bedlen = (size_t)(q - r->bedbuf.data());
            if (r->bedlen > (1u << 22)) {
                fwrite(r->bedbuf.data(), 1, r->bedlen, r->bed);
                r->bedlen = 0;
            }
        }
        const size_t l1 = (size_t)(tend[1] - tok[1]), l3 = (size_t)(tend[3] - tok[3]);
        if (r->inter_only && l1 == l3 && memcmp(tok[1], tok[3], l1) == 0) continue;    // ref != mref (1582)
        auto a = r->ids.find(std::string(tok[1], l1));
        auto b = r->ids.find(std::string(tok[3], l3));
        int32_t* o = rec + n * 4;
        o[0] = a == r->ids.end() ? -1 : a->second;
        o[1] = (int32_t)p1;
        o[2] = b == r->ids.end() ? -1 : b->second;
        o[3] = (int32_t)p2;
        ++n;
    }
    *n_out = n;
    return HH_OK;
}

extern "C" int hh_pairs_close(hh_pairs_reader* r) {
    if (!r) return HH_OK;
    if (r->gz) gzclose(r->gz);
    if (r->bed) {
        if (r->bedlen) fwrite(r->bedbuf.data(), 1, r->bedlen, r->bed);
        fclose(r->bed);
    }
    delete r;
    return HH_OK;
}

// ---------------------------------------------------------------------------------------------
// paired_links.clm: for every contig pair with >= 2 links, four lines (orientations ++ +- -+ --),
// each `{ci}{s} {cj}{s}\t{2*links}\t{d d d d ...}` with every ascending distance printed twice.
//   names_blob: NUL-separated contig names; key_i / key_j: contig ids per pair;
//   offsets[n_pairs+1]: start of each pair's block in dist (in links); dist: [4][total_links] int64,
//   each orientation's block of a pair already sorted ascending.
// ---------------------------------------------------------------------------------------------
extern "C" int hh_clm_write(const char* path, const char* names_blob, int32_t n_names, const int32_t* key_i, const int32_t* key_j,
                            int64_t n_pairs, const int64_t* offsets, const int64_t* dist, int64_t total_links) {
    if (!path || !names_blob || !offsets || (n_pairs > 0 && (!key_i || !key_j || !dist))) {
        hh_set_error("hh_clm_write: bad argument");
        return HH_ERR_ARG;
    }
    std::vector<const char*> name(n_names);
    std::vector<size_t> nlen(n_names);
    const char* p = names_blob;
    for (int32_t i = 0; i < n_names; ++i) {
        name[i] = p;
        nlen[i] = strlen(p);
        p += nlen[i] + 1;
    }
    FILE* f = fopen(path, "w");
    if (!f) {
        hh_set_error("hh_clm_write: cannot create %s", path);
        return HH_ERR_ARG;
    }
    setvbuf(f, nullptr, _IOFBF, 1 << 22);
    static const char sg[4][2] = {{'+', '+'}, {'+', '-'}, {'-', '+'}, {'-', '-'}};
    std::vector<char> line;
    for (int64_t e = 0; e < n_pairs; ++e) {
        const int64_t s = offsets[e], links = offsets[e + 1] - s;
        if (links < 2) continue;                                   // `if len(list_) < 8: continue`
        const int32_t a = key_i[e], b = key_j[e];
        if (a < 0 || a >= n_names || b < 0 || b >= n_names || s < 0 || s + links > total_links) {
            fclose(f);
            hh_set_error("hh_clm_write: pair %lld is out of range", (long long)e);
            return HH_ERR_ARG;
        }
        line.resize(nlen[a] + nlen[b] + 64 + (size_t)links * 2 * 21);
        for (int k = 0; k < 4; ++k) {
            char* q = line.data();
            memcpy(q, name[a], nlen[a]);
            q += nlen[a];
            *q++ = sg[k][0];
            *q++ = ' ';
            memcpy(q, name[b], nlen[b]);
            q += nlen[b];
            *q++ = sg[k][1];
            *q++ = '\t';
            q = put_i64(q, links * 2);
            *q++ = '\t';
            const int64_t* d = dist + (size_t)k * (size_t)total_links + s;
            for (int64_t t = 0; t < links; ++t) {
                if (t) *q++ = ' ';
                q = put_i64(q, d[t]);
                *q++ = ' ';
                q = put_i64(q, d[t]);
            }
            *q++ = '\n';
            fwrite(line.data(), 1, (size_t)(q - line.data()), f);
        }
    }
    if (fclose(f) != 0) {
        hh_set_error("hh_clm_write: write to %s failed", path);
        return HH_ERR_ARG;
    }
    return HH_OK;
}

// ---------------------------------------------------------------------------------------------
// BAM reader (bam_generator, scripts/HapHiC_cluster.py:1586-1593, with the htslib filters of 2855 / 2862):
// BGZF blocks are inflated by a small thread pool (blocks are independent deflate streams), the
// records are walked in the decompressed stream and one int32 record per read1 alignment is produced:
// (id(reference_name), reference_start, id(next_reference_name), next_reference_start).
// ---------------------------------------------------------------------------------------------
#include <atomic>
#include <thread>

struct hh_bam_reader {
    FILE* f = nullptr;
    std::vector<uint8_t> comp;        // compressed bytes not yet consumed
    size_t comp_pos = 0, comp_len = 0;
    bool file_eof = false;
    std::vector<uint8_t> raw;         // decompressed bytes not yet consumed
    size_t raw_pos = 0, raw_len = 0;
    std::string header_text;
    std::vector<int32_t> ref_to_id;   // BAM refID -> contig id (-1 = not in the FASTA)
    int inter_only = 1;
    int threads = 1;
    int64_t n_records = 0;
};

static inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct hh_bgzf_block {
    size_t in_off, in_len;            // deflate payload inside comp
    size_t out_off;                   // destination offset inside raw
    uint32_t isize, crc;
};

// decompress the next batch of whole BGZF blocks (up to ~target_bytes of compressed input); appends to raw.
// Returns 0 = appended something, 1 = end of file, <0 = error (message set).
static int bam_fill(hh_bam_reader* r, size_t target_bytes) {
    // compact the compressed buffer and read more of the file
    if (r->comp_pos > 0) {
        memmove(r->comp.data(), r->comp.data() + r->comp_pos, r->comp_len - r->comp_pos);
        r->comp_len -= r->comp_pos;
        r->comp_pos = 0;
    }
    if (!r->file_eof && r->comp_len < target_bytes) {
        if (r->comp.size() < target_bytes + (1u << 16)) r->comp.resize(target_bytes + (1u << 16));
        const size_t got = fread(r->comp.data() + r->comp_len, 1, r->comp.size() - r->comp_len, r->f);
        if (got == 0) r->file_eof = true;
        r->comp_len += got;
    }
    if (r->comp_len == 0) return 1;
    // block boundaries
    std::vector<hh_bgzf_block> blocks;
    size_t p = 0, out_total = 0;
    while (p + 18 <= r->comp_len) {
        const uint8_t* h = r->comp.data() + p;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) {
            hh_set_error("hh_bam: not a BGZF block at compressed offset (file is not a BAM / is corrupt)");
            return -1;
        }
        const size_t xlen = le16(h + 10);
        if (p + 12 + xlen > r->comp_len) break;
        size_t bsize = 0;
        for (size_t q = 0; q + 4 <= xlen;) {
            const uint8_t* sf = h + 12 + q;
            const size_t slen = le16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) bsize = (size_t)le16(sf + 4) + 1;
            q += 4 + slen;
        }
        if (bsize == 0 || bsize < 12 + xlen + 8) {
            hh_set_error("hh_bam: BGZF block without a valid BC subfield");
            return -1;
        }
        if (p + bsize > r->comp_len) break;          // incomplete block: wait for more input
        hh_bgzf_block b;
        b.in_off = p + 12 + xlen;
        b.in_len = bsize - 12 - xlen - 8;
        b.crc = le32(h + bsize - 8);
        b.isize = le32(h + bsize - 4);
        b.out_off = out_total;
        out_total += b.isize;
        blocks.push_back(b);
        p += bsize;
    }
    if (blocks.empty()) {
        if (r->file_eof) {
            hh_set_error("hh_bam: truncated BGZF block at the end of the file");
            return -1;
        }
        // a single block larger than what is buffered cannot happen (blocks are <= 64 KiB): read more
        return bam_fill(r, target_bytes * 2);
    }
    // make room in raw (keep the unconsumed tail)
    if (r->raw_pos > 0) {
        memmove(r->raw.data(), r->raw.data() + r->raw_pos, r->raw_len - r->raw_pos);
        r->raw_len -= r->raw_pos;
        r->raw_pos = 0;
    }
    if (r->raw.size() < r->raw_len + out_total) r->raw.resize(r->raw_len + out_total);
    uint8_t* out_base = r->raw.data() + r->raw_len;
    const uint8_t* in_base = r->comp.data();
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto work = [&]() {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            failed = 1;
            return;
        }
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= blocks.size()) break;
            const hh_bgzf_block& b = blocks[k];
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef*>(in_base + b.in_off);
            zs.avail_in = (uInt)b.in_len;
            zs.next_out = out_base + b.out_off;
            zs.avail_out = b.isize;
            const int rc = b.isize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
            if ((b.isize && rc != Z_STREAM_END) || zs.avail_out != 0 ||
                (uint32_t)crc32(crc32(0L, Z_NULL, 0), out_base + b.out_off, b.isize) != b.crc) {
                failed = 1;
                break;
            }
        }
        inflateEnd(&zs);
    };
    const int nt = (int)std::min<size_t>((size_t)std::max(1, r->threads), blocks.size());
    if (nt <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work);
        for (auto& t : pool) t.join();
    }
    if (failed) {
        hh_set_error("hh_bam: BGZF block failed to inflate (corrupt file)");
        return -1;
    }
    r->raw_len += out_total;
    r->comp_pos = p;
    return 0;
}

// make sure `need` decompressed bytes are available at raw_pos; returns 0 ok, 1 clean EOF (nothing left), -1 error
static int bam_need(hh_bam_reader* r, size_t need) {
    while (r->raw_len - r->raw_pos < need) {
        const int rc = bam_fill(r, 32u << 20);
        if (rc < 0) return -1;
        if (rc == 1) {
            if (r->raw_len == r->raw_pos) return 1;
            hh_set_error("hh_bam: truncated BAM record");
            return -1;
        }
    }
    return 0;
}

extern "C" int hh_bam_open(const char* path, const char* names_blob, int32_t n_names, int inter_only, int threads,
                           hh_bam_reader** out) {
    if (!path || !names_blob || !out || n_names < 0) {
        hh_set_error("hh_bam_open: bad argument");
        return HH_ERR_ARG;
    }
    *out = nullptr;
    hh_bam_reader* r = new hh_bam_reader();
    r->f = fopen(path, "rb");
    if (!r->f) {
        hh_set_error("hh_bam_open: cannot open %s", path);
        delete r;
        return HH_ERR_ARG;
    }
    r->inter_only = inter_only;
    r->threads = threads > 0 ? threads : 1;
    auto fail = [&](const char* msg) {
        if (msg) hh_set_error("%s", msg);
        fclose(r->f);
        delete r;
        return HH_ERR_ARG;
    };
    if (bam_need(r, 12) != 0) return fail(nullptr);
    const uint8_t* p = r->raw.data() + r->raw_pos;
    if (memcmp(p, "BAM\1", 4) != 0) return fail("hh_bam_open: not a BAM file");
    const size_t l_text = le32(p + 4);
    if (bam_need(r, 12 + l_text) != 0) return fail(nullptr);
    p = r->raw.data() + r->raw_pos;
    r->header_text.assign(reinterpret_cast<const char*>(p + 8), l_text);
    while (!r->header_text.empty() && r->header_text.back() == '\0') r->header_text.pop_back();
    const int32_t n_ref = (int32_t)le32(p + 8 + l_text);
    r->raw_pos += 12 + l_text;
    std::unordered_map<std::string, int32_t> ids;
    ids.reserve((size_t)n_names * 2);
    const char* q = names_blob;
    for (int32_t i = 0; i < n_names; ++i) {
        const size_t l = strlen(q);
        ids.emplace(std::string(q, l), i);
        q += l + 1;
    }
    r->ref_to_id.assign((size_t)(n_ref > 0 ? n_ref : 0), -1);
    for (int32_t k = 0; k < n_ref; ++k) {
        if (bam_need(r, 4) != 0) return fail("hh_bam_open: truncated BAM header");
        const size_t l_name = le32(r->raw.data() + r->raw_pos);
        if (bam_need(r, 8 + l_name) != 0) return fail("hh_bam_open: truncated BAM header");
        const char* nm = reinterpret_cast<const char*>(r->raw.data() + r->raw_pos + 4);
        auto it = ids.find(std::string(nm, l_name ? l_name - 1 : 0));
        if (it != ids.end()) r->ref_to_id[(size_t)k] = it->second;
        r->raw_pos += 8 + l_name;
    }
    *out = r;
    return HH_OK;
}

extern "C" int hh_bam_header_text(hh_bam_reader* r, const char** text, int64_t* len) {
    if (!r || !text || !len) {
        hh_set_error("hh_bam_header_text: bad argument");
        return HH_ERR_ARG;
    }
    *text = r->header_text.data();
    *len = (int64_t)r->header_text.size();
    return HH_OK;
}

extern "C" int hh_bam_next(hh_bam_reader* r, int32_t* rec, int64_t max_records, int64_t* n_out) {
    if (!r || !rec || !n_out || max_records <= 0) {
        hh_set_error("hh_bam_next: bad argument");
        return HH_ERR_ARG;
    }
    int64_t n = 0;
    *n_out = 0;
    const int32_t n_ref = (int32_t)r->ref_to_id.size();
    while (n < max_records) {
        int rc = bam_need(r, 4);
        if (rc == 1) break;
        if (rc < 0) return HH_ERR_ARG;
        const size_t bs = le32(r->raw.data() + r->raw_pos);
        if (bs < 32) {
            hh_set_error("hh_bam_next: corrupt BAM record (block_size %zu)", bs);
            return HH_ERR_ARG;
        }
        rc = bam_need(r, 4 + bs);
        if (rc != 0) {
            if (rc == 1) hh_set_error("hh_bam_next: truncated BAM record");
            return HH_ERR_ARG;
        }
        const uint8_t* p = r->raw.data() + r->raw_pos + 4;
        r->raw_pos += 4 + bs;
        r->n_records++;
        const int32_t refid = (int32_t)le32(p), pos = (int32_t)le32(p + 4);
        const uint16_t flag = le16(p + 14);
        const int32_t mrefid = (int32_t)le32(p + 20), mpos = (int32_t)le32(p + 24);
        if (!(flag & 0x40)) continue;                              // flag.read1
        if (r->inter_only && refid == mrefid) continue;            // refid != mrefid (2862)
        int32_t* o = rec + n * 4;
        o[0] = (refid >= 0 && refid < n_ref) ? r->ref_to_id[(size_t)refid] : -1;
        o[1] = pos;
        o[2] = (mrefid >= 0 && mrefid < n_ref) ? r->ref_to_id[(size_t)mrefid] : -1;
        o[3] = mpos;
        ++n;
    }
    *n_out = n;
    return HH_OK;
}

extern "C" int hh_bam_close(hh_bam_reader* r) {
    if (!r) return HH_OK;
    if (r->f) fclose(r->f);
    delete r;
    return HH_OK;
}
