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
