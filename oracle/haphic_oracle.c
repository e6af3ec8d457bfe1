/* CPU oracle for the link-counting loop of `haphic cluster` -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the per-read-pair loop of parse_alignments_for_ctgs
 * (scripts/HapHiC_cluster.py:1596-1655 with is_flank 299-307 and update_HT_link_dict 404-416) over integer records.
 * Only tests/, __graft_entry__ and bench.py's cpu_baseline leg may load it (through oracle/haphic_oracle.py,
 * which compiles it with gcc into oracle/_build/).  It is pinned by the same golden fixtures as the numpy
 * restatement (tests/test_oracle_golden.py) and serves as the single-core "what an optimised CPU port does"
 * number next to the Python-loop port that mirrors the reference's own speed.
 *
 * Records: int32 {ctg_a, pos_a, ctg_b, pos_b}, 0-based positions; ids outside [0, n) = names missing from the
 * FASTA (skipped, 1625); ctg_a == ctg_b is dropped (1582).  name_rank[c] = rank of contig c under Python's str
 * order (the reference sorts each pair by name, 1629).
 * Output, one entry per contig pair in dict-insertion (first-seen) order: key_i, key_j, full, flank,
 * first_full / first_flank (index of the first [flank-qualifying] record, -1 if none), ht[4] = {HH, HT, TH, TT}.
 * ctg_links[n] = per-contig flank-link totals (ctg_link_dict, 1638-1639).
 * Returns the number of entries, or -1 when `cap` entries are not enough. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline int is_flank(int64_t coord, int64_t length, int64_t flank) {   /* 299-307, coord is 1-based */
    return !flank || coord <= flank || coord > length - flank;
}

typedef struct {                 /* one cache line per contig pair */
    int32_t i, j;
    int64_t full, flank, first_full, first_flank;
    int64_t ht[4];
} hho_entry;

typedef struct {
    int32_t i, j;
    int64_t ci, cj;
    uint64_t slot;
    int ok;
} hho_item;

#define HHO_AHEAD 16             /* records decoded (and their table slots prefetched) ahead of the update */

int64_t hho_count_links(const int32_t* rec, int64_t n_rec, int32_t n, const int64_t* lengths, const int32_t* name_rank,
                        const uint8_t* in_nx, int64_t flank_bp, int64_t cap, int32_t* key_i, int32_t* key_j, int64_t* full,
                        int64_t* flank, int64_t* first_full, int64_t* first_flank, int64_t* ht, int64_t* ctg_links,
                        int64_t* n_used) {
    uint64_t slots = 16;
    while (slots < (uint64_t)cap * 2) slots <<= 1;
    int64_t* table = (int64_t*)calloc(slots, sizeof(int64_t));       /* entry index + 1, 0 = empty */
    hho_entry* ent = (hho_entry*)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(hho_entry));
    if (!table || !ent) {
        free(table);
        free(ent);
        return -1;
    }
    memset(ctg_links, 0, (size_t)n * sizeof(int64_t));
    const uint64_t mask = slots - 1;
    int64_t nnz = 0, used = 0;
    hho_item ring[HHO_AHEAD];
    /* software pipeline: record r + HHO_AHEAD is decoded and its table slot prefetched while record r is applied;
     * the order of the updates is the stream order, exactly as in the plain loop */
    for (int64_t r = -HHO_AHEAD; r < n_rec; ++r) {
        const int64_t q = r + HHO_AHEAD;
        hho_item nxt;
        nxt.ok = 0;
        if (q < n_rec) {
            const int32_t a = rec[4 * q], pa = rec[4 * q + 1], b = rec[4 * q + 2], pb = rec[4 * q + 3];
            if (a != b && a >= 0 && a < n && b >= 0 && b < n) {                  /* 1582, 1625 */
                nxt.ok = 1;
                if (name_rank[a] < name_rank[b]) {                               /* 1629 */
                    nxt.i = a; nxt.ci = (int64_t)pa + 1; nxt.j = b; nxt.cj = (int64_t)pb + 1;
                } else {
                    nxt.i = b; nxt.ci = (int64_t)pb + 1; nxt.j = a; nxt.cj = (int64_t)pa + 1;
                }
                nxt.slot = mix64(((uint64_t)(uint32_t)nxt.i << 32) | (uint32_t)nxt.j) & mask;
                __builtin_prefetch(table + nxt.slot, 1, 0);
            }
        }
        hho_item cur = ring[(r + HHO_AHEAD) % HHO_AHEAD];
        ring[(r + HHO_AHEAD) % HHO_AHEAD] = nxt;
        if (r < 0 || !cur.ok) continue;
        ++used;
        const int32_t i = cur.i, j = cur.j;
        const int64_t ci = cur.ci, cj = cur.cj;
        const int64_t li = lengths[i], lj = lengths[j];
        uint64_t s = cur.slot;
        int64_t e;
        for (;;) {
            e = table[s] - 1;
            if (e < 0) {
                if (nnz >= cap) {
                    free(table);
                    free(ent);
                    return -1;
                }
                e = nnz++;
                table[s] = e + 1;
                hho_entry* x = ent + e;
                x->i = i;
                x->j = j;
                x->full = 0;
                x->flank = 0;
                x->first_full = r;
                x->first_flank = -1;
                x->ht[0] = x->ht[1] = x->ht[2] = x->ht[3] = 0;
                break;
            }
            if (ent[e].i == i && ent[e].j == j) break;
            s = (s + 1) & mask;
        }
        hho_entry* x = ent + e;
        if (in_nx[i] && in_nx[j] && is_flank(ci, li, flank_bp) && is_flank(cj, lj, flank_bp)) {   /* 1636-1639 */
            if (x->flank++ == 0) x->first_flank = r;
            ctg_links[i]++;
            ctg_links[j]++;
        }
        const int ti = ci * 2 > li, tj = cj * 2 > lj;                            /* 404-416 */
        x->ht[2 * ti + tj]++;
        x->full++;                                                               /* 1649 */
    }
    for (int64_t e = 0; e < nnz; ++e) {
        key_i[e] = ent[e].i;
        key_j[e] = ent[e].j;
        full[e] = ent[e].full;
        flank[e] = ent[e].flank;
        first_full[e] = ent[e].first_full;
        first_flank[e] = ent[e].first_flank;
        memcpy(ht + 4 * e, ent[e].ht, 4 * sizeof(int64_t));
    }
    free(table);
    free(ent);
    *n_used = used;
    return nnz;
}
