// icl_host.cpp — native host helpers of the `-icl` path: m8 ingest, grading, collinearity-file writer.
// See include/icl_host.h for the reference lines each entry point replaces.  No CUDA here.
#include "../../include/icl_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

void set_err(char *err, int32_t errlen, const std::string &msg) {
    if (err && errlen > 0) {
        snprintf(err, (size_t)errlen, "%s", msg.c_str());
    }
}

template <class F>
void parallel_for(int64_t n, int threads, F f) {        // f(index) for index in [0, n), dynamic scheduling
    if (threads < 1) threads = 1;
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n) break;
            f(i);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

inline uint64_t hash_bytes(const char *p, size_t n) {   // FNV-1a, finalised
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

// open-addressing map  gene id (bytes) -> row in the gene table
struct IdMap {
    const char *ids = nullptr;
    const int64_t *off = nullptr;
    std::vector<int32_t> slot;
    uint64_t mask = 0;
    void build(const char *ids_, const int64_t *off_, int64_t n) {
        ids = ids_; off = off_;
        uint64_t cap = 16;
        while (cap < (uint64_t)n * 2 + 8) cap <<= 1;
        slot.assign(cap, -1);
        mask = cap - 1;
        for (int64_t i = 0; i < n; i++) {
            const char *p = ids + off[i];
            size_t len = (size_t)(off[i + 1] - off[i]);
            uint64_t h = hash_bytes(p, len) & mask;
            for (;;) {
                int32_t s = slot[h];
                if (s < 0) { slot[h] = (int32_t)i; break; }
                // duplicated gene ids: the first row wins (pandas would raise on a non-unique index)
                if ((size_t)(off[s + 1] - off[s]) == len && !memcmp(ids + off[s], p, len)) break;
                h = (h + 1) & mask;
            }
        }
    }
    int32_t find(const char *p, size_t len) const {
        uint64_t h = hash_bytes(p, len) & mask;
        for (;;) {
            int32_t s = slot[h];
            if (s < 0) return -1;
            if ((size_t)(off[s + 1] - off[s]) == len && !memcmp(ids + off[s], p, len)) return s;
            h = (h + 1) & mask;
        }
    }
};

inline bool is_int_literal(const char *p, const char *q) {
    if (p < q && (*p == '+' || *p == '-')) p++;
    if (p >= q) return false;
    for (; p < q; p++)
        if (*p < '0' || *p > '9') return false;
    return true;
}

inline double parse_double_slow(const char *p, const char *q, bool *ok) {
    char buf[64];
    size_t n = (size_t)(q - p);
    if (n == 0 || n >= sizeof(buf)) { *ok = false; return NAN; }
    memcpy(buf, p, n);
    buf[n] = 0;
    char *end = nullptr;
    double v = strtod(buf, &end);          // correctly rounded; equals pandas' parser for BLAST's short decimals
    *ok = end == buf + n;
    return v;
}

// [sign] digits [. digits] [e|E [sign] digits], nothing else, at most 19 significant digits:
// value = (neg ? -1 : 1) * w * 10^e10 with w exact.
inline bool parse_simple(const char *p, const char *q, uint64_t &w, int &e10, bool &neg) {
    neg = false;
    if (p < q && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
    w = 0; e10 = 0;
    int digits = 0, sig = 0;
    for (; p < q && *p >= '0' && *p <= '9'; p++) {
        digits++;
        if (sig > 0 || *p != '0') { if (++sig > 19) return false; }
        w = w * 10 + (uint64_t)(*p - '0');
    }
    if (p < q && *p == '.') {
        p++;
        for (; p < q && *p >= '0' && *p <= '9'; p++) {
            digits++;
            if (sig > 0 || *p != '0') { if (++sig > 19) return false; }
            w = w * 10 + (uint64_t)(*p - '0');
            e10--;
        }
    }
    if (digits == 0) return false;
    if (p < q && (*p == 'e' || *p == 'E')) {
        p++;
        bool eneg = false;
        if (p < q && (*p == '+' || *p == '-')) { eneg = *p == '-'; p++; }
        int e = 0, ed = 0;
        for (; p < q && *p >= '0' && *p <= '9'; p++) { if (++ed > 4) return false; e = e * 10 + (*p - '0'); }
        if (ed == 0) return false;
        e10 += eneg ? -e : e;
    }
    return p == q;
}

static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                  1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Exact double of a decimal token: the classic fast path (mantissa < 2^53, |exponent| <= 22: one correctly rounded
// multiplication or division of two exact doubles), strtod otherwise.
inline double parse_double(const char *p, const char *q, bool *ok) {
    uint64_t w; int e10; bool neg;
    if (parse_simple(p, q, w, e10, neg) && w < (1ull << 53) && e10 >= -22 && e10 <= 22) {
        double v = (double)w;
        v = e10 < 0 ? v / kPow10[-e10] : v * kPow10[e10];
        *ok = true;
        return neg ? -v : v;
    }
    return parse_double_slow(p, q, ok);
}

// Only the outcome of `value < limit` is needed for the e-value column (base.py:178): decide it from an extended
// precision estimate when the value is not within a hair of the limit, otherwise from the exact conversion.
struct LessThan {
    long double limit, lo, hi;
    long double p10[801];
    explicit LessThan(double lim) : limit(lim) {
        long double band = fabsl(limit) * 1e-15L;
        lo = limit - band; hi = limit + band;
        for (int k = -400; k <= 400; k++) p10[k + 400] = powl(10.0L, (long double)k);
    }
    // returns 1 / 0, or -1 when the token does not parse (NaN in pandas: comparison false)
    int test(const char *p, const char *q) const {
        uint64_t w; int e10; bool neg;
        if (parse_simple(p, q, w, e10, neg) && e10 >= -400 && e10 <= 400) {
            long double v = (long double)w * p10[e10 + 400];
            if (neg) v = -v;
            if (v < lo) return 1;
            if (v > hi) return 0;
        }
        bool ok = true;
        double v = parse_double_slow(p, q, &ok);
        if (!ok) return -1;
        return v < (double)limit ? 1 : 0;
    }
};

struct ChunkOut {
    std::vector<int64_t> row;
    std::vector<int32_t> g1, g2;
    std::vector<double> bits;
    int64_t n_rows = 0;
    bool all_int = true;
    std::string error;
};

}  // namespace

extern "C" void icl_host_free_hits(icl_hits *h) {
    if (!h) return;
    free(h->row); free(h->g1); free(h->g2); free(h->bits);
    memset(h, 0, sizeof(*h));
}

extern "C" int icl_host_read_m8(const char *path, int32_t reverse, double min_score, double max_evalue,
                                const char *ids1, const int64_t *off1, int64_t n1,
                                const char *ids2, const int64_t *off2, int64_t n2,
                                int32_t threads, icl_hits *out, char *err, int32_t errlen) {
    if (!path || !out || !off1 || !off2 || (n1 > 0 && !ids1) || (n2 > 0 && !ids2)) { set_err(err, errlen, "icl_host_read_m8: null argument"); return -1; }
    memset(out, 0, sizeof(*out));
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_err(err, errlen, std::string("cannot open ") + path + ": " + strerror(errno)); return -2; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); set_err(err, errlen, "fstat failed"); return -2; }
    const size_t size = (size_t)sb.st_size;
    const char *data = nullptr;
    if (size > 0) {
        data = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (data == MAP_FAILED) { close(fd); set_err(err, errlen, "mmap failed"); return -2; }
        madvise((void *)data, size, MADV_SEQUENTIAL);
    }
    IdMap m1, m2;
    m1.build(ids1, off1, n1);
    const bool same = ids1 == ids2 && off1 == off2 && n1 == n2;
    if (!same) m2.build(ids2, off2, n2);
    const IdMap &map1 = m1, &map2 = same ? m1 : m2;

    if (threads < 1) threads = 1;
    int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>((int64_t)threads * 4, (int64_t)(size >> 16) + 1));
    std::vector<size_t> start(n_chunks + 1, size);
    start[0] = 0;
    for (int64_t c = 1; c < n_chunks; c++) {            // chunk boundaries at line starts
        size_t p = size / n_chunks * c;
        while (p < size && data[p - 1] != '\n') p++;
        start[c] = p;
    }
    for (int64_t c = 1; c <= n_chunks; c++) start[c] = std::max(start[c], start[c - 1]);
    std::vector<ChunkOut> outs(n_chunks);
    const LessThan evalue_below(max_evalue);
    parallel_for(n_chunks, threads, [&](int64_t c) {
        ChunkOut &o = outs[c];
        const char *p = data + start[c], *end = data + start[c + 1];
        while (p < end) {
            const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
            const char *le = eol ? eol : end;
            const char *next = eol ? eol + 1 : end;
            if (le > p && le[-1] == '\r') le--;
            if (le == p) { p = next; continue; }        // blank line (pandas: skip_blank_lines)
            const char *f[13];
            int nf = 0;
            f[0] = p;
            for (const char *q = p; q < le && nf < 12; q++)
                if (*q == '\t') f[++nf] = q + 1;
            // f[k] = start of field k; field k ends at f[k+1]-1 (or le for the last one found)
            if (nf < 11) { o.error = "row with fewer than 12 tab-separated columns"; return; }
            const char *e11 = le;
            if (nf == 12) e11 = f[12] - 1;
            const int64_t row = o.n_rows++;
            const char *a0 = f[0], *a1 = f[1] - 1, *b0 = f[1], *b1 = f[2] - 1;
            if (reverse) { std::swap(a0, b0); std::swap(a1, b1); }
            bool ok2 = true;
            if (o.all_int && !is_int_literal(f[11], e11)) o.all_int = false;
            const int ev_ok = evalue_below.test(f[10], f[11] - 1);
            double bs = parse_double(f[11], e11, &ok2);
            p = next;
            if (ev_ok < 0 || !ok2) continue;             // unparsable -> NaN in pandas -> fails both comparisons
            if (!(bs >= min_score) || ev_ok == 0) continue;                               // base.py:178
            size_t la = (size_t)(a1 - a0), lb = (size_t)(b1 - b0);
            if (la == lb && !memcmp(a0, b0, la)) continue;                                // id1 != id2
            int32_t g1 = map1.find(a0, la);
            if (g1 < 0) continue;                                                         // base.py:179
            int32_t g2 = map2.find(b0, lb);
            if (g2 < 0) continue;
            o.row.push_back(row); o.g1.push_back(g1); o.g2.push_back(g2); o.bits.push_back(bs);
        }
    });
    if (data) munmap((void *)data, size);
    close(fd);
    int64_t total = 0, rows = 0;
    bool all_int = true;
    for (auto &o : outs) {
        if (!o.error.empty()) { set_err(err, errlen, std::string("icl_host_read_m8: ") + o.error); return -3; }
        total += (int64_t)o.row.size();
        all_int = all_int && o.all_int;
    }
    // keep='first' on (id1, id2) in file order (base.py:180).  The keys are split by hash over `parts` private tables;
    // every thread walks all rows in file order (cheap: two ints per row) and inserts only the keys of its part, so
    // the first occurrence in file order wins inside every part and the random probes stay inside one small table.
    std::vector<int64_t> cbase((size_t)n_chunks + 1, 0), rbase((size_t)n_chunks + 1, 0);
    for (int64_t c = 0; c < n_chunks; c++) {
        cbase[c + 1] = cbase[c] + (int64_t)outs[c].row.size();
        rbase[c + 1] = rbase[c] + outs[c].n_rows;
    }
    std::vector<uint8_t> dup((size_t)std::max<int64_t>(total, 1), 0);
    int parts = 1;
    while (parts < threads && parts < 64) parts <<= 1;
    auto key_of = [](int32_t a, int32_t b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
    auto mix = [](uint64_t key) { return key * 0x9E3779B97F4A7C15ull; };
    parallel_for(parts, threads, [&](int64_t part) {
        uint64_t cap = 16;
        while (cap < (uint64_t)total * 2 / (uint64_t)parts + 64) cap <<= 1;
        std::vector<uint64_t> seen(cap, ~0ull);
        const uint64_t mask = cap - 1;
        for (int64_t c = 0; c < n_chunks; c++) {
            const ChunkOut &o = outs[c];
            for (size_t i = 0; i < o.row.size(); i++) {
                const uint64_t key = key_of(o.g1[i], o.g2[i]), hm = mix(key);
                if ((int64_t)(hm >> 58) % parts != part) continue;
                uint64_t h = (hm >> 20) & mask;
                for (;;) {
                    if (seen[h] == ~0ull) { seen[h] = key; break; }
                    if (seen[h] == key) { dup[(size_t)cbase[c] + i] = 1; break; }
                    h = (h + 1) & mask;
                }
            }
        }
    });
    std::vector<int64_t> obase((size_t)n_chunks + 1, 0);
    for (int64_t c = 0; c < n_chunks; c++) {
        int64_t kept = 0;
        for (size_t i = 0; i < outs[c].row.size(); i++) kept += !dup[(size_t)cbase[c] + i];
        obase[c + 1] = obase[c] + kept;
    }
    const int64_t kept_total = obase[n_chunks];
    out->row = (int64_t *)malloc((size_t)std::max<int64_t>(kept_total, 1) * sizeof(int64_t));
    out->g1 = (int32_t *)malloc((size_t)std::max<int64_t>(kept_total, 1) * sizeof(int32_t));
    out->g2 = (int32_t *)malloc((size_t)std::max<int64_t>(kept_total, 1) * sizeof(int32_t));
    out->bits = (double *)malloc((size_t)std::max<int64_t>(kept_total, 1) * sizeof(double));
    if (!out->row || !out->g1 || !out->g2 || !out->bits) { icl_host_free_hits(out); set_err(err, errlen, "out of memory"); return -5; }
    parallel_for(n_chunks, threads, [&](int64_t c) {
        const ChunkOut &o = outs[c];
        int64_t k = obase[c];
        for (size_t i = 0; i < o.row.size(); i++) {
            if (dup[(size_t)cbase[c] + i]) continue;
            out->row[k] = rbase[c] + o.row[i]; out->g1[k] = o.g1[i]; out->g2[k] = o.g2[i]; out->bits[k] = o.bits[i];
            k++;
        }
    });
    const int64_t k = kept_total;
    rows = rbase[n_chunks];
    out->n = k;
    out->n_rows_file = rows;
    out->bits_is_int = all_int ? 1 : 0;
    return 0;
}

extern "C" int icl_host_grade(int64_t n, const int32_t *query, const double *bits, int32_t n_query,
                              int32_t multiple, int32_t repeat_number, const int32_t *grading3,
                              int32_t *grade_out, uint8_t *ambiguous, int32_t threads, char *err, int32_t errlen) {
    if (n < 0 || n_query < 0 || (n > 0 && (!query || !bits || !grade_out)) || !grading3 || (n_query > 0 && !ambiguous)) {
        set_err(err, errlen, "icl_host_grade: bad argument");
        return -1;
    }
    const int64_t rn = repeat_number;
    const int64_t red = std::min<int64_t>(multiple, rn), blue = std::min<int64_t>(4 + (int64_t)multiple, rn);   // :64,:67
    std::vector<int64_t> start((size_t)n_query + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        if (query[i] < 0 || query[i] >= n_query) { set_err(err, errlen, "icl_host_grade: query code out of range"); return -1; }
        start[(size_t)query[i] + 1]++;
    }
    for (int32_t q = 0; q < n_query; q++) start[(size_t)q + 1] += start[q];
    std::vector<int64_t> rows((size_t)n), fill(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < n; i++) rows[(size_t)fill[query[i]]++] = i;            // stable: file order inside a query
    if (n_query > 0) memset(ambiguous, 0, (size_t)n_query);
    const int64_t block = 256;
    parallel_for((n_query + block - 1) / block, threads, [&](int64_t blk) {
        int64_t q0 = blk * block, q1 = std::min<int64_t>(q0 + block, n_query);
        for (int64_t q = q0; q < q1; q++) {
            int64_t *b = rows.data() + start[q], *e = rows.data() + start[q + 1];
            const int64_t m = e - b;
            if (m == 0) continue;
            std::stable_sort(b, e, [&](int64_t x, int64_t y) { return bits[x] > bits[y]; });
            for (int64_t r = 0; r < m; r++)
                grade_out[b[r]] = r < red ? grading3[0] : r < blue ? grading3[1] : r < rn ? grading3[2] : 0;
            const int64_t bounds[3] = {red, blue, rn};
            for (int64_t bd : bounds)
                if (bd > 0 && bd < m && bits[b[bd - 1]] == bits[b[bd]]) ambiguous[q] = 1;
        }
    });
    return 0;
}

namespace {
inline char *put_int(char *p, int64_t v) {
    char tmp[24];
    int n = 0;
    bool neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-v) : (uint64_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}
inline int int_len(int64_t v) {
    int n = v < 0 ? 2 : 1;
    uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
    while (u >= 10) { u /= 10; n++; }
    return n;
}
}  // namespace

extern "C" int icl_host_write_blocks(const char *path, int64_t n_sel,
                                     const char *headers, const int64_t *header_off,
                                     const int64_t *mem_begin, const int64_t *mem_end,
                                     const int32_t *loc1, const int32_t *loc2,
                                     const int32_t *chr1_code, const int32_t *chr2_code,
                                     const int64_t *row_of1, int64_t maxo1, const char *names1, const int64_t *name_off1,
                                     const int32_t *strand_code1,
                                     const int64_t *row_of2, int64_t maxo2, const char *names2, const int64_t *name_off2,
                                     const int32_t *strand_code2,
                                     int32_t threads, char *err, int32_t errlen) {
    if (!path || n_sel < 0) { set_err(err, errlen, "icl_host_write_blocks: bad argument"); return -1; }
    std::vector<int64_t> size((size_t)n_sel + 1, 0);
    std::atomic<int> bad(0);
    const int64_t grain = 64;
    parallel_for((n_sel + grain - 1) / grain, threads, [&](int64_t blk) {
        for (int64_t b = blk * grain; b < std::min<int64_t>(n_sel, (blk + 1) * grain); b++) {
            int64_t s = header_off[b + 1] - header_off[b] + 1;
            for (int64_t k = mem_begin[b]; k < mem_end[b]; k++) {
                int64_t r1 = row_of1[(int64_t)chr1_code[b] * maxo1 + loc1[k]], r2 = row_of2[(int64_t)chr2_code[b] * maxo2 + loc2[k]];
                if (r1 < 0 || r2 < 0) { bad = 1; continue; }
                s += (name_off1[r1 + 1] - name_off1[r1]) + 1 + int_len(loc1[k]) + 1 + (name_off2[r2 + 1] - name_off2[r2]) + 1 +
                     int_len(loc2[k]) + 1 + (strand_code1[r1] == strand_code2[r2] ? 1 : 2) + 1;
            }
            size[(size_t)b + 1] = s;
        }
    });
    if (bad) { set_err(err, errlen, "icl_host_write_blocks: a block member has no gene at (chromosome, order)"); return -4; }
    for (int64_t b = 0; b < n_sel; b++) size[(size_t)b + 1] += size[(size_t)b];
    const int64_t total = size[(size_t)n_sel];
    std::vector<char> buf((size_t)std::max<int64_t>(total, 1));
    parallel_for((n_sel + grain - 1) / grain, threads, [&](int64_t blk) {
        for (int64_t b = blk * grain; b < std::min<int64_t>(n_sel, (blk + 1) * grain); b++) {
            char *p = buf.data() + size[(size_t)b];
            int64_t hl = header_off[b + 1] - header_off[b];
            memcpy(p, headers + header_off[b], (size_t)hl);
            p += hl;
            *p++ = '\n';
            for (int64_t k = mem_begin[b]; k < mem_end[b]; k++) {
                int64_t r1 = row_of1[(int64_t)chr1_code[b] * maxo1 + loc1[k]], r2 = row_of2[(int64_t)chr2_code[b] * maxo2 + loc2[k]];
                int64_t l1 = name_off1[r1 + 1] - name_off1[r1], l2 = name_off2[r2 + 1] - name_off2[r2];
                memcpy(p, names1 + name_off1[r1], (size_t)l1); p += l1; *p++ = ' ';
                p = put_int(p, loc1[k]); *p++ = ' ';
                memcpy(p, names2 + name_off2[r2], (size_t)l2); p += l2; *p++ = ' ';
                p = put_int(p, loc2[k]); *p++ = ' ';
                if (strand_code1[r1] == strand_code2[r2]) *p++ = '1';                 // run_colliearity.py:185-187
                else { *p++ = '-'; *p++ = '1'; }
                *p++ = '\n';
            }
        }
    });
    FILE *f = fopen(path, "wb");
    if (!f) { set_err(err, errlen, std::string("cannot write ") + path + ": " + strerror(errno)); return -2; }
    size_t w = total > 0 ? fwrite(buf.data(), 1, (size_t)total, f) : 0;
    int rc = fclose(f);
    if ((int64_t)w != total || rc != 0) { set_err(err, errlen, "short write"); return -2; }
    return 0;
}


// ------------------------------------------------------------------------------------------------
// host gather of a comparison sharded over GPUs: one rank's blocks to their place in the merged arrays
// ------------------------------------------------------------------------------------------------
extern "C" int icl_host_place_blocks(int64_t n_blocks, const int32_t *pair, const int32_t *global_pair,
                                     const int64_t *block_base, const int64_t *member_base, const int64_t *member_off,
                                     int32_t n_block_fields, const void *const *block_src, void *const *block_dst,
                                     const int32_t *block_bytes, int32_t n_member_fields, const void *const *member_src,
                                     void *const *member_dst, const int32_t *member_bytes, int32_t *out_pair,
                                     int64_t *out_member_off, const int32_t *member_idx, const int64_t *rows,
                                     int32_t *out_member_idx) {
    if (n_blocks < 0 || (n_blocks > 0 && (!pair || !global_pair || !block_base || !member_base || !member_off || !out_pair || !out_member_off)))
        return -1;
    int64_t first = 0;                                         // first block of the current pair's run
    for (int64_t b = 0; b < n_blocks; b++) {
        if (b > 0 && pair[b] != pair[b - 1]) first = b;
        const int32_t g = global_pair[pair[b]];
        const int64_t dest = block_base[g] + (b - first);
        const int64_t m0 = member_off[b], m1 = member_off[b + 1], mdest = member_base[g] + (m0 - member_off[first]);
        out_pair[dest] = g;
        out_member_off[dest] = mdest;
        for (int32_t k = 0; k < n_block_fields; k++) {
            const int32_t w = block_bytes[k];
            memcpy((char *)block_dst[k] + dest * w, (const char *)block_src[k] + b * w, (size_t)w);
        }
        for (int32_t k = 0; k < n_member_fields; k++) {
            const int32_t w = member_bytes[k];
            memcpy((char *)member_dst[k] + mdest * w, (const char *)member_src[k] + m0 * w, (size_t)(m1 - m0) * w);
        }
        if (member_idx && rows && out_member_idx)
            for (int64_t i = m0; i < m1; i++) out_member_idx[mdest + (i - m0)] = (int32_t)rows[member_idx[i]];
    }
    return 0;
}
