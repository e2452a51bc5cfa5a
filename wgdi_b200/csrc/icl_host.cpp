// icl_host.cpp — native host helpers of the `-icl` path: m8 ingest, grading, collinearity-file writer.
// See include/icl_host.h for the reference lines each entry point replaces.  No CUDA here.
#include "../../include/icl_host.h"

#include <fcntl.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <climits>
#include <new>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ICL_HOST_TIMING=1: phase times of the entry points on stderr (diagnostics)
struct PhaseClock {
    bool on;
    std::chrono::steady_clock::time_point t;
    PhaseClock() : on(getenv("ICL_HOST_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "icl_host: %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

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
    std::exception_ptr failure;                         // the first exception of any worker, rethrown in the caller
    std::mutex failure_lock;
    auto work = [&]() {
        try {
            for (;;) {
                int64_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(failure_lock);
            if (!failure) failure = std::current_exception();
            next.store(n);                              // nobody starts another index
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (failure) std::rethrow_exception(failure);
}

inline uint64_t hash_bytes(const char *p, size_t n) {   // 8 bytes per multiplication; ids of >= 8 bytes end with an overlapping word
    uint64_t h = 0x9E3779B97F4A7C15ull ^ ((uint64_t)n * 0xff51afd7ed558ccdull);
    if (n >= 8) {
        const char *last = p + n - 8;
        for (; p < last; p += 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            h ^= h >> 32;
        }
        uint64_t w;
        memcpy(&w, last, 8);
        h = (h ^ w) * 0x9FB21C651E98DF25ull;
    } else {
        uint64_t w = 0;
        for (size_t i = 0; i < n; i++) w |= (uint64_t)(unsigned char)p[i] << (8 * i);
        h = (h ^ w) * 0x9FB21C651E98DF25ull;
    }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

// open-addressing map  gene id (bytes) -> row in the gene table.  A slot holds the id itself when it has at most 26
// bytes (32-byte slots, two per cache line): a lookup touches ONE line instead of slot -> offsets -> id bytes, and
// `prefetch` asks for that line while the row's numbers are still being parsed.
struct IdMap {
    struct Slot { int32_t row; uint8_t len; char key[27]; };   // len 255: the id lives in `ids` only
    const char *ids = nullptr;
    const int64_t *off = nullptr;
    Slot *slot = nullptr;                  // 2 MB aligned and advised for huge pages: a lookup is a random access into
    size_t slot_bytes = 0;                 // megabytes, and under 4 KB pages it would pay a TLB miss before its cache miss
    uint64_t mask = 0;
    IdMap() = default;
    IdMap(const IdMap &) = delete;
    IdMap &operator=(const IdMap &) = delete;
    ~IdMap() { free(slot); }
    static bool same(const Slot &s, const char *p, size_t len, const char *ids, const int64_t *off) {
        if (s.len != 255) return s.len == len && !memcmp(s.key, p, len);
        return (size_t)(off[s.row + 1] - off[s.row]) == len && !memcmp(ids + off[s.row], p, len);
    }
    void build(const char *ids_, const int64_t *off_, int64_t n) {
        ids = ids_; off = off_;
        uint64_t cap = 16;
        while (cap < (uint64_t)n * 2 + 8) cap <<= 1;
        const size_t huge = (size_t)2 << 20;
        slot_bytes = (cap * sizeof(Slot) + huge - 1) / huge * huge;
        void *mem = nullptr;
        if (posix_memalign(&mem, huge, slot_bytes) != 0) throw std::bad_alloc();
        madvise(mem, slot_bytes, MADV_HUGEPAGE);
        slot = (Slot *)mem;
        memset(mem, 0, slot_bytes);
        for (uint64_t i = 0; i < cap; i++) slot[i].row = -1;
        mask = cap - 1;
        for (int64_t i = 0; i < n; i++) {
            const char *p = ids + off[i];
            size_t len = (size_t)(off[i + 1] - off[i]);
            uint64_t h = hash_bytes(p, len) & mask;
            for (;;) {
                Slot &s = slot[h];
                if (s.row < 0) {
                    s.row = (int32_t)i;
                    if (len <= 26) { s.len = (uint8_t)len; memcpy(s.key, p, len); }
                    else s.len = 255;
                    break;
                }
                // duplicated gene ids: the first row wins (pandas would raise on a non-unique index)
                if (same(s, p, len, ids, off)) break;
                h = (h + 1) & mask;
            }
        }
    }
    uint64_t hash(const char *p, size_t len) const { return hash_bytes(p, len) & mask; }
    void prefetch(uint64_t h) const { __builtin_prefetch(&slot[h]); }
    int32_t find_at(uint64_t h, const char *p, size_t len) const {
        for (;;) {
            const Slot &s = slot[h];
            if (s.row < 0) return -1;
            if (same(s, p, len, ids, off)) return s.row;
            h = (h + 1) & mask;
        }
    }
    int32_t find(const char *p, size_t len) const { return find_at(hash(p, len), p, len); }
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
inline bool parse_simple(const char *p, const char *q, uint64_t &w, int &e10, bool &neg, int *nsig = nullptr) {
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
    if (nsig) *nsig = sig;
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
    // decades that decide without any arithmetic: a positive value below 10^dec_below is below the limit and one of at
    // least 10^dec_above is not, each with a relative margin of 1e-12 (far more than the rounding of the conversion)
    int dec_below = -100000, dec_above = 100000;
    explicit LessThan(double lim) : limit(lim) {
        long double band = fabsl(limit) * 1e-15L;
        lo = limit - band; hi = limit + band;
        for (int k = -400; k <= 400; k++) p10[k + 400] = powl(10.0L, (long double)k);
        if (lim > 0 && std::isfinite(lim)) {
            for (int k = -400; k <= 400; k++) {
                if (p10[k + 400] * (1.0L + 1e-12L) <= limit) dec_below = k;
                if (dec_above == 100000 && p10[k + 400] * (1.0L - 1e-12L) > limit) dec_above = k;
            }
        }
    }
    // returns 1 / 0, or -1 when the token does not parse (NaN in pandas: comparison false)
    int test(const char *p, const char *q) const {
        uint64_t w; int e10; bool neg; int sig;
        if (parse_simple(p, q, w, e10, neg, &sig) && e10 >= -400 && e10 <= 400) {
            if (!neg && w != 0) {                          // value in [10^E, 10^(E+1))
                const int E = e10 + sig - 1;
                if (E + 1 <= dec_below) return 1;
                if (E >= dec_above) return 0;
            }
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

// One row: the end of line (NULL: the data ends first) and the starts of the fields behind the first 12 tabs.
// 16 bytes per step where SSE2 is there (every x86-64): a byte loop spends more on its mispredicted tab branches
// than everything else the parser does with the row.
inline const char *scan_row(const char *p, const char *end, const char **f, int &nf) {
    const char *q = p;
#if defined(__SSE2__)
    const __m128i tabs = _mm_set1_epi8('\t'), nls = _mm_set1_epi8('\n');
    while (q + 16 <= end) {
        const __m128i v = _mm_loadu_si128((const __m128i *)q);
        unsigned mt = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, tabs));
        const unsigned mn = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, nls));
        const char *eol = nullptr;
        if (mn) {
            const unsigned b = (unsigned)__builtin_ctz(mn);
            mt &= (1u << b) - 1u;
            eol = q + b;
        }
        while (mt) {
            const int b = __builtin_ctz(mt);
            mt &= mt - 1;
            if (nf < 12) f[++nf] = q + b + 1;
        }
        if (eol) return eol;
        q += 16;
    }
#endif
    for (; q < end; q++) {
        if (*q == '\n') return q;
        if (*q == '\t' && nf < 12) f[++nf] = q + 1;
    }
    return nullptr;
}

struct ChunkOut {
    std::vector<int64_t> row;
    std::vector<int32_t> g1, g2;
    std::vector<double> bits;
    int64_t n_rows = 0;
    int32_t g1min = INT32_MAX, g1max = -1;
    bool all_int = true;
    std::string error;
};

// Stable grouping of a flat index space by an integer key (a query gene): the keys come as consecutive segments
// (the parser's chunks, or one array).  The key range is cut into parts; a part's thread walks ALL keys twice
// (sequential reads) and builds the CSR of its own key range: start[k - lo] .. start[k - lo + 1] index `flat`
// (flat positions in input order = file order), `aux` and `val` (the segment's other arrays at the same positions).
struct Seg { const int32_t *key; const int32_t *aux; const double *val; int64_t n; int32_t kmin, kmax; };   // kmin..kmax: key range of the segment

template <class G>
void group_parts(const std::vector<Seg> &segs, int64_t n_keys, int threads, bool want_val, G per_part) {
    if (n_keys <= 0) return;
    const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(n_keys, (int64_t)std::max(threads, 1) * 2));
    parallel_for(parts, threads, [&](int64_t part) {
        const int64_t lo = n_keys * part / parts, hi = n_keys * (part + 1) / parts;
        const uint64_t width = (uint64_t)(hi - lo);
        std::vector<int64_t> start(width + 1, 0);
        // a BLAST table is grouped by query: most chunks hold a narrow key range and are skipped by most parts
        for (const Seg &sg : segs) {
            if (sg.kmax < lo || sg.kmin >= hi) continue;
            for (int64_t i = 0; i < sg.n; i++) {
                const uint64_t k = (uint64_t)((int64_t)sg.key[i] - lo);
                if (k < width) start[k + 1]++;
            }
        }
        for (uint64_t k = 0; k < width; k++) start[k + 1] += start[k];
        const int64_t total = start[width];
        std::vector<int64_t> flat((size_t)total), fill(start.begin(), start.end() - 1);
        std::vector<int32_t> aux;
        const bool want_aux = !segs.empty() && segs[0].aux != nullptr;
        if (want_aux) aux.resize((size_t)total);
        std::vector<double> val;
        if (want_val) val.resize((size_t)total);
        int64_t base = 0;
        for (const Seg &sg : segs) {
            if (sg.kmax < lo || sg.kmin >= hi) { base += sg.n; continue; }
            for (int64_t i = 0; i < sg.n; i++) {
                const uint64_t k = (uint64_t)((int64_t)sg.key[i] - lo);
                if (k < width) {
                    const int64_t pos = fill[k]++;
                    flat[(size_t)pos] = base + i;
                    if (want_aux) aux[(size_t)pos] = sg.aux[i];
                    if (want_val) val[(size_t)pos] = sg.val[i];
                }
            }
            base += sg.n;
        }
        per_part(lo, hi, start, flat, aux, val);
    });
}

}  // namespace

extern "C" void icl_host_free_hits(icl_hits *h) {
    if (!h) return;
    free(h->row); free(h->g1); free(h->g2); free(h->bits);
    memset(h, 0, sizeof(*h));
}

namespace {

// what the parser leaves behind: the rows that pass the row filters (base.py:178-179) in file order, chunk by chunk
struct Parsed {
    std::vector<ChunkOut> outs;
    std::vector<int64_t> cbase, rbase;     // first flat index / first file row of every chunk
    int64_t n_chunks = 0, total = 0, rows_file = 0;
    bool all_int = true;
    std::vector<Seg> segs() const {
        std::vector<Seg> s;
        for (int64_t c = 0; c < n_chunks; c++)
            s.push_back(Seg{outs[c].g1.data(), outs[c].g2.data(), outs[c].bits.data(), (int64_t)outs[c].g1.size(), outs[c].g1min,
                            outs[c].g1max});
        return s;
    }
};

int parse_m8(const char *who, const char *path, int32_t reverse, double min_score, double max_evalue,
             const char *ids1, const int64_t *off1, int64_t n1, const char *ids2, const int64_t *off2, int64_t n2,
             int32_t threads, bool want_row, Parsed &P, PhaseClock &clk, char *err, int32_t errlen) {
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
    clk.lap("m8 id maps");
    const IdMap &map1 = m1, &map2 = same ? m1 : m2;

    int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>((int64_t)threads * 4, (int64_t)(size >> 16) + 1));
    std::vector<size_t> start(n_chunks + 1, size);
    start[0] = 0;
    for (int64_t c = 1; c < n_chunks; c++) {            // chunk boundaries at line starts
        size_t p = size / n_chunks * c;
        while (p < size && data[p - 1] != '\n') p++;
        start[c] = p;
    }
    for (int64_t c = 1; c <= n_chunks; c++) start[c] = std::max(start[c], start[c - 1]);
    P.n_chunks = n_chunks;
    P.outs.assign((size_t)n_chunks, ChunkOut());
    std::vector<ChunkOut> &outs = P.outs;
    const LessThan evalue_below(max_evalue);
    parallel_for(n_chunks, threads, [&](int64_t c) {
        ChunkOut &o = outs[c];
        const char *p = data + start[c], *end = data + start[c + 1];
        {   // BLAST rows are 50-70 bytes: room for every row of a typical chunk (untouched pages cost nothing), growth otherwise
            const size_t room = (size_t)(end - p) / 40 + 1;
            if (want_row) o.row.reserve(room);
            o.g1.reserve(room); o.g2.reserve(room); o.bits.reserve(room);
        }
        // Rows that pass the number filters wait in a short queue with their subject's map slot prefetched: the slot
        // table of a large genome does not fit the L2 cache, and a lookup right after the hash would wait ~100 ns
        // for its line.  Consecutive rows of a BLAST table share their query: the last id looked up in map 1 is
        // remembered.
        struct Pending { const char *a0, *b0; uint32_t la, lb; uint64_t hb; int64_t row; double bs; };
        constexpr int kQueue = 24;
        Pending pend[kQueue];
        int n_pend = 0;
        const char *last_a = nullptr;
        size_t last_la = 0;
        int32_t last_g1 = -1;
        auto flush = [&]() {
            for (int i = 0; i < n_pend; i++) {
                const Pending &q = pend[i];
                int32_t g1;
                if (last_a && q.la == last_la && !memcmp(q.a0, last_a, q.la)) g1 = last_g1;
                else { g1 = map1.find(q.a0, q.la); last_a = q.a0; last_la = q.la; last_g1 = g1; }
                if (g1 < 0) continue;                                                     // base.py:179
                const int32_t g2 = map2.find_at(q.hb, q.b0, q.lb);
                if (g2 < 0) continue;
                if (want_row) o.row.push_back(q.row);
                o.g1min = std::min(o.g1min, g1); o.g1max = std::max(o.g1max, g1);
                o.g1.push_back(g1); o.g2.push_back(g2); o.bits.push_back(q.bs);
            }
            n_pend = 0;
        };
        while (p < end) {
            const char *f[13];
            int nf = 0;
            f[0] = p;
            const char *eol = scan_row(p, end, f, nf);
            const char *le = eol ? eol : end;
            const char *next = eol ? eol + 1 : end;
            if (le > p && le[-1] == '\r') le--;
            if (le == p) { p = next; continue; }        // blank line (pandas: skip_blank_lines)
            // f[k] = start of field k; field k ends at f[k+1]-1 (or le for the last one found)
            if (nf < 11) { o.error = "row with fewer than 12 tab-separated columns"; return; }
            const char *e11 = le;
            if (nf == 12) e11 = f[12] - 1;
            const int64_t row = o.n_rows++;
            const char *a0 = f[0], *a1 = f[1] - 1, *b0 = f[1], *b1 = f[2] - 1;
            if (reverse) { std::swap(a0, b0); std::swap(a1, b1); }
            const size_t la = (size_t)(a1 - a0), lb = (size_t)(b1 - b0);
            bool ok2 = true;
            if (o.all_int && !is_int_literal(f[11], e11)) o.all_int = false;
            const int ev_ok = evalue_below.test(f[10], f[11] - 1);
            double bs = parse_double(f[11], e11, &ok2);
            p = next;
            if (ev_ok < 0 || !ok2) continue;             // unparsable -> NaN in pandas -> fails both comparisons
            if (!(bs >= min_score) || ev_ok == 0) continue;                               // base.py:178
            if (la == lb && !memcmp(a0, b0, la)) continue;                                // id1 != id2
            const uint64_t hb = map2.hash(b0, lb);
            map2.prefetch(hb);
            pend[n_pend++] = Pending{a0, b0, (uint32_t)la, (uint32_t)lb, hb, row, bs};
            if (n_pend == kQueue) flush();
        }
        flush();
    });
    clk.lap("m8 parse");
    if (data) munmap((void *)data, size);
    close(fd);
    P.total = 0;
    P.all_int = true;
    for (auto &o : outs) {
        if (!o.error.empty()) { set_err(err, errlen, std::string(who) + ": " + o.error); return -3; }
        P.total += (int64_t)o.g1.size();
        P.all_int = P.all_int && o.all_int;
    }
    P.cbase.assign((size_t)n_chunks + 1, 0);
    P.rbase.assign((size_t)n_chunks + 1, 0);
    for (int64_t c = 0; c < n_chunks; c++) {
        P.cbase[c + 1] = P.cbase[c] + (int64_t)outs[c].g1.size();
        P.rbase[c + 1] = P.rbase[c] + outs[c].n_rows;
    }
    P.rows_file = P.rbase[n_chunks];
    return 0;
}

// keep='first' on (id1, id2) in file order (base.py:180) inside one query's rows (file order): marks the later rows
// of an id2 the group has already seen — a handful of comparisons for a short group, a small private table otherwise.
// `seen` is scratch.  Returns the number of rows marked.
template <class Mark>
inline int64_t mark_group_duplicates(const int32_t *sub, int64_t m, std::vector<int32_t> &seen, Mark mark) {
    int64_t n_dup = 0;
    if (m < 2) return 0;
    if (m <= 12) {
        for (int64_t j = 1; j < m; j++)
            for (int64_t i = 0; i < j; i++)
                if (sub[i] == sub[j]) { mark(j); n_dup++; break; }
        return n_dup;
    }
    uint64_t cap = 32;
    while (cap < (uint64_t)m * 2) cap <<= 1;
    seen.assign(cap, -1);
    const uint64_t mask = cap - 1;
    for (int64_t j = 0; j < m; j++) {
        const int32_t g = sub[j];
        uint64_t h = ((uint64_t)(uint32_t)g * 0x9E3779B97F4A7C15ull >> 32) & mask;
        for (;;) {
            if (seen[h] < 0) { seen[h] = g; break; }
            if (seen[h] == g) { mark(j); n_dup++; break; }
            h = (h + 1) & mask;
        }
    }
    return n_dup;
}

}  // namespace

extern "C" int icl_host_read_m8(const char *path, int32_t reverse, double min_score, double max_evalue,
                                const char *ids1, const int64_t *off1, int64_t n1,
                                const char *ids2, const int64_t *off2, int64_t n2,
                                int32_t threads, icl_hits *out, char *err, int32_t errlen) try {
    if (!path || !out || !off1 || !off2 || (n1 > 0 && !ids1) || (n2 > 0 && !ids2)) { set_err(err, errlen, "icl_host_read_m8: null argument"); return -1; }
    memset(out, 0, sizeof(*out));
    if (threads < 1) threads = 1;
    PhaseClock clk;
    Parsed P;
    if (int rc = parse_m8("icl_host_read_m8", path, reverse, min_score, max_evalue, ids1, off1, n1, ids2, off2, n2, threads, true, P,
                          clk, err, errlen))
        return rc;
    const int64_t n_chunks = P.n_chunks, total = P.total;
    const std::vector<ChunkOut> &outs = P.outs;
    const std::vector<int64_t> &cbase = P.cbase, &rbase = P.rbase;
    // duplicates share their id1, so the rows are grouped by gene 1 (group_parts: stable, file order inside a group)
    std::vector<uint8_t> dup((size_t)std::max<int64_t>(total, 1), 0);
    group_parts(P.segs(), n1, threads, false, [&](int64_t lo, int64_t hi, const std::vector<int64_t> &gs, const std::vector<int64_t> &flat,
                                                  const std::vector<int32_t> &sub, const std::vector<double> &) {
        std::vector<int32_t> seen;
        for (int64_t q = lo; q < hi; q++) {
            const int64_t b = gs[(size_t)(q - lo)], m = gs[(size_t)(q - lo) + 1] - b;
            mark_group_duplicates(sub.data() + b, m, seen, [&](int64_t j) { dup[(size_t)flat[(size_t)(b + j)]] = 1; });
        }
    });
    clk.lap("read_m8 dedup");
    std::vector<int64_t> obase((size_t)n_chunks + 1, 0);
    for (int64_t c = 0; c < n_chunks; c++) {
        int64_t kept = 0;
        for (size_t i = 0; i < outs[c].g1.size(); i++) kept += !dup[(size_t)cbase[c] + i];
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
        for (size_t i = 0; i < o.g1.size(); i++) {
            if (dup[(size_t)cbase[c] + i]) continue;
            out->row[k] = rbase[c] + o.row[i]; out->g1[k] = o.g1[i]; out->g2[k] = o.g2[i]; out->bits[k] = o.bits[i];
            k++;
        }
    });
    clk.lap("read_m8 compact");
    out->n = kept_total;
    out->n_rows_file = P.rows_file;
    out->bits_is_int = P.all_int ? 1 : 0;
    return 0;
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_read_m8: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_read_m8: ") + e.what());
    return -6;
}

// ------------------------------------------------------------------------------------------------
// fused ingest of `comparison = genomes`: m8 table -> graded points, nothing per hit crosses into Python
// ------------------------------------------------------------------------------------------------
struct icl_ingest {
    Parsed P;
    int32_t threads = 1;
    int64_t n1 = 0, n2 = 0, n_hits = 0;
    int64_t red = 0, blue = 0, rn = 0;
    std::vector<int8_t> cls;               // per flat index: 0 dropped (duplicate or beyond repeat_number), 1 / 2 / 3
    std::vector<int64_t> amb_off;          // CSR over the ambiguous queries
    std::vector<int64_t> amb_flat;         // flat index of every hit of an ambiguous query (file order, duplicates removed)
    std::vector<double> amb_bits;
    std::vector<int64_t> kept_base;        // per chunk: points before it (filled by icl_host_ingest_count)
};

extern "C" void icl_host_ingest_close(icl_ingest *h) { delete h; }

extern "C" int icl_host_ingest_open(const char *path, int32_t reverse, double min_score, double max_evalue,
                                    const char *ids1, const int64_t *off1, int64_t n1,
                                    const char *ids2, const int64_t *off2, int64_t n2,
                                    int32_t multiple, int32_t repeat_number, int32_t blue_extra,
                                    int32_t threads, icl_ingest **out, icl_ingest_info *info, char *err, int32_t errlen) try {
    if (!path || !out || !info || !off1 || !off2 || (n1 > 0 && !ids1) || (n2 > 0 && !ids2)) { set_err(err, errlen, "icl_host_ingest_open: null argument"); return -1; }
    *out = nullptr;
    memset(info, 0, sizeof(*info));
    if (threads < 1) threads = 1;
    PhaseClock clk;
    std::unique_ptr<icl_ingest> owner(new icl_ingest());    // released to the caller on success only
    icl_ingest *h = owner.get();
    h->threads = threads; h->n1 = n1; h->n2 = n2;
    if (int rc = parse_m8("icl_host_ingest_open", path, reverse, min_score, max_evalue, ids1, off1, n1, ids2, off2, n2, threads, false,
                          h->P, clk, err, errlen))
        return rc;
    const int64_t rn = repeat_number;
    const int64_t red = std::min<int64_t>(multiple, rn), blue = std::min<int64_t>((int64_t)blue_extra + (int64_t)multiple, rn);  // :64,:67
    h->red = red; h->blue = blue; h->rn = rn;
    h->cls.assign((size_t)std::max<int64_t>(h->P.total, 1), 0);
    // one pass over the queries: duplicates out (base.py:180), then the per-query rank by bit score descending
    // (run_colliearity.py:66-73) — stable here; a query whose classes depend on the order of tied scores is set aside
    struct PartAmb { std::vector<int64_t> count, flat; std::vector<double> bits; };
    const int64_t max_parts = (int64_t)threads * 2 + 1;
    std::vector<PartAmb> pamb((size_t)max_parts);
    std::vector<int64_t> part_of_lo((size_t)max_parts, -1), part_dups((size_t)max_parts, 0);
    std::atomic<int64_t> part_counter(0);
    int8_t *cls = h->cls.data();
    group_parts(h->P.segs(), n1, threads, true, [&](int64_t lo, int64_t hi, const std::vector<int64_t> &gs, const std::vector<int64_t> &flat,
                                                    const std::vector<int32_t> &sub, const std::vector<double> &val) {
        const int64_t slot = part_counter.fetch_add(1);
        part_of_lo[(size_t)slot] = lo;
        PartAmb &pa = pamb[(size_t)slot];
        std::vector<int32_t> seen, idx;
        std::vector<uint8_t> isdup;
        int64_t dups = 0;
        for (int64_t q = lo; q < hi; q++) {
            const int64_t b = gs[(size_t)(q - lo)], m = gs[(size_t)(q - lo) + 1] - b;
            if (m == 0) continue;
            isdup.assign((size_t)m, 0);
            dups += mark_group_duplicates(sub.data() + b, m, seen, [&](int64_t j) { isdup[(size_t)j] = 1; });
            idx.clear();
            for (int64_t j = 0; j < m; j++)
                if (!isdup[(size_t)j]) idx.push_back((int32_t)j);
            const int64_t k = (int64_t)idx.size();
            const double *v = val.data() + b;
            bool amb = false;
            const int64_t bounds[3] = {red, blue, rn};
            std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return v[x] > v[y]; });
            for (int64_t bd : bounds)
                if (bd > 0 && bd < k && v[idx[(size_t)bd - 1]] == v[idx[(size_t)bd]]) amb = true;
            if (!amb) {
                for (int64_t r = 0; r < k && r < rn; r++)
                    cls[flat[(size_t)(b + idx[(size_t)r])]] = (int8_t)(r < red ? 1 : r < blue ? 2 : 3);
            } else {                                        // the caller ranks it with numpy (SURVEY.md hazard H2)
                pa.count.push_back(k);
                for (int64_t j = 0; j < m; j++)
                    if (!isdup[(size_t)j]) { pa.flat.push_back(flat[(size_t)(b + j)]); pa.bits.push_back(v[j]); }
            }
        }
        part_dups[(size_t)slot] = dups;
    });
    // ambiguous queries in query order (parts by their key range)
    std::vector<int64_t> order;
    for (int64_t s = 0; s < part_counter.load(); s++) order.push_back(s);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return part_of_lo[(size_t)a] < part_of_lo[(size_t)b]; });
    h->amb_off.assign(1, 0);
    int64_t dups = 0;
    for (int64_t s : order) {
        const PartAmb &pa = pamb[(size_t)s];
        for (int64_t c : pa.count) h->amb_off.push_back(h->amb_off.back() + c);
        h->amb_flat.insert(h->amb_flat.end(), pa.flat.begin(), pa.flat.end());
        h->amb_bits.insert(h->amb_bits.end(), pa.bits.begin(), pa.bits.end());
        dups += part_dups[(size_t)s];
    }
    h->n_hits = h->P.total - dups;
    clk.lap("ingest dedup + grade");
    info->n_hits = h->n_hits;
    info->n_rows_file = h->P.rows_file;
    info->bits_is_int = h->P.all_int ? 1 : 0;
    info->n_amb_queries = (int64_t)h->amb_off.size() - 1;
    info->n_amb_hits = (int64_t)h->amb_flat.size();
    *out = owner.release();
    return 0;
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_ingest_open: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_ingest_open: ") + e.what());
    return -6;
}

extern "C" int icl_host_ingest_ambiguous(const icl_ingest *h, int64_t *amb_off, double *amb_bits) try {
    if (!h || !amb_off) return -1;
    memcpy(amb_off, h->amb_off.data(), h->amb_off.size() * sizeof(int64_t));
    if (!h->amb_bits.empty()) {
        if (!amb_bits) return -1;
        memcpy(amb_bits, h->amb_bits.data(), h->amb_bits.size() * sizeof(double));
    }
    return 0;
} catch (const std::exception &) {
    return -5;
}

extern "C" int icl_host_ingest_set_classes(icl_ingest *h, const int8_t *amb_cls) try {
    if (!h || (!h->amb_flat.empty() && !amb_cls)) return -1;
    for (size_t j = 0; j < h->amb_flat.size(); j++) {
        if (amb_cls[j] < 0 || amb_cls[j] > 3) return -1;
        h->cls[(size_t)h->amb_flat[j]] = amb_cls[j];
    }
    return 0;
} catch (const std::exception &) {
    return -5;
}

extern "C" int icl_host_ingest_count(icl_ingest *h, const int32_t *chr1, const int32_t *chr2, int32_t nc1, int32_t nc2,
                                     uint8_t *present, int64_t *n_points) try {
    if (!h || !chr1 || !chr2 || !present || !n_points || nc1 < 1 || nc2 < 1) return -1;
    PhaseClock clk;
    const Parsed &P = h->P;
    const int64_t ncode = (int64_t)nc1 * nc2;
    memset(present, 0, (size_t)ncode);
    h->kept_base.assign((size_t)P.n_chunks + 1, 0);
    std::atomic<int> bad(0);
    parallel_for(P.n_chunks, h->threads, [&](int64_t c) {
        const ChunkOut &o = P.outs[c];
        const int8_t *cls = h->cls.data() + P.cbase[c];
        int64_t kept = 0;
        for (size_t i = 0; i < o.g1.size(); i++) {
            if (!cls[i]) continue;
            const int32_t a = chr1[o.g1[i]], b = chr2[o.g2[i]];
            if (a < 0 || a >= nc1 || b < 0 || b >= nc2) { bad = 1; continue; }
            // shared table, every writer stores the same value; read first, so that after the first hit of a pair the
            // line is only read (no table per chunk: scaffold-level assemblies have millions of chromosome pairs)
            uint8_t *cell = present + (size_t)a * nc2 + b;
            if (!__atomic_load_n(cell, __ATOMIC_RELAXED)) __atomic_store_n(cell, (uint8_t)1, __ATOMIC_RELAXED);
            kept++;
        }
        h->kept_base[(size_t)c + 1] = kept;
    });
    if (bad) return -4;
    for (int64_t c = 0; c < P.n_chunks; c++) h->kept_base[(size_t)c + 1] += h->kept_base[(size_t)c];
    *n_points = h->kept_base[(size_t)P.n_chunks];
    clk.lap("ingest count");
    return 0;
} catch (const std::exception &) {
    return -5;
}

extern "C" int icl_host_ingest_points(const icl_ingest *h, const int32_t *chr1, const int64_t *order1, const int32_t *chr2,
                                      const int64_t *order2, int32_t nc2, const int32_t *pair_rank, const int32_t *grading3,
                                      int32_t *pair_id, int32_t *loc1, int32_t *loc2, int32_t *grading, int32_t *grade_class) try {
    if (!h || !chr1 || !chr2 || !order1 || !order2 || !pair_rank || !grading3 || h->kept_base.empty()) return -1;
    const Parsed &P = h->P;
    if (h->kept_base[(size_t)P.n_chunks] > 0 && (!pair_id || !loc1 || !loc2 || !grading || !grade_class)) return -1;
    PhaseClock clk;
    parallel_for(P.n_chunks, h->threads, [&](int64_t c) {
        const ChunkOut &o = P.outs[c];
        const int8_t *cls = h->cls.data() + P.cbase[c];
        int64_t k = h->kept_base[(size_t)c];
        for (size_t i = 0; i < o.g1.size(); i++) {
            const int8_t cl = cls[i];
            if (!cl) continue;
            const int32_t g1 = o.g1[i], g2 = o.g2[i];
            pair_id[k] = pair_rank[(size_t)chr1[g1] * nc2 + chr2[g2]];
            loc1[k] = (int32_t)order1[g1];
            loc2[k] = (int32_t)order2[g2];
            grading[k] = grading3[cl - 1];
            grade_class[k] = cl - 1;
            k++;
        }
    });
    clk.lap("ingest points");
    return 0;
} catch (const std::exception &) {
    return -5;
}

extern "C" int icl_host_grade(int64_t n, const int32_t *query, const double *bits, int32_t n_query,
                              int32_t multiple, int32_t repeat_number, const int32_t *grading3,
                              int32_t *grade_out, uint8_t *ambiguous, int32_t threads, char *err, int32_t errlen) try {
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
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_grade: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_grade: ") + e.what());
    return -6;
}

// Per-query rank for `-bi` (wgdi/block_info.py:104-105): 0-based rank by bit score descending among a query's hits, -1 beyond
// the first repeat_number; stable here, and a query is flagged when a run of tied scores touches a rank below `class_bound`
// (the homo classes switch there) or the cut at repeat_number — the caller re-ranks exactly those with numpy (hazard H2).
extern "C" int icl_host_rank(int64_t n, const int32_t *query, const double *bits, int32_t n_query, int32_t repeat_number,
                             int32_t class_bound, int64_t *rank_out, uint8_t *ambiguous, int32_t threads,
                             char *err, int32_t errlen) try {
    if (n < 0 || n_query < 0 || (n > 0 && (!query || !bits || !rank_out)) || (n_query > 0 && !ambiguous)) {
        set_err(err, errlen, "icl_host_rank: bad argument");
        return -1;
    }
    for (int64_t i = 0; i < n; i++)
        if (query[i] < 0 || query[i] >= n_query) { set_err(err, errlen, "icl_host_rank: query code out of range"); return -1; }
    if (n_query > 0) memset(ambiguous, 0, (size_t)n_query);
    const int64_t rn = repeat_number, tie_below = std::min<int64_t>(class_bound, rn);
    std::vector<Seg> segs{Seg{query, nullptr, bits, n, 0, n_query > 0 ? n_query - 1 : 0}};
    group_parts(segs, n_query, threads, true, [&](int64_t lo, int64_t hi, const std::vector<int64_t> &gs, const std::vector<int64_t> &flat,
                                                  const std::vector<int32_t> &, const std::vector<double> &val) {
        std::vector<int32_t> idx;
        for (int64_t q = lo; q < hi; q++) {
            const int64_t b = gs[(size_t)(q - lo)], m = gs[(size_t)(q - lo) + 1] - b;
            if (m == 0) continue;
            const double *v = val.data() + b;
            idx.resize((size_t)m);
            for (int64_t j = 0; j < m; j++) idx[(size_t)j] = (int32_t)j;
            std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return v[x] > v[y]; });
            for (int64_t r = 0; r < m; r++) rank_out[flat[(size_t)(b + idx[(size_t)r])]] = r < rn ? r : -1;
            for (int64_t r = 0; r + 1 < m; r++)             // element at rank r ties with the next one
                if (v[idx[(size_t)r]] == v[idx[(size_t)r + 1]] && (r < tie_below || r == rn - 1)) { ambiguous[q] = 1; break; }
        }
    });
    return 0;
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_rank: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_rank: ") + e.what());
    return -6;
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
                                     int32_t threads, char *err, int32_t errlen) try {
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
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_write_blocks: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_write_blocks: ") + e.what());
    return -6;
}


// Headers assembled here instead of one Python f-string per block (run_colliearity.py:143-151, :194-195).
extern "C" int icl_host_write_collinearity(const char *path, int64_t n_sel, const int64_t *score,
                                           const char *ptext, const int64_t *ptext_off, const int32_t *pidx,
                                           const char *pair_text, const int64_t *pair_text_off, const int32_t *pair,
                                           const int64_t *mem_begin, const int64_t *mem_end,
                                           const int32_t *loc1, const int32_t *loc2,
                                           const int32_t *chr1_code, const int32_t *chr2_code,
                                           const int64_t *row_of1, int64_t maxo1, const char *names1, const int64_t *name_off1,
                                           const int32_t *strand_code1,
                                           const int64_t *row_of2, int64_t maxo2, const char *names2, const int64_t *name_off2,
                                           const int32_t *strand_code2,
                                           int32_t threads, char *err, int32_t errlen) try {
    if (!path || n_sel < 0 || (n_sel > 0 && (!score || !ptext || !ptext_off || !pidx || !pair_text || !pair_text_off || !pair ||
                                             !mem_begin || !mem_end || !loc2))) {
        set_err(err, errlen, "icl_host_write_collinearity: bad argument");
        return -1;
    }
    std::string headers;
    std::vector<int64_t> off((size_t)n_sel + 1, 0);
    headers.reserve((size_t)n_sel * 72);
    std::string tail;
    char num[32];
    for (int64_t b = 0; b < n_sel; b++) {
        if (mem_end[b] - mem_begin[b] < 2) { set_err(err, errlen, "icl_host_write_collinearity: a block of fewer than two gene pairs"); return -1; }
        tail.assign(" score=");
        tail.append(num, (size_t)(put_int(num, score[b]) - num));
        tail.append(" pvalue=");
        tail.append(ptext + ptext_off[pidx[b]], (size_t)(ptext_off[pidx[b] + 1] - ptext_off[pidx[b]]));
        tail.append(" N=");
        tail.append(num, (size_t)(put_int(num, mem_end[b] - mem_begin[b]) - num));
        tail.push_back(' ');
        tail.append(pair_text + pair_text_off[pair[b]], (size_t)(pair_text_off[pair[b] + 1] - pair_text_off[pair[b]]));
        tail.append(loc2[mem_begin[b]] < loc2[mem_begin[b] + 1] ? " plus" : " minus");                   // :194-195
        const size_t colon = tail.find(':');                  // :147-151 re-cuts the header at ':' (a chromosome name with one)
        headers.append("# Alignment ");
        headers.append(num, (size_t)(put_int(num, b + 1) - num));
        headers.push_back(':');
        headers.append(tail, 0, colon);
        off[(size_t)b + 1] = (int64_t)headers.size();
    }
    return icl_host_write_blocks(path, n_sel, headers.data(), off.data(), mem_begin, mem_end, loc1, loc2, chr1_code, chr2_code,
                                 row_of1, maxo1, names1, name_off1, strand_code1, row_of2, maxo2, names2, name_off2, strand_code2,
                                 threads, err, errlen);
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_write_collinearity: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_write_collinearity: ") + e.what());
    return -6;
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
                                     int32_t *out_member_idx) try {
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
} catch (const std::exception &) {
    return -5;
}

// ------------------------------------------------------------------------------------------------
// `-bi`: the collinearity file back into arrays (wgdi/base.py:120-141 as wgdi/block_info.py:18-75 consumes it)
// ------------------------------------------------------------------------------------------------
namespace {
struct TextColumn {
    std::string text;
    std::vector<int64_t> off{0};
    void add(const char *p, size_t n) { text.append(p, n); off.push_back((int64_t)text.size()); }
    void close() { off.push_back((int64_t)text.size()); }
};
template <class T>
T *copy_out(const std::vector<T> &v) {
    T *p = (T *)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}
char *copy_out(const std::string &s) {
    char *p = (char *)malloc(std::max<size_t>(s.size(), 1));
    if (p && !s.empty()) memcpy(p, s.data(), s.size());
    return p;
}
}  // namespace

extern "C" void icl_host_free_collinearity(icl_collinearity *c) {
    if (!c) return;
    free(c->id_text); free(c->id_off); free(c->pv_text); free(c->pv_off); free(c->off); free(c->g1); free(c->g2);
    free(c->loc1); free(c->loc2); free(c->block1_text); free(c->block1_off); free(c->block2_text); free(c->block2_off);
    memset(c, 0, sizeof(*c));
}

extern "C" int icl_host_read_collinearity(const char *path, const char *ids1, const int64_t *off1, int64_t n1,
                                          const char *ids2, const int64_t *off2, int64_t n2,
                                          icl_collinearity *out, char *err, int32_t errlen) try {
    if (!path || !out || !off1 || !off2 || (n1 > 0 && !ids1) || (n2 > 0 && !ids2)) { set_err(err, errlen, "icl_host_read_collinearity: null argument"); return -1; }
    memset(out, 0, sizeof(*out));
    PhaseClock clk;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_err(err, errlen, std::string("cannot open ") + path + ": " + strerror(errno)); return -2; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); set_err(err, errlen, "fstat failed"); return -2; }
    const size_t size = (size_t)sb.st_size;
    const char *data = nullptr;
    if (size > 0) {
        data = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (data == MAP_FAILED) { close(fd); set_err(err, errlen, "mmap failed"); return -2; }
    }
    struct Unmap {
        const char *d; size_t n; int fd;
        ~Unmap() { if (d) munmap((void *)d, n); close(fd); }
    } unmap{data, size, fd};
    // only the plain case is handled here: printable ASCII, ' ' and '\n'.  Anything else (tabs, '\r', other white space the
    // reference's re.split(r"\s") would cut at, non-ASCII text) is left to the Python reader: return 1.
    for (size_t i = 0; i < size; i++) {
        const unsigned char ch = (unsigned char)data[i];
        if (ch != '\n' && (ch < 0x20 || ch > 0x7e)) return 1;
    }
    IdMap m1, m2;
    m1.build(ids1, off1, n1);
    const bool same = ids1 == ids2 && off1 == off2 && n1 == n2;
    if (!same) m2.build(ids2, off2, n2);
    const IdMap &map1 = m1, &map2 = same ? m1 : m2;
    TextColumn idc, pvc, b1, b2;
    std::vector<int64_t> boff{0};
    std::vector<int32_t> g1, g2;
    std::vector<double> l1, l2;
    std::string cur_id, cur_pv;
    bool in_blocks = false;
    int64_t count = 0;
    char num[32];
    auto close_block = [&]() {
        if (!count) return;
        idc.add(cur_id.data(), cur_id.size());
        pvc.add(cur_pv.data(), cur_pv.size());
        boff.push_back(boff.back() + count);
        b1.close(); b2.close();
        count = 0;
    };
    auto plain_int = [](const char *p, const char *q, int64_t &v) {      // what float() -> int() -> str() leaves unchanged
        if (p >= q || q - p > 15) return false;
        v = 0;
        for (; p < q; p++) {
            if (*p < '0' || *p > '9') return false;
            v = v * 10 + (*p - '0');
        }
        return true;
    };
    const char *p = data, *end = data + size;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        const char *next = eol ? eol + 1 : end;
        const char *a = p, *b = le;
        while (a < b && *a == ' ') a++;                                 // line.strip()
        while (b > a && b[-1] == ' ') b--;
        p = next;
        if (b - a >= 11 && !memcmp(a, "# Alignment", 11)) {              // re.match(r"# Alignment", line)
            in_blocks = true;
            close_block();
            // re.findall(r'[\.\d+]+', line): tokens 0 and 2
            int tok = 0;
            cur_id.clear(); cur_pv.clear();
            for (const char *q = a; q < b;) {
                auto is_tok = [](char ch) { return (ch >= '0' && ch <= '9') || ch == '.' || ch == '+'; };
                if (!is_tok(*q)) { q++; continue; }
                const char *t = q;
                while (q < b && is_tok(*q)) q++;
                if (tok == 0) cur_id.assign(t, (size_t)(q - t));
                if (tok == 2) { cur_pv.assign(t, (size_t)(q - t)); tok++; break; }
                tok++;
            }
            if (tok < 3) return 1;                                      // the Python reader raises what the reference raises
            continue;
        }
        if (!in_blocks) continue;
        // name1 loc1 name2 loc2 ...: str.split(' ') — consecutive blanks give empty fields, as in the reference
        const char *f[5];
        int nf = 0;
        f[0] = a;
        for (const char *q = a; q < b && nf < 4; q++)
            if (*q == ' ') f[++nf] = q + 1;
        if (nf < 3) return 1;
        const char *e3 = nf >= 4 ? f[4] - 1 : b;
        int64_t v1, v2;
        if (!plain_int(f[1], f[2] - 1, v1) || !plain_int(f[3], e3, v2)) return 1;
        g1.push_back(map1.find(f[0], (size_t)(f[1] - 1 - f[0])));
        g2.push_back(map2.find(f[2], (size_t)(f[3] - 1 - f[2])));
        l1.push_back((double)v1);
        l2.push_back((double)v2);
        if (count) { b1.text.push_back('_'); b2.text.push_back('_'); }
        b1.text.append(num, (size_t)(put_int(num, v1) - num));
        b2.text.append(num, (size_t)(put_int(num, v2) - num));
        count++;
    }
    close_block();
    out->n_blocks = (int64_t)boff.size() - 1;
    out->n_members = (int64_t)g1.size();
    out->id_text = copy_out(idc.text); out->id_off = copy_out(idc.off);
    out->pv_text = copy_out(pvc.text); out->pv_off = copy_out(pvc.off);
    out->off = copy_out(boff);
    out->g1 = copy_out(g1); out->g2 = copy_out(g2);
    out->loc1 = copy_out(l1); out->loc2 = copy_out(l2);
    out->block1_text = copy_out(b1.text); out->block1_off = copy_out(b1.off);
    out->block2_text = copy_out(b2.text); out->block2_off = copy_out(b2.off);
    if (!out->id_text || !out->id_off || !out->pv_text || !out->pv_off || !out->off || !out->g1 || !out->g2 || !out->loc1 ||
        !out->loc2 || !out->block1_text || !out->block1_off || !out->block2_text || !out->block2_off) {
        icl_host_free_collinearity(out);
        set_err(err, errlen, "out of memory");
        return -5;
    }
    clk.lap("read_collinearity");
    return 0;
} catch (const std::bad_alloc &) {
    set_err(err, errlen, "icl_host_read_collinearity: out of memory");
    return -5;
} catch (const std::exception &e) {
    set_err(err, errlen, std::string("icl_host_read_collinearity: ") + e.what());
    return -6;
}
