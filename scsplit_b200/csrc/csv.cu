// Streaming parser for the reference's dense count CSVs (written by DataFrame.to_csv at scSplit:462-463, read back with
// pd.read_csv(header=0, index_col=0) + csr_matrix at scSplit:544-546): header "SNV,<barcode>,...", then one row per SNV
// "<CHROM:POS>,<int>,<int>,...".  The reference materialises a dense int64 V x N frame first (4.8 GB at 30k x 20k);
// this goes straight to CSR with one pass over an mmap, rows split across host threads.  Host-side I/O only.
// At the other end of the run, scs_csv_write_f64 renders the float matrices the reference hands to DataFrame.to_csv
// (scSplit:596, 599-600) in the same text form, byte for byte.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// the non-zero fields one parser thread found in its share of the rows (grown by realloc: big blocks are remapped, not copied)
struct Entries {
    int32_t* idx = nullptr;
    int32_t* val = nullptr;
    size_t n = 0, cap = 0;
    Entries() = default;
    Entries(const Entries&) = delete;
    Entries& operator=(const Entries&) = delete;
    ~Entries() { free(idx); free(val); }
    bool reserve(size_t extra) {
        if (n + extra <= cap) return true;
        const size_t want = std::max<size_t>(cap * 2, std::max<size_t>(n + extra, (size_t)1 << 16));
        int32_t* a = (int32_t*)realloc(idx, want * sizeof(int32_t));
        if (a) idx = a;
        int32_t* b = (int32_t*)realloc(val, want * sizeof(int32_t));
        if (b) val = b;
        if (!a || !b) return false;
        cap = want;
        return true;
    }
};

struct Chunk {
    const char* b = nullptr;
    const char* e = nullptr;
    std::vector<int64_t> row_nnz;
    Entries ent;
    std::string ids;
    int err = 0;          // 1 = not a plain non-negative integer field, 2 = wrong number of fields, 3 = count overflow, 4 = out of memory
    int64_t err_row = 0;
};

}  // namespace

struct scs_csv {
    int64_t V = 0, N = 0, nnz = 0;
    std::string header_first;            // name of the index column ("SNV")
    std::string barcodes;                // '\n'-joined
    std::string ids;                     // '\n'-joined
    std::vector<Chunk> chunks;           // the rows in file order, one share per parser thread; merged by scs_csv_fetch
};

namespace {

void parse_chunk(Chunk* c, int64_t ncols) {
    const char* p = c->b;
    int64_t row = 0;
    while (p < c->e) {
        const char* line_end = (const char*)memchr(p, '\n', (size_t)(c->e - p));
        if (!line_end) line_end = c->e;
        const char* le = line_end;
        if (le > p && le[-1] == '\r') --le;
        if (le == p) { p = line_end + 1; continue; }            // blank line
        const char* q = (const char*)memchr(p, ',', (size_t)(le - p));
        if (!q) { c->err = 2; c->err_row = row; return; }
        c->ids.append(p, (size_t)(q - p));
        c->ids.push_back('\n');
        ++q;
        int64_t col = 0;
        const size_t n0 = c->ent.n;
        while (true) {
            // Count matrices are ~95 % zeros and nearly all counts have one digit.  Sixteen bytes = eight fields "d," at a time:
            // one comparison says whether the block has that shape, a bit mask says which of its fields are not "0".  Rows of
            // other shapes, and always the last field of a line, take the general code below.
#if defined(__SSE2__)
            while (q + 16 < le) {
                const __m128i v = _mm_loadu_si128((const __m128i*)q);
                const int commas = _mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8(',')));
                const __m128i d = _mm_sub_epi8(v, _mm_set1_epi8('0'));
                const int digits = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_subs_epu8(d, _mm_set1_epi8(9)), _mm_setzero_si128()));
                if ((commas & 0xAAAA) != 0xAAAA || (digits & 0x5555) != 0x5555) break;
                int nzm = ~_mm_movemask_epi8(_mm_cmpeq_epi8(d, _mm_setzero_si128())) & 0x5555;
                if (nzm) {
                    if (!c->ent.reserve(8)) { c->err = 4; c->err_row = row; return; }
                    do {
                        const int b = __builtin_ctz((unsigned)nzm);
                        c->ent.idx[c->ent.n] = (int32_t)(col + (b >> 1));
                        c->ent.val[c->ent.n++] = (int32_t)(q[b] - '0');
                        nzm &= nzm - 1;
                    } while (nzm);
                }
                col += 8;
                q += 16;
            }
#endif
            if (q + 2 < le && q[1] == ',' && q[0] >= '0' && q[0] <= '9') {       // one single-digit field
                if (q[0] != '0') {
                    if (!c->ent.reserve(1)) { c->err = 4; c->err_row = row; return; }
                    c->ent.idx[c->ent.n] = (int32_t)col;
                    c->ent.val[c->ent.n++] = (int32_t)(q[0] - '0');
                }
                ++col;
                q += 2;
                continue;
            }
            while (q < le && (*q == ' ' || *q == '\t')) ++q;
            int64_t v = 0;
            const char* s = q;
            while (q < le && *q >= '0' && *q <= '9') {
                v = v * 10 + (*q - '0');
                if (v > 2147483647LL) { c->err = 3; c->err_row = row; return; }
                ++q;
            }
            if (q == s) { c->err = 1; c->err_row = row; return; }   // empty, negative, float, quoted ...
            while (q < le && (*q == ' ' || *q == '\t')) ++q;
            if (q < le && *q != ',') { c->err = 1; c->err_row = row; return; }
            if (v) {
                if (!c->ent.reserve(1)) { c->err = 4; c->err_row = row; return; }
                c->ent.idx[c->ent.n] = (int32_t)col;
                c->ent.val[c->ent.n++] = (int32_t)v;
            }
            ++col;
            if (q >= le) break;
            ++q;                                                   // the comma
            if (q >= le) { c->err = 1; c->err_row = row; return; }  // trailing comma = empty last field
        }
        if (col != ncols) { c->err = 2; c->err_row = row; return; }
        c->row_nnz.push_back((int64_t)(c->ent.n - n0));
        ++row;
        p = line_end + 1;
    }
}

// One float64 as numpy / Python print it (what DataFrame.to_csv writes when no float_format is given): the shortest
// digit string that reads back as the same double, positional for 1e-4 <= |x| < 1e16 (always with a fractional part:
// "1.0"), otherwise d[.ddd]e+XX with at least two exponent digits; "inf" / "-inf"; nothing at all for NaN (na_rep='').
// std::to_chars produces the shortest digits; only their layout is decided here.
char* put_shortest(char* p, double x) {
    if (x != x) return p;
    if (std::isinf(x)) {
        if (x < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    char tmp[48];
    const auto r = std::to_chars(tmp, tmp + sizeof tmp, x, std::chars_format::scientific);
    const char* q = tmp;
    if (*q == '-') *p++ = *q++;
    char dig[24];
    int nd = 0;
    for (; q < r.ptr && *q != 'e'; ++q)
        if (*q != '.') dig[nd++] = *q;
    ++q;                                          // 'e'
    const bool eneg = *q == '-';
    ++q;                                          // sign (to_chars always writes one)
    int ex = 0;
    for (; q < r.ptr; ++q) ex = ex * 10 + (*q - '0');
    const int decpt = (eneg ? -ex : ex) + 1;      // value = 0.d1d2... * 10^decpt
    if (decpt <= -4 || decpt > 16) {
        *p++ = dig[0];
        if (nd > 1) {
            *p++ = '.';
            memcpy(p, dig + 1, (size_t)nd - 1);
            p += nd - 1;
        }
        *p++ = 'e';
        int e10 = decpt - 1;
        *p++ = e10 < 0 ? '-' : '+';
        if (e10 < 0) e10 = -e10;
        if (e10 >= 100) { *p++ = (char)('0' + e10 / 100); e10 %= 100; }
        *p++ = (char)('0' + e10 / 10);
        *p++ = (char)('0' + e10 % 10);
    } else if (decpt <= 0) {
        *p++ = '0';
        *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        memcpy(p, dig, (size_t)nd);
        p += nd;
    } else if (decpt >= nd) {
        memcpy(p, dig, (size_t)nd);
        p += nd;
        for (int i = nd; i < decpt; ++i) *p++ = '0';
        *p++ = '.';
        *p++ = '0';
    } else {
        memcpy(p, dig, (size_t)decpt);
        p += decpt;
        *p++ = '.';
        memcpy(p, dig + decpt, (size_t)(nd - decpt));
        p += nd - decpt;
    }
    return p;
}

// "%.0f" (the float_format of scSplit:599-600): the P/A matrices hold 0, 1 and NaN only, anything else goes through printf
char* put_rounded(char* p, double x) {
    if (x != x) return p;
    if (x == 0.0 && !std::signbit(x)) { *p++ = '0'; return p; }
    if (x == 1.0) { *p++ = '1'; return p; }
    return p + snprintf(p, 330, "%.0f", x);
}

}  // namespace

using namespace scs;

extern "C" {

int scs_csv_parse(const char* path, int threads, scs_csv** out) {
    SCS_CHECK(path && out, "null argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    SCS_CHECK(fd >= 0, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); set_error("%s is empty or unreadable", path); return 1; }
    const size_t size = (size_t)st.st_size;
    const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    SCS_CHECK(base != MAP_FAILED, "mmap of %s failed", path);
    madvise((void*)base, size, MADV_SEQUENTIAL);
    scs_csv* c = new scs_csv();
    int rc = 0;
    do {
        const char* end = base + size;
        const char* hl = (const char*)memchr(base, '\n', size);
        if (!hl) hl = end;
        const char* he = hl;
        if (he > base && he[-1] == '\r') --he;
        if (memchr(base, '"', (size_t)(he - base))) { set_error("quoted header fields are not supported by the fast parser"); rc = 2; break; }
        const char* q = (const char*)memchr(base, ',', (size_t)(he - base));
        if (!q) { set_error("header of %s has no columns", path); rc = 1; break; }
        c->header_first.assign(base, (size_t)(q - base));
        ++q;
        int64_t ncols = 0;
        while (q <= he) {
            const char* n = (const char*)memchr(q, ',', (size_t)(he - q));
            if (!n) n = he;
            c->barcodes.append(q, (size_t)(n - q));
            c->barcodes.push_back('\n');
            ++ncols;
            if (n >= he) break;
            q = n + 1;
        }
        c->N = ncols;
        const char* body = hl < end ? hl + 1 : end;
        int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min(nt, 64));
        if ((size_t)(end - body) < (size_t)(1 << 20)) nt = 1;
        std::vector<Chunk>& chunks = c->chunks;
        chunks = std::vector<Chunk>((size_t)nt);
        const char* cur = body;
        for (int i = 0; i < nt; ++i) {
            const char* stop = (i == nt - 1) ? end : body + (size_t)(end - body) / nt * (i + 1);
            if (stop < cur) stop = cur;
            if (i != nt - 1) {
                const char* nl = (const char*)memchr(stop, '\n', (size_t)(end - stop));
                stop = nl ? nl + 1 : end;
            }
            chunks[i].b = cur;
            chunks[i].e = stop;
            cur = stop;
        }
        std::vector<std::thread> pool;
        for (int i = 1; i < nt; ++i) pool.emplace_back(parse_chunk, &chunks[i], ncols);
        parse_chunk(&chunks[0], ncols);
        for (auto& t : pool) t.join();
        int64_t V = 0, nnz = 0;
        for (auto& ch : chunks) {
            if (ch.err) {
                set_error("%s: data row %lld: %s", path, (long long)(V + ch.err_row + 1),
                          ch.err == 1 ? "field is not a plain non-negative integer" : ch.err == 2 ? "wrong number of fields"
                          : ch.err == 3 ? "count exceeds int32" : "out of memory");
                rc = ch.err == 1 ? 2 : 1;   // 2 = declined (a general CSV reader may still accept it), 1 = malformed
                break;
            }
            V += (int64_t)ch.row_nnz.size();
            nnz += (int64_t)ch.ent.n;
            ch.b = ch.e = nullptr;          // the mapping goes away below
        }
        if (rc) break;
        c->V = V;
        c->nnz = nnz;
        for (auto& ch : chunks) {
            c->ids += ch.ids;
            std::string().swap(ch.ids);
        }
    } while (false);
    munmap((void*)base, size);
    if (rc) { delete c; return rc; }
    *out = c;
    return 0;
}

int scs_csv_shape(scs_csv* c, int64_t out[5]) {
    SCS_CHECK(c && out, "null argument");
    out[0] = c->V; out[1] = c->N; out[2] = c->nnz; out[3] = (int64_t)c->ids.size(); out[4] = (int64_t)c->barcodes.size();
    return 0;
}

int scs_csv_fetch(scs_csv* c, int64_t* indptr, int32_t* indices, int32_t* data, char* ids, char* barcodes) {
    SCS_CHECK(c, "null argument");
    // every parser thread's share goes straight into the caller's arrays (its rows and entries start where the shares before it end)
    std::vector<int64_t> row0(c->chunks.size() + 1, 0), ent0(c->chunks.size() + 1, 0);
    for (size_t i = 0; i < c->chunks.size(); ++i) {
        row0[i + 1] = row0[i] + (int64_t)c->chunks[i].row_nnz.size();
        ent0[i + 1] = ent0[i] + (int64_t)c->chunks[i].ent.n;
    }
    auto place = [&](size_t i) {
        const Chunk& ch = c->chunks[i];
        if (indptr) {
            int64_t at = ent0[i];
            int64_t* o = indptr + row0[i];
            for (int64_t k : ch.row_nnz) { *o++ = at; at += k; }
        }
        if (indices && ch.ent.n) memcpy(indices + ent0[i], ch.ent.idx, ch.ent.n * 4);
        if (data && ch.ent.n) memcpy(data + ent0[i], ch.ent.val, ch.ent.n * 4);
    };
    std::vector<std::thread> pool;
    for (size_t i = 1; i < c->chunks.size(); ++i) pool.emplace_back(place, i);
    if (!c->chunks.empty()) place(0);
    for (auto& t : pool) t.join();
    if (indptr) indptr[c->V] = c->nnz;
    if (ids) memcpy(ids, c->ids.data(), c->ids.size());
    if (barcodes) memcpy(barcodes, c->barcodes.data(), c->barcodes.size());
    return 0;
}

int scs_csv_write_f64(const char* path, const char* header, const char* labels, int64_t rows, int64_t cols, const double* values,
                      int format) {
    SCS_CHECK(path && header && (labels || rows == 0) && (values || rows * cols == 0), "null argument");
    SCS_CHECK(rows >= 0 && cols >= 0, "negative shape");
    SCS_CHECK(format == 0 || format == 1, "unknown float format %d", format);
    FILE* f = fopen(path, "wb");
    SCS_CHECK(f, "cannot open %s for writing", path);
    std::vector<char> buf((size_t)4 << 20);
    const size_t field = format == 0 ? 32 : 336;            // longest rendering of one value, separator included
    char* p = buf.data();
    bool ok = true;
    auto flush = [&]() {
        if (p != buf.data() && fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) != (size_t)(p - buf.data())) ok = false;
        p = buf.data();
    };
    const size_t hl = strlen(header);
    if (fwrite(header, 1, hl, f) != hl || fputc('\n', f) == EOF) ok = false;
    const char* lab = labels;
    for (int64_t r = 0; r < rows && ok; ++r) {
        const char* le = strchr(lab, '\n');
        if (!le) { fclose(f); set_error("row label %lld of %lld is missing", (long long)r, (long long)rows); return 1; }
        const size_t need = (size_t)(le - lab) + (size_t)cols * field + 2;
        if (need > buf.size()) { flush(); buf.resize(need); p = buf.data(); }
        if ((size_t)(buf.data() + buf.size() - p) < need) flush();
        memcpy(p, lab, (size_t)(le - lab));
        p += le - lab;
        lab = le + 1;
        const double* row = values + r * cols;
        for (int64_t c = 0; c < cols; ++c) {
            *p++ = ',';
            p = format == 0 ? put_shortest(p, row[c]) : put_rounded(p, row[c]);
        }
        *p++ = '\n';
    }
    flush();
    if (fclose(f) != 0) ok = false;
    SCS_CHECK(ok, "write to %s failed", path);
    return 0;
}

int scs_csv_free(scs_csv* c) {
    delete c;
    return 0;
}

}  // extern "C"
