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
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct scs_csv {
    int64_t V = 0, N = 0, nnz = 0;
    std::string header_first;            // name of the index column ("SNV")
    std::string barcodes;                // '\n'-joined
    std::string ids;                     // '\n'-joined
    std::vector<int64_t> indptr;
    std::vector<int32_t> indices;
    std::vector<int32_t> data;
};

namespace {

struct Chunk {
    const char* b;
    const char* e;
    std::vector<int64_t> row_nnz;
    std::vector<int32_t> idx, val;
    std::string ids;
    int err = 0;          // 1 = not a plain non-negative integer field, 2 = wrong number of fields, 3 = count overflow
    int64_t err_row = 0;
};

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
        int64_t col = 0, nz = 0;
        while (true) {
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
            if (v) { c->idx.push_back((int32_t)col); c->val.push_back((int32_t)v); ++nz; }
            ++col;
            if (q >= le) break;
            ++q;                                                   // the comma
            if (q >= le) { c->err = 1; c->err_row = row; return; }  // trailing comma = empty last field
        }
        if (col != ncols) { c->err = 2; c->err_row = row; return; }
        c->row_nnz.push_back(nz);
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
        std::vector<Chunk> chunks((size_t)nt);
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
                          ch.err == 1 ? "field is not a plain non-negative integer" : ch.err == 2 ? "wrong number of fields" : "count exceeds int32");
                rc = ch.err == 1 ? 2 : 1;   // 2 = declined (a general CSV reader may still accept it), 1 = malformed
                break;
            }
            V += (int64_t)ch.row_nnz.size();
            nnz += (int64_t)ch.idx.size();
        }
        if (rc) break;
        c->V = V;
        c->nnz = nnz;
        c->indptr.resize((size_t)V + 1);
        c->indices.resize((size_t)nnz);
        c->data.resize((size_t)nnz);
        int64_t r = 0, z = 0;
        c->indptr[0] = 0;
        for (auto& ch : chunks) {
            for (int64_t k : ch.row_nnz) { c->indptr[(size_t)r + 1] = c->indptr[(size_t)r] + k; ++r; }
            if (!ch.idx.empty()) {
                memcpy(c->indices.data() + z, ch.idx.data(), ch.idx.size() * 4);
                memcpy(c->data.data() + z, ch.val.data(), ch.val.size() * 4);
                z += (int64_t)ch.idx.size();
            }
            c->ids += ch.ids;
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
    if (indptr) memcpy(indptr, c->indptr.data(), c->indptr.size() * 8);
    if (indices && c->nnz) memcpy(indices, c->indices.data(), (size_t)c->nnz * 4);
    if (data && c->nnz) memcpy(data, c->data.data(), (size_t)c->nnz * 4);
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
