// Native counterparts of the reference's preprocessing scripts, byte-compatible with
// their outputs:
//   preprocessing/map_items.py:21-96      -> cu2b_prep_map      (sequential ids in first-appearance
//                                            order, rows grouped by ascending user, input order kept)
//   preprocessing/map_netflix.py:9-27     -> cu2b_prep_map with a second file (shared mappings,
//                                            rows of unknown users / items skipped), delimiter ' ',
//                                            rating in column 3
//   preprocessing/sort_ratings.py:29-37   -> cu2b_prep_sort     (by user, then item, stable)
//   preprocessing/split_to_test_train.py:39-49,71-82 -> cu2b_prep_split (split_true: one shuffle of
//                                            all rows with Python's random.seed(seed) /
//                                            random.shuffle stream, cut, stable sort by user)
//   preprocessing/create_config.py:10-19  -> cu2b_prep_create_config
//   preprocessing/convert_to_np.py:6-13   -> cu2b_prep_convert_to_np (np.genfromtxt(delimiter=',')
//                                            -> np.save: float64 .npy, version 1.0 header)
// The scripts hold every row as Python objects (minutes and tens of GB at the Netflix size); here the
// file is mmap-ed, parsed once, and written with a block formatter. Ratings are parsed to double and
// printed as Python prints a float (shortest round-trip repr, ".0" for integral values), ids as
// 64-bit integers. Host code only; nothing here touches the GPU.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <omp.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "cu2b_internal.h"

#define CU2B_TRY_STATUS(expr)               \
    do {                                    \
        cu2b_status s__ = (expr);           \
        if (s__ != CU2B_OK) return s__;     \
    } while (0)

namespace {

struct Row {
    int64_t user, item;
    double rating;
};

struct Mapped {
    const char *base = nullptr;
    size_t size = 0;
    int fd = -1;
    ~Mapped() {
        if (base && size) munmap((void *)base, size);
        if (fd >= 0) close(fd);
    }
    cu2b_status open_file(const char *path) {
        fd = open(path, O_RDONLY);
        if (fd < 0) return cu2b_fail(CU2B_ERR_IO, "cannot open %s", path);
        struct stat st;
        if (fstat(fd, &st) != 0) return cu2b_fail(CU2B_ERR_IO, "cannot stat %s", path);
        size = (size_t)st.st_size;
        if (size == 0) return CU2B_OK;
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { size = 0; return cu2b_fail(CU2B_ERR_IO, "mmap failed for %s", path); }
        base = (const char *)p;
        return CU2B_OK;
    }
};

// Splits one line into fields the way csv.reader does for these files (no quoting in them):
// every occurrence of the delimiter separates two fields, so "a  b" with ' ' has an empty field.
// Fields are [begin, end) pointers into the line.
int split_fields(const char *p, const char *end, char delim, const char **fb, const char **fe, int cap) {
    int n = 0;
    const char *s = p;
    for (const char *q = p;; ++q) {
        if (q == end || *q == delim) {
            if (n < cap) { fb[n] = s; fe[n] = q; }
            ++n;
            if (q == end) break;
            s = q + 1;
        }
    }
    return n;
}

// int(str): optional surrounding whitespace, optional sign, digits.
bool parse_int(const char *b, const char *e, int64_t *out) {
    while (b < e && (*b == ' ' || *b == '\t')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) --e;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; ++b; }
    if (b == e) return false;
    int64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
    }
    *out = neg ? -v : v;
    return true;
}

bool parse_float(const char *b, const char *e, double *out) {
    while (b < e && (*b == ' ' || *b == '\t')) ++b;
    while (e > b && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r')) --e;
    if (b == e) return false;
    char buf[64];
    const size_t n = std::min<size_t>((size_t)(e - b), sizeof buf - 1);
    memcpy(buf, b, n);
    buf[n] = 0;
    char *endp = nullptr;
    *out = strtod(buf, &endp);
    return endp == buf + n;
}

// Reads "user <d> item <d> ... rating ..." rows. rating_col is the 0-based field index of the rating.
cu2b_status read_rows(const char *path, char delim, bool has_header, int rating_col, std::vector<Row> *rows) {
    Mapped f;
    CU2B_TRY_STATUS(f.open_file(path));
    const char *p = f.base, *end = f.base + f.size;
    int64_t line_no = 0;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const char *next = nl ? nl + 1 : end;
        ++line_no;
        if (has_header && line_no == 1) { p = next; continue; }
        const char *lt = le;
        if (lt > p && lt[-1] == '\r') --lt;
        if (lt == p) { p = next; continue; }  // csv.reader yields [] for a blank line; the scripts never see one
        const char *fb[8], *fe[8];
        const int nf = split_fields(p, lt, delim, fb, fe, 8);
        Row r;
        if (nf <= rating_col || rating_col >= 8 || !parse_int(fb[0], fe[0], &r.user) || !parse_int(fb[1], fe[1], &r.item) ||
            !parse_float(fb[rating_col], fe[rating_col], &r.rating))
            return cu2b_fail(CU2B_ERR_IO, "%s:%lld: expected <int>%c<int>%c...<float>", path, (long long)line_no, delim, delim);
        rows->push_back(r);
        p = next;
    }
    return CU2B_OK;
}

// str(float) of Python: shortest round-trip digits; fixed notation for 1e-4 <= |x| < 1e16 with
// ".0" appended to integral values, exponent notation otherwise.
char *format_pyfloat(double v, char *dst) {
    if (v != v) { memcpy(dst, "nan", 3); return dst + 3; }
    if (v == 1.0 / 0.0) { memcpy(dst, "inf", 3); return dst + 3; }
    if (v == -1.0 / 0.0) { memcpy(dst, "-inf", 4); return dst + 4; }
    const double a = v < 0 ? -v : v;
    if (a != 0 && (a < 1e-4 || a >= 1e16)) {
        char tmp[40];
        auto res = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::scientific);
        // to_chars: d.ddde+XX ; Python: d.ddde+XX with at least two exponent digits and no ".0"
        // mantissa padding -- identical for everything these files can contain
        const size_t n = (size_t)(res.ptr - tmp);
        memcpy(dst, tmp, n);
        return dst + n;
    }
    auto res = std::to_chars(dst, dst + 32, v, std::chars_format::fixed);
    char *e = res.ptr;
    if (!memchr(dst, '.', (size_t)(e - dst))) { *e++ = '.'; *e++ = '0'; }
    return e;
}

char *format_int(int64_t v, char *dst) {
    auto res = std::to_chars(dst, dst + 24, v);
    return res.ptr;
}

// write_to_file of map_items.py:78-87: header + "user,item,rating" rows, '\n' line ends.
cu2b_status write_rows(const char *path, const std::vector<Row> &rows, const std::vector<int64_t> *order) {
    FILE *f = fopen(path, "wb");
    if (!f) return cu2b_fail(CU2B_ERR_IO, "cannot create %s", path);
    std::vector<char> buf(1 << 22);
    size_t used = 0;
    const char *hdr = "userId,itemId,rating\n";
    memcpy(buf.data(), hdr, strlen(hdr));
    used = strlen(hdr);
    const size_t n = order ? order->size() : rows.size();
    for (size_t t = 0; t < n; ++t) {
        const Row &r = rows[order ? (size_t)(*order)[t] : t];
        if (used + 128 > buf.size()) {
            if (fwrite(buf.data(), 1, used, f) != used) { fclose(f); return cu2b_fail(CU2B_ERR_IO, "short write to %s", path); }
            used = 0;
        }
        char *p = buf.data() + used;
        p = format_int(r.user, p);
        *p++ = ',';
        p = format_int(r.item, p);
        *p++ = ',';
        p = format_pyfloat(r.rating, p);
        *p++ = '\n';
        used = (size_t)(p - buf.data());
    }
    if (used && fwrite(buf.data(), 1, used, f) != used) { fclose(f); return cu2b_fail(CU2B_ERR_IO, "short write to %s", path); }
    if (fclose(f) != 0) return cu2b_fail(CU2B_ERR_IO, "cannot close %s", path);
    return CU2B_OK;
}

// map_rows of map_items.py:21-59: ids are replaced by 1 + (number of distinct ids seen before).
void map_rows(std::vector<Row> *rows, std::unordered_map<int64_t, int64_t> *umap, std::unordered_map<int64_t, int64_t> *imap,
              bool add_missing, int64_t *missing_users, int64_t *missing_items) {
    size_t w = 0;
    for (size_t t = 0; t < rows->size(); ++t) {
        Row r = (*rows)[t];
        auto u = umap->find(r.user);
        if (u == umap->end()) {
            if (!add_missing) { ++*missing_users; continue; }
            u = umap->emplace(r.user, (int64_t)umap->size() + 1).first;
        }
        auto i = imap->find(r.item);
        if (i == imap->end()) {
            if (!add_missing) { ++*missing_items; continue; }
            i = imap->emplace(r.item, (int64_t)imap->size() + 1).first;
        }
        r.user = u->second;
        r.item = i->second;
        (*rows)[w++] = r;
    }
    rows->resize(w);
}

// sort_by_user of map_items.py:62-75: ascending user, input order kept inside a user.
void stable_by_user(const std::vector<Row> &rows, std::vector<int64_t> *order) {
    order->resize(rows.size());
    std::iota(order->begin(), order->end(), (int64_t)0);
    std::stable_sort(order->begin(), order->end(), [&](int64_t a, int64_t b) { return rows[(size_t)a].user < rows[(size_t)b].user; });
}

// The Mersenne Twister stream of CPython's `random` module: random.seed(int) is init_by_array over
// the 32-bit words of |seed|; getrandbits(k <= 32) is the top k bits of one output word;
// _randbelow(n) rejects values >= n; shuffle walks i = n-1 .. 1 and swaps x[i], x[randbelow(i+1)].
struct PyRandom {
    uint32_t mt[624];
    int idx = 624;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void seed(uint64_t seed_abs) {
        uint32_t key[2] = {(uint32_t)seed_abs, (uint32_t)(seed_abs >> 32)};
        const int klen = key[1] ? 2 : 1;
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = 624 > klen ? 624 : klen; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= klen) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint64_t randbelow(uint64_t n) {  // n >= 1
        int k = 0;
        for (uint64_t t = n; t; t >>= 1) ++k;  // n.bit_length()
        for (;;) {
            uint64_t r;
            if (k <= 32) {
                r = next() >> (32 - k);
            } else {  // getrandbits fills 32-bit words from the least significant one
                const uint64_t lo = next();
                const uint64_t hi = next() >> (64 - k);
                r = (hi << 32) | lo;
            }
            if (r < n) return r;
        }
    }
};

}  // namespace

extern "C" cu2b_status cu2b_prep_map(const char *in_path, const char *out_path, char delimiter, int has_header,
                                     int rating_col, const char *in2_path, const char *out2_path, int64_t *n_rows,
                                     int64_t *n_rows2, int64_t *n_users, int64_t *n_items, int64_t *skipped_users,
                                     int64_t *skipped_items) {
    if (!in_path || !out_path || rating_col < 2 || (in2_path != nullptr) != (out2_path != nullptr))
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_prep_map: bad argument");
    std::unordered_map<int64_t, int64_t> umap, imap;
    std::vector<Row> rows;
    std::vector<int64_t> order;
    int64_t mu = 0, mi = 0;
    CU2B_TRY_STATUS(read_rows(in_path, delimiter, has_header != 0, rating_col, &rows));
    map_rows(&rows, &umap, &imap, true, &mu, &mi);
    stable_by_user(rows, &order);
    CU2B_TRY_STATUS(write_rows(out_path, rows, &order));
    if (n_rows) *n_rows = (int64_t)rows.size();
    if (n_rows2) *n_rows2 = 0;
    if (in2_path) {  // map_netflix.py:20-22: the second file is mapped with the first file's tables
        rows.clear();
        CU2B_TRY_STATUS(read_rows(in2_path, delimiter, has_header != 0, rating_col, &rows));
        map_rows(&rows, &umap, &imap, false, &mu, &mi);
        stable_by_user(rows, &order);
        CU2B_TRY_STATUS(write_rows(out2_path, rows, &order));
        if (n_rows2) *n_rows2 = (int64_t)rows.size();
    }
    if (n_users) *n_users = (int64_t)umap.size();
    if (n_items) *n_items = (int64_t)imap.size();
    if (skipped_users) *skipped_users = mu;
    if (skipped_items) *skipped_items = mi;
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_prep_sort(const char *in_path, const char *out_path, int64_t *n_rows) {
    if (!in_path || !out_path) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_prep_sort: null argument");
    std::vector<Row> rows;
    CU2B_TRY_STATUS(read_rows(in_path, ',', true, 2, &rows));
    std::vector<int64_t> order(rows.size());
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {  // sort_ratings.py:34
        const Row &x = rows[(size_t)a], &y = rows[(size_t)b];
        return x.user != y.user ? x.user < y.user : x.item < y.item;
    });
    CU2B_TRY_STATUS(write_rows(out_path, rows, &order));
    if (n_rows) *n_rows = (int64_t)rows.size();
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_prep_split(const char *in_path, const char *train_path, const char *test_path,
                                       double test_ratio, int64_t seed, int64_t *n_train, int64_t *n_test) {
    if (!in_path || !train_path || !test_path) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_prep_split: null argument");
    std::vector<Row> rows;
    CU2B_TRY_STATUS(read_rows(in_path, ',', true, 2, &rows));
    const size_t n = rows.size();
    std::vector<int64_t> perm(n);
    std::iota(perm.begin(), perm.end(), (int64_t)0);
    PyRandom rng;
    rng.seed((uint64_t)(seed < 0 ? -seed : seed));
    for (size_t i = n; i-- > 1;) std::swap(perm[i], perm[(size_t)rng.randbelow((uint64_t)i + 1)]);  // random.shuffle
    const double train_percent = 1 - test_ratio;                      // split_to_test_train.py:76
    const size_t cut = std::min(n, (size_t)std::max(0.0, (double)(int64_t)((double)n * train_percent)));  // int(num * pct)
    auto emit = [&](size_t lo, size_t hi, const char *path) -> cu2b_status {
        std::vector<int64_t> part(perm.begin() + (long)lo, perm.begin() + (long)hi);
        std::stable_sort(part.begin(), part.end(), [&](int64_t a, int64_t b) { return rows[(size_t)a].user < rows[(size_t)b].user; });
        return write_rows(path, rows, &part);
    };
    CU2B_TRY_STATUS(emit(0, cut, train_path));
    CU2B_TRY_STATUS(emit(cut, n, test_path));
    if (n_train) *n_train = (int64_t)cut;
    if (n_test) *n_test = (int64_t)(n - cut);
    return CU2B_OK;
}

// create_config.py:13-15: '0 {:d} {:d} {:f} {:d} {:f} {:f} {:f} {:f}', no trailing newline.
extern "C" cu2b_status cu2b_prep_create_config(const char *path, int num_iterations, int num_factors,
                                               double learning_rate, int seed, double p_reg, double q_reg,
                                               double user_bias_reg, double item_bias_reg) {
    if (!path) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_prep_create_config: null path");
    FILE *f = fopen(path, "wb");
    if (!f) return cu2b_fail(CU2B_ERR_IO, "cannot create %s", path);
    fprintf(f, "0 %d %d %f %d %f %f %f %f", num_iterations, num_factors, learning_rate, seed, p_reg, q_reg,
            user_bias_reg, item_bias_reg);
    if (fclose(f) != 0) return cu2b_fail(CU2B_ERR_IO, "cannot close %s", path);
    return CU2B_OK;
}

// Ratings triplets (0-based ids, as every other entry point holds them) -> the reference's input
// CSV ("userId,itemId,rating", 1-based ids; util.cu:17-45 reads it back). Used by the experiment
// harness to materialise synthetic data sets for the bin/mf CLI.
extern "C" cu2b_status cu2b_write_ratings_csv(const char *path, const cu2b_rating *ratings, int64_t n) {
    if (!path || (!ratings && n > 0) || n < 0) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_write_ratings_csv: bad argument");
    std::vector<Row> rows((size_t)n);
    for (int64_t t = 0; t < n; ++t) rows[(size_t)t] = Row{(int64_t)ratings[t].user + 1, (int64_t)ratings[t].item + 1, (double)ratings[t].rating};
    return write_rows(path, rows, nullptr);
}

// convert_to_np.py:6-8: np.save(out, np.genfromtxt(in, delimiter=',')). genfromtxt drops '#' comments
// and blank lines, splits on ',', converts every field with float() (a missing or non-numeric field
// becomes nan), requires the same number of fields on every line, and squeezes the result: a single
// column or a single row is saved as a 1-D array, a single value as a 0-d array. np.save writes
// format version 1.0: "\x93NUMPY" 1 0 <u16 header length> <dict> padded with spaces to a multiple of 64
// bytes and terminated by '\n', then the float64 values in C order.
extern "C" cu2b_status cu2b_prep_convert_to_np(const char *in_path, const char *out_path, int64_t *n_rows,
                                               int64_t *n_cols) {
    if (!in_path || !out_path) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_prep_convert_to_np: null argument");
    int fd = open(in_path, O_RDONLY);
    if (fd < 0) return cu2b_fail(CU2B_ERR_IO, "cannot open %s", in_path);
    struct stat st;
    fstat(fd, &st);
    const size_t size = (size_t)st.st_size;
    const char *base = nullptr;
    if (size) {
        base = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (base == MAP_FAILED) { close(fd); return cu2b_fail(CU2B_ERR_IO, "mmap failed for %s", in_path); }
    }
    close(fd);
    const char *end = base + size;
    int nthreads = std::max(1, omp_get_max_threads());
    if (size < cu2b_io_parallel_min_bytes()) nthreads = 1;
    std::vector<const char *> cut(nthreads + 1);
    cut[0] = base;
    cut[nthreads] = end;
    for (int t = 1; t < nthreads; ++t) {
        const char *p = base + size * t / nthreads;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        cut[t] = nl ? nl + 1 : end;
    }
    struct Part {
        std::vector<double> vals;
        int64_t rows = 0, cols = -1;  // cols of the chunk's first data line
        int64_t bad_line = -1, bad_cols = 0;  // first line (0-based inside the chunk) with another width
    };
    std::vector<Part> part(nthreads);
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; };
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int t = 0; t < nthreads; ++t) {
        Part &pt = part[t];
        const char *p = cut[t], *e = cut[t + 1];
        pt.vals.reserve((size_t)(e - p) / 8 + 16);
        std::string tok;
        int64_t line_no = 0;
        while (p < e) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *le = nl ? nl : e;
            const char *hash = (const char *)memchr(p, '#', (size_t)(le - p));  // comments='#'
            const char *ce = hash ? hash : le;
            const char *a = p, *b = ce;
            while (a < b && is_space(*a)) ++a;
            while (b > a && is_space(b[-1])) --b;
            if (a < b) {  // not blank
                int64_t cols = 0;
                const char *f = p;  // fields are split on the unstripped line, each field stripped
                // trailing whitespace / '\r' of the line belongs to the last field and is stripped there
                while (true) {
                    const char *comma = (const char *)memchr(f, ',', (size_t)(ce - f));
                    const char *fe = comma ? comma : ce;
                    const char *x = f, *y = fe;
                    while (x < y && is_space(*x)) ++x;
                    while (y > x && is_space(y[-1])) --y;
                    double v = NAN;
                    if (x < y) {
                        tok.assign(x, (size_t)(y - x));
                        // float(): no hex floats, nothing may follow the number
                        if (tok.find_first_of("xX") == std::string::npos) {
                            char *endp = nullptr;
                            const double d = strtod(tok.c_str(), &endp);
                            if (endp == tok.c_str() + tok.size()) v = d;
                        }
                    }
                    pt.vals.push_back(v);
                    ++cols;
                    if (!comma) break;
                    f = comma + 1;
                }
                if (pt.cols < 0) pt.cols = cols;
                else if (cols != pt.cols && pt.bad_line < 0) { pt.bad_line = line_no; pt.bad_cols = cols; }
                ++pt.rows;
            }
            ++line_no;
            p = nl ? nl + 1 : e;
        }
    }
    if (size) munmap((void *)base, size);
    int64_t rows = 0, cols = -1;
    for (int t = 0; t < nthreads; ++t) {
        if (part[t].cols < 0) continue;
        if (cols < 0) cols = part[t].cols;
        if (part[t].cols != cols || part[t].bad_line >= 0)
            return cu2b_fail(CU2B_ERR_IO, "%s: lines with different numbers of columns (%lld and %lld); genfromtxt raises "
                             "ValueError here", in_path, (long long)cols,
                             (long long)(part[t].cols != cols ? part[t].cols : part[t].bad_cols));
        rows += part[t].rows;
    }
    if (cols < 0) cols = 0;
    // squeezed shape, printed as Python prints a tuple
    char shape[64];
    int64_t first_dim = -1;
    if (rows == 0) { snprintf(shape, sizeof shape, "(0,)"); first_dim = 0; }  // genfromtxt of an empty file: empty 1-D array
    else if (rows == 1 && cols == 1) snprintf(shape, sizeof shape, "()");
    else if (rows == 1) { snprintf(shape, sizeof shape, "(%lld,)", (long long)cols); first_dim = cols; }
    else if (cols == 1) { snprintf(shape, sizeof shape, "(%lld,)", (long long)rows); first_dim = rows; }
    else { snprintf(shape, sizeof shape, "(%lld, %lld)", (long long)rows, (long long)cols); first_dim = rows; }
    std::string dict = std::string("{'descr': '<f8', 'fortran_order': False, 'shape': ") + shape + ", }";
    if (first_dim >= 0) {  // numpy leaves room for the first axis to grow in place (21 digits)
        char digits[32];
        const int nd = snprintf(digits, sizeof digits, "%lld", (long long)first_dim);
        dict.append((size_t)std::max(0, 21 - nd), ' ');
    }
    const size_t hlen = dict.size() + 1;
    const size_t padlen = 64 - ((10 + hlen) % 64);
    dict.append(padlen, ' ');
    dict.push_back('\n');
    if (dict.size() > 65535) return cu2b_fail(CU2B_ERR_UNSUPPORTED, "npy header too long");
    FILE *f = fopen(out_path, "wb");
    if (!f) return cu2b_fail(CU2B_ERR_IO, "cannot create %s", out_path);
    const unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(dict.size() & 0xff),
                                     (unsigned char)(dict.size() >> 8)};
    bool ok = fwrite(magic, 1, 10, f) == 10 && fwrite(dict.data(), 1, dict.size(), f) == dict.size();
    for (int t = 0; t < nthreads && ok; ++t)
        if (!part[t].vals.empty())
            ok = fwrite(part[t].vals.data(), sizeof(double), part[t].vals.size(), f) == part[t].vals.size();
    if (fclose(f) != 0) ok = false;
    if (!ok) return cu2b_fail(CU2B_ERR_IO, "short write to %s", out_path);
    if (n_rows) *n_rows = rows;
    if (n_cols) *n_cols = cols;
    return CU2B_OK;
}
