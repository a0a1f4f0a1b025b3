// CsvFileStream (execution/file_stream.rs:10-335) as a batch producer for the device pipeline — SURVEY.md 8(f) rank 3.
//
// The reference reads one line at a time into a String, splits it into a Vec<&str>, parses every field into a ParsedValue enum,
// pushes those into per-column Vecs and finally rebuilds typed arrays from them.  Here the file is read in 4 MiB blocks and every
// field is parsed straight into the column's Arrow buffers (8-byte values, LSB-first bitmaps, int32 offsets + bytes), which are
// reused from batch to batch and handed to RecordBatch::try_new / rvl_stream_push as they are.  Two stages: the caller's thread
// cuts the file into batches of `batch_size` data lines (memchr for the newlines, blank lines dropped, line numbers kept) and
// stays a few batches ahead; worker threads parse whole batches concurrently; batches — and errors — come out in file order, so
// a consumer that stops early (LIMIT) never sees an error from a batch the reference would not have read.  Files under 8 MiB
// are parsed by the caller itself.  Same observable behaviour:
// header line always skipped (:134-151), blank lines skipped and not counted (:168-170), fields trimmed (:43), "" / "null" = NULL,
// Rust's i64 / f64 / bool text rules (:60-110), the same error text with the same 1-based line numbers, batches of `batch_size`
// data lines (default: 8 MiB worth of estimated row bytes, clamped to 1 000..100 000, :346-369).
//
// Validity of Int64 / Float64 columns: the reference passes its `nulls` vector (true = NULL field) where PrimitiveArray::new
// expects a validity vector (:233-239, :265-271 vs primitive.rs:31-33), so a column holding a null comes out inverted.  Default
// here is the evident intent (null fields are null); set_csv_reference_validity(true) reproduces the reference bit for bit.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "rivulus.hpp"

namespace rivulus {

static bool g_csv_reference_validity = false;
void set_csv_reference_validity(bool on) { g_csv_reference_validity = on; }
bool csv_reference_validity() { return g_csv_reference_validity; }

size_t calculate_adaptive_batch_size(const Schema& schema) {  // file_stream.rs:346-369
    size_t row = 0;
    for (const auto& f : schema.fields) switch (f.data_type) {
        case ExecType::Int64: case ExecType::Float64: row += 8; break;
        case ExecType::Boolean: row += 1; break;
        case ExecType::String: row += 32; break;
        case ExecType::Null: break;
    }
    if (row == 0) return 10000;
    const size_t t = (8u * 1024 * 1024) / row;
    return std::min<size_t>(std::max<size_t>(t, 1000), 100000);
}

namespace {
using sv = std::string_view;

inline bool is_ws(uint32_t c) {  // char::is_whitespace
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
           c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
inline uint32_t decode(const char* p, size_t* len) {  // p points at the lead byte of a valid UTF-8 sequence
    const unsigned char b = (unsigned char)p[0];
    if (b < 0x80) { *len = 1; return b; }
    if (b < 0xE0) { *len = 2; return ((b & 0x1Fu) << 6) | ((unsigned char)p[1] & 0x3Fu); }
    if (b < 0xF0) { *len = 3; return ((b & 0x0Fu) << 12) | (((unsigned char)p[1] & 0x3Fu) << 6) | ((unsigned char)p[2] & 0x3Fu); }
    *len = 4;
    return ((b & 0x07u) << 18) | (((unsigned char)p[1] & 0x3Fu) << 12) | (((unsigned char)p[2] & 0x3Fu) << 6) | ((unsigned char)p[3] & 0x3Fu);
}
sv trim(sv s) {  // str::trim; ASCII fast path, multi-byte white space (NBSP, U+2000.., U+3000 ...) decoded
    const char* b = s.data();
    const char* e = b + s.size();
    while (b < e) {
        const unsigned char c = (unsigned char)*b;
        if (c < 0x80) { if (is_ws(c)) { ++b; continue; } break; }
        size_t n; if (!is_ws(decode(b, &n))) break; b += n;
    }
    while (e > b) {
        const unsigned char c = (unsigned char)e[-1];
        if (c < 0x80) { if (is_ws(c)) { --e; continue; } break; }
        const char* p = e - 1;
        while (p > b && ((unsigned char)*p & 0xC0) == 0x80) --p;
        size_t n; if (!is_ws(decode(p, &n))) break; e = p;
    }
    return sv(b, (size_t)(e - b));
}
bool valid_utf8(const char* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        // eight ASCII bytes at a time
        if (i + 8 <= n) { uint64_t w; std::memcpy(&w, s + i, 8); if ((w & 0x8080808080808080ull) == 0) { i += 8; continue; } }
        const unsigned char b = (unsigned char)s[i];
        if (b < 0x80) { ++i; continue; }
        size_t len; uint32_t min;
        if (b >= 0xC2 && b <= 0xDF) { len = 2; min = 0x80; }
        else if (b >= 0xE0 && b <= 0xEF) { len = 3; min = 0x800; }
        else if (b >= 0xF0 && b <= 0xF4) { len = 4; min = 0x10000; }
        else return false;
        if (i + len > n) return false;
        for (size_t k = 1; k < len; ++k) if (((unsigned char)s[i + k] & 0xC0) != 0x80) return false;
        size_t dl; const uint32_t cp = decode(s + i, &dl);
        if (cp < min || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += len;
    }
    return true;
}
bool parse_i64(sv s, int64_t* out) {  // i64::from_str: [+-]? digit+, overflow is an error
    const char* b = s.data();
    const char* e = b + s.size();
    if (b == e) return false;
    const bool plus = *b == '+';
    if (plus) ++b;                                   // from_chars takes '-' but not '+'
    if (b == e || (plus && *b == '-')) return false;
    const char* d = *b == '-' ? b + 1 : b;
    if (d == e) return false;
    for (const char* q = d; q < e; ++q) if (*q < '0' || *q > '9') return false;
    const auto r = std::from_chars(b, e, *out, 10);
    return r.ec == std::errc() && r.ptr == e;
}
bool ieq(sv s, const char* lit) {
    const size_t n = std::strlen(lit);
    if (s.size() != n) return false;
    for (size_t k = 0; k < n; ++k) { char c = s[k]; if (c >= 'A' && c <= 'Z') c = (char)(c + 32); if (c != lit[k]) return false; }
    return true;
}
bool parse_f64(sv s, double* out) {  // f64::from_str (dec2flt): [+-]? ( inf | infinity | nan | digits[.digits][e[+-]digits] ), correctly rounded
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i == s.size()) return false;
    const sv body = s.substr(i);
    if (ieq(body, "inf") || ieq(body, "infinity")) { *out = neg ? -HUGE_VAL : HUGE_VAL; return true; }
    if (ieq(body, "nan")) { const uint64_t bits = 0x7ff8000000000000ull | (neg ? 0x8000000000000000ull : 0); std::memcpy(out, &bits, 8); return true; }
    size_t k = i, digits = 0;
    while (k < s.size() && s[k] >= '0' && s[k] <= '9') { ++k; ++digits; }
    if (k < s.size() && s[k] == '.') { ++k; while (k < s.size() && s[k] >= '0' && s[k] <= '9') { ++k; ++digits; } }
    if (digits == 0) return false;
    if (k < s.size() && (s[k] == 'e' || s[k] == 'E')) {
        ++k;
        if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
        size_t ed = 0;
        while (k < s.size() && s[k] >= '0' && s[k] <= '9') { ++k; ++ed; }
        if (ed == 0) return false;
    }
    if (k != s.size()) return false;
    double v = 0.0;
    const auto r = std::from_chars(body.data(), body.data() + body.size(), v, std::chars_format::general);
    if (r.ec == std::errc::result_out_of_range) {  // from_chars leaves the value alone; strtod says which way (inf or 0 / subnormal)
        const std::string z(body);
        v = std::strtod(z.c_str(), nullptr);
    }
    *out = neg ? -v : v;
    return true;
}
int parse_bool(sv s) {  // to_lowercase() then "true"|"t"|"1" / "false"|"f"|"0" (:96-98); -1 = error
    if (ieq(s, "true") || ieq(s, "t") || s == "1") return 1;
    if (ieq(s, "false") || ieq(s, "f") || s == "0") return 0;
    return -1;
}

struct alignas(128) ColumnBuf {  // one column of the batch being parsed, Arrow layout, reused across batches (own cache lines: see CsvJob)
    ExecType type = ExecType::Null;
    std::vector<int64_t> i64; std::vector<double> f64; std::vector<uint8_t> bits, validity, data; std::vector<int32_t> offsets;
    bool any_null = false;
    void begin(size_t cap) {
        any_null = false;
        const size_t nb = (cap + 7) / 8;
        validity.assign(nb, 0);
        switch (type) {
            case ExecType::Int64: i64.resize(cap); break;
            case ExecType::Float64: f64.resize(cap); break;
            case ExecType::Boolean: bits.assign(nb, 0); break;
            case ExecType::String: offsets.resize(cap + 1); offsets[0] = 0; data.clear(); break;
            case ExecType::Null: break;
        }
    }
};
}  // namespace

// One batch on its way through the reader: the text of its data lines (stage 1, the caller's thread) and the parsed columns
// (stage 2, a worker thread or the caller itself).
// (alignas + locals in the parse loop: two jobs that share a cache line — one worker's per-row progress store next to the vector
// headers another worker reads on every row — doubled the parse time of concurrent batches)
struct alignas(128) CsvJob {
    std::vector<char> text;            // the file bytes the batch was cut from (whole buffer spans)
    std::vector<uint32_t> starts, lens;   // offset and length in `text` of every data line, terminator excluded
    std::vector<uint64_t> line_no;     // 1-based number of every line in the file (header = 1, blank lines count)
    std::string read_error;            // set when reading the file failed after these lines: surfaces instead of the batch
    std::vector<ColumnBuf> cols;
    size_t rows = 0;
    std::string parse_error;           // first bad line in file order
    bool done = false;
};

// Parsed-batch buffers outlive their reader: a fresh multi-megabyte buffer costs a page fault per 4 KB on first touch (about as much
// as parsing into it, measured), so finished readers park their jobs here and the next reader starts with warm ones.
std::mutex g_job_pool_mu;
std::vector<std::unique_ptr<CsvJob>> g_job_pool;
constexpr size_t kJobPoolMax = 24;
std::unique_ptr<CsvJob> pooled_job() {
    std::lock_guard<std::mutex> g(g_job_pool_mu);
    if (g_job_pool.empty()) return std::make_unique<CsvJob>();
    auto j = std::move(g_job_pool.back());
    g_job_pool.pop_back();
    return j;
}
void park_job(std::unique_ptr<CsvJob> j) {
    if (!j) return;
    std::lock_guard<std::mutex> g(g_job_pool_mu);
    if (g_job_pool.size() < kJobPoolMax) g_job_pool.push_back(std::move(j));
}

// What the parse workers read on every field.  It lives in its own allocation, written once: the reader's cursor (line counter,
// buffer positions) changes on every line in the caller's thread, and sharing a cache line with it made every worker ~4x slower.
struct alignas(128) CsvParseConfig {
    std::string delim = ",";
    std::vector<ExecType> types;
};

struct CsvBatchReader::Impl {
    int fd = -1;
    std::shared_ptr<const CsvParseConfig> cfg;
    SchemaRef schema;
    size_t batch_size = 0, current_line = 0;
    bool finished = false, eof = false, failed = false;
    std::vector<char> buf; size_t pos = 0, end = 0;   // unread bytes are buf[pos, end)
    // pipeline
    std::vector<std::thread> workers;
    std::mutex mu; std::condition_variable work_cv, done_cv;
    std::deque<CsvJob*> todo;                         // jobs waiting for a worker
    std::deque<std::unique_ptr<CsvJob>> order;        // jobs in file order, head = next batch the caller gets
    std::vector<std::unique_ptr<CsvJob>> spare;       // finished jobs whose buffers are reused
    std::unique_ptr<CsvJob> current;                  // the batch last returned by read_batch
    bool stop = false;
    size_t depth = 1;
    size_t last_text_bytes = 0;

    ~Impl() {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        work_cv.notify_all();
        for (auto& t : workers) t.join();
        if (fd >= 0) ::close(fd);
        park_job(std::move(current));
        for (auto& j : order) park_job(std::move(j));
        for (auto& j : spare) park_job(std::move(j));
    }

    // a complete line inside the buffer, including its '\n' (BufRead::read_line), or the unterminated rest of the file once EOF is seen
    bool find_line(const char** p, size_t* n) {
        if (pos < end) {
            const char* nl = (const char*)std::memchr(buf.data() + pos, '\n', end - pos);
            if (nl) { *p = buf.data() + pos; *n = (size_t)(nl - (buf.data() + pos)) + 1; pos += *n; return true; }
            if (eof) { *p = buf.data() + pos; *n = end - pos; pos = end; return true; }   // last line without a newline
        }
        return false;
    }
    // more bytes behind the partial line; false when the file is exhausted.  Invalidates pointers into the buffer.
    // Throws std::string on an I/O error.
    bool refill() {
        if (eof) return false;
        if (pos > 0) { std::memmove(buf.data(), buf.data() + pos, end - pos); end -= pos; pos = 0; }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        for (;;) {
            const ssize_t got = ::read(fd, buf.data() + end, buf.size() - end);
            if (got < 0) {
                if (errno == EINTR) continue;
                const int e = errno;
                throw std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")";
            }
            if (got == 0) eof = true; else end += (size_t)got;
            return true;
        }
    }
    bool next_line(const char** p, size_t* n) {
        for (;;) {
            if (find_line(p, n)) return true;
            if (!refill()) return false;
        }
    }

    // `line.trim().is_empty()` (:168) decided on raw bytes; a line that is not valid UTF-8 is never blank (stage 2 reports it)
    static bool is_blank(const char* p, size_t n) {
        for (size_t i = 0; i < n; ++i) {
            const unsigned char c = (unsigned char)p[i];
            if (c < 0x80) { if (!is_ws(c)) return false; continue; }
            return valid_utf8(p, n) && trim(sv(p, n)).empty();
        }
        return true;
    }

    // stage 1: the text of the next `batch_size` data lines.  False when the file is exhausted and the job holds nothing.
    bool fill(CsvJob& j) {
        j.text.clear(); j.starts.clear(); j.lens.clear(); j.line_no.clear(); j.read_error.clear(); j.parse_error.clear(); j.rows = 0; j.done = false;
        if (finished) return false;
        // size a fresh job like the last one: growing a multi-megabyte vector step by step is a chain of mremap calls, each a TLB
        // shootdown across every thread of the process
        if (j.text.capacity() == 0 && last_text_bytes > 0) { j.text.reserve(last_text_bytes + last_text_bytes / 8); j.starts.reserve(batch_size); j.lens.reserve(batch_size); j.line_no.reserve(batch_size); }
        // Lines are recorded as (offset, length) into `text`, which receives the buffer in whole spans — one copy per buffer refill
        // instead of one per line; blank lines and line terminators simply are not referenced.
        try {
            size_t span_lo = pos;                              // first buffer byte not yet appended to `text`
            auto flush = [&] { j.text.insert(j.text.end(), buf.data() + span_lo, buf.data() + pos); span_lo = pos; };
            const char* p; size_t n;
            while (j.line_no.size() < batch_size) {
                if (!find_line(&p, &n)) {
                    flush();
                    if (!refill()) { finished = true; break; }
                    span_lo = pos;
                    continue;
                }
                ++current_line;
                if (n > 0 && p[n - 1] == '\n') { --n; if (n > 0 && p[n - 1] == '\r') --n; }
                if (is_blank(p, n)) continue;
                j.starts.push_back((uint32_t)(j.text.size() + (size_t)(p - (buf.data() + span_lo))));
                j.lens.push_back((uint32_t)n);
                j.line_no.push_back(current_line);
                if (j.text.size() + (size_t)(buf.data() + pos - (buf.data() + span_lo)) > (size_t)UINT32_MAX - (64u << 20)) break;   // 4 GiB of text: close the batch
            }
            flush();
        } catch (const std::string& io) {
            j.read_error = "Stream execution error: Failed to read line " + std::to_string(current_line + 1) + ": " + io;
            finished = true;
        }
        last_text_bytes = std::max(last_text_bytes, j.text.size());
        return !j.line_no.empty() || !j.read_error.empty();
    }

    // stage 2: every line of the job into the job's columns; stops at the first bad line
    static void parse(const CsvParseConfig& cfg, CsvJob& j) {
        const size_t nf = cfg.types.size();
        j.cols.resize(nf);   // a pooled job keeps whatever buffers its columns already own
        for (size_t i = 0; i < nf; ++i) j.cols[i].type = cfg.types[i];
        const size_t nlines = j.line_no.size();
        static const bool trace = std::getenv("RVL_CSV_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        for (auto& c : j.cols) { c.begin(nlines); if (c.type == ExecType::String && c.data.capacity() == 0) c.data.reserve(j.text.size() / 2); }
        const auto t1 = std::chrono::steady_clock::now();
        struct Tr { bool on; std::chrono::steady_clock::time_point a, b; size_t n; ~Tr() { if (on) std::fprintf(stderr, "[csv] job of %zu lines: buffers %.2f ms, parse %.2f ms\n", n,
            std::chrono::duration<double, std::milli>(b - a).count(), std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b).count()); } } tr{trace, t0, t1, nlines};
        const char* const text = j.text.data();
        const uint32_t* const starts = j.starts.data();
        const uint32_t* const lens = j.lens.data();
        size_t rows = 0;
        for (size_t r = 0; r < nlines; ++r) {
            const char* p = text + starts[r];
            const size_t n = lens[r];
            if (!valid_utf8(p, n)) {   // read_line fails before the line is counted: "line {current_line + 1}" is this line's number
                j.parse_error = "Stream execution error: Failed to read line " + std::to_string(j.line_no[r]) + ": stream did not contain valid UTF-8";
                break;
            }
            if (!parse_line(cfg, j, r, sv(p, n))) break;
            rows = r + 1;
        }
        j.rows = rows;
    }

    static bool parse_line(const CsvParseConfig& cfg, CsvJob& j, size_t r, sv line) {  // :42-121, writing row r of every column
        const std::string& delim = cfg.delim;
        const size_t nf = j.cols.size();
        size_t start = 0;
        // the field-count check precedes any parsing (:45-52)
        size_t count = 1;
        if (delim.size() == 1) { for (const char c : line) count += c == delim[0]; }
        else for (size_t q = line.find(delim); q != sv::npos; q = line.find(delim, q + delim.size())) ++count;
        if (count != nf) {
            j.parse_error = "Stream execution error: Parse error: Line " + std::to_string(j.line_no[r]) + ": Expected " + std::to_string(nf) + " fields, found " +
                            std::to_string(count);
            return false;
        }
        for (size_t field = 0; field < nf; ++field) {
            size_t q = delim.size() == 1 ? line.find(delim[0], start) : line.find(delim, start);
            if (q == sv::npos) q = line.size();
            const sv s = trim(line.substr(start, q - start));
            start = q + delim.size();
            ColumnBuf& c = j.cols[field];
            const bool null = s.empty() || s == "null";
            const char* bad = nullptr;
            bool valid = !null;
            switch (c.type) {
                case ExecType::Int64:
                    if (null) c.i64[r] = 0; else if (!parse_i64(s, &c.i64[r])) bad = "Int64";
                    break;
                case ExecType::Float64:
                    if (null) c.f64[r] = 0.0; else if (!parse_f64(s, &c.f64[r])) bad = "Float64";
                    break;
                case ExecType::String:
                    if (!null) {
                        if (c.data.size() + s.size() > (size_t)INT32_MAX) { j.parse_error = "Stream execution error: string column exceeds the 2 GiB offset range in one batch"; return false; }
                        c.data.insert(c.data.end(), s.begin(), s.end());
                    }
                    c.offsets[r + 1] = (int32_t)c.data.size();
                    break;
                case ExecType::Boolean:
                    if (!null) { const int b = parse_bool(s); if (b < 0) bad = "Boolean"; else if (b) c.bits[r >> 3] |= (uint8_t)(1u << (r & 7)); }
                    break;
                case ExecType::Null: valid = false; break;
            }
            if (bad) {
                j.parse_error = "Stream execution error: Parse error: Line " + std::to_string(j.line_no[r]) + ", field " + std::to_string(field) + ": Cannot parse '" +
                                std::string(s) + "' as " + bad;
                return false;
            }
            if (valid) c.validity[r >> 3] |= (uint8_t)(1u << (r & 7)); else c.any_null = true;
        }
        return true;
    }

    void worker_loop() {
        const std::shared_ptr<const CsvParseConfig> wcfg = cfg;   // the worker's own handle: nothing of *this is read per field
        for (;;) {
            CsvJob* j = nullptr;
            {
                std::unique_lock<std::mutex> g(mu);
                work_cv.wait(g, [&] { return stop || !todo.empty(); });
                if (stop) return;
                j = todo.front(); todo.pop_front();
            }
            if (j->read_error.empty()) parse(*wcfg, *j);
            { std::lock_guard<std::mutex> g(mu); j->done = true; }
            done_cv.notify_all();
        }
    }
};

static int g_csv_threads = -1;
void set_csv_threads(int n) { g_csv_threads = n; }

CsvBatchReader::CsvBatchReader(const std::string& path, SchemaRef schema, std::optional<size_t> batch_size, std::optional<std::string> delimiter)
    : impl_(std::make_unique<Impl>()) {  // CsvFileStream::new :20-40
    impl_->fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (impl_->fd < 0) { const int e = errno; throw Error(std::string("Failed to open file: ") + std::strerror(e) + " (os error " + std::to_string(e) + ")"); }
#ifdef POSIX_FADV_SEQUENTIAL
    ::posix_fadvise(impl_->fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    impl_->schema = std::move(schema);
    impl_->batch_size = batch_size ? *batch_size : calculate_adaptive_batch_size(*impl_->schema);
    {
        auto cfg = std::make_shared<CsvParseConfig>();
        if (delimiter && !delimiter->empty()) cfg->delim = *delimiter;
        for (const auto& f : impl_->schema->fields) cfg->types.push_back(f.data_type);
        impl_->cfg = std::move(cfg);
    }
    impl_->buf.resize(4u << 20);
    // parse workers: only for files worth it (several batches of text); small files are parsed by the caller
    struct stat st;
    const bool big = ::fstat(impl_->fd, &st) == 0 && (!S_ISREG(st.st_mode) || st.st_size >= (8 << 20));
    int threads = g_csv_threads;
    if (threads < 0) threads = big ? (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 0;
    if (threads > 0 && impl_->batch_size > 0) {
        impl_->depth = (size_t)threads + 2;
        for (int i = 0; i < threads; ++i) impl_->workers.emplace_back([m = impl_.get()] { m->worker_loop(); });
    }
}
CsvBatchReader::~CsvBatchReader() = default;
const SchemaRef& CsvBatchReader::schema() const { return impl_->schema; }
size_t CsvBatchReader::batch_size() const { return impl_->batch_size; }

size_t CsvBatchReader::read_batch() {  // read_batch :123-199
    Impl& m = *impl_;
    if (m.current) { m.spare.push_back(std::move(m.current)); }
    if (m.failed) return 0;
    if (m.current_line == 0 && !m.finished) {  // the first line is the header, whatever it holds (:134-151)
        const char* p; size_t n;
        bool got = false;
        try { got = m.next_line(&p, &n); }
        catch (const std::string& io) { m.failed = true; throw Error("Stream execution error: Failed to read header: " + io); }
        if (!got) { m.finished = true; return 0; }
        if (!valid_utf8(p, n)) { m.failed = true; throw Error("Stream execution error: Failed to read header: stream did not contain valid UTF-8"); }
        ++m.current_line;
    }
    // keep `depth` batches of text in the pipeline (stage 1 runs here, ahead of the workers)
    static const bool trace = std::getenv("RVL_CSV_TRACE") != nullptr;
    const auto tt0 = std::chrono::steady_clock::now();
    while (m.order.size() < m.depth && !m.finished) {
        std::unique_ptr<CsvJob> j;
        if (!m.spare.empty()) { j = std::move(m.spare.back()); m.spare.pop_back(); } else j = pooled_job();
        if (!m.fill(*j)) { m.spare.push_back(std::move(j)); break; }
        CsvJob* raw = j.get();
        m.order.push_back(std::move(j));
        if (!m.workers.empty()) {
            { std::lock_guard<std::mutex> g(m.mu); m.todo.push_back(raw); }
            m.work_cv.notify_one();
        }
    }
    if (m.order.empty()) return 0;
    std::unique_ptr<CsvJob> j = std::move(m.order.front());
    m.order.pop_front();
    const auto tt1 = std::chrono::steady_clock::now();
    if (m.workers.empty()) { if (j->read_error.empty()) Impl::parse(*m.cfg, *j); }
    else { std::unique_lock<std::mutex> g(m.mu); m.done_cv.wait(g, [&] { return j->done; }); }
    if (trace) std::fprintf(stderr, "[csv] read_batch: fill %.2f ms, wait/parse %.2f ms\n", std::chrono::duration<double, std::milli>(tt1 - tt0).count(),
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tt1).count());
    // errors surface in file order: a bad line first (it precedes the failed read), then the read error
    if (!j->parse_error.empty()) { m.failed = true; const std::string e = j->parse_error; m.spare.push_back(std::move(j)); throw Error(e); }
    if (!j->read_error.empty()) { m.failed = true; const std::string e = j->read_error; m.spare.push_back(std::move(j)); throw Error(e); }
    m.current = std::move(j);
    return m.current->rows;
}

std::vector<rvl_column> CsvBatchReader::columns() const {  // build_record_batch :201-326, as host views of the parsed buffers
    const Impl& im = *impl_;
    if (!im.current) return {};
    const CsvJob& m = *im.current;
    std::vector<rvl_column> out(m.cols.size());
    for (size_t i = 0; i < m.cols.size(); ++i) {
        const ColumnBuf& c = m.cols[i];
        rvl_column& r = out[i];
        std::memset(&r, 0, sizeof r);
        r.dtype = (int32_t)c.type; r.location = RVL_HOST; r.length = (int64_t)m.rows; r.offset = 0;
        switch (c.type) {
            case ExecType::Int64: r.values = c.i64.data(); break;
            case ExecType::Float64: r.values = c.f64.data(); break;
            case ExecType::Boolean: r.values = c.bits.data(); break;
            case ExecType::String: r.offsets = c.offsets.data(); r.data = c.data.data(); r.data_len = (int64_t)c.data.size(); break;
            case ExecType::Null: break;
        }
        // bitmap iff the batch holds a null (:233-239 `nulls.iter().any`, string.rs:41-45, boolean.rs:37-41)
        if (c.any_null && c.type != ExecType::Null) r.validity = c.validity.data();
    }
    return out;
}

void CsvBatchReader::apply_reference_validity() {
    // PrimitiveArray::new(values, Some(nulls)) with nulls[i] = true for a NULL field: valid exactly where the field was null
    if (!impl_->current) return;
    CsvJob& m = *impl_->current;
    for (auto& c : m.cols) {
        if (!c.any_null || (c.type != ExecType::Int64 && c.type != ExecType::Float64)) continue;
        const size_t nb = (m.rows + 7) / 8;
        for (size_t b = 0; b < nb; ++b) c.validity[b] = (uint8_t)~c.validity[b];
        if (m.rows & 7) c.validity[nb - 1] &= (uint8_t)((1u << (m.rows & 7)) - 1);
    }
}

namespace {
struct CsvFileStream : DataStream {
    ContextRef ctx; CsvBatchReader reader;
    CsvFileStream(ContextRef c, const std::string& path, SchemaRef s, std::optional<size_t> b, std::optional<std::string> d)
        : ctx(std::move(c)), reader(path, std::move(s), b, std::move(d)) {}
    SchemaRef schema() const override { return reader.schema(); }
    std::optional<RecordBatch> next_batch() override {
        if (reader.read_batch() == 0) return std::nullopt;
        if (csv_reference_validity()) reader.apply_reference_validity();
        try { return RecordBatch::try_new(ctx, reader.schema(), reader.columns()); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream execution error: Failed to create RecordBatch: ") + e.what()); }
    }
};
}  // namespace

DataStreamRef make_csv_file_stream(const ContextRef& ctx, const std::string& path, SchemaRef schema, std::optional<size_t> batch_size,
                                   std::optional<std::string> delimiter) {
    return std::make_unique<CsvFileStream>(ctx ? ctx : Context::shared(0), path, std::move(schema), batch_size, std::move(delimiter));
}

}  // namespace rivulus
