// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's CSV source (execution/file_stream.rs), part of the oracle.
// Nothing under rivulus_b200/ may include, link or call this file.
//
// Follows /root/reference/src/execution/file_stream.rs line by line: one line at a time (read_line), split on the delimiter, trim,
// parse per schema type into ParsedValue, build one RecordBatch per `batch_size` non-blank lines.  Pinned by the reference's own four
// tests (file_stream.rs:372-461, ported in tests/test_oracle_golden.py) and the main.rs demo query (main.rs:233-256).  The text ->
// number rules are Rust's `str::parse::<i64>` / `::<f64>` (core::num, core::num::dec2flt), restated from their documented grammar:
// "parity unpinned — code reading only" for inputs the reference's tests do not hold.
//
// The reference builds Int64 / Float64 columns with `PrimitiveArray::new(values, Some(nulls))` where nulls[i] = true marks a NULL
// field (:213-240, :245-272) while `new` takes a VALIDITY vector (primitive.rs:31-33): whenever such a column holds a null, its
// validity comes out inverted (null fields valid with value 0, every parsed number null).  No reference test observes it
// (:432-445 stops at num_rows).  set_csv_reference_validity(true) reproduces that behaviour bit for bit; the default builds the
// validity the code evidently means ("and fix the validity inversion").
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rivulus_oracle.hpp"

namespace orc {

static bool g_csv_reference_validity = false;
void set_csv_reference_validity(bool on) { g_csv_reference_validity = on; }
bool csv_reference_validity() { return g_csv_reference_validity; }

size_t calculate_adaptive_batch_size(const Schema& schema) {  // file_stream.rs:346-369
    const size_t target = 8u * 1024 * 1024;
    size_t row = 0;
    for (const auto& f : schema.fields) switch (f.data_type) {
        case ExecType::Int64: case ExecType::Float64: row += 8; break;
        case ExecType::Boolean: row += 1; break;
        case ExecType::String: row += 32; break;
        case ExecType::Null: break;
    }
    if (row == 0) return 10000;
    const size_t t = target / row;
    return t < 1000 ? 1000 : (t > 100000 ? 100000 : t);
}

namespace {
// char::is_whitespace = Unicode White_Space
bool is_ws(uint32_t c) {
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
           c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
// decode the code point starting at s[i] (s is valid UTF-8); returns its byte length
size_t decode_at(const std::string& s, size_t i, uint32_t& cp) {
    const unsigned char b = (unsigned char)s[i];
    if (b < 0x80) { cp = b; return 1; }
    if (b < 0xE0) { cp = ((b & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu); return 2; }
    if (b < 0xF0) { cp = ((b & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) | ((unsigned char)s[i + 2] & 0x3Fu); return 3; }
    cp = ((b & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) | (((unsigned char)s[i + 2] & 0x3Fu) << 6) | ((unsigned char)s[i + 3] & 0x3Fu);
    return 4;
}
std::string trim(const std::string& s) {  // str::trim
    size_t b = 0, e = s.size();
    while (b < e) { uint32_t cp; const size_t n = decode_at(s, b, cp); if (!is_ws(cp)) break; b += n; }
    while (e > b) {
        size_t p = e - 1;
        while (p > b && ((unsigned char)s[p] & 0xC0) == 0x80) --p;
        uint32_t cp; decode_at(s, p, cp);
        if (!is_ws(cp)) break;
        e = p;
    }
    return s.substr(b, e - b);
}
bool valid_utf8(const std::string& s) {  // what read_line's from_utf8 check accepts
    size_t i = 0, n = s.size();
    while (i < n) {
        const unsigned char b = (unsigned char)s[i];
        size_t len; uint32_t min;
        if (b < 0x80) { ++i; continue; }
        else if (b >= 0xC2 && b <= 0xDF) { len = 2; min = 0x80; }
        else if (b >= 0xE0 && b <= 0xEF) { len = 3; min = 0x800; }
        else if (b >= 0xF0 && b <= 0xF4) { len = 4; min = 0x10000; }
        else return false;
        if (i + len > n) return false;
        for (size_t k = 1; k < len; ++k) if (((unsigned char)s[i + k] & 0xC0) != 0x80) return false;
        uint32_t cp; decode_at(s, i, cp);
        if (cp < min || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += len;
    }
    return true;
}
bool parse_i64(const std::string& s, int64_t& out) {  // i64::from_str (core::num::from_str_radix, radix 10)
    size_t i = 0;
    bool neg = false;
    if (s.empty()) return false;
    if (s[0] == '+' || s[0] == '-') { if (s.size() == 1) return false; neg = s[0] == '-'; i = 1; }
    __int128 v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > ((__int128)1 << 63)) return false;
    }
    if (neg) v = -v;
    if (v > (__int128)INT64_MAX || v < (__int128)INT64_MIN) return false;
    out = (int64_t)v;
    return true;
}
bool ieq(const std::string& s, size_t from, const char* lit) {
    size_t n = std::strlen(lit);
    if (s.size() - from != n) return false;
    for (size_t k = 0; k < n; ++k) { char c = s[from + k]; if (c >= 'A' && c <= 'Z') c = (char)(c + 32); if (c != lit[k]) return false; }
    return true;
}
bool parse_f64(const std::string& s, double& out) {  // f64::from_str (core::num::dec2flt): Sign? (inf|infinity|nan|Number), correctly rounded
    size_t i = 0;
    if (s.empty()) return false;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i == s.size()) return false;
    if (ieq(s, i, "inf") || ieq(s, i, "infinity")) { out = neg ? -HUGE_VAL : HUGE_VAL; return true; }
    if (ieq(s, i, "nan")) { uint64_t bits = 0x7ff8000000000000ull | (neg ? 0x8000000000000000ull : 0); std::memcpy(&out, &bits, 8); return true; }
    size_t digits = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') { ++i; ++digits; }
    if (i < s.size() && s[i] == '.') { ++i; while (i < s.size() && s[i] >= '0' && s[i] <= '9') { ++i; ++digits; } }
    if (digits == 0) return false;
    if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
        size_t ed = 0;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') { ++i; ++ed; }
        if (ed == 0) return false;
    }
    if (i != s.size()) return false;
    out = std::strtod(s.c_str(), nullptr);  // glibc: correctly rounded, overflow -> inf, underflow -> 0 / subnormal, like dec2flt
    return true;
}
std::string ascii_lower(std::string s) { for (auto& c : s) if (c >= 'A' && c <= 'Z') c = (char)(c + 32); return s; }

struct CsvFileStream : DataStream {  // file_stream.rs:10-335
    FILE* f = nullptr; SchemaRef schema_; size_t batch_size = 0, current_line = 0; bool finished = false; std::string delim = ",";
    ~CsvFileStream() override { if (f) std::fclose(f); }
    SchemaRef schema() const override { return schema_; }

    // BufRead::read_line: bytes up to and including '\n'; 0 = EOF; invalid UTF-8 is an io::Error
    bool read_line(std::string& line, bool& eof) {
        line.clear(); eof = false;
        int c;
        while ((c = std::fgetc(f)) != EOF) { line.push_back((char)c); if (c == '\n') break; }
        if (line.empty()) { eof = true; return true; }
        return valid_utf8(line);
    }
    std::vector<AnyValue> parse_line(const std::string& line) {  // :42-121
        std::vector<std::string> fields;
        size_t pos = 0;
        for (;;) {
            const size_t q = line.find(delim, pos);
            fields.push_back(trim(line.substr(pos, q == std::string::npos ? std::string::npos : q - pos)));
            if (q == std::string::npos) break;
            pos = q + delim.size();
        }
        const size_t nf = schema_->fields.size();
        if (fields.size() != nf)
            throw OracleError("Line " + std::to_string(current_line) + ": Expected " + std::to_string(nf) + " fields, found " + std::to_string(fields.size()));
        std::vector<AnyValue> values;
        for (size_t i = 0; i < nf; ++i) {
            const std::string& s = fields[i];
            const bool null = s.empty() || s == "null";
            auto bad = [&](const char* ty) {
                return OracleError("Line " + std::to_string(current_line) + ", field " + std::to_string(i) + ": Cannot parse '" + s + "' as " + ty);
            };
            switch (schema_->fields[i].data_type) {
                case ExecType::Int64: { int64_t v; if (null) values.push_back(AnyValue::Null()); else if (parse_i64(s, v)) values.push_back(AnyValue::Int64(v)); else throw bad("Int64"); break; }
                case ExecType::Float64: { double v; if (null) values.push_back(AnyValue::Null()); else if (parse_f64(s, v)) values.push_back(AnyValue::Float64(v)); else throw bad("Float64"); break; }
                case ExecType::String: values.push_back(null ? AnyValue::Null() : AnyValue::String(s)); break;
                case ExecType::Boolean: {
                    if (null) { values.push_back(AnyValue::Null()); break; }
                    const std::string l = ascii_lower(s);  // to_lowercase(): no non-ASCII character lowers to one of t r u e f a l s 0 1
                    if (l == "true" || l == "t" || l == "1") values.push_back(AnyValue::Boolean(true));
                    else if (l == "false" || l == "f" || l == "0") values.push_back(AnyValue::Boolean(false));
                    else throw bad("Boolean");
                    break;
                }
                case ExecType::Null: values.push_back(AnyValue::Null()); break;
            }
        }
        return values;
    }
    std::optional<RecordBatch> next_batch() override {  // read_batch :123-199
        if (finished) return std::nullopt;
        const size_t nf = schema_->fields.size();
        std::vector<std::vector<AnyValue>> data(nf);
        size_t lines_read = 0;
        std::string line;
        bool eof;
        if (current_line == 0) {  // the first line is a header, always (:134-151)
            if (!read_line(line, eof)) throw OracleError("Stream execution error: Failed to read header: stream did not contain valid UTF-8");
            if (eof) { finished = true; return std::nullopt; }
            ++current_line;
        }
        while (lines_read < batch_size) {
            if (!read_line(line, eof))
                throw OracleError("Stream execution error: Failed to read line " + std::to_string(current_line + 1) + ": stream did not contain valid UTF-8");
            if (eof) { finished = true; break; }
            ++current_line;
            if (!line.empty() && line.back() == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); }
            if (trim(line).empty()) continue;
            std::vector<AnyValue> values;
            try { values = parse_line(line); }
            catch (const OracleError& e) { throw OracleError(std::string("Stream execution error: Parse error: ") + e.what()); }
            for (size_t c = 0; c < nf; ++c) data[c].push_back(std::move(values[c]));
            ++lines_read;
        }
        if (lines_read == 0) return std::nullopt;
        return build_record_batch(data, lines_read);
    }
    RecordBatch build_record_batch(const std::vector<std::vector<AnyValue>>& data, size_t num_rows) {  // :201-326
        std::vector<ArrayRef> cols;
        for (size_t c = 0; c < data.size(); ++c) {
            switch (schema_->fields[c].data_type) {
                case ExecType::Int64: case ExecType::Float64: {
                    const bool is_i = schema_->fields[c].data_type == ExecType::Int64;
                    std::vector<int64_t> vi; std::vector<double> vf; std::vector<bool> nulls;
                    bool any = false;
                    for (const auto& v : data[c]) {
                        const bool n = v.is_null();
                        if (is_i) vi.push_back(n ? 0 : v.i); else vf.push_back(n ? 0.0 : v.f);
                        nulls.push_back(n); any = any || n;
                    }
                    std::optional<std::vector<bool>> validity;
                    if (any) {
                        // :233-239 passes `nulls` as the validity vector; the corrected form passes its complement
                        if (!g_csv_reference_validity) nulls.flip();
                        validity = std::move(nulls);
                    }
                    if (is_i) cols.push_back(PrimitiveArray<int64_t>::make(std::move(vi), std::move(validity)));
                    else cols.push_back(PrimitiveArray<double>::make(std::move(vf), std::move(validity)));
                    break;
                }
                case ExecType::String: {
                    std::vector<std::optional<std::string>> v;
                    for (const auto& x : data[c]) v.push_back(x.is_null() ? std::nullopt : std::optional<std::string>(x.s));
                    cols.push_back(StringArray::make(v));
                    break;
                }
                case ExecType::Boolean: {
                    std::vector<std::optional<bool>> v;
                    for (const auto& x : data[c]) v.push_back(x.is_null() ? std::nullopt : std::optional<bool>(x.b));
                    cols.push_back(BooleanArray::make(v));
                    break;
                }
                case ExecType::Null: cols.push_back(std::make_shared<NullArray>(num_rows)); break;
            }
        }
        try { return RecordBatch::try_new(schema_, std::move(cols)); }
        catch (const OracleError& e) { throw OracleError(std::string("Stream execution error: Failed to create RecordBatch: ") + e.what()); }
    }
};
}  // namespace

DataStreamRef csv_file_stream(const std::string& path, SchemaRef schema, std::optional<size_t> batch_size, std::optional<std::string> delimiter) {  // :20-40
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { const int e = errno; throw OracleError(std::string("Failed to open file: ") + std::strerror(e) + " (os error " + std::to_string(e) + ")"); }
    auto s = std::make_unique<CsvFileStream>();
    s->f = f; s->schema_ = schema;
    s->batch_size = batch_size ? *batch_size : calculate_adaptive_batch_size(*schema);
    if (delimiter) s->delim = *delimiter;
    return s;
}

}  // namespace orc
