// rivulus.cpp — host layer over the C ABI (see rivulus.hpp).  Pure host logic (plan building, optimizer, validation,
// lowering, dtype inference, error text) follows the reference line by line; every data-path step is a call into
// librivulus_gpu.so.  Nothing here computes a query result on the CPU.
#include "rivulus.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

namespace rivulus {

// ------------------------------------------------------------------------------------------------- page-locked host buffers
namespace {
constexpr size_t kPinMin = 64 * 1024;          // smaller buffers are not worth a pinned block
constexpr size_t kPoolKeepBytes = 2ull << 30;  // cached (free) pinned bytes kept for reuse
struct PinnedPool {
    std::mutex mu;
    std::map<size_t, std::vector<void*>> free_by_class;   // size class (power of two) -> free blocks
    std::set<void*> pinned;                               // every live or cached pinned block
    size_t cached = 0;
    bool unavailable = false;                             // cudaHostAlloc failed once (no usable device): stop asking
    ~PinnedPool() {}                                      // blocks are returned to the driver at process exit
};
PinnedPool& pool() { static PinnedPool* p = new PinnedPool(); return *p; }   // intentionally leaked: buffers may outlive static destruction
size_t size_class(size_t bytes) { size_t c = kPinMin; while (c < bytes) c <<= 1; return c; }
}  // namespace

void* host_buffer_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    if (bytes >= kPinMin) {
        PinnedPool& P = pool();
        const size_t cls = size_class(bytes);
        {
            std::lock_guard<std::mutex> g(P.mu);
            auto it = P.free_by_class.find(cls);
            if (it != P.free_by_class.end() && !it->second.empty()) {
                void* p = it->second.back(); it->second.pop_back(); P.cached -= cls;
                return p;
            }
            if (P.unavailable) goto pageable;
        }
        void* p = nullptr;
        if (rvl_host_alloc(cls, &p) == RVL_OK && p != nullptr) {
            std::lock_guard<std::mutex> g(P.mu);
            P.pinned.insert(p);
            return p;
        }
        { std::lock_guard<std::mutex> g(P.mu); P.unavailable = true; }
    }
pageable:
    void* p = std::malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
}

void host_buffer_free(void* p, size_t bytes) {
    if (!p) return;
    if (bytes >= kPinMin) {
        PinnedPool& P = pool();
        std::unique_lock<std::mutex> g(P.mu);
        if (P.pinned.count(p)) {
            const size_t cls = size_class(bytes);
            if (P.cached + cls <= kPoolKeepBytes) { P.free_by_class[cls].push_back(p); P.cached += cls; return; }
            P.pinned.erase(p);
            g.unlock();
            rvl_host_free(p);
            return;
        }
    }
    std::free(p);
}

void host_buffer_pool_trim() {
    PinnedPool& P = pool();
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> g(P.mu);
        for (auto& kv : P.free_by_class) { for (void* p : kv.second) { drop.push_back(p); P.pinned.erase(p); } kv.second.clear(); }
        P.cached = 0;
    }
    for (void* p : drop) rvl_host_free(p);
}

// ------------------------------------------------------------------------------------------------- helpers
static void check(int32_t rc) {
    if (rc != RVL_OK) {
        const char* m = rvl_last_error();
        throw Error(m ? m : "rivulus_gpu: unknown error", rc == RVL_OUT_OF_BOUNDS);
    }
}
template <class V> static inline bool get_bit(const V& b, size_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
template <class V> static inline void set_bit(V& b, size_t i) { b[i >> 3] |= (uint8_t)(1u << (i & 7)); }
template <class V> static size_t count_valid(const V& bits, size_t n) {
    size_t c = 0;
    for (size_t i = 0; i < n; ++i) c += get_bit(bits, i);
    return c;
}

// ------------------------------------------------------------------------------------------------- datatypes/series.rs
const char* dtype_name(DataType d) {
    switch (d) {
        case DataType::Int64: return "Int64";
        case DataType::Float64: return "Float64";
        case DataType::String: return "String";
        case DataType::Boolean: return "Boolean";
        case DataType::Null: return "Null";
    }
    return "?";
}

DataType AnyValue::data_type() const {
    switch (tag) {
        case kNull: return DataType::Null;
        case kInt64: return DataType::Int64;
        case kFloat64: return DataType::Float64;
        case kString: return DataType::String;
        case kBoolean: return DataType::Boolean;
    }
    return DataType::Null;
}

static std::string f64_text(double v, bool debug) {  // Rust `{}` / `{:?}` of an f64
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    std::string s(buf, r.ptr);
    if (debug && s.find('.') == std::string::npos) s += ".0";
    return s;
}
static std::string quoted(const std::string& s) {  // `{:?}` of a str
    std::string o = "\"";
    for (char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\t') o += "\\t";
        else if (c == '\r') o += "\\r";
        else o += c;
    }
    return o + "\"";
}
std::string AnyValue::display() const {
    switch (tag) {
        case kNull: return "null";
        case kInt64: return std::to_string(i);
        case kFloat64: return f64_text(f, false);
        case kString: return s;
        case kBoolean: return b ? "true" : "false";
    }
    return "";
}
std::string AnyValue::debug() const {
    switch (tag) {
        case kNull: return "Null";
        case kInt64: return "Int64(" + std::to_string(i) + ")";
        case kFloat64: return "Float64(" + f64_text(f, true) + ")";
        case kString: return "String(" + quoted(s) + ")";
        case kBoolean: return std::string("Boolean(") + (b ? "true" : "false") + ")";
    }
    return "";
}

static bool types_compatible(DataType e, DataType f) {  // series.rs:255-264
    if (e == f) return true;
    return (e == DataType::Int64 && f == DataType::Float64) || (e == DataType::Float64 && f == DataType::Int64);
}

Series Series::make(const std::string& name, const std::vector<AnyValue>& data) {  // series.rs:185-221
    if (data.empty()) throw Error("Empty series not allowed");
    std::optional<DataType> first;
    for (const auto& v : data)
        if (!v.is_null()) { first = v.data_type(); break; }
    DataType dtype = first.value_or(DataType::Null);
    bool saw_int = false;
    for (const auto& v : data) {
        if (v.is_null()) continue;
        const DataType cur = v.data_type();
        if (!types_compatible(dtype, cur))
            throw Error(std::string("Mixed types in series: expected ") + dtype_name(dtype) + ", found " + dtype_name(cur));
        if (dtype == DataType::Int64 && cur == DataType::Float64) dtype = DataType::Float64;
        saw_int |= cur == DataType::Int64;
    }
    Series s;
    s.name_ = name; s.dtype_ = dtype; s.len_ = data.size();
    const size_t n = data.size();
    bool any_null = false;
    for (const auto& v : data) any_null |= v.is_null();
    if (dtype == DataType::Null) return s;
    if (any_null) {
        s.b_->validity_.assign((n + 7) / 8, 0);
        for (size_t i = 0; i < n; ++i) if (!data[i].is_null()) set_bit(s.b_->validity_, i);
    }
    switch (dtype) {
        case DataType::Int64:
            s.b_->i64_.resize(n);
            for (size_t i = 0; i < n; ++i) s.b_->i64_[i] = data[i].tag == AnyValue::kInt64 ? data[i].i : 0;
            break;
        case DataType::Float64: {
            s.b_->f64_.resize(n);
            const bool mixed = saw_int;
            if (mixed) s.b_->int_tag_.assign((n + 7) / 8, 0);
            for (size_t i = 0; i < n; ++i) {
                if (data[i].tag == AnyValue::kFloat64) s.b_->f64_[i] = data[i].f;
                else {
                    s.b_->f64_[i] = 0.0;
                    if (data[i].tag == AnyValue::kInt64) { set_bit(s.b_->int_tag_, i); std::memcpy(&s.b_->f64_[i], &data[i].i, 8); }
                }
            }
            break;
        }
        case DataType::Boolean:
            s.b_->bits_.assign((n + 7) / 8, 0);
            for (size_t i = 0; i < n; ++i) if (data[i].tag == AnyValue::kBoolean && data[i].b) set_bit(s.b_->bits_, i);
            break;
        case DataType::String: {
            s.b_->offsets_.resize(n + 1);
            s.b_->offsets_[0] = 0;
            size_t total = 0;
            for (size_t i = 0; i < n; ++i) { if (data[i].tag == AnyValue::kString) total += data[i].s.size(); }
            if (total > (size_t)INT32_MAX) throw Error("String data exceeds int32 offsets");
            s.b_->data_.reserve(total);
            for (size_t i = 0; i < n; ++i) {
                if (data[i].tag == AnyValue::kString) s.b_->data_.insert(s.b_->data_.end(), data[i].s.begin(), data[i].s.end());
                s.b_->offsets_[i + 1] = (int32_t)s.b_->data_.size();
            }
            break;
        }
        default: break;
    }
    return s;
}

Series Series::empty(const std::string& name, DataType dtype) {  // series.rs:223-229
    Series s; s.name_ = name; s.dtype_ = dtype; s.len_ = 0;
    if (dtype == DataType::String) s.b_->offsets_.assign(1, 0);
    return s;
}

void Series::infer_from_validity() {
    if (len_ == 0) throw Error("Empty series not allowed");
    if (!b_->validity_.empty()) {
        const size_t valid = count_valid(b_->validity_, len_);
        if (valid == len_) b_->validity_.clear();
        else if (valid == 0) {  // every value null -> DataType::Null (series.rs:190-198)
            dtype_ = DataType::Null;
            b_->i64_.clear(); b_->f64_.clear(); b_->bits_.clear(); b_->offsets_.clear(); b_->data_.clear(); b_->validity_.clear();
        }
    }
}
Series Series::from_i64(const std::string& name, const int64_t* v, size_t n, const uint8_t* validity_bits) {
    Series s; s.name_ = name; s.dtype_ = DataType::Int64; s.len_ = n;
    s.b_->i64_.assign(v, v + n);
    if (validity_bits) s.b_->validity_.assign(validity_bits, validity_bits + (n + 7) / 8);
    if (!s.b_->validity_.empty()) for (size_t i = 0; i < n; ++i) if (!get_bit(s.b_->validity_, i)) s.b_->i64_[i] = 0;
    s.infer_from_validity();
    return s;
}
Series Series::from_f64(const std::string& name, const double* v, size_t n, const uint8_t* validity_bits) {
    Series s; s.name_ = name; s.dtype_ = DataType::Float64; s.len_ = n;
    s.b_->f64_.assign(v, v + n);
    if (validity_bits) s.b_->validity_.assign(validity_bits, validity_bits + (n + 7) / 8);
    if (!s.b_->validity_.empty()) for (size_t i = 0; i < n; ++i) if (!get_bit(s.b_->validity_, i)) s.b_->f64_[i] = 0.0;
    s.infer_from_validity();
    return s;
}
Series Series::from_bool_bits(const std::string& name, const uint8_t* value_bits, size_t n, const uint8_t* validity_bits) {
    Series s; s.name_ = name; s.dtype_ = DataType::Boolean; s.len_ = n;
    s.b_->bits_.assign(value_bits, value_bits + (n + 7) / 8);
    if (validity_bits) s.b_->validity_.assign(validity_bits, validity_bits + (n + 7) / 8);
    if (n & 7) s.b_->bits_.back() &= (uint8_t)((1u << (n & 7)) - 1u);
    if (!s.b_->validity_.empty()) for (size_t i = 0; i < (n + 7) / 8; ++i) s.b_->bits_[i] &= s.b_->validity_[i];
    s.infer_from_validity();
    return s;
}
Series Series::from_strings(const std::string& name, const int32_t* offsets, size_t n, const uint8_t* data, const uint8_t* validity_bits) {
    Series s; s.name_ = name; s.dtype_ = DataType::String; s.len_ = n;
    s.b_->offsets_.assign(offsets, offsets + n + 1);
    s.b_->data_.assign(data, data + offsets[n]);
    if (validity_bits) s.b_->validity_.assign(validity_bits, validity_bits + (n + 7) / 8);
    s.infer_from_validity();
    return s;
}
static const uint8_t* checked_validity(const std::vector<uint8_t>& validity, size_t n) {
    if (validity.empty()) return nullptr;
    if (validity.size() < (n + 7) / 8) throw Error("validity bitmap shorter than the column");
    return validity.data();
}
Series Series::from_i64(const std::string& name, std::vector<int64_t> v, std::vector<uint8_t> validity_bits) {
    return from_i64(name, v.data(), v.size(), checked_validity(validity_bits, v.size()));
}
Series Series::from_f64(const std::string& name, std::vector<double> v, std::vector<uint8_t> validity_bits) {
    return from_f64(name, v.data(), v.size(), checked_validity(validity_bits, v.size()));
}
Series Series::from_bool_bits(const std::string& name, std::vector<uint8_t> value_bits, size_t n, std::vector<uint8_t> validity_bits) {
    if (value_bits.size() < (n + 7) / 8) throw Error("value bitmap shorter than the column");
    return from_bool_bits(name, value_bits.data(), n, checked_validity(validity_bits, n));
}
Series Series::from_strings(const std::string& name, std::vector<int32_t> offsets, std::vector<uint8_t> data, std::vector<uint8_t> validity_bits) {
    if (offsets.empty()) throw Error("Empty series not allowed");
    if (data.size() < (size_t)offsets.back()) throw Error("string data shorter than the offsets say");
    return from_strings(name, offsets.data(), offsets.size() - 1, data.data(), checked_validity(validity_bits, offsets.size() - 1));
}

AnyValue Series::at(size_t i) const {  // series.rs:273-288
    if (i >= len_) throw Error("Index " + std::to_string(i) + " out of bounds for series of length " + std::to_string(len_), true);
    if (!is_valid(i)) return AnyValue::Null();
    switch (dtype_) {
        case DataType::Int64: return AnyValue::Int64(b_->i64_[i]);
        case DataType::Float64:
            if (!b_->int_tag_.empty() && get_bit(b_->int_tag_, i)) { int64_t v; std::memcpy(&v, &b_->f64_[i], 8); return AnyValue::Int64(v); }
            return AnyValue::Float64(b_->f64_[i]);
        case DataType::Boolean: return AnyValue::Boolean(get_bit(b_->bits_, i));
        case DataType::String: return AnyValue::String(std::string(b_->data_.begin() + b_->offsets_[i], b_->data_.begin() + b_->offsets_[i + 1]));
        case DataType::Null: return AnyValue::Null();
    }
    return AnyValue::Null();
}
std::optional<AnyValue> Series::get(size_t i) const { return i < len_ ? std::optional<AnyValue>(at(i)) : std::nullopt; }
std::vector<AnyValue> Series::to_values() const {
    std::vector<AnyValue> v; v.reserve(len_);
    for (size_t i = 0; i < len_; ++i) v.push_back(at(i));
    return v;
}
std::string Series::display() const { return std::string("Series: numbers [") + dtype_name(dtype_) + "; " + std::to_string(len_) + "]"; }
size_t Series::null_count() const {
    if (dtype_ == DataType::Null) return len_;
    if (b_->validity_.empty()) return 0;
    return len_ - count_valid(b_->validity_, len_);
}

static int32_t rvl_dtype_of(DataType d) {
    switch (d) {
        case DataType::Int64: return RVL_INT64;
        case DataType::Float64: return RVL_FLOAT64;
        case DataType::String: return RVL_STRING;
        case DataType::Boolean: return RVL_BOOLEAN;
        case DataType::Null: return RVL_NULL;
    }
    return RVL_NULL;
}
static ExecType exec_type_of(DataType d) { return (ExecType)rvl_dtype_of(d); }
static DataType series_type_of(int32_t rvl) {
    switch (rvl) {
        case RVL_INT64: return DataType::Int64;
        case RVL_FLOAT64: return DataType::Float64;
        case RVL_STRING: return DataType::String;
        case RVL_BOOLEAN: return DataType::Boolean;
        default: return DataType::Null;
    }
}

rvl_column Series::as_column(size_t offset, size_t length, bool flatten_nulls) const {
    rvl_column c{};
    c.dtype = rvl_dtype_of(dtype_);
    c.location = RVL_HOST;
    c.length = (int64_t)length;
    c.offset = (int64_t)offset;
    const bool keep_validity = !b_->validity_.empty() && !(flatten_nulls && dtype_ != DataType::String);
    c.validity = keep_validity ? b_->validity_.data() : nullptr;
    switch (dtype_) {
        case DataType::Int64: c.values = b_->i64_.data(); break;
        case DataType::Float64: c.values = b_->f64_.data(); break;
        case DataType::Boolean: c.values = b_->bits_.data(); break;
        case DataType::String: c.offsets = b_->offsets_.data(); c.data = b_->data_.data(); c.data_len = (int64_t)b_->data_.size(); break;
        case DataType::Null: break;
    }
    return c;
}

rvl_column Series::tag_column(size_t offset, size_t length) const {
    rvl_column c{};
    c.dtype = RVL_BOOLEAN; c.location = RVL_HOST; c.length = (int64_t)length; c.offset = (int64_t)offset;
    c.values = b_->int_tag_.data();
    return c;
}

Series Series::from_mixed(const std::string& name, const rvl_column& v, const rvl_column& t, DataType dtype) {
    const size_t n = (size_t)v.length;
    if (n == 0 || dtype == DataType::Null) { rvl_column c = v; if (dtype == DataType::Null) { c.dtype = RVL_NULL; c.null_count = c.length; } return from_column(name, c, dtype); }
    Series s; s.name_ = name; s.len_ = n; s.dtype_ = dtype;
    if (v.validity != nullptr && v.null_count > 0) s.b_->validity_.assign(v.validity, v.validity + (n + 7) / 8);
    const uint8_t* tags = (const uint8_t*)t.values;
    if (dtype == DataType::Int64) {  // only Int64 survivors: an ordinary Int64 series
        s.b_->i64_.assign((const int64_t*)v.values, (const int64_t*)v.values + n);
        return s;
    }
    s.b_->f64_.assign((const double*)v.values, (const double*)v.values + n);
    bool any_int = false;
    for (size_t i = 0; i < (n + 7) / 8; ++i) any_int |= tags[i] != 0;
    if (any_int) s.b_->int_tag_.assign(tags, tags + (n + 7) / 8);
    return s;
}

Series Series::from_column(const std::string& name, const rvl_column& c, DataType dtype_if_empty) {
    // the caller downloaded into buffers it owns and passes them here through `c` (host pointers, offset 0)
    const size_t n = (size_t)c.length;
    if (n == 0) return Series::empty(name, dtype_if_empty);
    Series s; s.name_ = name; s.len_ = n; s.dtype_ = series_type_of(c.dtype);
    if (c.dtype == RVL_NULL || c.null_count == c.length) { s.dtype_ = DataType::Null; return s; }  // series.rs:190-198
    if (c.validity != nullptr && c.null_count > 0) s.b_->validity_.assign(c.validity, c.validity + (n + 7) / 8);
    switch (c.dtype) {
        case RVL_INT64: s.b_->i64_.assign((const int64_t*)c.values, (const int64_t*)c.values + n); break;
        case RVL_FLOAT64: s.b_->f64_.assign((const double*)c.values, (const double*)c.values + n); break;
        case RVL_BOOLEAN: s.b_->bits_.assign((const uint8_t*)c.values, (const uint8_t*)c.values + (n + 7) / 8); break;
        case RVL_STRING:
            s.b_->offsets_.assign(c.offsets, c.offsets + n + 1);
            s.b_->data_.assign(c.data, c.data + c.data_len);
            break;
        default: break;
    }
    return s;
}

Series Series::from_array(const std::string& name, ArrayData&& a, DataType dtype) {
    // the buffers were downloaded straight into (page-locked) HostVecs: they are moved in, not copied
    const size_t n = (size_t)a.length;
    if (n == 0) return Series::empty(name, dtype);
    Series s; s.name_ = name; s.len_ = n; s.dtype_ = series_type_of((int32_t)a.dtype);
    if (dtype == DataType::Null || a.dtype == ExecType::Null || a.null_count == a.length) { s.dtype_ = DataType::Null; return s; }  // series.rs:190-198
    if (a.has_validity && a.null_count > 0) s.b_->validity_ = std::move(a.validity);
    switch (a.dtype) {
        case ExecType::Int64: s.b_->i64_ = std::move(a.i64); break;
        case ExecType::Float64: s.b_->f64_ = std::move(a.f64); break;
        case ExecType::Boolean: s.b_->bits_ = std::move(a.bits); break;
        case ExecType::String: s.b_->offsets_ = std::move(a.offsets); s.b_->data_ = std::move(a.data); break;
        default: break;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------- datatypes/dataframe.rs
DataFrame DataFrame::make(std::vector<Series> columns) {  // dataframe.rs:29-56
    DataFrame df;
    if (columns.empty()) return df;
    std::set<std::string> seen;
    for (const auto& c : columns)
        if (!seen.insert(c.name()).second) throw Error("Duplicate column name: '" + c.name() + "'");
    const size_t expected = columns.front().len();
    for (const auto& c : columns)
        if (c.len() != expected)
            throw Error("Column lengths mismatch: expected " + std::to_string(expected) + ", found " + std::to_string(c.len()) +
                        " for column '" + c.name() + "'");
    df.columns_ = std::move(columns);
    return df;
}
const Series* DataFrame::column(const std::string& name) const {
    for (const auto& s : columns_) if (s.name() == name) return &s;
    return nullptr;
}
std::vector<std::string> DataFrame::column_names() const {
    std::vector<std::string> n;
    for (const auto& s : columns_) n.push_back(s.name());
    return n;
}
DataFrame DataFrame::select(const std::vector<std::string>& names) const {  // dataframe.rs:96-110
    std::vector<Series> cols;
    for (const auto& n : names) {
        const Series* s = column(n);
        if (!s) throw Error("Column not found: '" + n + "'");
        cols.push_back(*s);
    }
    return DataFrame::unchecked(std::move(cols));
}
const Series& DataFrame::operator[](const std::string& name) const {
    const Series* s = column(name);
    if (!s) throw Error("Column '" + name + "' not found", true);
    return *s;
}

// ------------------------------------------------------------------------------------------------- expressions/expr.rs
const char* op_name(BinaryOperator op) {
    static const char* n[] = {"Plus", "Minus", "Multiply", "Divide", "Eq", "NotEq", "Lt", "Gt", "LtEq", "GtEq", "And", "Or"};
    return n[(int)op];
}
std::string Expr::debug() const {
    switch (kind) {
        case Column: return "Column(" + quoted(name) + ")";
        case Literal: return "Literal(" + value.debug() + ")";
        case Alias: return "Alias(" + left->debug() + ", " + quoted(name) + ")";
        case Binary: return "BinaryExpr { left: " + left->debug() + ", op: " + op_name(op) + ", right: " + right->debug() + " }";
    }
    return "";
}

// ------------------------------------------------------------------------------------------------- context
Context::Context(int device) { check(rvl_ctx_create(device, &ctx_)); }
Context::~Context() { if (ctx_) rvl_ctx_destroy(ctx_); }
std::shared_ptr<Context> Context::shared(int device) {
    static std::mutex mu;
    static std::map<int, std::shared_ptr<Context>> all;
    std::lock_guard<std::mutex> g(mu);
    auto& c = all[device];
    if (!c) c = std::make_shared<Context>(device);
    return c;
}
int64_t launch_count(int device) {
    int64_t n = 0;
    check(rvl_ctx_launch_count(Context::shared(device)->handle(), &n));
    return n;
}

// ------------------------------------------------------------------------------------------------- schema
const char* exec_type_name(ExecType t) {
    switch (t) {
        case ExecType::Null: return "Null";
        case ExecType::Boolean: return "Boolean";
        case ExecType::Int64: return "Int64";
        case ExecType::Float64: return "Float64";
        case ExecType::String: return "String";
    }
    return "?";
}
std::optional<size_t> Schema::index_of(const std::string& n) const {
    for (size_t i = 0; i < fields.size(); ++i) if (fields[i].name == n) return i;
    return std::nullopt;
}
const Field* Schema::field_by_name(const std::string& n) const {
    auto i = index_of(n);
    return i ? &fields[*i] : nullptr;
}

AnyValue ArrayData::value(size_t i) const {
    if (i >= (size_t)length) throw Error("Index " + std::to_string(i) + " out of bounds", true);
    if (dtype == ExecType::Null) return AnyValue::Null();
    if (has_validity && !get_bit(validity, i)) return AnyValue::Null();
    switch (dtype) {
        case ExecType::Int64: return AnyValue::Int64(i64[i]);
        case ExecType::Float64: return AnyValue::Float64(f64[i]);
        case ExecType::Boolean: return AnyValue::Boolean(get_bit(bits, i));
        case ExecType::String: return AnyValue::String(std::string(data.begin() + offsets[i], data.begin() + offsets[i + 1]));
        default: return AnyValue::Null();
    }
}

// ------------------------------------------------------------------------------------------------- RecordBatch
RecordBatch::Handle::~Handle() { if (b) rvl_batch_release(b); }

RecordBatch RecordBatch::adopt(const ContextRef& ctx, SchemaRef schema, rvl_batch* handle) {
    RecordBatch rb; rb.ctx_ = ctx; rb.schema_ = std::move(schema);
    rb.h_ = std::make_shared<Handle>(); rb.h_->b = handle;
    return rb;
}

RecordBatch RecordBatch::try_new(const ContextRef& ctx, SchemaRef schema, const std::vector<rvl_column>& cols) {  // record_batch.rs:16-58
    if (schema->fields.size() != cols.size())
        throw Error("Schema has " + std::to_string(schema->fields.size()) + " fields but " + std::to_string(cols.size()) + " columns provided");
    const int64_t n = cols.empty() ? 0 : cols[0].length;
    for (size_t i = 0; i < cols.size(); ++i)
        if (cols[i].length != n)
            throw Error("Column " + std::to_string(i) + " has length " + std::to_string(cols[i].length) + " but expected " + std::to_string(n));
    for (size_t i = 0; i < cols.size(); ++i)
        if ((int32_t)schema->fields[i].data_type != cols[i].dtype)
            throw Error("Column " + std::to_string(i) + " has type " + exec_type_name((ExecType)cols[i].dtype) + " but schema expects " +
                        exec_type_name(schema->fields[i].data_type));
    rvl_batch* b = nullptr;
    check(rvl_batch_upload(ctx->handle(), cols.data(), (int32_t)cols.size(), &b));
    return adopt(ctx, std::move(schema), b);
}

static std::string batch_mismatch(const Schema& schema, const std::vector<std::pair<int64_t, int32_t>>& cols /* (length, dtype) */, size_t num_rows) {
    // validate() :348-378
    if (schema.fields.size() != cols.size())
        return "Schema has " + std::to_string(schema.fields.size()) + " fields but " + std::to_string(cols.size()) + " columns present";
    for (size_t i = 0; i < cols.size(); ++i) {
        if ((size_t)cols[i].first != num_rows)
            return "Column " + std::to_string(i) + " has length " + std::to_string(cols[i].first) + " but expected " + std::to_string(num_rows);
        if ((int32_t)schema.fields[i].data_type != cols[i].second)
            return std::string("Column ") + std::to_string(i) + " has type " + exec_type_name((ExecType)cols[i].second) + " but schema expects " +
                   exec_type_name(schema.fields[i].data_type);
    }
    return "";
}
RecordBatch RecordBatch::new_unchecked(const ContextRef& ctx, SchemaRef schema, const std::vector<rvl_column>& cols, size_t num_rows) {  // :60-66
    std::vector<std::pair<int64_t, int32_t>> shape;
    for (const auto& c : cols) shape.emplace_back(c.length, c.dtype);
    const std::string why = batch_mismatch(*schema, shape, num_rows);
    if (why.empty()) return try_new(ctx, std::move(schema), cols);
    RecordBatch r; r.ctx_ = ctx; r.schema_ = std::move(schema); r.invalid_ = why;
    return r;
}
void RecordBatch::validate() const {  // record_batch.rs:348-378
    if (!invalid_.empty()) throw Error(invalid_);
    std::vector<std::pair<int64_t, int32_t>> shape;
    const size_t nc = num_columns();
    for (size_t i = 0; i < nc; ++i) {
        rvl_column v{};
        check(rvl_batch_column(handle(), (int32_t)i, &v));
        shape.emplace_back(v.length, v.dtype);
    }
    const std::string why = batch_mismatch(*schema_, shape, num_rows());
    if (!why.empty()) throw Error(why);
}
size_t RecordBatch::memory_size() const {  // record_batch.rs:380-400: size_of_val(Schema) = 24, size_of::<Vec<ArrayRef>>() = 24, ArrayRef = 16
    const size_t n = num_rows();
    size_t total = 24 + 24 + schema_->fields.size() * 16;
    for (const auto& f : schema_->fields) switch (f.data_type) {
        case ExecType::Int64: case ExecType::Float64: total += n * 8; break;
        case ExecType::Boolean: total += (n + 7) / 8; break;
        case ExecType::String: total += n * 20; break;
        case ExecType::Null: total += 16; break;
    }
    return total;
}
std::optional<ArrayData> RecordBatch::column_by_name(const std::string& name) const {  // record_batch.rs:84-86
    auto i = schema_->index_of(name);
    if (!i) return std::nullopt;
    return column_data(*i);
}
void RecordBatchBuilder::add_column(const rvl_column& c) {  // record_batch.rs:518-546
    if (columns_.size() >= schema_->fields.size()) throw Error("Cannot add more columns than schema defines");
    const Field& f = schema_->fields[columns_.size()];
    if (c.dtype != (int32_t)f.data_type)
        throw Error(std::string("Column type ") + exec_type_name((ExecType)c.dtype) + " doesn't match expected type " + exec_type_name(f.data_type));
    if (!columns_.empty() && c.length != columns_[0].length)
        throw Error("Column length " + std::to_string(c.length) + " doesn't match expected length " + std::to_string(columns_[0].length));
    columns_.push_back(c);
}
RecordBatch RecordBatchBuilder::finish() const {  // record_batch.rs:548-558
    if (columns_.size() != schema_->fields.size())
        throw Error("Expected " + std::to_string(schema_->fields.size()) + " columns but only " + std::to_string(columns_.size()) + " provided");
    return RecordBatch::try_new(ctx_ ? ctx_ : Context::shared(0), schema_, columns_);
}

RecordBatch RecordBatch::empty(const ContextRef& ctx, SchemaRef schema) {  // record_batch.rs:402-421
    std::vector<rvl_column> cols;
    static const int32_t zero_off[1] = {0};
    for (const auto& f : schema->fields) {
        rvl_column c{};
        c.dtype = (int32_t)f.data_type; c.location = RVL_HOST; c.length = 0;
        if (f.data_type == ExecType::String) c.offsets = zero_off;
        cols.push_back(c);
    }
    rvl_batch* b = nullptr;
    check(rvl_batch_upload(ctx->handle(), cols.data(), (int32_t)cols.size(), &b));
    return adopt(ctx, std::move(schema), b);
}

size_t RecordBatch::num_rows() const {
    if (!handle()) return 0;
    int64_t n = 0; check(rvl_batch_num_rows(handle(), &n));
    return (size_t)n;
}
size_t RecordBatch::num_columns() const {
    if (!handle()) return 0;
    int32_t n = 0; check(rvl_batch_num_columns(handle(), &n));
    return (size_t)n;
}

RecordBatch RecordBatch::slice(size_t offset, size_t length) const {  // record_batch.rs:92-106
    if (!(offset + length <= num_rows())) throw Error("Slice out of bounds", true);
    rvl_batch* v = nullptr;
    check(rvl_batch_slice(handle(), (int64_t)offset, (int64_t)length, &v));
    return adopt(ctx_, schema_, v);
}

RecordBatch RecordBatch::take(const std::vector<size_t>& indices) const {  // record_batch.rs:108-129
    std::vector<int64_t> idx(indices.begin(), indices.end());
    rvl_batch* out = nullptr;
    const int32_t rc = rvl_batch_take(ctx_->handle(), handle(), idx.data(), (int64_t)idx.size(), &out);
    if (rc == RVL_OUT_OF_BOUNDS) throw Error(rvl_last_error());  // an Err(String) in the reference, not a panic
    check(rc);
    return adopt(ctx_, schema_, out);
}

RecordBatch RecordBatch::select_columns(const std::vector<size_t>& indices) const {  // record_batch.rs:180-206
    const size_t nc = num_columns();
    auto out_schema = std::make_shared<Schema>();
    std::vector<int32_t> idx;
    for (size_t i : indices) {
        if (i >= nc) throw Error("Column index " + std::to_string(i) + " out of bounds for " + std::to_string(nc) + " columns");
        out_schema->fields.push_back(schema_->fields[i]);
        idx.push_back((int32_t)i);
    }
    rvl_batch* v = nullptr;
    check(rvl_batch_select(handle(), idx.data(), (int32_t)idx.size(), &v));
    return adopt(ctx_, out_schema, v);
}

RecordBatch RecordBatch::select_columns_by_name(const std::vector<std::string>& names) const {  // record_batch.rs:208-219
    std::vector<size_t> idx;
    for (const auto& n : names) {
        auto i = schema_->index_of(n);
        if (!i) throw Error("Column '" + n + "' not found");
        idx.push_back(*i);
    }
    return select_columns(idx);
}

RecordBatch RecordBatch::filter(const RecordBatch& pb, size_t pc) const {  // record_batch.rs:221-243
    const size_t plen = pb.num_rows(), n = num_rows();
    if (plen != n) throw Error("Predicate length " + std::to_string(plen) + " doesn't match batch length " + std::to_string(n));
    if (pc >= pb.num_columns() || pb.schema_->fields[pc].data_type != ExecType::Boolean) throw Error("Predicate must be a BooleanArray");
    // one fused launch over [this batch's columns..., predicate column]: mask mode keeps rows whose mask is Some(true)
    const size_t nc = num_columns();
    std::vector<rvl_batch*> owned;
    const rvl_batch* input = handle();
    int32_t mask_index;
    if (pb.handle() == handle()) {
        mask_index = (int32_t)pc;
    } else {
        // the predicate lives in another batch: build a zero-copy view [columns of this batch + the predicate column]
        // by concatenating column views is not expressible in the ABI, so upload-free path = select on each and re-wrap
        std::vector<rvl_column> views(nc + 1);
        for (size_t i = 0; i < nc; ++i) check(rvl_batch_column(handle(), (int32_t)i, &views[i]));
        check(rvl_batch_column(pb.handle(), (int32_t)pc, &views[nc]));
        rvl_batch* joined = nullptr;
        check(rvl_batch_wrap_device(ctx_->handle(), views.data(), (int32_t)views.size(), &joined));
        owned.push_back(joined);
        input = joined;
        mask_index = (int32_t)nc;
    }
    rvl_predicate pred{};
    pred.mode = RVL_PRED_BOOL_COLUMN; pred.column = mask_index;
    std::vector<int32_t> proj(nc);
    for (size_t i = 0; i < nc; ++i) proj[i] = (int32_t)i;
    rvl_batch* out = nullptr;
    const int32_t rc = rvl_filter_project(ctx_->handle(), input, &pred, proj.data(), (int32_t)nc, -1, &out);
    for (auto* b : owned) rvl_batch_release(b);
    check(rc);
    return adopt(ctx_, schema_, out);
}

RecordBatch RecordBatch::concat(const std::vector<RecordBatch>& batches) {  // record_batch.rs:245-275
    if (batches.empty()) throw Error("Cannot concatenate empty batch list");
    for (size_t i = 1; i < batches.size(); ++i)
        if (!(*batches[i].schema_ == *batches[0].schema_)) throw Error("All batches must have the same schema");
    std::vector<const rvl_batch*> hs;
    for (const auto& b : batches) hs.push_back(b.handle());
    rvl_batch* out = nullptr;
    check(rvl_batch_concat(batches[0].ctx_->handle(), hs.data(), (int32_t)hs.size(), &out));
    return adopt(batches[0].ctx_, batches[0].schema_, out);
}

int64_t RecordBatch::column_null_count(size_t i) const {
    rvl_column v{};
    check(rvl_batch_column(handle(), (int32_t)i, &v));
    return v.null_count;
}

ArrayData RecordBatch::column_data(size_t i) const {
    rvl_column v{};
    check(rvl_batch_column(handle(), (int32_t)i, &v));
    ArrayData a;
    a.dtype = (ExecType)v.dtype; a.length = v.length; a.null_count = v.null_count;
    const size_t n = (size_t)v.length;
    rvl_column dst{};
    dst.dtype = v.dtype; dst.location = RVL_HOST; dst.length = v.length;
    a.has_validity = v.validity != nullptr && v.dtype != RVL_NULL;
    if (a.has_validity) { a.validity.assign((n + 7) / 8 + 1, 0); dst.validity = a.validity.data(); }
    switch (v.dtype) {
        case RVL_INT64: a.i64.resize(n + 1); dst.values = a.i64.data(); break;     // not zero-filled: the download overwrites it
        case RVL_FLOAT64: a.f64.resize(n + 1); dst.values = a.f64.data(); break;
        case RVL_BOOLEAN: a.bits.assign((n + 7) / 8 + 1, 0); dst.values = a.bits.data(); break;
        case RVL_STRING: {
            a.offsets.resize(n + 1); a.offsets[0] = 0; dst.offsets = a.offsets.data();
            // the view's data_len is the span of the viewed window
            a.data.resize((size_t)std::max<int64_t>(v.data_len, 0) + 1); dst.data = a.data.data(); dst.data_len = v.data_len;
            break;
        }
        default: break;
    }
    check(rvl_batch_download_column(ctx_->handle(), handle(), (int32_t)i, &dst));
    switch (v.dtype) {
        case RVL_INT64: a.i64.resize(n); break;
        case RVL_FLOAT64: a.f64.resize(n); break;
        case RVL_BOOLEAN: a.bits.resize((n + 7) / 8); break;
        case RVL_STRING: a.data.resize(n ? (size_t)a.offsets[n] : 0); break;
        default: break;
    }
    if (a.has_validity) a.validity.resize((n + 7) / 8);
    return a;
}

// ------------------------------------------------------------------------------------------------- streams
std::vector<RecordBatch> DataStream::collect() {
    std::vector<RecordBatch> out;
    while (auto b = next_batch()) out.push_back(*b);
    return out;
}

namespace {
struct MemoryStream : DataStream {  // stream.rs:58-114
    SchemaRef schema_; std::vector<RecordBatch> batches; size_t cur = 0;
    SchemaRef schema() const override { return schema_; }
    std::optional<RecordBatch> next_batch() override {
        if (cur < batches.size()) return batches[cur++];
        return std::nullopt;
    }
};
struct FilterStream : DataStream {  // stream.rs:116-163
    DataStreamRef input; std::string col;
    SchemaRef schema() const override { return input->schema(); }
    std::optional<RecordBatch> next_batch() override {
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        auto idx = b->schema()->index_of(col);
        if (!idx) throw Error("Stream execution error: Column '" + col + "' not found in schema");
        if (b->schema()->fields[*idx].data_type != ExecType::Boolean)
            throw Error("Stream execution error: Predicate column '" + col + "' is not of boolean type");
        try { return b->filter(*b, *idx); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream execution error: ") + e.what()); }
    }
};
struct SelectStream : DataStream {  // stream.rs:165-213
    DataStreamRef input; std::vector<std::string> names; SchemaRef out_schema;
    SchemaRef schema() const override { return out_schema; }
    std::optional<RecordBatch> next_batch() override {
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        try { return b->select_columns_by_name(names); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream execution error: ") + e.what()); }
    }
};
struct LimitStream : DataStream {  // streaming.rs:246-288
    DataStreamRef input; size_t limit = 0, rows_returned = 0;
    SchemaRef schema() const override { return input->schema(); }
    std::optional<RecordBatch> next_batch() override {
        if (rows_returned >= limit) return std::nullopt;  // :269-271: no upstream pull once the limit is reached
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        const size_t remaining = limit - rows_returned, rows = b->num_rows();
        if (rows <= remaining) { rows_returned += rows; return b; }
        auto lb = b->slice(0, remaining);
        rows_returned += remaining;
        return lb;
    }
};
}  // namespace

DataStreamRef make_memory_stream(const ContextRef&, SchemaRef schema, std::vector<RecordBatch> batches) {  // stream.rs:66-81
    for (const auto& b : batches)
        if (!(*b.schema() == *schema)) throw Error("Schema mismatch: expected <schema>, found <schema>");
    auto s = std::make_unique<MemoryStream>(); s->schema_ = std::move(schema); s->batches = std::move(batches);
    return s;
}
DataStreamRef make_filter_stream(DataStreamRef in, std::string predicate_column) {
    auto s = std::make_unique<FilterStream>(); s->input = std::move(in); s->col = std::move(predicate_column); return s;
}
DataStreamRef make_select_stream(DataStreamRef in, std::vector<std::string> columns) {  // stream.rs:173-194
    auto in_schema = in->schema();
    auto out = std::make_shared<Schema>();
    for (const auto& c : columns) {
        auto i = in_schema->index_of(c);
        if (!i) throw Error("Stream execution error: Column '" + c + "' not found in schema");
        out->fields.push_back(in_schema->fields[*i]);
    }
    auto s = std::make_unique<SelectStream>(); s->input = std::move(in); s->names = std::move(columns); s->out_schema = out;
    return s;
}
DataStreamRef make_limit_stream(DataStreamRef in, size_t limit) {
    auto s = std::make_unique<LimitStream>(); s->input = std::move(in); s->limit = limit; return s;
}

std::vector<RecordBatch> collect_all_batches(DataStream& s) {  // streaming.rs:335-341
    std::vector<RecordBatch> out;
    while (auto b = s.next_batch()) out.push_back(*b);
    return out;
}
RecordBatch collect_stream_batches(const ContextRef& ctx, DataStream& s) {  // streaming.rs:343-352
    auto schema = s.schema();
    auto batches = collect_all_batches(s);
    if (batches.empty()) return RecordBatch::empty(ctx, schema);
    try { return RecordBatch::concat(batches); }
    catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Conversion error: ") + e.what()); }
}

std::vector<RecordBatch> dataframe_to_batches(const ContextRef& ctx, const DataFrame& df, size_t batch_size) {  // streaming.rs:135-233
    std::vector<RecordBatch> batches;
    if (df.is_empty()) return batches;
    const size_t num_rows = df.height();
    const size_t num_batches = (num_rows + batch_size - 1) / batch_size;
    auto schema = std::make_shared<Schema>();
    for (const auto& s : df.columns()) schema->fields.push_back(Field{s.name(), exec_type_of(s.dtype()), true});  // :148-161
    for (size_t bi = 0; bi < num_batches; ++bi) {
        const size_t start = bi * batch_size, end = std::min((bi + 1) * batch_size, num_rows);
        std::vector<rvl_column> cols;
        for (const auto& s : df.columns()) {
            // a Float64 series holding AnyValue::Int64 panics in the reference (streaming.rs:189)
            if (s.is_mixed())
                for (size_t i = start; i < end; ++i)
                    if (s.at(i).tag == AnyValue::kInt64) throw Error("Type mismatch in Float64 series", true);
            cols.push_back(s.as_column(start, end - start, /*flatten_nulls=*/true));
        }
        try { batches.push_back(RecordBatch::try_new(ctx, schema, cols)); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Conversion error: ") + e.what()); }
    }
    return batches;
}

// ------------------------------------------------------------------------------------------------- StreamingPhysicalPlan
StreamingPhysicalPlan StreamingPhysicalPlan::memory_source(std::vector<RecordBatch> b) {
    StreamingPhysicalPlan p; p.kind = MemorySource;
    if (!b.empty()) p.ctx = b[0].context();
    p.batches = std::move(b);
    return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::dataframe_source(DataFrame df, size_t batch_size, ContextRef ctx) {
    StreamingPhysicalPlan p; p.kind = DataFrameSource; p.df = std::move(df); p.batch_size = batch_size;
    p.ctx = std::move(ctx);  // resolved lazily (Context::shared(0)) when the plan executes
    return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::csv_file_source(std::string path, SchemaRef schema, std::optional<size_t> batch_size,
                                                             std::optional<std::string> delimiter, ContextRef ctx) {  // streaming.rs:299-311
    StreamingPhysicalPlan p; p.kind = CsvFileSource; p.csv_path = std::move(path); p.csv_schema = std::move(schema);
    p.csv_batch_size = batch_size; p.csv_delimiter = std::move(delimiter); p.ctx = std::move(ctx);
    return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::filter(std::string col) const {
    StreamingPhysicalPlan p; p.kind = Filter; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.predicate_column = std::move(col); p.ctx = ctx; return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::filter_expr(Expr pred) const {
    StreamingPhysicalPlan p; p.kind = FilterExpr; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.predicate = std::move(pred); p.ctx = ctx; return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::select(std::vector<std::string> cols) const {
    StreamingPhysicalPlan p; p.kind = Select; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.columns = std::move(cols); p.ctx = ctx; return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::limit(size_t n_) const {
    StreamingPhysicalPlan p; p.kind = Limit; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.n = n_; p.ctx = ctx; return p;
}

static DataStreamRef build_stream(const StreamingPhysicalPlan& p, size_t min_batch_rows) {  // streaming.rs:70-133
    using K = StreamingPhysicalPlan;
    switch (p.kind) {
        case K::MemorySource: {
            if (p.batches.empty()) throw Error("Invalid operation: Cannot create stream from empty batch list");
            try { return make_memory_stream(p.ctx, p.batches[0].schema(), p.batches); }
            catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream error: ") + e.what()); }
        }
        case K::DataFrameSource: {
            const ContextRef ctx = p.ctx ? p.ctx : Context::shared(0);
            auto b = dataframe_to_batches(ctx, p.df, std::max(p.batch_size, min_batch_rows));
            SchemaRef s = b.empty() ? std::make_shared<Schema>() : b[0].schema();
            return make_memory_stream(ctx, s, std::move(b));
        }
        case K::CsvFileSource: {  // :96-105
            try { return make_csv_file_stream(p.ctx, p.csv_path, p.csv_schema, p.csv_batch_size, p.csv_delimiter); }
            catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Invalid operation: ") + e.what()); }
        }
        case K::Filter: return make_filter_stream(build_stream(*p.input, min_batch_rows), p.predicate_column);
        case K::FilterExpr: return make_filter_expr_stream(build_stream(*p.input, min_batch_rows), p.predicate);
        case K::Select: {
            auto in = build_stream(*p.input, min_batch_rows);
            try { return make_select_stream(std::move(in), p.columns); }
            catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream error: ") + e.what()); }
        }
        case K::Limit: return make_limit_stream(build_stream(*p.input, min_batch_rows), p.n);
        case K::HashJoin: throw Error("not yet implemented: Streaming hash join not yet implemented", true);   // streaming.rs:128-131
    }
    return nullptr;
}
DataStreamRef StreamingPhysicalPlan::execute() const { return build_stream(*this, 0); }

static std::string wrap_stream_err(const std::string& w) {
    // a StreamError raised inside next_batch becomes StreamingExecutionError::Stream via `?` (streaming.rs:12-13)
    if (w.rfind("Stream execution error:", 0) == 0) return "Stream error: " + w;
    return w;
}
// collect() of [Limit] -> [Select] -> [Filter] -> DataFrame source as ONE device pipeline (rvl_stream_*): batches are pushed
// straight from the Series' host buffers through pinned multi-slot staging, H2D overlaps the fused Filter + Select + Limit
// kernel of the previous batch, the running row count stays on the device and a reached LIMIT stops further transfers
// (LimitStream, streaming.rs:269-271).  Rows, order, nulls and schema are those of the operator chain below (execute());
// any plan shape or error case outside the pattern takes that chain, so every reference error surfaces unchanged.
static bool g_stream_fusion = true;
void set_stream_fusion(bool on) { g_stream_fusion = on; }

static std::optional<RecordBatch> try_fused_collect(const StreamingPhysicalPlan& top) {
    using K = StreamingPhysicalPlan;
    if (!g_stream_fusion) return std::nullopt;
    const K* p = &top;
    std::optional<size_t> limit;
    const std::vector<std::string>* sel = nullptr;
    const std::string* pred = nullptr;
    if (p->kind == K::Limit) { if (p->n == 0) return std::nullopt; limit = p->n; p = p->input.get(); }
    if (p->kind == K::Select) { sel = &p->columns; p = p->input.get(); }
    const Expr* cmp = nullptr;   // extension: a single `column <op> literal` predicate runs as the fused comparison operator
    if (p->kind == K::Filter) { pred = &p->predicate_column; p = p->input.get(); }
    else if (p->kind == K::FilterExpr) {
        if (p->predicate.kind != Expr::Binary || p->predicate.op == BinaryOperator::And || p->predicate.op == BinaryOperator::Or) return std::nullopt;
        cmp = &p->predicate; p = p->input.get();
    }
    const bool from_csv = p->kind == K::CsvFileSource;
    if (!from_csv && (p->kind != K::DataFrameSource || p->df.is_empty() || p->batch_size == 0)) return std::nullopt;
    if (!limit && !sel && !pred && !cmp) return std::nullopt;
    const DataFrame& df = p->df;
    auto in_schema = std::make_shared<Schema>();
    std::vector<int32_t> dtypes;
    if (from_csv) {
        if (!p->csv_schema || p->csv_schema->fields.empty()) return std::nullopt;
        *in_schema = *p->csv_schema;
        for (const auto& f : in_schema->fields) {
            if (f.data_type == ExecType::Null) return std::nullopt;
            dtypes.push_back((int32_t)f.data_type);
        }
    } else {
        for (const auto& s : df.columns()) {
            in_schema->fields.push_back(Field{s.name(), exec_type_of(s.dtype()), true});
            dtypes.push_back(rvl_dtype_of(s.dtype()));
            if (s.dtype() == DataType::Null) return std::nullopt;
        }
    }
    const size_t ncols = in_schema->fields.size();
    rvl_predicate rp{};
    rp.mode = RVL_PRED_TRUE;
    if (pred) {
        auto i = in_schema->index_of(*pred);
        if (!i || in_schema->fields[*i].data_type != ExecType::Boolean) return std::nullopt;  // the chain raises the reference's error
        rp.mode = RVL_PRED_BOOL_COLUMN; rp.column = (int32_t)*i;
    }
    std::string lit_keep;
    if (cmp) {
        if (cmp->left->kind != Expr::Column || cmp->right->kind != Expr::Literal) return std::nullopt;
        auto i = in_schema->index_of(cmp->left->name);
        if (!i) return std::nullopt;                       // the chain raises the error
        const AnyValue& lit = cmp->right->value;
        lit_keep = lit.s;
        rp.mode = RVL_PRED_CMP_LITERAL; rp.column = (int32_t)*i; rp.op = (int32_t)cmp->op;
        rp.lit_dtype = rvl_dtype_of(lit.data_type());
        rp.lit_i64 = lit.i; rp.lit_f64 = lit.f; rp.lit_bool = lit.b ? 1 : 0;
        rp.lit_str = (const uint8_t*)lit_keep.data(); rp.lit_str_len = (int64_t)lit_keep.size();
    }
    std::vector<int32_t> proj;
    auto out_schema = std::make_shared<Schema>();
    if (sel) {
        std::set<std::string> seen;
        for (const auto& name : *sel) {
            auto i = in_schema->index_of(name);
            if (!i || !seen.insert(name).second) return std::nullopt;
            proj.push_back((int32_t)*i);
            out_schema->fields.push_back(in_schema->fields[*i]);
        }
        if (proj.empty()) return std::nullopt;
    } else {
        for (size_t i = 0; i < ncols; ++i) proj.push_back((int32_t)i);
        out_schema = in_schema;
    }
    const ContextRef ctx = p->ctx ? p->ctx : Context::shared(0);
    rvl_stream_config cfg{};
    cfg.n_staging = 3; cfg.transfer = RVL_TRANSFER_AUTO;
    struct Closer { rvl_stream* s = nullptr; ~Closer() { if (s) rvl_stream_close(s); } } closer;
    auto open = [&](size_t slot_rows) {
        cfg.batch_rows = (int64_t)slot_rows;
        check(rvl_stream_open(ctx->handle(), dtypes.data(), (int32_t)ncols, &rp, proj.data(), (int32_t)proj.size(),
                              limit ? (int64_t)*limit : -1, &cfg, &closer.s));
    };
    auto finish = [&]() {
        rvl_batch* out = nullptr;
        check(rvl_stream_collect(closer.s, &out));
        return RecordBatch::adopt(ctx, out_schema, out);
    };

    if (from_csv) {
        // CsvFileStream -> [Filter] -> [Select] -> [Limit] (file_stream.rs + stream.rs) as one pipeline: the parser fills reusable host
        // buffers, rvl_stream_push copies them into the stream's pinned staging slot before it returns (so the parser can go on with
        // the next batch at once) and appends small batches to one operator launch; parsing overlaps H2D and the kernels of the
        // batches before.  A reached LIMIT stops the reading (LimitStream pulls nothing further, streaming.rs:269-271).
        std::unique_ptr<CsvBatchReader> reader;
        try { reader = std::make_unique<CsvBatchReader>(p->csv_path, p->csv_schema, p->csv_batch_size, p->csv_delimiter); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Invalid operation: ") + e.what()); }   // streaming.rs:102-103
        // String batches cannot be appended to a group: one launch per batch, slot = batch; otherwise groups of up to 1 Mi rows
        bool has_string = false;
        for (const auto& f : in_schema->fields) has_string |= f.data_type == ExecType::String;
        const size_t bs = std::max<size_t>(reader->batch_size(), 1);
        open(has_string ? bs : std::max(bs, K::kCollectBatchRows));
        static const bool trace = std::getenv("RVL_HOST_TRACE") != nullptr;
        using clk = std::chrono::steady_clock;
        double t_read = 0, t_push = 0;
        auto t_mark = clk::now();
        auto lap = [&](double& acc) { const auto now = clk::now(); acc += std::chrono::duration<double, std::milli>(now - t_mark).count(); t_mark = now; };
        struct Report { const bool on; double& r; double& p; clk::time_point t0; ~Report() { if (on) std::fprintf(stderr, "[host] csv pipeline: read_batch %.1f ms, push %.1f ms, total %.1f ms\n", r, p, std::chrono::duration<double, std::milli>(clk::now() - t0).count()); } } report{trace, t_read, t_push, t_mark};
        for (;;) {
            size_t rows = 0;
            try { rows = reader->read_batch(); lap(t_read); }
            catch (const Error& e) {
                if (e.panic) throw;
                // The reference reads batch j only while fewer than `limit` rows came out of batches < j: an error in a batch it
                // would never have pulled must not surface.  Collect what is in flight and see.
                if (limit) { RecordBatch done = finish(); if (done.num_rows() >= *limit) return done; }
                throw Error(wrap_stream_err(e.what()));
            }
            if (rows == 0) break;
            if (csv_reference_validity()) reader->apply_reference_validity();
            const std::vector<rvl_column> cols = reader->columns();
            int32_t accepted = 0;
            check(rvl_stream_push(closer.s, cols.data(), (int32_t)ncols, &accepted));
            lap(t_push);
            if (!accepted) break;
        }
        return finish();
    }

    // dataframe_to_batches converts every batch before the first one is pulled: a Float64 series holding an Int64 panics (streaming.rs:189)
    for (const auto& s : df.columns())
        if (s.is_mixed())
            for (size_t i = 0; i < s.len(); ++i)
                if (s.at(i).tag == AnyValue::kInt64) throw Error("Type mismatch in Float64 series", true);

    const size_t rows = df.height(), batch = std::max(p->batch_size, K::kCollectBatchRows);
    open(std::min(batch, rows));
    std::vector<rvl_column> cols(ncols);
    for (size_t start = 0; start < rows; start += batch) {
        const size_t len = std::min(batch, rows - start);
        for (size_t c = 0; c < ncols; ++c) cols[c] = df.columns()[c].as_column(start, len, /*flatten_nulls=*/true);
        int32_t accepted = 0;
        check(rvl_stream_push(closer.s, cols.data(), (int32_t)ncols, &accepted));
        if (!accepted) break;  // LIMIT reached on the device: nothing more is transferred
    }
    return finish();
}

RecordBatch StreamingPhysicalPlan::collect() const {  // streaming.rs:235-238
    if (batch_size == 0 && kind == DataFrameSource) throw Error("attempt to divide by zero", true);
    if (auto fused = try_fused_collect(*this)) return *fused;
    auto s = build_stream(*this, kCollectBatchRows);
    try { return collect_stream_batches(ctx ? ctx : Context::shared(0), *s); }
    catch (const Error& e) { if (e.panic) throw; throw Error(wrap_stream_err(e.what())); }
}
std::vector<RecordBatch> StreamingPhysicalPlan::collect_batches() const {  // streaming.rs:240-243
    auto s = execute();
    try { return collect_all_batches(*s); }
    catch (const Error& e) { if (e.panic) throw; throw Error(wrap_stream_err(e.what())); }
}

// ------------------------------------------------------------------------------------------------- logical_plan/plan.rs
using SchemaVec = std::vector<std::pair<std::string, DataType>>;

static std::pair<std::string, DataType> resolve_expr_schema(const Expr& e, const SchemaVec& in) {  // logical_plan/plan.rs:204-262
    switch (e.kind) {
        case Expr::Column:
            for (const auto& p : in) if (p.first == e.name) return {e.name, p.second};
            return {e.name, DataType::Null};
        case Expr::Alias: return {e.name, resolve_expr_schema(*e.left, in).second};
        case Expr::Binary: {
            const auto l = resolve_expr_schema(*e.left, in), r = resolve_expr_schema(*e.right, in);
            DataType res = DataType::Null;
            switch (e.op) {
                case BinaryOperator::Eq: case BinaryOperator::NotEq: case BinaryOperator::Lt: case BinaryOperator::Gt:
                case BinaryOperator::LtEq: case BinaryOperator::GtEq: case BinaryOperator::And: case BinaryOperator::Or:
                    res = DataType::Boolean; break;
                default:
                    if (l.second == DataType::Float64 || r.second == DataType::Float64) res = DataType::Float64;
                    else if (l.second == DataType::Int64 && r.second == DataType::Int64) res = DataType::Int64;
                    else if (l.second == DataType::Null) res = r.second;
                    else if (r.second == DataType::Null) res = l.second;
            }
            return {l.first, res};
        }
        case Expr::Literal: return {"literal", e.value.data_type()};
    }
    return {"", DataType::Null};
}

SchemaVec LogicalPlan::schema() const {  // logical_plan/plan.rs:63-113
    switch (kind) {
        case DataFrameSource: case CsvFileSource: return src_schema;   // :65-66
        case Select: {
            const auto in = input->schema();
            SchemaVec out;
            for (const auto& e : expressions) out.push_back(resolve_expr_schema(e, in));
            return out;
        }
        case Filter: case Limit: return input->schema();
        case Join: {  // :79-111: the left schema, then the right columns except the right key, "_right" on a name the left side has
            const auto left = input->schema();
            SchemaVec out = left;
            for (const auto& rc : right->schema()) {
                if (rc.first == right_key) continue;
                bool clash = false;
                for (const auto& lc : left) clash = clash || lc.first == rc.first;
                out.emplace_back(clash ? rc.first + "_right" : rc.first, rc.second);
            }
            return out;
        }
    }
    return {};
}

bool dtype_is_numeric(DataType d) { return d == DataType::Int64 || d == DataType::Float64; }   // series.rs:136-142
bool dtype_is_comparable_with(DataType a, DataType b) {  // series.rs:144-159
    if (a == b) return true;
    if ((a == DataType::Int64 && b == DataType::Float64) || (a == DataType::Float64 && b == DataType::Int64)) return true;
    return a == DataType::Null || b == DataType::Null;
}
static bool is_comparable_with(DataType a, DataType b) { return dtype_is_comparable_with(a, b); }

static void validate_expr_columns(const Expr& e, const SchemaVec& schema) {  // logical_plan/plan.rs:264-286
    switch (e.kind) {
        case Expr::Column: {
            bool found = false;
            for (const auto& p : schema) found |= p.first == e.name;
            if (!found) throw Error("Logical plan error: Column not found: '" + e.name + "'");
            break;
        }
        case Expr::Binary: validate_expr_columns(*e.left, schema); validate_expr_columns(*e.right, schema); break;
        case Expr::Alias: validate_expr_columns(*e.left, schema); break;
        case Expr::Literal: break;
    }
}

void LogicalPlan::validate() const {  // logical_plan/plan.rs:115-202
    switch (kind) {
        case DataFrameSource:
            for (const auto& p : src_schema)
                if (!df.column(p.first)) throw Error("Logical plan error: Column not found: '" + p.first + "'");
            break;
        case Select: {
            input->validate();
            const auto in = input->schema();
            for (const auto& e : expressions) validate_expr_columns(e, in);
            break;
        }
        case Filter: input->validate(); validate_expr_columns(predicate, input->schema()); break;
        case Limit: input->validate(); break;
        case CsvFileSource: break;   // :128
        case Join: {  // :156-200 — the right side is validated first
            right->validate();
            input->validate();
            const auto ls = input->schema(), rs = right->schema();
            const DataType* lt = nullptr; const DataType* rt = nullptr;
            for (const auto& c : ls) if (c.first == left_key && !lt) lt = &c.second;
            if (!lt) throw Error("Logical plan error: Column not found: '" + left_key + "'");
            for (const auto& c : rs) if (c.first == right_key && !rt) rt = &c.second;
            if (!rt) throw Error("Logical plan error: Column not found: '" + right_key + "'");
            if (!is_comparable_with(*lt, *rt))
                throw Error(std::string("Logical plan error: Incompatible join key types: '") + dtype_name(*lt) + "' and '" + dtype_name(*rt) + "'");
            break;
        }
    }
}

std::string LogicalPlan::shape() const {
    switch (kind) {
        case DataFrameSource: return "Source";
        case CsvFileSource: return "CsvSource";
        case Join: return "Join(" + input->shape() + ", " + right->shape() + ")";
        case Select: return "Select(" + input->shape() + ")";
        case Filter: return "Filter(" + input->shape() + ")";
        case Limit: return "Limit(" + input->shape() + ")";
    }
    return "";
}

std::string LogicalPlan::describe() const {
    switch (kind) {
        case DataFrameSource: return "DataFrameSource";
        case CsvFileSource: return "CsvFileSource { path: \"" + csv_path + "\" }";
        case Join:
            return "Join { left: " + input->describe() + ", right: " + right->describe() + ", left_key: \"" + left_key + "\", right_key: \"" + right_key +
                   "\", join_type: Inner }";
        case Select: {
            std::string e;
            for (size_t i = 0; i < expressions.size(); ++i) e += (i ? ", " : "") + expressions[i].debug();
            return "Select { input: " + input->describe() + ", expressions: [" + e + "] }";
        }
        case Filter: return "Filter { input: " + input->describe() + ", predicate: " + predicate.debug() + " }";
        case Limit: return "Limit { input: " + input->describe() + ", n: " + std::to_string(n) + " }";
    }
    return "";
}

// ------------------------------------------------------------------------------------------------- logical_plan/optimizer.rs
static void extract_column_names(const Expr& e, std::vector<std::string>& out) {  // optimizer.rs:76-87
    switch (e.kind) {
        case Expr::Column: out.push_back(e.name); break;
        case Expr::Binary: extract_column_names(*e.left, out); extract_column_names(*e.right, out); break;
        case Expr::Alias: extract_column_names(*e.left, out); break;
        case Expr::Literal: break;
    }
}
static bool predicate_uses_only_selected_columns(const Expr& pred, const std::vector<Expr>& exprs) {  // optimizer.rs:66-74,89-100
    std::vector<std::string> pc, sc;
    extract_column_names(pred, pc);
    for (const auto& e : exprs) {
        if (e.kind == Expr::Column) sc.push_back(e.name);
        else if (e.kind == Expr::Alias && e.left->kind == Expr::Column) sc.push_back(e.left->name);
    }
    for (const auto& c : pc) if (std::find(sc.begin(), sc.end(), c) == sc.end()) return false;
    return true;
}
LogicalPlan optimize(LogicalPlan plan) {  // optimizer.rs:15-64 (push_predicates_down)
    switch (plan.kind) {
        case LogicalPlan::Select: {
            LogicalPlan& in = *plan.input;
            if (in.kind == LogicalPlan::Filter) {  // Select(Filter(x)) -> Filter(Select(x)) when the predicate only reads selected columns
                if (predicate_uses_only_selected_columns(in.predicate, plan.expressions)) {
                    LogicalPlan sel; sel.kind = LogicalPlan::Select; sel.input = in.input; sel.expressions = plan.expressions;
                    LogicalPlan fil; fil.kind = LogicalPlan::Filter; fil.predicate = in.predicate;
                    fil.input = std::make_shared<LogicalPlan>(std::move(sel));
                    return fil;
                }
                return plan;
            }
            LogicalPlan out = plan;
            out.input = std::make_shared<LogicalPlan>(optimize(in));
            return out;
        }
        case LogicalPlan::Filter: {
            LogicalPlan out = plan;
            out.input = std::make_shared<LogicalPlan>(optimize(*plan.input));
            return out;
        }
        case LogicalPlan::Join: {  // optimizer.rs:50-61
            LogicalPlan out = plan;
            out.input = std::make_shared<LogicalPlan>(optimize(*plan.input));
            out.right = std::make_shared<LogicalPlan>(optimize(*plan.right));
            return out;
        }
        default: return plan;  // Limit and sources are left untouched (optimizer.rs:62)
    }
}

// ------------------------------------------------------------------------------------------------- planner.rs + plan.rs (eager engine, on the GPU)
static std::pair<std::string, std::string> convert_select_expr(const Expr& e) {  // planner.rs:113-132
    switch (e.kind) {
        case Expr::Column: return {e.name, e.name};
        case Expr::Alias:
            if (e.left->kind == Expr::Column) return {e.left->name, e.name};
            throw Error("Unsupported expression: " + quoted(e.debug()));
        case Expr::Binary: throw Error("Unsupported expression: " + quoted(e.debug()));
        case Expr::Literal: throw Error("Select expression must be a column or alias, found: " + quoted(e.debug()));
    }
    return {};
}
struct FilterSpec { std::string column; AnyValue value; BinaryOperator op; };
static FilterSpec convert_filter_predicate(const Expr& p) {  // planner.rs:134-189
    if (p.kind != Expr::Binary) {
        const char* t = p.kind == Expr::Column ? "Column" : (p.kind == Expr::Literal ? "Literal" : "Alias");
        throw Error(std::string("Filter must be a binary comparison, found: ") + t);
    }
    switch (p.op) {
        case BinaryOperator::Eq: case BinaryOperator::NotEq: case BinaryOperator::Lt:
        case BinaryOperator::Gt: case BinaryOperator::LtEq: case BinaryOperator::GtEq: break;
        case BinaryOperator::And: case BinaryOperator::Or:
            throw Error("Unsupported filter: only simple column comparisons supported, found: " + quoted(p.debug()));
        default: throw Error(std::string("Unsupported binary operator in filter: ") + op_name(p.op));
    }
    if (p.left->kind != Expr::Column) throw Error("Filter left side must be a column reference, found: " + quoted(p.left->debug()));
    if (p.right->kind != Expr::Literal) throw Error("Filter right side must be a literal value, found: " + quoted(p.right->debug()));
    return {p.left->name, p.right->value, p.op};
}
// ---- extension: And / Or over comparison leaves (rivulus.hpp: set_extensions)
static bool g_extensions = false;
void set_extensions(bool on) { g_extensions = on; }
bool extensions_enabled() { return g_extensions; }
static bool is_compound(const Expr& p) { return p.kind == Expr::Binary && (p.op == BinaryOperator::And || p.op == BinaryOperator::Or); }
// every leaf must have the shape the reference accepts for a whole predicate (planner.rs:152-186): same checks, same errors
static void check_predicate_tree(const Expr& p) {
    if (is_compound(p)) { check_predicate_tree(*p.left); check_predicate_tree(*p.right); return; }
    convert_filter_predicate(p);
}
static void leaf_columns(const Expr& p, std::vector<std::string>& out) {
    if (is_compound(p)) { leaf_columns(*p.left, out); leaf_columns(*p.right, out); return; }
    out.push_back(p.left->name);
}

static void check_lowering(const LogicalPlan& p) {  // logical_to_physical runs over the whole tree before execution
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return;
        case LogicalPlan::CsvFileSource:   // planner.rs:45-49
            throw Error("Conversion failed: CSV file source not supported in non-streaming physical planner. Use streaming planner instead.");
        case LogicalPlan::Select: check_lowering(*p.input); for (const auto& e : p.expressions) convert_select_expr(e); return;
        case LogicalPlan::Filter:
            check_lowering(*p.input);
            if (g_extensions && is_compound(p.predicate)) check_predicate_tree(p.predicate);
            else convert_filter_predicate(p.predicate);
            return;
        case LogicalPlan::Limit: check_lowering(*p.input); return;
        case LogicalPlan::Join: check_lowering(*p.input); check_lowering(*p.right); return;   // planner.rs:97-98: left, then right
    }
}

namespace {
// A DataFrame living on the device: the eager engine's intermediate result.
struct Frame {
    RecordBatch rb;                    // default-constructed when the frame has no columns; visible columns first, then hidden tags
    std::vector<std::string> names;
    std::vector<DataType> dtypes;      // eager dtype of each column (may be Null while the device array is typed)
    std::vector<int> tag_col;          // per visible column: index in `rb` of its hidden Int64-tag column (mixed series), or -1
    size_t rows = 0;
};

Frame upload_frame(const ContextRef& ctx, const DataFrame& df) {
    Frame f;
    f.rows = df.height();
    if (df.is_empty()) return f;
    auto schema = std::make_shared<Schema>();
    std::vector<rvl_column> cols;
    for (const auto& s : df.columns()) {
        schema->fields.push_back(Field{s.name(), exec_type_of(s.dtype()), true});
        cols.push_back(s.as_column(0, s.len(), /*flatten_nulls=*/false));
        f.names.push_back(s.name());
        f.dtypes.push_back(s.dtype());
        f.tag_col.push_back(-1);
    }
    for (size_t i = 0; i < df.columns().size(); ++i) {  // hidden tag columns of mixed Float64/Int64 series
        const Series& s = df.columns()[i];
        if (!s.is_mixed()) continue;
        f.tag_col[i] = (int)cols.size();
        schema->fields.push_back(Field{"__int64_tag(" + s.name() + ")", ExecType::Boolean, true});
        cols.push_back(s.tag_column(0, s.len()));
    }
    f.rb = RecordBatch::try_new(ctx, schema, cols);
    return f;
}

DataFrame download_frame(const Frame& f) {
    std::vector<Series> out;
    for (size_t i = 0; i < f.names.size(); ++i) {
        ArrayData a = f.rb.column_data(i);
        rvl_column c{};
        c.dtype = (int32_t)a.dtype; c.location = RVL_HOST; c.length = a.length; c.null_count = a.null_count;
        c.validity = a.has_validity ? a.validity.data() : nullptr;
        switch (a.dtype) {
            case ExecType::Int64: c.values = a.i64.data(); break;
            case ExecType::Float64: c.values = a.f64.data(); break;
            case ExecType::Boolean: c.values = a.bits.data(); break;
            case ExecType::String: c.offsets = a.offsets.data(); c.data = a.data.data(); c.data_len = (int64_t)a.data.size(); break;
            default: break;
        }
        if (f.dtypes[i] == DataType::Null && a.length > 0) { c.dtype = RVL_NULL; c.null_count = a.length; }
        if (f.tag_col[i] >= 0 && a.length > 0) {
            ArrayData t = f.rb.column_data((size_t)f.tag_col[i]);
            rvl_column tc{};
            tc.dtype = RVL_BOOLEAN; tc.location = RVL_HOST; tc.length = t.length; tc.values = t.bits.data();
            out.push_back(Series::from_mixed(f.names[i], c, tc, f.dtypes[i]));
            continue;
        }
        out.push_back(Series::from_array(f.names[i], std::move(a), f.dtypes[i]));
    }
    return DataFrame::unchecked(std::move(out));
}

// dtype the reference's Series::new would infer for column i of `rb` holding `rows` (> 0) rows
DataType inferred_dtype(const RecordBatch& rb, size_t i, DataType current, size_t rows, int tag = -1) {
    if (current == DataType::Null) return DataType::Null;
    const size_t nulls = (size_t)rb.column_null_count(i);
    if (nulls == rows) return DataType::Null;
    if (tag >= 0) {  // Int64 + Float64 survivors => Float64; only Int64 survivors => Int64 (series.rs:190-214)
        int64_t ints = 0;
        if (rvl_batch_count_true(rb.context()->handle(), rb.handle(), tag, &ints) != RVL_OK) throw Error(rvl_last_error());
        return (size_t)ints + nulls == rows ? DataType::Int64 : DataType::Float64;
    }
    return current;
}

rvl_predicate to_predicate(int32_t column, BinaryOperator op, const AnyValue& lit) {
    rvl_predicate p{};
    p.mode = RVL_PRED_CMP_LITERAL; p.column = column; p.op = (int32_t)op;
    p.lit_dtype = rvl_dtype_of(lit.data_type());
    p.lit_i64 = lit.i; p.lit_f64 = lit.f; p.lit_bool = lit.b ? 1 : 0;
    p.lit_str = (const uint8_t*)lit.s.data(); p.lit_str_len = (int64_t)lit.s.size();
    return p;
}

// Filter arm (plan.rs:97-150), optionally fused with the Select (:68-96) and Limit (:151-173) arms right above it.
struct BatchGuard {
    rvl_batch* b = nullptr;
    BatchGuard() = default;
    explicit BatchGuard(rvl_batch* x) : b(x) {}
    BatchGuard(const BatchGuard&) = delete;
    BatchGuard& operator=(const BatchGuard&) = delete;
    ~BatchGuard() { if (b) rvl_batch_release(b); }
    rvl_batch* release() { rvl_batch* x = b; b = nullptr; return x; }
};

// leaf `column <op> literal` over frame `in`: the kernel predicate (mixed Float64 / Int64 series compare row by row with their own type)
rvl_predicate leaf_predicate(const Frame& in, const FilterSpec& f) {
    int32_t pc = -1;
    for (size_t i = 0; i < in.names.size(); ++i) if (in.names[i] == f.column) { pc = (int32_t)i; break; }
    if (pc < 0) throw Error("Column not found: '" + f.column + "'");  // plan.rs:104-110
    rvl_predicate pred = to_predicate(pc, f.op, f.value);
    if (in.tag_col[(size_t)pc] >= 0) pred.tag_column = in.tag_col[(size_t)pc] + 1;
    return pred;
}

// extension: selection mask of an And / Or tree — one predicate-kernel launch per leaf, BooleanArray::{and, or} on the device per node
rvl_batch* tree_mask(const ContextRef& ctx, const Frame& in, const Expr& p) {
    if (is_compound(p)) {
        BatchGuard l(tree_mask(ctx, in, *p.left)), r(tree_mask(ctx, in, *p.right));
        rvl_batch* out = nullptr;
        check(rvl_boolean_op(ctx->handle(), p.op == BinaryOperator::And ? RVL_BOOL_AND : RVL_BOOL_OR, l.b, 0, r.b, 0, &out));
        return out;
    }
    const FilterSpec f = convert_filter_predicate(p);
    const rvl_predicate pred = leaf_predicate(in, f);
    rvl_batch* mask = nullptr;
    check(rvl_predicate_mask(ctx->handle(), in.rb.handle(), &pred, &mask));
    return mask;
}

Frame exec_filter(const ContextRef& ctx, const Frame& in, const Expr& predicate, const std::vector<std::pair<std::string, std::string>>* select,
                  std::optional<size_t> limit) {
    // the kernel predicate and the batch it runs over: a comparison against the frame itself, or (extension) the mask of an
    // And / Or tree riding behind the frame's columns in a zero-copy wrapped batch
    rvl_predicate pred{};
    const rvl_batch* src = in.rb.handle();
    BatchGuard mask, wrapped;
    if (g_extensions && is_compound(predicate)) {
        std::vector<std::string> cols;
        leaf_columns(predicate, cols);
        for (const auto& c : cols) {
            bool found = false;
            for (const auto& n : in.names) found |= n == c;
            if (!found) throw Error("Column not found: '" + c + "'");
        }
        mask.b = tree_mask(ctx, in, predicate);
        int32_t ncols = 0;
        check(rvl_batch_num_columns(src, &ncols));
        std::vector<rvl_column> views((size_t)ncols + 1);
        for (int32_t i = 0; i < ncols; ++i) check(rvl_batch_column(src, i, &views[(size_t)i]));
        check(rvl_batch_column(mask.b, 0, &views[(size_t)ncols]));
        check(rvl_batch_wrap_device(ctx->handle(), views.data(), ncols + 1, &wrapped.b));
        pred.mode = RVL_PRED_BOOL_COLUMN; pred.column = ncols;
        src = wrapped.b;
    } else {
        pred = leaf_predicate(in, convert_filter_predicate(predicate));
    }
    // projected columns: every input column (plain Filter) or the Select list, by source name
    std::vector<int32_t> proj;
    std::vector<std::string> out_names;
    if (select) {
        for (const auto& pr : *select) {
            int32_t idx = -1;
            for (size_t i = 0; i < in.names.size(); ++i) if (in.names[i] == pr.first) { idx = (int32_t)i; break; }
            if (idx < 0) throw Error("Column not found: '" + pr.first + "'");  // plan.rs:72-78
            proj.push_back(idx); out_names.push_back(pr.second);
        }
    } else {
        for (size_t i = 0; i < in.names.size(); ++i) { proj.push_back((int32_t)i); out_names.push_back(in.names[i]); }
    }
    // hidden tag columns of the projected mixed columns ride along behind the visible ones
    const size_t nvis = proj.size();
    Frame r;
    r.tag_col.assign(nvis, -1);
    for (size_t j = 0; j < nvis; ++j) {
        const int t = in.tag_col[(size_t)proj[j]];
        if (t >= 0) { r.tag_col[j] = (int)proj.size(); proj.push_back(t); }
    }
    rvl_batch* out = nullptr;
    check(rvl_filter_project(ctx->handle(), src, &pred, proj.data(), (int32_t)proj.size(), limit ? (int64_t)*limit : -1, &out));
    auto schema = std::make_shared<Schema>();
    for (size_t j = 0; j < proj.size(); ++j)
        schema->fields.push_back(Field{j < nvis ? out_names[j] : in.rb.schema()->fields[(size_t)proj[j]].name, in.rb.schema()->fields[(size_t)proj[j]].data_type, true});
    r.rb = RecordBatch::adopt(ctx, schema, out);
    r.rows = r.rb.num_rows();
    r.names = out_names;
    if (r.rows == 0 && (select || limit)) throw Error("Series error: Empty series not allowed");  // Series::new on no data: plan.rs:89-91, :167-169
    for (size_t j = 0; j < nvis; ++j) {
        const DataType cur = in.dtypes[(size_t)proj[j]];
        r.dtypes.push_back(r.rows == 0 ? cur : inferred_dtype(r.rb, j, cur, r.rows, r.tag_col[j]));  // Series::empty keeps the dtype (plan.rs:140-141)
    }
    if (select) {  // DataFrame::new over the renamed series (plan.rs:95): duplicate final names
        std::set<std::string> seen;
        for (const auto& n : r.names) if (!seen.insert(n).second) throw Error("DataFrame error: Duplicate column name: '" + n + "'");
    }
    return r;
}

// extension: selection mask of a predicate tree over a RecordBatch (the streaming engine's arrays: numeric / Boolean nulls were
// already flattened by dataframe_to_batches, strings keep theirs)
rvl_batch* batch_tree_mask(const RecordBatch& b, const Expr& p) {
    const ContextRef& ctx = b.context();
    if (is_compound(p)) {
        BatchGuard l(batch_tree_mask(b, *p.left)), r(batch_tree_mask(b, *p.right));
        rvl_batch* out = nullptr;
        check(rvl_boolean_op(ctx->handle(), p.op == BinaryOperator::And ? RVL_BOOL_AND : RVL_BOOL_OR, l.b, 0, r.b, 0, &out));
        return out;
    }
    const FilterSpec f = convert_filter_predicate(p);
    auto idx = b.schema()->index_of(f.column);
    if (!idx) throw Error("Stream execution error: Column '" + f.column + "' not found in schema");
    const rvl_predicate pred = to_predicate((int32_t)*idx, f.op, f.value);
    rvl_batch* mask = nullptr;
    check(rvl_predicate_mask(ctx->handle(), b.handle(), &pred, &mask));
    return mask;
}

struct FilterExprStream : DataStream {
    DataStreamRef input; Expr pred;
    SchemaRef schema() const override { return input->schema(); }
    std::optional<RecordBatch> next_batch() override {
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        auto ms = std::make_shared<Schema>();
        ms->fields.push_back(Field{"mask", ExecType::Boolean, true});
        RecordBatch mask = RecordBatch::adopt(b->context(), ms, batch_tree_mask(*b, pred));
        try { return b->filter(mask, 0); }
        catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Stream execution error: ") + e.what()); }
    }
};

Frame exec_node(const ContextRef& ctx, const LogicalPlan& p) {  // physical_plan/plan.rs:65-173
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return upload_frame(ctx, p.df);
        case LogicalPlan::CsvFileSource: throw Error("unreachable: rejected by check_lowering");
        case LogicalPlan::Filter: {
            Frame in = exec_node(ctx, *p.input);
            if (in.names.empty()) { std::vector<std::string> c; leaf_columns(p.predicate, c); throw Error("Column not found: '" + c[0] + "'"); }
            return exec_filter(ctx, in, p.predicate, nullptr, std::nullopt);
        }
        case LogicalPlan::Select: {
            std::vector<std::pair<std::string, std::string>> sel;
            for (const auto& e : p.expressions) sel.push_back(convert_select_expr(e));
            if (p.input->kind == LogicalPlan::Filter && !sel.empty()) {  // Select(Filter(x)): one fused launch
                Frame in = exec_node(ctx, *p.input->input);
                if (!in.names.empty()) return exec_filter(ctx, in, p.input->predicate, &sel, std::nullopt);
                std::vector<std::string> c; leaf_columns(p.input->predicate, c);
                throw Error("Column not found: '" + c[0] + "'");
            }
            Frame in = exec_node(ctx, *p.input);
            std::vector<size_t> idx;
            for (const auto& pr : sel) {
                size_t k = in.names.size();
                for (size_t i = 0; i < in.names.size(); ++i) if (in.names[i] == pr.first) { k = i; break; }
                if (k == in.names.size()) throw Error("Column not found: '" + pr.first + "'");
                idx.push_back(k);
            }
            Frame r;
            if (sel.empty()) return r;  // DataFrame::new(vec![]) = empty frame
            if (in.rows == 0) throw Error("Series error: Empty series not allowed");  // plan.rs:89-91
            r.rows = in.rows;
            const size_t nvis = idx.size();
            r.tag_col.assign(nvis, -1);
            std::vector<size_t> all = idx;
            for (size_t j = 0; j < nvis; ++j)
                if (in.tag_col[idx[j]] >= 0) { r.tag_col[j] = (int)all.size(); all.push_back((size_t)in.tag_col[idx[j]]); }
            r.rb = in.rb.select_columns(all);
            auto schema = std::make_shared<Schema>();
            std::set<std::string> seen;
            for (size_t j = 0; j < nvis; ++j) {
                r.names.push_back(sel[j].second);
                r.dtypes.push_back(inferred_dtype(r.rb, j, in.dtypes[idx[j]], r.rows, r.tag_col[j]));
                schema->fields.push_back(Field{sel[j].second, in.rb.schema()->fields[idx[j]].data_type, true});
            }
            for (size_t j = nvis; j < all.size(); ++j) schema->fields.push_back(in.rb.schema()->fields[all[j]]);
            for (const auto& n : r.names) if (!seen.insert(n).second) throw Error("DataFrame error: Duplicate column name: '" + n + "'");
            r.rb = r.rb.with_schema(schema);
            return r;
        }
        case LogicalPlan::Limit: {
            // Limit(Filter) / Limit(Select(Filter)) with n > 0: the limit goes into the fused launch (early termination)
            if (p.n > 0) {
                const LogicalPlan* c = p.input.get();
                if (c->kind == LogicalPlan::Filter) {
                    Frame in = exec_node(ctx, *c->input);
                    if (in.names.empty()) { std::vector<std::string> cn; leaf_columns(c->predicate, cn); throw Error("Column not found: '" + cn[0] + "'"); }
                    return exec_filter(ctx, in, c->predicate, nullptr, p.n);
                }
                if (c->kind == LogicalPlan::Select && c->input->kind == LogicalPlan::Filter && !c->expressions.empty()) {
                    std::vector<std::pair<std::string, std::string>> sel;
                    for (const auto& e : c->expressions) sel.push_back(convert_select_expr(e));
                    Frame in = exec_node(ctx, *c->input->input);
                    if (in.names.empty()) { std::vector<std::string> cn; leaf_columns(c->input->predicate, cn); throw Error("Column not found: '" + cn[0] + "'"); }
                    return exec_filter(ctx, in, c->input->predicate, &sel, p.n);
                }
            }
            Frame in = exec_node(ctx, *p.input);
            if (in.names.empty()) return in;  // width 0: DataFrame::new(vec![])
            Frame r;
            r.names = in.names;
            r.tag_col = in.tag_col;
            if (p.n == 0) {  // plan.rs:154-161: empty series keep their dtypes
                r.rows = 0; r.dtypes = in.dtypes; r.rb = in.rb.slice(0, 0);
                return r;
            }
            if (in.rows == 0) throw Error("Series error: Empty series not allowed");  // plan.rs:167-169
            const size_t lim = std::min(p.n, in.rows);
            r.rows = lim;
            r.rb = in.rb.slice(0, lim);
            for (size_t j = 0; j < in.names.size(); ++j) r.dtypes.push_back(inferred_dtype(r.rb, j, in.dtypes[j], lim, in.tag_col[j]));
            return r;
        }
        case LogicalPlan::Join: {  // HashJoin, plan.rs:174-284: build = left, probe = right (planner.rs:100-108)
            const Frame build = exec_node(ctx, *p.input);
            int bkey = -1, pkey = -1;
            for (size_t i = 0; i < build.names.size(); ++i) if (build.names[i] == p.left_key) { bkey = (int)i; break; }
            if (bkey < 0) throw Error("called `Option::unwrap()` on a `None` value", true);   // :183
            const Frame probe = exec_node(ctx, *p.right);
            for (size_t i = 0; i < probe.names.size(); ++i) if (probe.names[i] == p.right_key) { pkey = (int)i; break; }
            if (pkey < 0) throw Error("called `Option::unwrap()` on a `None` value", true);   // :196
            // device columns asked of each side: the visible ones (the build side without its key), then their hidden tag columns
            std::vector<int32_t> pproj, bproj;
            std::vector<size_t> pvis, bvis;
            for (size_t i = 0; i < probe.names.size(); ++i) { pvis.push_back(i); pproj.push_back((int32_t)i); }
            for (size_t i = 0; i < build.names.size(); ++i) if ((int)i != bkey) { bvis.push_back(i); bproj.push_back((int32_t)i); }
            std::vector<int> ptag(pvis.size(), -1), btag(bvis.size(), -1);   // position of the tag column inside pproj / bproj
            for (size_t j = 0; j < pvis.size(); ++j) if (probe.tag_col[pvis[j]] >= 0) { ptag[j] = (int)pproj.size(); pproj.push_back(probe.tag_col[pvis[j]]); }
            for (size_t j = 0; j < bvis.size(); ++j) if (build.tag_col[bvis[j]] >= 0) { btag[j] = (int)bproj.size(); bproj.push_back(build.tag_col[bvis[j]]); }
            rvl_batch* out = nullptr;
            int64_t n_pairs = 0;
            check(rvl_hash_join_inner(ctx->handle(), build.rb.handle(), bkey, build.tag_col[(size_t)bkey] >= 0 ? build.tag_col[(size_t)bkey] + 1 : 0,
                                      probe.rb.handle(), pkey, probe.tag_col[(size_t)pkey] >= 0 ? probe.tag_col[(size_t)pkey] + 1 : 0,
                                      pproj.data(), (int32_t)pproj.size(), bproj.data(), (int32_t)bproj.size(), &out, &n_pairs));
            // joined batch = [probe visible | probe tags | build visible | build tags]  ->  frame order [visible ... | tags ...]
            auto jschema = std::make_shared<Schema>();
            for (int32_t c : pproj) jschema->fields.push_back(probe.rb.schema()->fields[(size_t)c]);
            for (int32_t c : bproj) jschema->fields.push_back(build.rb.schema()->fields[(size_t)c]);
            RecordBatch joined = RecordBatch::adopt(ctx, jschema, out);
            Frame r;
            r.rows = (size_t)n_pairs;
            std::vector<size_t> order;
            std::vector<std::pair<size_t, int>> tags;   // (visible position, column of `joined`)
            const size_t pbase = 0, bbase = pproj.size();
            for (size_t j = 0; j < pvis.size(); ++j) {
                r.names.push_back(probe.names[pvis[j]]); r.dtypes.push_back(probe.dtypes[pvis[j]]);
                order.push_back(pbase + j);
                if (ptag[j] >= 0) tags.emplace_back(r.names.size() - 1, (int)(pbase + (size_t)ptag[j]));
            }
            for (size_t j = 0; j < bvis.size(); ++j) {
                bool clash = false;   // :237-241: "_right" on a build column whose name the probe side has
                for (const auto& n : probe.names) clash = clash || n == build.names[bvis[j]];
                r.names.push_back(clash ? build.names[bvis[j]] + "_right" : build.names[bvis[j]]); r.dtypes.push_back(build.dtypes[bvis[j]]);
                order.push_back(bbase + j);
                if (btag[j] >= 0) tags.emplace_back(r.names.size() - 1, (int)(bbase + (size_t)btag[j]));
            }
            r.tag_col.assign(r.names.size(), -1);
            for (const auto& t : tags) { r.tag_col[t.first] = (int)order.size(); order.push_back((size_t)t.second); }
            r.rb = joined.select_columns(order);
            auto schema = std::make_shared<Schema>();
            for (size_t j = 0; j < order.size(); ++j) {
                Field f = jschema->fields[order[j]];
                if (j < r.names.size()) f.name = r.names[j];
                schema->fields.push_back(f);
            }
            r.rb = r.rb.with_schema(schema);
            // Series::new re-infers every dtype from the gathered values (:225, :243); no pairs: Series::empty keeps them (:256-283)
            if (n_pairs > 0)
                for (size_t j = 0; j < r.names.size(); ++j) r.dtypes[j] = inferred_dtype(r.rb, j, r.dtypes[j], r.rows, r.tag_col[j]);
            std::set<std::string> seen;   // DataFrame::new (:253)
            for (const auto& n : r.names) if (!seen.insert(n).second) throw Error("DataFrame error: Duplicate column name: '" + n + "'");
            return r;
        }
    }
    return Frame();
}
}  // namespace

DataStreamRef make_filter_expr_stream(DataStreamRef in, Expr predicate) {
    auto st = std::make_unique<FilterExprStream>(); st->input = std::move(in); st->pred = std::move(predicate); return st;
}

DataFrame execute_eager(const LogicalPlan& optimized, const ContextRef& ctx) {
    check_lowering(optimized);  // logical_to_physical fails before anything executes (builder.rs:99-100)
    const ContextRef c = ctx ? ctx : Context::shared(0);
    return download_frame(exec_node(c, optimized));
}

// ------------------------------------------------------------------------------------------------- streaming_planner.rs
StreamingPhysicalPlan logical_to_streaming(const LogicalPlan& plan, const ContextRef& ctx) {  // streaming_planner.rs:29-100
    switch (plan.kind) {
        case LogicalPlan::DataFrameSource: return StreamingPhysicalPlan::dataframe_source(plan.df, 1024, ctx);  // :31-33
        case LogicalPlan::CsvFileSource: {  // :35-62 — every field nullable
            auto schema = std::make_shared<Schema>();
            for (const auto& f : plan.src_schema) schema->fields.push_back(Field{f.first, exec_type_of(f.second), true});
            return StreamingPhysicalPlan::csv_file_source(plan.csv_path, schema, plan.csv_batch_size, plan.csv_delimiter, ctx);
        }
        case LogicalPlan::Select: {  // :65-69, 102-135
            auto in = logical_to_streaming(*plan.input, ctx);
            std::vector<std::string> names;
            for (const auto& e : plan.expressions) {
                if (e.kind == Expr::Column) names.push_back(e.name);
                else if (e.kind == Expr::Alias) {
                    if (e.left->kind == Expr::Column) names.push_back(e.left->name);  // alias dropped (:110-113)
                    else throw Error("Streaming planner error: Expression conversion error: Complex expressions with aliases not yet supported: " + e.debug());
                } else
                    throw Error("Streaming planner error: Expression conversion error: Complex expressions not yet supported in streaming mode: " + e.debug());
            }
            return in.select(names);
        }
        case LogicalPlan::Filter: {  // :71-75, 137-168
            auto in = logical_to_streaming(*plan.input, ctx);
            const Expr& p = plan.predicate;
            if (p.kind == Expr::Column) return in.filter(p.name);
            if (p.kind == Expr::Binary && g_extensions) {
                // extension: comparison leaves and And / Or over them; a malformed leaf raises the eager planner's message
                try { check_predicate_tree(p); }
                catch (const Error& e) { throw Error(std::string("Streaming planner error: Expression conversion error: ") + e.what()); }
                return in.filter_expr(p);
            }
            if (p.kind == Expr::Binary) {
                if (p.left->kind == Expr::Column)
                    throw Error("Streaming planner error: Expression conversion error: Binary expressions not yet supported in streaming mode. "
                                "Found expression on column '" + p.left->name + "'. "
                                "Currently only simple boolean column references are supported (e.g., .filter(col('is_active')))");
                throw Error("Streaming planner error: Expression conversion error: Complex binary expressions not supported in streaming mode");
            }
            throw Error("Streaming planner error: Expression conversion error: Unsupported filter expression type: " + p.debug());
        }
        case LogicalPlan::Limit: return logical_to_streaming(*plan.input, ctx).limit(plan.n);  // :76-79
        case LogicalPlan::Join: {  // :81-98; executing the node is `todo!()` in the reference (streaming.rs:128-131)
            (void)logical_to_streaming(*plan.input, ctx); (void)logical_to_streaming(*plan.right, ctx);
            StreamingPhysicalPlan sp; sp.kind = StreamingPhysicalPlan::HashJoin; sp.ctx = ctx;
            return sp;
        }
    }
    return StreamingPhysicalPlan();
}

// ------------------------------------------------------------------------------------------------- logical_plan/builder.rs
LazyFrame LazyFrame::from_dataframe(const DataFrame& df, ContextRef ctx) {  // builder.rs:27-39
    LazyFrame lf;
    lf.plan_.kind = LogicalPlan::DataFrameSource;
    lf.plan_.df = df;
    for (const auto& s : df.columns()) lf.plan_.src_schema.emplace_back(s.name(), s.dtype());
    lf.ctx_ = std::move(ctx);
    return lf;
}
LazyFrame LazyFrame::from_csv(std::string path, std::vector<std::pair<std::string, DataType>> schema, std::optional<size_t> batch_size,
                              std::optional<std::string> delimiter, ContextRef ctx) {  // builder.rs:41-55
    LazyFrame lf;
    lf.plan_.kind = LogicalPlan::CsvFileSource;
    lf.plan_.csv_path = std::move(path); lf.plan_.src_schema = std::move(schema);
    lf.plan_.csv_batch_size = batch_size; lf.plan_.csv_delimiter = std::move(delimiter);
    lf.ctx_ = std::move(ctx);
    return lf;
}
LazyFrame LazyFrame::select(std::vector<Expr> e) const {
    LazyFrame lf; lf.ctx_ = ctx_; lf.plan_.kind = LogicalPlan::Select; lf.plan_.input = std::make_shared<LogicalPlan>(plan_); lf.plan_.expressions = std::move(e); return lf;
}
LazyFrame LazyFrame::filter(Expr p) const {
    LazyFrame lf; lf.ctx_ = ctx_; lf.plan_.kind = LogicalPlan::Filter; lf.plan_.input = std::make_shared<LogicalPlan>(plan_); lf.plan_.predicate = std::move(p); return lf;
}
LazyFrame LazyFrame::limit(size_t n) const {
    LazyFrame lf; lf.ctx_ = ctx_; lf.plan_.kind = LogicalPlan::Limit; lf.plan_.input = std::make_shared<LogicalPlan>(plan_); lf.plan_.n = n; return lf;
}

LazyFrame LazyFrame::inner_join(const LazyFrame& right, std::string left_key, std::string right_key) const {  // builder.rs:84-94
    LazyFrame lf; lf.ctx_ = ctx_; lf.plan_.kind = LogicalPlan::Join; lf.plan_.input = std::make_shared<LogicalPlan>(plan_);
    lf.plan_.right = std::make_shared<LogicalPlan>(right.plan_); lf.plan_.left_key = std::move(left_key); lf.plan_.right_key = std::move(right_key);
    return lf;
}

DataFrame LazyFrame::collect() const {  // builder.rs:96-104
    LogicalPlan opt = optimize(plan_);
    opt.validate();  // "Logical plan error: …" passes through typed (builder.rs:98)
    try { return execute_eager(opt, ctx_); }
    catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Execution error: ") + e.what()); }
}

RecordBatch LazyFrame::collect_streaming() const {  // builder.rs:106-113
    LogicalPlan opt = optimize(plan_);
    opt.validate();
    StreamingPhysicalPlan sp = logical_to_streaming(opt, ctx_);  // "Streaming planner error: …"
    try { return sp.collect(); }
    catch (const Error& e) { if (e.panic) throw; throw Error(std::string("Execution error: ") + e.what()); }
}

}  // namespace rivulus
