// rivulus_oracle.cpp — TEST INFRASTRUCTURE ONLY (see rivulus_oracle.hpp header comment).
// CPU restatement of the reference's filter/project/limit path; every function cites the
// reference file:line (relative to /root/reference/src) it follows.
#include "rivulus_oracle.hpp"

#include <map>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <set>
#include <sstream>

namespace orc {

// ------------------------------------------------------------------------------------------
// datatypes/series.rs
// ------------------------------------------------------------------------------------------

const char* dtype_name(DataType d) {  // series.rs:162-172
    switch (d) {
        case DataType::Int64: return "Int64";
        case DataType::Float64: return "Float64";
        case DataType::String: return "String";
        case DataType::Boolean: return "Boolean";
        case DataType::Null: return "Null";
    }
    return "?";
}

DataType AnyValue::data_type() const {  // series.rs:20-28
    switch (tag) {
        case kNull: return DataType::Null;
        case kInt64: return DataType::Int64;
        case kFloat64: return DataType::Float64;
        case kString: return DataType::String;
        case kBoolean: return DataType::Boolean;
    }
    return DataType::Null;
}

static std::string f64_display(double v, bool debug) {
    // Rust `{}` / `{:?}` for f64: shortest round-trip; Debug (and Display for integral values
    // below 1e16) keep a trailing ".0".
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    std::string s(buf, r.ptr);
    // to_chars(fixed) on shortest repr never prints exponent; make sure of a fractional part for Debug
    if (s.find('.') == std::string::npos && debug) s += ".0";
    return s;
}

std::string AnyValue::display() const {  // series.rs:61-71
    switch (tag) {
        case kNull: return "null";
        case kInt64: return std::to_string(i);
        case kFloat64: return f64_display(f, false);
        case kString: return s;
        case kBoolean: return b ? "true" : "false";
    }
    return "";
}

static std::string str_debug(const std::string& s) {
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

std::string AnyValue::debug() const {
    switch (tag) {
        case kNull: return "Null";
        case kInt64: return "Int64(" + std::to_string(i) + ")";
        case kFloat64: return "Float64(" + f64_display(f, true) + ")";
        case kString: return "String(" + str_debug(s) + ")";
        case kBoolean: return std::string("Boolean(") + (b ? "true" : "false") + ")";
    }
    return "";
}

bool any_eq(const AnyValue& a, const AnyValue& b) {  // series.rs:87-98
    if (a.tag == AnyValue::kNull && b.tag == AnyValue::kNull) return true;
    if (a.tag == AnyValue::kInt64 && b.tag == AnyValue::kInt64) return a.i == b.i;
    if (a.tag == AnyValue::kFloat64 && b.tag == AnyValue::kFloat64) return a.f == b.f;
    if (a.tag == AnyValue::kString && b.tag == AnyValue::kString) return a.s == b.s;
    if (a.tag == AnyValue::kBoolean && b.tag == AnyValue::kBoolean) return a.b == b.b;
    return false;
}

std::optional<int> any_partial_cmp(const AnyValue& a, const AnyValue& b) {  // series.rs:100-117
    using T = AnyValue;
    if (a.tag == T::kNull && b.tag == T::kNull) return 0;   // :105
    if (a.tag == T::kNull) return -1;                       // :106
    if (b.tag == T::kNull) return 1;                        // :107
    if (a.tag == T::kInt64 && b.tag == T::kInt64) return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);  // :109
    if (a.tag == T::kFloat64 && b.tag == T::kFloat64) {     // :110 (f64::partial_cmp: NaN -> None)
        if (a.f < b.f) return -1;
        if (a.f > b.f) return 1;
        if (a.f == b.f) return 0;
        return std::nullopt;
    }
    if (a.tag == T::kString && b.tag == T::kString) {       // :111 (str ordering = bytewise lexicographic)
        size_t n = std::min(a.s.size(), b.s.size());
        int c = n ? std::memcmp(a.s.data(), b.s.data(), n) : 0;
        if (c != 0) return c < 0 ? -1 : 1;
        return a.s.size() < b.s.size() ? -1 : (a.s.size() > b.s.size() ? 1 : 0);
    }
    if (a.tag == T::kBoolean && b.tag == T::kBoolean) return (int)a.b - (int)b.b;  // :112 (false < true)
    return std::nullopt;                                    // :114
}

static bool are_types_compatible(DataType e, DataType f) {  // series.rs:255-264
    if (e == f) return true;
    return (e == DataType::Int64 && f == DataType::Float64) || (e == DataType::Float64 && f == DataType::Int64);
}

Series Series::make(const std::string& name, std::vector<AnyValue> data) {  // series.rs:185-221
    if (data.empty()) throw OracleError("Empty series not allowed");  // :186-188, text :178
    std::optional<DataType> first;
    for (const auto& v : data)                                         // :191-196
        if (!v.is_null()) { first = v.data_type(); break; }
    DataType dtype = first.value_or(DataType::Null);                   // :198
    for (const auto& v : data) {                                       // :200-214
        if (!v.is_null()) {
            DataType cur = v.data_type();
            if (!are_types_compatible(dtype, cur))
                throw OracleError(std::string("Mixed types in series: expected ") + dtype_name(dtype) +
                                  ", found " + dtype_name(cur));       // :176
            if (dtype == DataType::Int64 && cur == DataType::Float64) dtype = DataType::Float64;  // :210-212
        }
    }
    Series s; s.name_ = name; s.data_ = std::move(data); s.dtype_ = dtype;
    return s;
}

Series Series::empty(const std::string& name, DataType dtype) {  // series.rs:223-229
    Series s; s.name_ = name; s.dtype_ = dtype; return s;
}

const AnyValue& Series::at(size_t i) const {  // series.rs:273-288
    if (i >= data_.size())
        throw Panic("Index " + std::to_string(i) + " out of bounds for series of length " + std::to_string(data_.size()));
    return data_[i];
}

// ------------------------------------------------------------------------------------------
// datatypes/dataframe.rs
// ------------------------------------------------------------------------------------------

DataFrame DataFrame::make(std::vector<Series> columns) {  // dataframe.rs:29-56
    DataFrame df;
    if (columns.empty()) return df;                        // :30-34
    std::set<std::string> seen;                            // :36-41
    for (const auto& c : columns)
        if (!seen.insert(c.name()).second) throw OracleError("Duplicate column name: '" + c.name() + "'");
    size_t expected = columns.front().len();               // :43-53
    for (const auto& c : columns)
        if (c.len() != expected)
            throw OracleError("Column lengths mismatch: expected " + std::to_string(expected) + ", found " +
                              std::to_string(c.len()) + " for column '" + c.name() + "'");
    df.columns_ = std::move(columns);
    return df;
}

const Series* DataFrame::column(const std::string& name) const {  // dataframe.rs:84-86
    for (const auto& s : columns_) if (s.name() == name) return &s;
    return nullptr;
}

DataFrame DataFrame::select(const std::vector<std::string>& names) const {  // dataframe.rs:96-110
    std::vector<Series> cols;
    for (const auto& n : names) {
        const Series* s = column(n);
        if (!s) throw OracleError("Column not found: '" + n + "'");
        cols.push_back(*s);  // clone
    }
    return DataFrame::unchecked(std::move(cols));  // no duplicate/length re-check (:109)
}

// ------------------------------------------------------------------------------------------
// expressions/expr.rs
// ------------------------------------------------------------------------------------------

const char* op_name(BinaryOperator op) {
    static const char* n[] = {"Plus", "Minus", "Multiply", "Divide", "Eq", "NotEq", "Lt", "Gt", "LtEq", "GtEq", "And", "Or"};
    return n[(int)op];
}

std::string Expr::debug() const {  // #[derive(Debug)] on expr.rs:3-13
    switch (kind) {
        case Column: return "Column(" + str_debug(name) + ")";
        case Literal: return "Literal(" + value.debug() + ")";
        case Alias: return "Alias(" + left->debug() + ", " + str_debug(name) + ")";
        case Binary:
            return "BinaryExpr { left: " + left->debug() + ", op: " + op_name(op) + ", right: " + right->debug() + " }";
    }
    return "";
}

// ------------------------------------------------------------------------------------------
// execution/schema.rs + arrays
// ------------------------------------------------------------------------------------------

const char* exec_type_name(ExecType t) {
    static const char* n[] = {"Null", "Boolean", "Int64", "Float64", "String"};
    return n[(int)t];
}

std::optional<size_t> Schema::index_of(const std::string& n) const {  // schema.rs:66-68
    for (size_t i = 0; i < fields.size(); ++i) if (fields[i].name == n) return i;
    return std::nullopt;
}

BitMap BitMap::zeros(size_t n) {  // bitmap.rs:11-19
    BitMap b; b.buffer = std::make_shared<std::vector<uint8_t>>((n + 7) / 8, 0); b.bit_count = n; return b;
}

BitMap BitMap::all_true(size_t n) {  // bitmap.rs:21-38
    BitMap b; b.buffer = std::make_shared<std::vector<uint8_t>>((n + 7) / 8, 0xFF); b.bit_count = n;
    if (n % 8 != 0 && !b.buffer->empty()) b.buffer->back() = (uint8_t)((1u << (n % 8)) - 1);
    return b;
}

BitMap BitMap::from_bools(const std::vector<bool>& v) {  // bitmap.rs:44-59
    BitMap b = zeros(v.size());
    for (size_t i = 0; i < v.size(); ++i) (*b.buffer)[i / 8] |= (uint8_t)((v[i] ? 1 : 0) << (i % 8));
    return b;
}

bool BitMap::get_bit(size_t index) const {  // bitmap.rs:61-68
    if (!(index < bit_count)) throw Panic("assertion failed: index < self.bit_count");
    size_t i = (index + offset) / 8, j = (index + offset) % 8;
    return (((*buffer)[i] >> j) & 1) != 0;
}

size_t BitMap::count(uint8_t value, size_t off, size_t len) const {  // bitmap.rs:74-86
    size_t c = 0;
    for (size_t i = off; i < off + len; ++i)
        if ((((*buffer)[i / 8] >> (i % 8)) & 1) == value) ++c;
    return c;
}

BitMap BitMap::slice(size_t off, size_t len) const {  // bitmap.rs:104-112
    if (!(off + len <= bit_count)) throw Panic("assertion failed: offset + length <= self.bit_count");
    BitMap b; b.buffer = buffer; b.bit_count = len; b.offset = off + offset; return b;
}

void BitmapBuilder::append(bool v) {  // bitmap.rs:142-155
    if (v) current_byte |= (uint8_t)(1u << current_bit_pos);
    ++current_bit_pos; ++bit_count;
    if (current_bit_pos == 8) { buffer.push_back(current_byte); current_byte = 0; current_bit_pos = 0; }
}

bool BitmapBuilder::has_nulls() const {  // bitmap.rs:157-176
    if (bit_count == 0) return false;
    for (uint8_t byte : buffer) if (byte != 0xFF) return true;
    if (current_bit_pos > 0) {
        uint8_t expected = (uint8_t)((1u << current_bit_pos) - 1);
        if (current_byte != expected) return true;
    }
    return false;
}

BitMap BitmapBuilder::finish() {  // bitmap.rs:178-188
    if (current_bit_pos > 0) buffer.push_back(current_byte);
    BitMap b; b.buffer = std::make_shared<std::vector<uint8_t>>(std::move(buffer)); b.bit_count = bit_count; b.offset = 0;
    return b;
}

template <typename T>
std::shared_ptr<PrimitiveArray<T>> PrimitiveArray<T>::make(std::vector<T> v, std::optional<std::vector<bool>> validity) {
    auto a = std::make_shared<PrimitiveArray<T>>();  // primitive.rs:31-42
    a->length = v.size();
    if (validity) a->null_bitmap = BitMap::from_bools(*validity);
    a->values = std::make_shared<std::vector<T>>(std::move(v));
    return a;
}
template <typename T> std::optional<T> PrimitiveArray<T>::value(size_t index) const {  // primitive.rs:48-60
    if (!(index < length)) throw Panic("Index " + std::to_string(index) + " out of bounds");
    size_t li = offset + index;
    if (null_bitmap && !null_bitmap->get_bit(li)) return std::nullopt;
    return (*values)[li];
}
template <> ExecType PrimitiveArray<int64_t>::data_type() const { return ExecType::Int64; }
template <> ExecType PrimitiveArray<double>::data_type() const { return ExecType::Float64; }
template <typename T> size_t PrimitiveArray<T>::null_count() const {  // primitive.rs:91-105
    return null_bitmap ? null_bitmap->count(0, offset, length) : 0;
}
template <typename T> ArrayRef PrimitiveArray<T>::slice(size_t off, size_t len) const {  // primitive.rs:107-117
    if (!(off + len <= length)) throw Panic("assertion failed: offset + length <= self.length");
    auto a = std::make_shared<PrimitiveArray<T>>(*this);
    a->offset = offset + off; a->length = len;
    return a;
}
template <typename T> std::shared_ptr<PrimitiveArray<T>> PrimitiveArrayBuilder<T>::finish() {  // primitive.rs:180-197
    auto a = std::make_shared<PrimitiveArray<T>>();
    if (null_builder.has_nulls()) a->null_bitmap = null_builder.finish();
    a->length = values.size();
    a->values = std::make_shared<std::vector<T>>(std::move(values));
    return a;
}
template struct PrimitiveArray<int64_t>;
template struct PrimitiveArray<double>;
template struct PrimitiveArrayBuilder<int64_t>;
template struct PrimitiveArrayBuilder<double>;

std::shared_ptr<BooleanArray> BooleanArray::make(const std::vector<std::optional<bool>>& v) {  // boolean.rs:19-50
    BitmapBuilder vb, nb;
    for (const auto& o : v) {
        if (o) { vb.append(*o); nb.append(true); }
        else { vb.append(false); nb.append(false); }
    }
    auto a = std::make_shared<BooleanArray>();
    a->values = vb.finish();
    if (nb.has_nulls()) a->null_bitmap = nb.finish();
    a->length = v.size();
    return a;
}
std::optional<bool> BooleanArray::value(size_t index) const {  // boolean.rs:91-103
    if (!(index < length)) throw Panic("Index " + std::to_string(index) + " out of bounds");
    size_t li = offset + index;
    if (null_bitmap && !null_bitmap->get_bit(li)) return std::nullopt;
    return values.get_bit(li);
}
size_t BooleanArray::null_count() const { return null_bitmap ? null_bitmap->count(0, offset, length) : 0; }  // :191-205
ArrayRef BooleanArray::slice(size_t off, size_t len) const {  // boolean.rs:207-217
    if (!(off + len <= length)) throw Panic("Slice out of bounds");
    auto a = std::make_shared<BooleanArray>(*this);
    a->offset = offset + off; a->length = len;
    return a;
}

static bool utf8_valid(const uint8_t* p, size_t n) {  // std::str::from_utf8 (string.rs:141)
    size_t i = 0;
    while (i < n) {
        uint8_t c = p[i];
        if (c < 0x80) { ++i; continue; }
        size_t need; uint32_t cp;
        if ((c & 0xE0) == 0xC0) { need = 1; cp = c & 0x1F; if (c < 0xC2) return false; }
        else if ((c & 0xF0) == 0xE0) { need = 2; cp = c & 0x0F; }
        else if ((c & 0xF8) == 0xF0) { need = 3; cp = c & 0x07; if (c > 0xF4) return false; }
        else return false;
        if (i + need >= n) return false;  // truncated sequence
        for (size_t k = 1; k <= need; ++k) {
            uint8_t cc = p[i + k];
            if ((cc & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (cc & 0x3F);
        }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += need + 1;
    }
    return true;
}

static void validate_utf8(const std::vector<uint8_t>& data, const std::vector<int32_t>& offsets) {  // string.rs:131-146
    for (size_t w = 0; w + 1 < offsets.size(); ++w) {
        size_t start = (size_t)offsets[w], end = (size_t)offsets[w + 1];
        if (end > data.size()) throw Panic("Invalid UTF-8 in string data: Offset out of bounds");
        if (!utf8_valid(data.data() + start, end - start)) throw Panic("Invalid UTF-8 in string data: Invalid UTF-8 sequence");
    }
}

std::shared_ptr<StringArray> StringArray::make(const std::vector<std::optional<std::string>>& v) {  // string.rs:19-58
    std::vector<int32_t> offsets; offsets.reserve(v.size() + 1);
    std::vector<uint8_t> data;
    BitmapBuilder nb;
    offsets.push_back(0);
    for (const auto& o : v) {
        if (o) { nb.append(true); data.insert(data.end(), o->begin(), o->end()); offsets.push_back((int32_t)data.size()); }
        else { nb.append(false); offsets.push_back((int32_t)data.size()); }
    }
    auto a = std::make_shared<StringArray>();
    if (nb.has_nulls()) a->null_bitmap = nb.finish();
    a->length = v.size();
    validate_utf8(data, offsets);  // :55
    a->offsets = std::make_shared<std::vector<int32_t>>(std::move(offsets));
    a->data = std::make_shared<std::vector<uint8_t>>(std::move(data));
    return a;
}
std::optional<std::string> StringArray::value(size_t index) const {  // string.rs:80-97
    if (!(index < length)) throw Panic("Index " + std::to_string(index) + " out of bounds");
    size_t li = offset + index;
    if (null_bitmap && !null_bitmap->get_bit(li)) return std::nullopt;
    size_t s = (size_t)(*offsets)[li], e = (size_t)(*offsets)[li + 1];
    return std::string((const char*)data->data() + s, e - s);
}
size_t StringArray::null_count() const { return null_bitmap ? null_bitmap->count(0, offset, length) : 0; }  // :158-172
ArrayRef StringArray::slice(size_t off, size_t len) const {  // string.rs:174-185
    if (!(off + len <= length)) throw Panic("Slice out of bounds");
    auto a = std::make_shared<StringArray>(*this);
    a->offset = offset + off; a->length = len;
    return a;
}

ArrayRef NullArray::slice(size_t off, size_t len) const {  // null.rs:55-62
    if (!(off + len <= length)) throw Panic("Slice out of bounds");
    auto a = std::make_shared<NullArray>(len); a->offset = offset + off; return a;
}

// ------------------------------------------------------------------------------------------
// execution/record_batch.rs
// ------------------------------------------------------------------------------------------

RecordBatch RecordBatch::try_new(SchemaRef schema, std::vector<ArrayRef> cols) {  // record_batch.rs:16-58
    if (schema->fields.size() != cols.size())
        throw OracleError("Schema has " + std::to_string(schema->fields.size()) + " fields but " +
                          std::to_string(cols.size()) + " columns provided");
    size_t n = cols.empty() ? 0 : cols[0]->len();
    for (size_t i = 0; i < cols.size(); ++i)
        if (cols[i]->len() != n)
            throw OracleError("Column " + std::to_string(i) + " has length " + std::to_string(cols[i]->len()) +
                              " but expected " + std::to_string(n));
    for (size_t i = 0; i < cols.size(); ++i)
        if (schema->fields[i].data_type != cols[i]->data_type())
            throw OracleError("Column " + std::to_string(i) + " has type " + exec_type_name(cols[i]->data_type()) +
                              " but schema expects " + exec_type_name(schema->fields[i].data_type));
    RecordBatch b; b.schema = std::move(schema); b.columns = std::move(cols); b.num_rows = n;
    return b;
}

RecordBatch RecordBatch::new_unchecked(SchemaRef schema, std::vector<ArrayRef> cols, size_t num_rows) {  // record_batch.rs:60-66
    RecordBatch b; b.schema = std::move(schema); b.columns = std::move(cols); b.num_rows = num_rows;
    return b;
}
void RecordBatch::validate() const {  // record_batch.rs:348-378
    if (schema->fields.size() != columns.size())
        throw OracleError("Schema has " + std::to_string(schema->fields.size()) + " fields but " + std::to_string(columns.size()) + " columns present");
    for (size_t i = 0; i < columns.size(); ++i) {
        if (columns[i]->len() != num_rows)
            throw OracleError("Column " + std::to_string(i) + " has length " + std::to_string(columns[i]->len()) + " but expected " + std::to_string(num_rows));
        if (schema->fields[i].data_type != columns[i]->data_type())
            throw OracleError("Column " + std::to_string(i) + " has type " + exec_type_name(columns[i]->data_type()) + " but schema expects " +
                              exec_type_name(schema->fields[i].data_type));
    }
}
size_t RecordBatch::memory_size() const {  // record_batch.rs:380-400: size_of_val(Schema) = 24, size_of::<Vec<ArrayRef>>() = 24, ArrayRef = 16
    size_t total = 24 + 24 + columns.size() * 16;
    for (const auto& c : columns) switch (c->data_type()) {
        case ExecType::Int64: case ExecType::Float64: total += c->len() * 8; break;
        case ExecType::Boolean: total += (c->len() + 7) / 8; break;
        case ExecType::String: total += c->len() * 20; break;
        case ExecType::Null: total += 16; break;
    }
    return total;
}
ArrayRef RecordBatch::column_by_name(const std::string& name) const {  // record_batch.rs:84-86
    auto i = schema->index_of(name);
    return i ? columns[*i] : nullptr;
}
void RecordBatchBuilder::add_column(ArrayRef column) {  // record_batch.rs:518-546
    if (columns.size() >= schema->fields.size()) throw OracleError("Cannot add more columns than schema defines");
    const Field& f = schema->fields[columns.size()];
    if (column->data_type() != f.data_type)
        throw OracleError(std::string("Column type ") + exec_type_name(column->data_type()) + " doesn't match expected type " + exec_type_name(f.data_type));
    if (!columns.empty() && column->len() != columns[0]->len())
        throw OracleError("Column length " + std::to_string(column->len()) + " doesn't match expected length " + std::to_string(columns[0]->len()));
    columns.push_back(std::move(column));
}
RecordBatch RecordBatchBuilder::finish() const {  // record_batch.rs:548-558
    if (columns.size() != schema->fields.size())
        throw OracleError("Expected " + std::to_string(schema->fields.size()) + " columns but only " + std::to_string(columns.size()) + " provided");
    return RecordBatch::try_new(schema, columns);
}

RecordBatch RecordBatch::slice(size_t off, size_t len) const {  // record_batch.rs:92-106
    if (!(off + len <= num_rows)) throw Panic("Slice out of bounds");
    RecordBatch b; b.schema = schema; b.num_rows = len;
    for (const auto& c : columns) b.columns.push_back(c->slice(off, len));
    return b;
}

ArrayRef take_array(const ArrayRef& array, const std::vector<size_t>& indices) {  // record_batch.rs:131-178
    switch (array->data_type()) {
        case ExecType::Int64: {  // :135-148
            auto* src = dynamic_cast<const PrimitiveArray<int64_t>*>(array.get());
            PrimitiveArrayBuilder<int64_t> b;
            for (size_t i : indices) { auto v = src->value(i); if (v) b.append_value(*v); else b.append_null(0); }
            return b.finish();
        }
        case ExecType::Float64: {  // :149-162
            auto* src = dynamic_cast<const PrimitiveArray<double>*>(array.get());
            PrimitiveArrayBuilder<double> b;
            for (size_t i : indices) { auto v = src->value(i); if (v) b.append_value(*v); else b.append_null(0.0); }
            return b.finish();
        }
        case ExecType::String: {  // :163-170
            auto* src = dynamic_cast<const StringArray*>(array.get());
            std::vector<std::optional<std::string>> vals;
            for (size_t i : indices) vals.push_back(src->value(i));
            return StringArray::make(vals);
        }
        case ExecType::Boolean: {  // :171-175
            auto* src = dynamic_cast<const BooleanArray*>(array.get());
            std::vector<std::optional<bool>> vals;
            for (size_t i : indices) vals.push_back(src->value(i));
            return BooleanArray::make(vals);
        }
        case ExecType::Null: return std::make_shared<NullArray>(indices.size());  // :176
    }
    return nullptr;
}

RecordBatch RecordBatch::take(const std::vector<size_t>& indices) const {  // record_batch.rs:108-129
    for (size_t i : indices)
        if (i >= num_rows)
            throw OracleError("Index " + std::to_string(i) + " out of bounds for " + std::to_string(num_rows) + " rows");
    RecordBatch b; b.schema = schema; b.num_rows = indices.size();
    for (const auto& c : columns) b.columns.push_back(take_array(c, indices));
    return b;
}

RecordBatch RecordBatch::select_columns(const std::vector<size_t>& idx) const {  // record_batch.rs:180-206
    for (size_t i : idx)
        if (i >= columns.size())
            throw OracleError("Column index " + std::to_string(i) + " out of bounds for " + std::to_string(columns.size()) + " columns");
    auto s = std::make_shared<Schema>();
    RecordBatch b; b.num_rows = num_rows;
    for (size_t i : idx) { s->fields.push_back(schema->fields[i]); b.columns.push_back(columns[i]); }
    b.schema = s;
    return b;
}

RecordBatch RecordBatch::select_columns_by_name(const std::vector<std::string>& names) const {  // record_batch.rs:208-219
    std::vector<size_t> idx;
    for (const auto& n : names) {
        auto i = schema->index_of(n);
        if (!i) throw OracleError("Column '" + n + "' not found");
        idx.push_back(*i);
    }
    return select_columns(idx);
}

RecordBatch RecordBatch::filter(const ArrayRef& predicate) const {  // record_batch.rs:221-243
    if (predicate->len() != num_rows)
        throw OracleError("Predicate length " + std::to_string(predicate->len()) + " doesn't match batch length " +
                          std::to_string(num_rows));
    auto* ba = dynamic_cast<const BooleanArray*>(predicate.get());
    if (!ba) throw OracleError("Predicate must be a BooleanArray");
    std::vector<size_t> sel;
    for (size_t i = 0; i < ba->len(); ++i) {  // :235-240  only Some(true) survives
        auto v = ba->value(i);
        if (v && *v) sel.push_back(i);
    }
    return take(sel);
}

ArrayRef concat_arrays(const std::vector<ArrayRef>& arrays) {  // record_batch.rs:277-342
    if (arrays.empty()) throw OracleError("Cannot concatenate empty array list");
    switch (arrays[0]->data_type()) {
        case ExecType::Int64: {
            PrimitiveArrayBuilder<int64_t> b;
            for (const auto& a : arrays) {
                auto* p = dynamic_cast<const PrimitiveArray<int64_t>*>(a.get());
                for (size_t i = 0; i < p->len(); ++i) { auto v = p->value(i); if (v) b.append_value(*v); else b.append_null(0); }
            }
            return b.finish();
        }
        case ExecType::Float64: {
            PrimitiveArrayBuilder<double> b;
            for (const auto& a : arrays) {
                auto* p = dynamic_cast<const PrimitiveArray<double>*>(a.get());
                for (size_t i = 0; i < p->len(); ++i) { auto v = p->value(i); if (v) b.append_value(*v); else b.append_null(0.0); }
            }
            return b.finish();
        }
        case ExecType::String: {
            std::vector<std::optional<std::string>> all;
            for (const auto& a : arrays) {
                auto* p = dynamic_cast<const StringArray*>(a.get());
                for (size_t i = 0; i < p->len(); ++i) all.push_back(p->value(i));
            }
            return StringArray::make(all);
        }
        case ExecType::Boolean: {
            std::vector<std::optional<bool>> all;
            for (const auto& a : arrays) {
                auto* p = dynamic_cast<const BooleanArray*>(a.get());
                for (size_t i = 0; i < p->len(); ++i) all.push_back(p->value(i));
            }
            return BooleanArray::make(all);
        }
        case ExecType::Null: {
            size_t total = 0;
            for (const auto& a : arrays) total += a->len();
            return std::make_shared<NullArray>(total);
        }
    }
    return nullptr;
}

RecordBatch RecordBatch::concat(const std::vector<RecordBatch>& batches) {  // record_batch.rs:245-275
    if (batches.empty()) throw OracleError("Cannot concatenate empty batch list");
    for (size_t i = 1; i < batches.size(); ++i)
        if (!(*batches[i].schema == *batches[0].schema)) throw OracleError("All batches must have the same schema");
    RecordBatch out; out.schema = batches[0].schema;
    for (const auto& b : batches) out.num_rows += b.num_rows;
    for (size_t c = 0; c < out.schema->fields.size(); ++c) {
        std::vector<ArrayRef> arrs;
        for (const auto& b : batches) arrs.push_back(b.columns[c]);
        out.columns.push_back(concat_arrays(arrs));
    }
    return out;
}

RecordBatch RecordBatch::empty(SchemaRef schema) {  // record_batch.rs:402-421
    RecordBatch b; b.schema = schema;
    for (const auto& f : schema->fields) {
        switch (f.data_type) {
            case ExecType::Int64: b.columns.push_back(PrimitiveArray<int64_t>::make({}, std::nullopt)); break;
            case ExecType::Float64: b.columns.push_back(PrimitiveArray<double>::make({}, std::nullopt)); break;
            case ExecType::String: b.columns.push_back(StringArray::make({})); break;
            case ExecType::Boolean: b.columns.push_back(BooleanArray::make({})); break;
            case ExecType::Null: b.columns.push_back(std::make_shared<NullArray>(0)); break;
        }
    }
    return b;
}

// ------------------------------------------------------------------------------------------
// execution/stream.rs + physical_plan/streaming.rs streams
// ------------------------------------------------------------------------------------------

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
        auto idx = b->schema->index_of(col);
        if (!idx) throw OracleError("Stream execution error: Column '" + col + "' not found in schema");
        const ArrayRef& pa = b->columns[*idx];
        if (pa->data_type() != ExecType::Boolean)
            throw OracleError("Stream execution error: Predicate column '" + col + "' is not of boolean type");
        try { return b->filter(pa); }
        catch (const OracleError& e) { throw OracleError(std::string("Stream execution error: ") + e.what()); }
    }
};
struct SelectStream : DataStream {  // stream.rs:165-213
    DataStreamRef input; std::vector<std::string> names; SchemaRef out_schema;
    SchemaRef schema() const override { return out_schema; }
    std::optional<RecordBatch> next_batch() override {
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        try { return b->select_columns_by_name(names); }
        catch (const OracleError& e) { throw OracleError(std::string("Stream execution error: ") + e.what()); }
    }
};
struct LimitStream : DataStream {  // streaming.rs:246-288
    DataStreamRef input; size_t limit = 0, rows_returned = 0;
    SchemaRef schema() const override { return input->schema(); }
    std::optional<RecordBatch> next_batch() override {
        if (rows_returned >= limit) return std::nullopt;  // :269-271 (no upstream pull)
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        size_t remaining = limit - rows_returned;
        if (b->num_rows <= remaining) { rows_returned += b->num_rows; return b; }
        auto lb = b->slice(0, remaining);
        rows_returned += remaining;
        return lb;
    }
};
}  // namespace

DataStreamRef memory_stream(SchemaRef schema, std::vector<RecordBatch> batches) {  // stream.rs:66-81
    for (const auto& b : batches)
        if (!(*b.schema == *schema)) throw OracleError("Schema mismatch: expected <schema>, found <schema>");
    auto s = std::make_unique<MemoryStream>(); s->schema_ = std::move(schema); s->batches = std::move(batches);
    return s;
}
DataStreamRef filter_stream(DataStreamRef in, std::string predicate_column) {  // stream.rs:123-128
    auto s = std::make_unique<FilterStream>(); s->input = std::move(in); s->col = std::move(predicate_column); return s;
}
DataStreamRef select_stream(DataStreamRef in, std::vector<std::string> columns) {  // stream.rs:173-194
    auto in_schema = in->schema();
    auto out = std::make_shared<Schema>();
    for (const auto& c : columns) {
        auto i = in_schema->index_of(c);
        if (!i) throw OracleError("Stream execution error: Column '" + c + "' not found in schema");
        out->fields.push_back(in_schema->fields[*i]);
    }
    auto s = std::make_unique<SelectStream>(); s->input = std::move(in); s->names = std::move(columns); s->out_schema = out;
    return s;
}
DataStreamRef limit_stream(DataStreamRef in, size_t limit) {  // streaming.rs:254-260
    auto s = std::make_unique<LimitStream>(); s->input = std::move(in); s->limit = limit; return s;
}

std::vector<RecordBatch> collect_all_batches(DataStream& s) {  // streaming.rs:335-341
    std::vector<RecordBatch> out;
    while (auto b = s.next_batch()) out.push_back(*b);
    return out;
}

RecordBatch collect_stream_batches(DataStream& s) {  // streaming.rs:343-352
    auto schema = s.schema();
    auto batches = collect_all_batches(s);
    if (batches.empty()) return RecordBatch::empty(schema);
    try { return RecordBatch::concat(batches); }
    catch (const OracleError& e) { throw OracleError(std::string("Conversion error: ") + e.what()); }
}

std::vector<RecordBatch> dataframe_to_batches(const DataFrame& df, size_t batch_size) {  // streaming.rs:135-233
    std::vector<RecordBatch> batches;
    if (df.is_empty()) return batches;  // :140-142
    size_t num_rows = df.height();
    size_t num_batches = (num_rows + batch_size - 1) / batch_size;
    auto schema = std::make_shared<Schema>();
    for (const auto& s : df.columns()) {  // :148-161
        ExecType t = ExecType::Null;
        switch (s.dtype()) {
            case DataType::Int64: t = ExecType::Int64; break;
            case DataType::Float64: t = ExecType::Float64; break;
            case DataType::String: t = ExecType::String; break;
            case DataType::Boolean: t = ExecType::Boolean; break;
            case DataType::Null: t = ExecType::Null; break;
        }
        schema->fields.push_back(Field{s.name(), t, true});
    }
    for (size_t bi = 0; bi < num_batches; ++bi) {  // :164-230
        size_t start = bi * batch_size, end = std::min((bi + 1) * batch_size, num_rows);
        std::vector<ArrayRef> arrays;
        for (const auto& s : df.columns()) {
            switch (s.dtype()) {
                case DataType::Int64: {  // :173-182  (Null -> 0, no validity)
                    std::vector<int64_t> v;
                    for (size_t i = start; i < end; ++i) {
                        const auto& a = s.at(i);
                        if (a.tag == AnyValue::kInt64) v.push_back(a.i);
                        else if (a.tag == AnyValue::kNull) v.push_back(0);
                        else throw Panic("Type mismatch in Int64 series");
                    }
                    arrays.push_back(PrimitiveArray<int64_t>::make(std::move(v), std::nullopt));
                    break;
                }
                case DataType::Float64: {  // :184-193
                    std::vector<double> v;
                    for (size_t i = start; i < end; ++i) {
                        const auto& a = s.at(i);
                        if (a.tag == AnyValue::kFloat64) v.push_back(a.f);
                        else if (a.tag == AnyValue::kNull) v.push_back(0.0);
                        else throw Panic("Type mismatch in Float64 series");
                    }
                    arrays.push_back(PrimitiveArray<double>::make(std::move(v), std::nullopt));
                    break;
                }
                case DataType::String: {  // :195-206 (nulls kept)
                    std::vector<std::optional<std::string>> v;
                    for (size_t i = start; i < end; ++i) {
                        const auto& a = s.at(i);
                        if (a.tag == AnyValue::kString) v.push_back(a.s);
                        else if (a.tag == AnyValue::kNull) v.push_back(std::nullopt);
                        else throw Panic("Type mismatch in String series");
                    }
                    arrays.push_back(StringArray::make(v));
                    break;
                }
                case DataType::Boolean: {  // :208-217 (Null -> false)
                    std::vector<std::optional<bool>> v;
                    for (size_t i = start; i < end; ++i) {
                        const auto& a = s.at(i);
                        if (a.tag == AnyValue::kBoolean) v.push_back(a.b);
                        else if (a.tag == AnyValue::kNull) v.push_back(false);
                        else throw Panic("Type mismatch in Boolean series");
                    }
                    arrays.push_back(BooleanArray::make(v));
                    break;
                }
                case DataType::Null: arrays.push_back(std::make_shared<NullArray>(end - start)); break;  // :219-221
            }
        }
        try { batches.push_back(RecordBatch::try_new(schema, std::move(arrays))); }
        catch (const OracleError& e) { throw OracleError(std::string("Conversion error: ") + e.what()); }
    }
    return batches;
}

// ------------------------------------------------------------------------------------------
// logical_plan/plan.rs
// ------------------------------------------------------------------------------------------

static std::pair<std::string, DataType> resolve_expr_schema(const Expr& e,
        const std::vector<std::pair<std::string, DataType>>& in) {  // logical_plan/plan.rs:204-233
    switch (e.kind) {
        case Expr::Column: {
            for (const auto& p : in) if (p.first == e.name) return {e.name, p.second};
            return {e.name, DataType::Null};
        }
        case Expr::Alias: return {e.name, resolve_expr_schema(*e.left, in).second};
        case Expr::Binary: {
            auto l = resolve_expr_schema(*e.left, in);
            auto r = resolve_expr_schema(*e.right, in);
            DataType res = DataType::Null;  // :235-262
            switch (e.op) {
                case BinaryOperator::Eq: case BinaryOperator::NotEq: case BinaryOperator::Lt: case BinaryOperator::Gt:
                case BinaryOperator::LtEq: case BinaryOperator::GtEq: case BinaryOperator::And: case BinaryOperator::Or:
                    res = DataType::Boolean; break;
                default:
                    if (l.second == DataType::Float64 || r.second == DataType::Float64) res = DataType::Float64;
                    else if (l.second == DataType::Int64 && r.second == DataType::Int64) res = DataType::Int64;
                    else if (l.second == DataType::Null) res = r.second;
                    else if (r.second == DataType::Null) res = l.second;
                    else res = DataType::Null;
            }
            return {l.first, res};
        }
        case Expr::Literal: return {"literal", e.value.data_type()};
    }
    return {"", DataType::Null};
}

std::vector<std::pair<std::string, DataType>> LogicalPlan::schema() const {  // logical_plan/plan.rs:63-113
    switch (kind) {
        case DataFrameSource: case CsvFileSource: return src_schema;  // :65-66
        case Select: {
            auto in = input->schema();
            std::vector<std::pair<std::string, DataType>> out;
            for (const auto& e : expressions) out.push_back(resolve_expr_schema(e, in));
            return out;
        }
        case Filter: case Limit: return input->schema();
        case Join: {  // :79-111: the left schema, then the right columns except the right key, "_right" on a name the left side has
            auto left = input->schema();
            auto out = left;
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

static void validate_expr_columns(const Expr& e, const std::vector<std::pair<std::string, DataType>>& schema) {
    switch (e.kind) {  // logical_plan/plan.rs:264-286
        case Expr::Column: {
            bool found = false;
            for (const auto& p : schema) if (p.first == e.name) found = true;
            if (!found) throw OracleError("Logical plan error: Column not found: '" + e.name + "'");
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
                if (!df.column(p.first)) throw OracleError("Logical plan error: Column not found: '" + p.first + "'");
            break;
        case Select: {
            input->validate();
            auto in = input->schema();
            for (const auto& e : expressions) validate_expr_columns(e, in);
            break;
        }
        case Filter: {
            input->validate();
            validate_expr_columns(predicate, input->schema());
            break;
        }
        case Limit: input->validate(); break;
        case CsvFileSource: break;  // :128
        case Join: {  // :156-200 — the right side is validated first
            right->validate();
            input->validate();
            const auto ls = input->schema(), rs = right->schema();
            const DataType* lt = nullptr; const DataType* rt = nullptr;
            for (const auto& c : ls) if (c.first == left_key && !lt) lt = &c.second;
            if (!lt) throw OracleError("Logical plan error: Column not found: '" + left_key + "'");
            for (const auto& c : rs) if (c.first == right_key && !rt) rt = &c.second;
            if (!rt) throw OracleError("Logical plan error: Column not found: '" + right_key + "'");
            if (!is_comparable_with(*lt, *rt))
                throw OracleError(std::string("Logical plan error: Incompatible join key types: '") + dtype_name(*lt) + "' and '" + dtype_name(*rt) + "'");
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------
// logical_plan/optimizer.rs
// ------------------------------------------------------------------------------------------

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
            if (in.kind == LogicalPlan::Filter) {  // :17-40 — note: no recursion below this pair
                if (predicate_uses_only_selected_columns(in.predicate, plan.expressions)) {
                    LogicalPlan sel; sel.kind = LogicalPlan::Select; sel.input = in.input; sel.expressions = plan.expressions;
                    LogicalPlan fil; fil.kind = LogicalPlan::Filter; fil.predicate = in.predicate;
                    fil.input = std::make_shared<LogicalPlan>(std::move(sel));
                    return fil;
                }
                return plan;
            }
            LogicalPlan out = plan;  // :41-44
            out.input = std::make_shared<LogicalPlan>(optimize(in));
            return out;
        }
        case LogicalPlan::Filter: {  // :46-49
            LogicalPlan out = plan;
            out.input = std::make_shared<LogicalPlan>(optimize(*plan.input));
            return out;
        }
        case LogicalPlan::Join: {  // :50-61
            LogicalPlan out = plan;
            out.input = std::make_shared<LogicalPlan>(optimize(*plan.input));
            out.right = std::make_shared<LogicalPlan>(optimize(*plan.right));
            return out;
        }
        default: return plan;  // :62 (Limit, sources: untouched)
    }
}

// ------------------------------------------------------------------------------------------
// physical_plan/planner.rs + physical_plan/plan.rs (eager engine)
// ------------------------------------------------------------------------------------------

bool eval_cmp(const AnyValue& row, BinaryOperator op, const AnyValue& lit) {  // physical_plan/plan.rs:114-120
    switch (op) {
        case BinaryOperator::Eq: return any_eq(row, lit);
        case BinaryOperator::NotEq: return !any_eq(row, lit);
        case BinaryOperator::Lt: { auto c = any_partial_cmp(row, lit); return c && *c < 0; }
        case BinaryOperator::Gt: { auto c = any_partial_cmp(row, lit); return c && *c > 0; }
        case BinaryOperator::LtEq: { auto c = any_partial_cmp(row, lit); return c && *c <= 0; }
        case BinaryOperator::GtEq: { auto c = any_partial_cmp(row, lit); return c && *c >= 0; }
        default: throw OracleError("invalid comparison operator");
    }
}

static std::pair<std::string, std::string> convert_select_expr(const Expr& e) {  // planner.rs:113-132
    switch (e.kind) {
        case Expr::Column: return {e.name, e.name};
        case Expr::Alias:
            if (e.left->kind == Expr::Column) return {e.left->name, e.name};
            throw OracleError("Unsupported expression: " + str_debug(e.debug()));
        case Expr::Binary: throw OracleError("Unsupported expression: " + str_debug(e.debug()));
        case Expr::Literal: throw OracleError("Select expression must be a column or alias, found: " + str_debug(e.debug()));
    }
    return {};
}

struct FilterSpec { std::string column; AnyValue value; BinaryOperator op; };
static FilterSpec convert_filter_predicate(const Expr& p) {  // planner.rs:134-189
    if (p.kind != Expr::Binary) {
        const char* t = p.kind == Expr::Column ? "Column" : (p.kind == Expr::Literal ? "Literal" : "Alias");
        throw OracleError(std::string("Filter must be a binary comparison, found: ") + t);
    }
    switch (p.op) {
        case BinaryOperator::Eq: case BinaryOperator::NotEq: case BinaryOperator::Lt:
        case BinaryOperator::Gt: case BinaryOperator::LtEq: case BinaryOperator::GtEq: break;
        case BinaryOperator::And: case BinaryOperator::Or:
            throw OracleError("Unsupported filter: only simple column comparisons supported, found: " + str_debug(p.debug()));
        default: throw OracleError(std::string("Unsupported binary operator in filter: ") + op_name(p.op));
    }
    if (p.left->kind != Expr::Column) throw OracleError("Filter left side must be a column reference, found: " + str_debug(p.left->debug()));
    if (p.right->kind != Expr::Literal) throw OracleError("Filter right side must be a literal value, found: " + str_debug(p.right->debug()));
    return {p.left->name, p.right->value, p.op};
}

// ---- extension: And / Or over comparison leaves (see rivulus_oracle.hpp)
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
// value_of(column) -> AnyValue of the current row
template <class F>
static bool eval_predicate_tree(const Expr& p, F&& value_of) {
    if (is_compound(p)) {
        const bool l = eval_predicate_tree(*p.left, value_of), r = eval_predicate_tree(*p.right, value_of);
        return p.op == BinaryOperator::And ? (l && r) : (l || r);
    }
    return eval_cmp(value_of(p.left->name), p.op, p.right->value);
}

// Lowering checks happen for the whole tree before any execution (planner.rs:41-111 runs first).
static void check_lowering(const LogicalPlan& p) {
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return;
        case LogicalPlan::CsvFileSource:   // planner.rs:45-49
            throw OracleError("Conversion failed: CSV file source not supported in non-streaming physical planner. Use streaming planner instead.");
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

// HashMap<AnyValue, Vec<usize>> key of the join (series.rs:73-98): same variant and same value; f64 hashed by to_bits, NaN == nothing.
// The reference holds no join test: pinned only by the main.rs demo queries (:170-196); everything else here is
// "parity unpinned — code reading only".
// 0.0 and -0.0 are equal but hash differently, so they meet only when the table's random state happens to collide: kept apart here.
struct JoinKey {
    int tag; uint64_t bits; std::string s;
    bool operator<(const JoinKey& o) const { return tag != o.tag ? tag < o.tag : (bits != o.bits ? bits < o.bits : s < o.s); }
};
static bool join_key_of(const AnyValue& v, JoinKey* k) {
    k->tag = (int)v.tag; k->bits = 0; k->s.clear();
    switch (v.tag) {
        case AnyValue::kNull: break;
        case AnyValue::kInt64: k->bits = (uint64_t)v.i; break;
        case AnyValue::kFloat64: if (v.f != v.f) return false; std::memcpy(&k->bits, &v.f, 8); break;
        case AnyValue::kString: k->s = v.s; break;
        case AnyValue::kBoolean: k->bits = v.b ? 1 : 0; break;
    }
    return true;
}

static DataFrame exec_node(const LogicalPlan& p) {  // physical_plan/plan.rs:65-173
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return p.df;  // :67
        case LogicalPlan::CsvFileSource: throw OracleError("unreachable: rejected by check_lowering");
        case LogicalPlan::Select: {                      // :68-96
            DataFrame in = exec_node(*p.input);
            std::vector<std::string> cols, finals;
            for (const auto& e : p.expressions) { auto pr = convert_select_expr(e); cols.push_back(pr.first); finals.push_back(pr.second); }
            for (const auto& c : cols) if (!in.column(c)) throw OracleError("Column not found: '" + c + "'");
            DataFrame sel;
            try { sel = in.select(cols); } catch (const OracleError& e) { throw OracleError(std::string("DataFrame error: ") + e.what()); }
            std::vector<Series> renamed;
            for (size_t i = 0; i < sel.columns().size() && i < finals.size(); ++i) {
                try { renamed.push_back(Series::make(finals[i], sel.columns()[i].data())); }  // :88-93 (EmptyData on 0 rows)
                catch (const OracleError& e) { throw OracleError(std::string("Series error: ") + e.what()); }
            }
            try { return DataFrame::make(std::move(renamed)); }
            catch (const OracleError& e) { throw OracleError(std::string("DataFrame error: ") + e.what()); }
        }
        case LogicalPlan::Filter: {                      // :97-150
            DataFrame in = exec_node(*p.input);
            std::vector<bool> mask; mask.reserve(in.height());
            if (g_extensions && is_compound(p.predicate)) {
                std::vector<std::string> cols;
                leaf_columns(p.predicate, cols);
                for (const auto& c : cols) if (!in.column(c)) throw OracleError("Column not found: '" + c + "'");
                for (size_t i = 0; i < in.height(); ++i)
                    mask.push_back(eval_predicate_tree(p.predicate, [&](const std::string& c) -> const AnyValue& { return in.column(c)->data()[i]; }));
            } else {
            FilterSpec f = convert_filter_predicate(p.predicate);
            const Series* fs = in.column(f.column);
            if (!fs) throw OracleError("Column not found: '" + f.column + "'");
            for (const auto& rv : fs->data()) mask.push_back(eval_cmp(rv, f.op, f.value));  // :112-130
            }
            std::vector<Series> out;
            for (const auto& s : in.columns()) {         // :132-147
                std::vector<AnyValue> kept;
                for (size_t i = 0; i < s.len(); ++i) if (mask[i]) kept.push_back(s.data()[i]);
                if (kept.empty()) out.push_back(Series::empty(s.name(), s.dtype()));
                else {
                    try { out.push_back(Series::make(s.name(), std::move(kept))); }
                    catch (const OracleError& e) { throw OracleError(std::string("Series error: ") + e.what()); }
                }
            }
            try { return DataFrame::make(std::move(out)); }
            catch (const OracleError& e) { throw OracleError(std::string("DataFrame error: ") + e.what()); }
        }
        case LogicalPlan::Limit: {                       // :151-173
            DataFrame in = exec_node(*p.input);
            std::vector<Series> out;
            if (p.n == 0 || in.is_empty()) {
                for (const auto& s : in.columns()) out.push_back(Series::empty(s.name(), s.dtype()));
                return DataFrame::make(std::move(out));
            }
            size_t lim = std::min(p.n, in.height());
            for (const auto& s : in.columns()) {
                std::vector<AnyValue> d(s.data().begin(), s.data().begin() + lim);
                try { out.push_back(Series::make(s.name(), std::move(d))); }  // :169 (EmptyData if height 0)
                catch (const OracleError& e) { throw OracleError(std::string("Series error: ") + e.what()); }
            }
            return DataFrame::make(std::move(out));
        }
        case LogicalPlan::Join: {                        // :174-284, build = left, probe = right (planner.rs:100-108)
            DataFrame build = exec_node(*p.input);
            const Series* bks = build.column(p.left_key);
            if (!bks) throw Panic("called `Option::unwrap()` on a `None` value");
            std::map<JoinKey, std::vector<size_t>> table;   // :186-193
            JoinKey k;
            for (size_t i = 0; i < bks->len(); ++i) if (join_key_of(bks->data()[i], &k)) table[k].push_back(i);
            DataFrame probe = exec_node(*p.right);
            const Series* pks = probe.column(p.right_key);
            if (!pks) throw Panic("called `Option::unwrap()` on a `None` value");
            std::vector<std::pair<size_t, size_t>> pairs;   // :197-205
            for (size_t i = 0; i < pks->len(); ++i) {
                if (!join_key_of(pks->data()[i], &k)) continue;
                auto it = table.find(k);
                if (it != table.end()) for (size_t b : it->second) pairs.emplace_back(i, b);
            }
            std::vector<Series> out;                         // materialize_join_result :208-254 / create_empty_join_result :256-283
            auto build_name = [&](const Series& s) { return probe.column(s.name()) ? s.name() + "_right" : s.name(); };
            try {
                for (const auto& s : probe.columns()) {
                    if (pairs.empty()) { out.push_back(Series::empty(s.name(), s.dtype())); continue; }
                    std::vector<AnyValue> d; d.reserve(pairs.size());
                    for (const auto& pr : pairs) d.push_back(s.data()[pr.first]);
                    out.push_back(Series::make(s.name(), std::move(d)));
                }
                for (const auto& s : build.columns()) {
                    if (s.name() == p.left_key) continue;
                    if (pairs.empty()) { out.push_back(Series::empty(build_name(s), s.dtype())); continue; }
                    std::vector<AnyValue> d; d.reserve(pairs.size());
                    for (const auto& pr : pairs) d.push_back(s.data()[pr.second]);
                    out.push_back(Series::make(build_name(s), std::move(d)));
                }
            } catch (const OracleError& e) { throw OracleError(std::string("Series error: ") + e.what()); }
            try { return DataFrame::make(std::move(out)); }
            catch (const OracleError& e) { throw OracleError(std::string("DataFrame error: ") + e.what()); }
        }
    }
    return DataFrame();
}

DataFrame execute_eager(const LogicalPlan& optimized) {
    check_lowering(optimized);
    return exec_node(optimized);
}

// ------------------------------------------------------------------------------------------
// physical_plan/streaming.rs + streaming_planner.rs
// ------------------------------------------------------------------------------------------

StreamingPhysicalPlan StreamingPhysicalPlan::memory_source(std::vector<RecordBatch> b) {
    StreamingPhysicalPlan p; p.kind = MemorySource; p.batches = std::move(b); return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::dataframe_source(DataFrame df, size_t batch_size) {
    StreamingPhysicalPlan p; p.kind = DataFrameSource; p.df = std::move(df); p.batch_size = batch_size; return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::csv_file_source(std::string path, SchemaRef schema, std::optional<size_t> batch_size,
                                                             std::optional<std::string> delimiter) {  // streaming.rs:299-311
    StreamingPhysicalPlan p; p.kind = CsvFileSource; p.csv_path = std::move(path); p.csv_schema = std::move(schema);
    p.csv_batch_size = batch_size; p.csv_delimiter = std::move(delimiter); return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::filter(std::string col) const {
    StreamingPhysicalPlan p; p.kind = Filter; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.predicate_column = std::move(col); return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::filter_expr(Expr predicate) const {
    StreamingPhysicalPlan p; p.kind = FilterExpr; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.predicate = std::move(predicate); return p;
}
StreamingPhysicalPlan StreamingPhysicalPlan::select(std::vector<std::string> cols) const {
    StreamingPhysicalPlan p; p.kind = Select; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.columns = std::move(cols); return p;
}

namespace {
// extension: FilterStream over a predicate tree instead of a Boolean column; the mask it builds has no nulls
struct FilterExprStream : DataStream {
    DataStreamRef input; Expr pred;
    SchemaRef schema() const override { return input->schema(); }
    std::optional<RecordBatch> next_batch() override {
        auto b = input->next_batch();
        if (!b) return std::nullopt;
        std::vector<std::string> cols;
        leaf_columns(pred, cols);
        for (const auto& c : cols)
            if (!b->schema->index_of(c)) throw OracleError("Stream execution error: Column '" + c + "' not found in schema");
        std::vector<std::optional<bool>> mask;
        mask.reserve(b->num_rows);
        for (size_t i = 0; i < b->num_rows; ++i)
            mask.push_back(eval_predicate_tree(pred, [&](const std::string& c) { return array_value(*b->columns[*b->schema->index_of(c)], i); }));
        try { return b->filter(BooleanArray::make(mask)); }
        catch (const OracleError& e) { throw OracleError(std::string("Stream execution error: ") + e.what()); }
    }
};
}  // namespace
StreamingPhysicalPlan StreamingPhysicalPlan::limit(size_t n) const {
    StreamingPhysicalPlan p; p.kind = Limit; p.input = std::make_shared<StreamingPhysicalPlan>(*this); p.n = n; return p;
}

DataStreamRef StreamingPhysicalPlan::execute() const {  // streaming.rs:70-133
    switch (kind) {
        case MemorySource: {
            if (batches.empty()) throw OracleError("Invalid operation: Cannot create stream from empty batch list");
            try { return memory_stream(batches[0].schema, batches); }
            catch (const OracleError& e) { throw OracleError(std::string("Stream error: ") + e.what()); }
        }
        case DataFrameSource: {
            auto b = dataframe_to_batches(df, batch_size);
            SchemaRef s = b.empty() ? std::make_shared<Schema>() : b[0].schema;
            return memory_stream(s, std::move(b));
        }
        case CsvFileSource: {  // :96-105
            try { return csv_file_stream(csv_path, csv_schema, csv_batch_size, csv_delimiter); }
            catch (const OracleError& e) { throw OracleError(std::string("Invalid operation: ") + e.what()); }
        }
        case Filter: return filter_stream(input->execute(), predicate_column);
        case FilterExpr: { auto st = std::make_unique<FilterExprStream>(); st->input = input->execute(); st->pred = predicate; return st; }
        case Select: {
            auto in = input->execute();
            try { return select_stream(std::move(in), columns); }
            catch (const OracleError& e) { throw OracleError(std::string("Stream error: ") + e.what()); }
        }
        case Limit: return limit_stream(input->execute(), n);
        case HashJoin: throw Panic("not yet implemented: Streaming hash join not yet implemented");   // streaming.rs:128-131
    }
    return nullptr;
}

static std::string wrap_stream_err(const std::string& w) {
    // StreamError raised inside next_batch is converted via `?` into StreamingExecutionError::Stream
    if (w.rfind("Stream execution error:", 0) == 0) return "Stream error: " + w;
    return w;
}

RecordBatch StreamingPhysicalPlan::collect() const {  // streaming.rs:235-238
    auto s = execute();
    try { return collect_stream_batches(*s); }
    catch (const OracleError& e) { throw OracleError(wrap_stream_err(e.what())); }
}
std::vector<RecordBatch> StreamingPhysicalPlan::collect_batches() const {  // streaming.rs:240-243
    auto s = execute();
    try { return collect_all_batches(*s); }
    catch (const OracleError& e) { throw OracleError(wrap_stream_err(e.what())); }
}

StreamingPhysicalPlan logical_to_streaming(const LogicalPlan& plan) {  // streaming_planner.rs:29-100
    switch (plan.kind) {
        case LogicalPlan::DataFrameSource: return StreamingPhysicalPlan::dataframe_source(plan.df, 1024);  // :31-33
        case LogicalPlan::CsvFileSource: {  // :35-62 — every field nullable
            auto schema = std::make_shared<Schema>();
            for (const auto& p : plan.src_schema) {
                ExecType t = ExecType::Null;
                switch (p.second) {
                    case DataType::Int64: t = ExecType::Int64; break;
                    case DataType::Float64: t = ExecType::Float64; break;
                    case DataType::String: t = ExecType::String; break;
                    case DataType::Boolean: t = ExecType::Boolean; break;
                    case DataType::Null: t = ExecType::Null; break;
                }
                schema->fields.push_back(Field{p.first, t, true});
            }
            return StreamingPhysicalPlan::csv_file_source(plan.csv_path, schema, plan.csv_batch_size, plan.csv_delimiter);
        }
        case LogicalPlan::Select: {  // :65-69, 102-135
            auto in = logical_to_streaming(*plan.input);
            std::vector<std::string> names;
            for (const auto& e : plan.expressions) {
                if (e.kind == Expr::Column) names.push_back(e.name);
                else if (e.kind == Expr::Alias) {
                    if (e.left->kind == Expr::Column) names.push_back(e.left->name);  // alias dropped :110-113
                    else throw OracleError("Streaming planner error: Expression conversion error: Complex expressions with aliases not yet supported: " + e.debug());
                } else
                    throw OracleError("Streaming planner error: Expression conversion error: Complex expressions not yet supported in streaming mode: " + e.debug());
            }
            return in.select(names);
        }
        case LogicalPlan::Filter: {  // :71-75, 137-168
            auto in = logical_to_streaming(*plan.input);
            const Expr& p = plan.predicate;
            if (p.kind == Expr::Column) return in.filter(p.name);
            if (p.kind == Expr::Binary && g_extensions) {
                // extension: comparison leaves and And / Or over them; a malformed leaf raises the eager planner's message
                try { check_predicate_tree(p); }
                catch (const OracleError& e) { throw OracleError(std::string("Streaming planner error: Expression conversion error: ") + e.what()); }
                return in.filter_expr(p);
            }
            if (p.kind == Expr::Binary) {
                if (p.left->kind == Expr::Column)
                    throw OracleError("Streaming planner error: Expression conversion error: Binary expressions not yet supported in streaming mode. "
                                      "Found expression on column '" + p.left->name + "'. "
                                      "Currently only simple boolean column references are supported (e.g., .filter(col('is_active')))");
                throw OracleError("Streaming planner error: Expression conversion error: Complex binary expressions not supported in streaming mode");
            }
            throw OracleError("Streaming planner error: Expression conversion error: Unsupported filter expression type: " + p.debug());
        }
        case LogicalPlan::Limit: return logical_to_streaming(*plan.input).limit(plan.n);  // :76-79
        case LogicalPlan::Join: {  // :81-98, then streaming.rs:128-131: execute() is `todo!()`
            (void)logical_to_streaming(*plan.input); (void)logical_to_streaming(*plan.right);
            StreamingPhysicalPlan sp; sp.kind = StreamingPhysicalPlan::HashJoin;
            return sp;
        }
    }
    return StreamingPhysicalPlan();
}

// ------------------------------------------------------------------------------------------
// logical_plan/builder.rs
// ------------------------------------------------------------------------------------------

LazyFrame LazyFrame::from_dataframe(const DataFrame& df) {  // builder.rs:27-39
    LazyFrame lf; lf.plan.kind = LogicalPlan::DataFrameSource; lf.plan.df = df;  // clone
    for (const auto& s : df.columns()) lf.plan.src_schema.emplace_back(s.name(), s.dtype());
    return lf;
}
LazyFrame LazyFrame::from_csv(std::string path, std::vector<std::pair<std::string, DataType>> schema, std::optional<size_t> batch_size,
                              std::optional<std::string> delimiter) {  // builder.rs:41-55
    LazyFrame lf; lf.plan.kind = LogicalPlan::CsvFileSource; lf.plan.csv_path = std::move(path); lf.plan.src_schema = std::move(schema);
    lf.plan.csv_batch_size = batch_size; lf.plan.csv_delimiter = std::move(delimiter);
    return lf;
}
LazyFrame LazyFrame::select(std::vector<Expr> e) const {  // builder.rs:57-64
    LazyFrame lf; lf.plan.kind = LogicalPlan::Select; lf.plan.input = std::make_shared<LogicalPlan>(plan); lf.plan.expressions = std::move(e); return lf;
}
LazyFrame LazyFrame::filter(Expr p) const {  // builder.rs:66-73
    LazyFrame lf; lf.plan.kind = LogicalPlan::Filter; lf.plan.input = std::make_shared<LogicalPlan>(plan); lf.plan.predicate = std::move(p); return lf;
}
LazyFrame LazyFrame::limit(size_t n) const {  // builder.rs:75-82
    LazyFrame lf; lf.plan.kind = LogicalPlan::Limit; lf.plan.input = std::make_shared<LogicalPlan>(plan); lf.plan.n = n; return lf;
}

LazyFrame LazyFrame::inner_join(const LazyFrame& right, std::string left_key, std::string right_key) const {  // builder.rs:84-94
    LazyFrame lf; lf.plan.kind = LogicalPlan::Join; lf.plan.input = std::make_shared<LogicalPlan>(plan);
    lf.plan.right = std::make_shared<LogicalPlan>(right.plan); lf.plan.left_key = std::move(left_key); lf.plan.right_key = std::move(right_key);
    return lf;
}

DataFrame LazyFrame::collect() const {  // builder.rs:96-104
    LogicalPlan opt = optimize(plan);
    opt.validate();  // throws "Logical plan error: ..."
    try { return execute_eager(opt); }
    catch (const OracleError& e) { throw OracleError(std::string("Execution error: ") + e.what()); }
}

RecordBatch LazyFrame::collect_streaming() const {  // builder.rs:106-113
    LogicalPlan opt = optimize(plan);
    opt.validate();
    StreamingPhysicalPlan sp = logical_to_streaming(opt);  // throws "Streaming planner error: ..."
    try { return sp.collect(); }
    catch (const OracleError& e) { throw OracleError(std::string("Execution error: ") + e.what()); }
}

// ------------------------------------------------------------------------------------------
// Fused-operator oracle (extension; composed only of reference pieces)
// ------------------------------------------------------------------------------------------

AnyValue array_value(const Array& a, size_t i) {
    switch (a.data_type()) {
        case ExecType::Int64: { auto v = dynamic_cast<const PrimitiveArray<int64_t>&>(a).value(i); return v ? AnyValue::Int64(*v) : AnyValue::Null(); }
        case ExecType::Float64: { auto v = dynamic_cast<const PrimitiveArray<double>&>(a).value(i); return v ? AnyValue::Float64(*v) : AnyValue::Null(); }
        case ExecType::String: { auto v = dynamic_cast<const StringArray&>(a).value(i); return v ? AnyValue::String(*v) : AnyValue::Null(); }
        case ExecType::Boolean: { auto v = dynamic_cast<const BooleanArray&>(a).value(i); return v ? AnyValue::Boolean(*v) : AnyValue::Null(); }
        case ExecType::Null: return AnyValue::Null();
    }
    return AnyValue::Null();
}

static RecordBatch finish_fused(const RecordBatch& filtered, const std::vector<size_t>& proj, int64_t limit) {
    RecordBatch sel = filtered.select_columns(proj);                 // SelectStream: record_batch.rs:180-206
    if (limit >= 0 && (size_t)limit < sel.num_rows) sel = sel.slice(0, (size_t)limit);  // LimitStream: streaming.rs:279-282
    return RecordBatch::concat({sel});                               // collect: streaming.rs:351 (offset 0, bitmap iff nulls)
}

RecordBatch filter_project_cmp(const RecordBatch& in, size_t pred_col, BinaryOperator op, const AnyValue& lit,
                               const std::vector<size_t>& proj, int64_t limit) {
    if (pred_col >= in.columns.size()) throw OracleError("Column index out of bounds");
    const Array& pc = *in.columns[pred_col];
    std::vector<std::optional<bool>> mask;
    mask.reserve(in.num_rows);
    for (size_t i = 0; i < in.num_rows; ++i) mask.push_back(eval_cmp(array_value(pc, i), op, lit));  // plan.rs:112-130
    RecordBatch filtered = in.filter(BooleanArray::make(mask));      // record_batch.rs:221-243
    return finish_fused(filtered, proj, limit);
}

RecordBatch filter_project_mask(const RecordBatch& in, size_t mask_col, const std::vector<size_t>& proj, int64_t limit) {
    if (mask_col >= in.columns.size()) throw OracleError("Column index out of bounds");
    RecordBatch filtered = in.filter(in.columns[mask_col]);          // FilterStream: stream.rs:136-162
    return finish_fused(filtered, proj, limit);
}

}  // namespace orc
