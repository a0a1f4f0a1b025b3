// rivulus.hpp — C++17 host layer of the B200-native filter / project / limit path.
//
// Mirrors the reference's public API for this path (same names, argument meaning and error text), and drives the
// GPU exclusively through the C ABI of include/rivulus_gpu.h — exactly what a Rust crate binding that header would
// do (the image has no rustc, so the host side is C++; see INTEGRATION.md for the Rust-side binding).
//
//   reference (under /root/reference/src)                      here
//   datatypes/series.rs      AnyValue, DataType, Series        rivulus::AnyValue, DataType, Series (columnar storage)
//   datatypes/dataframe.rs   DataFrame                         rivulus::DataFrame
//   expressions/expr.rs      Expr, BinaryOperator              rivulus::Expr, BinaryOperator
//   logical_plan/*           LogicalPlan, QueryOptimizer       rivulus::LogicalPlan, optimize()
//   logical_plan/builder.rs  LazyFrame                         rivulus::LazyFrame
//   physical_plan/planner.rs + plan.rs (eager executor)        execute_eager(): Filter+Select fused into one rvl_filter_project
//   execution/record_batch.rs RecordBatch                      rivulus::RecordBatch (device-resident)
//   execution/stream.rs      DataStream, Memory/Filter/Select  rivulus::DataStream, MemoryStream, FilterStream, SelectStream
//   physical_plan/streaming.rs StreamingPhysicalPlan, LimitStream  rivulus::StreamingPhysicalPlan, LimitStream
//
// Storage differs on purpose: a Series holds Arrow-layout column buffers (values / LSB-first validity / int32
// offsets + bytes) instead of Vec<AnyValue>, so handing a DataFrame to the device is a plain copy.
// There is no CPU execution path: every query runs the CUDA kernels; without a GPU the calls fail loudly.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rivulus_gpu.h"

namespace rivulus {

// thrown for every reference `Err(..)`; what() is the reference's Display text.  `panic` marks what the reference
// reports by panicking (slice / index out of bounds).
struct Error : std::runtime_error {
    bool panic = false;
    explicit Error(const std::string& m, bool p = false) : std::runtime_error(m), panic(p) {}
};

// ---------------------------------------------------------------------------------------- datatypes/series.rs
enum class DataType : int { Int64 = 0, Float64 = 1, String = 2, Boolean = 3, Null = 4 };  // series.rs:126-133
const char* to_string(DataType d);

struct AnyValue {  // series.rs:6-13
    enum Tag : uint8_t { Null = 0, Int64 = 1, Float64 = 2, String = 3, Boolean = 4 };
    Tag tag = Null;
    int64_t i = 0;
    double f = 0.0;
    bool b = false;
    std::string s;
    AnyValue() = default;
    AnyValue(int64_t v) : tag(Int64), i(v) {}           // From<i64>    series.rs:31-35
    AnyValue(int v) : tag(Int64), i(v) {}
    AnyValue(double v) : tag(Float64), f(v) {}           // From<f64>    :37-41
    AnyValue(const char* v) : tag(String), s(v) {}       // From<&str>   :49-53
    AnyValue(std::string v) : tag(String), s(std::move(v)) {}
    AnyValue(bool v) : tag(Boolean), b(v) {}             // From<bool>   :55-59
    bool is_null() const { return tag == Null; }
    DataType data_type() const;                          // :20-28
    std::string display() const;                         // :61-71
    std::string debug() const;
    bool operator==(const AnyValue& o) const;            // PartialEq :87-98
};
std::optional<int> partial_cmp(const AnyValue& a, const AnyValue& b);  // series.rs:100-117

// One column in Arrow layout.  `validity` empty = no nulls.
class Series {
  public:
    static Series make(const std::string& name, const std::vector<AnyValue>& data);  // Series::new  series.rs:185-221
    static Series empty(const std::string& name, DataType dtype);                    // :223-229
    const std::string& name() const { return name_; }
    size_t len() const { return len_; }
    bool is_empty() const { return len_ == 0; }
    const DataType& dtype() const { return dtype_; }
    AnyValue at(size_t i) const;              // Index<usize> :273-288 (throws Error{panic} out of bounds)
    std::optional<AnyValue> get(size_t i) const;
    std::vector<AnyValue> to_values() const;
    std::string display() const;              // "Series: numbers [{dtype}; {len}]" :267-271
    Series renamed(const std::string& n) const { Series s = *this; s.name_ = n; return s; }
    size_t null_count() const;

    // raw buffers (rvl_column view over host memory)
    rvl_column as_column() const;
    static Series from_host_column(const std::string& name, const rvl_column& c, DataType dtype_hint);

  private:
    friend class RecordBatch;
    std::string name_;
    DataType dtype_ = DataType::Null;
    size_t len_ = 0;
    std::vector<int64_t> i64_;
    std::vector<double> f64_;
    std::vector<uint8_t> bits_;      // Boolean values, LSB-first
    std::vector<int32_t> offsets_;   // String
    std::vector<uint8_t> data_;      // String
    std::vector<uint8_t> validity_;  // LSB-first, empty = all valid
    std::vector<uint8_t> int_tag_;   // Float64-dtype series that also holds Int64 values (series.rs:210-212): per-row "is Int64"
    std::vector<int64_t> int_vals_;
};

// ---------------------------------------------------------------------------------------- datatypes/dataframe.rs
class DataFrame {
  public:
    static DataFrame make(std::vector<Series> columns);  // DataFrame::new :29-56
    static DataFrame empty() { return DataFrame(); }
    size_t height() const { return columns_.empty() ? 0 : columns_[0].len(); }
    size_t width() const { return columns_.size(); }
    std::pair<size_t, size_t> shape() const { return {height(), width()}; }
    bool is_empty() const { return columns_.empty(); }
    const Series* column(const std::string& name) const;
    std::vector<std::string> column_names() const;
    const std::vector<Series>& columns() const { return columns_; }
    DataFrame select(const std::vector<std::string>& names) const;  // :96-110
    const Series& operator[](const std::string& name) const;        // Index<&str> :136-149
    static DataFrame unchecked(std::vector<Series> c) { DataFrame d; d.columns_ = std::move(c); return d; }

  private:
    std::vector<Series> columns_;
};

// ---------------------------------------------------------------------------------------- expressions/expr.rs
enum class BinaryOperator : int { Plus = 0, Minus, Multiply, Divide, Eq, NotEq, Lt, Gt, LtEq, GtEq, And, Or };
const char* to_string(BinaryOperator op);

struct Expr {
    enum Kind { Column, Literal, BinaryExpr, Alias } kind = Column;
    std::string name;
    AnyValue value;
    BinaryOperator op = BinaryOperator::Eq;
    std::shared_ptr<Expr> left, right;  // Alias: inner = left

    static Expr col(const std::string& n) { Expr e; e.kind = Column; e.name = n; return e; }
    static Expr lit(AnyValue v) { Expr e; e.kind = Literal; e.value = std::move(v); return e; }
    Expr alias(const std::string& n) const { Expr e; e.kind = Alias; e.name = n; e.left = std::make_shared<Expr>(*this); return e; }
    Expr binary(BinaryOperator o, const Expr& r) const {
        Expr e; e.kind = BinaryExpr; e.op = o; e.left = std::make_shared<Expr>(*this); e.right = std::make_shared<Expr>(r); return e;
    }
    Expr add(const Expr& o) const { return binary(BinaryOperator::Plus, o); }
    Expr sub(const Expr& o) const { return binary(BinaryOperator::Minus, o); }
    Expr mul(const Expr& o) const { return binary(BinaryOperator::Multiply, o); }
    Expr div(const Expr& o) const { return binary(BinaryOperator::Divide, o); }
    Expr eq(const Expr& o) const { return binary(BinaryOperator::Eq, o); }
    Expr neq(const Expr& o) const { return binary(BinaryOperator::NotEq, o); }
    Expr lt(const Expr& o) const { return binary(BinaryOperator::Lt, o); }
    Expr gt(const Expr& o) const { return binary(BinaryOperator::Gt, o); }
    Expr lte(const Expr& o) const { return binary(BinaryOperator::LtEq, o); }
    Expr gte(const Expr& o) const { return binary(BinaryOperator::GtEq, o); }
    Expr and_(const Expr& o) const { return binary(BinaryOperator::And, o); }
    Expr or_(const Expr& o) const { return binary(BinaryOperator::Or, o); }
    std::string debug() const;
};

// ---------------------------------------------------------------------------------------- device context (one per GPU)
class Context {
  public:
    explicit Context(int device = 0);
    ~Context();
    Context(const Context&) = delete;
    rvl_ctx* handle() const { return ctx_; }
    static std::shared_ptr<Context> shared(int device = 0);  // process-wide default context per device

  private:
    rvl_ctx* ctx_ = nullptr;
};
using ContextRef = std::shared_ptr<Context>;

// ---------------------------------------------------------------------------------------- execution/schema.rs
enum class ExecType : int { Null = 0, Boolean = 1, Int64 = 2, Float64 = 3, String = 4 };  // schema.rs:1-8 (= rvl_dtype)
struct Field {
    std::string name; ExecType data_type; bool nullable = true;
    bool operator==(const Field& o) const { return name == o.name && data_type == o.data_type && nullable == o.nullable; }
};
struct Schema {
    std::vector<Field> fields;
    std::optional<size_t> index_of(const std::string& n) const;
    const Field* field_by_name(const std::string& n) const;
    size_t num_fields() const { return fields.size(); }
    bool operator==(const Schema& o) const { return fields == o.fields; }
};
using SchemaRef = std::shared_ptr<Schema>;

// host copy of one device column (what `value(i)` reads in the reference)
struct ArrayData {
    ExecType dtype = ExecType::Null;
    int64_t length = 0, null_count = 0;
    std::vector<int64_t> i64; std::vector<double> f64; std::vector<uint8_t> bits;
    std::vector<int32_t> offsets; std::vector<uint8_t> data; std::vector<uint8_t> validity;  // validity empty = bitmap absent
    AnyValue value(size_t i) const;
};

// ---------------------------------------------------------------------------------------- execution/record_batch.rs
class RecordBatch {
  public:
    RecordBatch() = default;
    // try_new over host columns: uploads (record_batch.rs:16-58)
    static RecordBatch try_new(const ContextRef& ctx, SchemaRef schema, const std::vector<rvl_column>& host_columns);
    static RecordBatch from_series(const ContextRef& ctx, const std::vector<Series>& cols, bool flatten_nulls);
    static RecordBatch adopt(const ContextRef& ctx, SchemaRef schema, rvl_batch* handle);
    const SchemaRef& schema() const { return schema_; }
    size_t num_rows() const;
    size_t num_columns() const { return schema_ ? schema_->fields.size() : 0; }
    bool is_empty() const { return num_rows() == 0; }
    RecordBatch slice(size_t offset, size_t length) const;                           // :92-106
    RecordBatch select_columns(const std::vector<size_t>& indices) const;            // :180-206
    RecordBatch select_columns_by_name(const std::vector<std::string>& names) const; // :208-219
    RecordBatch filter(const RecordBatch& predicate_batch, size_t predicate_column) const;  // :221-243 (mask = a Boolean column)
    RecordBatch filter_by_column(size_t mask_column) const;
    static RecordBatch concat(const std::vector<RecordBatch>& batches);              // :245-275
    static RecordBatch empty(const ContextRef& ctx, SchemaRef schema);               // :402-421
    ArrayData column_data(size_t i) const;   // download
    rvl_batch* handle() const { return h_ ? h_->b : nullptr; }
    const ContextRef& context() const { return ctx_; }

  private:
    struct Handle { rvl_batch* b = nullptr; ~Handle(); };
    ContextRef ctx_;
    SchemaRef schema_;
    std::shared_ptr<Handle> h_;
};

// ---------------------------------------------------------------------------------------- execution/stream.rs, streaming.rs
class DataStream {  // trait DataStream  stream.rs:25-54
  public:
    virtual ~DataStream() = default;
    virtual SchemaRef schema() const = 0;
    virtual std::optional<RecordBatch> next_batch() = 0;
    std::vector<RecordBatch> collect();     // :30-39
    RecordBatch concatenate();              // :41-53
};
using DataStreamRef = std::unique_ptr<DataStream>;
DataStreamRef make_memory_stream(SchemaRef schema, std::vector<RecordBatch> batches);        // MemoryStream::new :66-81
DataStreamRef make_filter_stream(DataStreamRef input, std::string predicate_column);         // FilterStream::new :123-128
DataStreamRef make_select_stream(DataStreamRef input, std::vector<std::string> columns);     // SelectStream::new :173-194
DataStreamRef make_limit_stream(DataStreamRef input, size_t limit);                          // LimitStream::new streaming.rs:254-260

struct StreamingPhysicalPlan {  // streaming.rs:28-68, 290-333
    enum Kind { MemorySource, DataFrameSource, Filter, Select, Limit } kind = MemorySource;
    std::vector<RecordBatch> batches;
    DataFrame df; size_t batch_size = 0;
    std::shared_ptr<StreamingPhysicalPlan> input;
    std::string predicate_column;
    std::vector<std::string> columns;
    size_t n = 0;
    ContextRef ctx;
    static StreamingPhysicalPlan memory_source(std::vector<RecordBatch> b);
    static StreamingPhysicalPlan dataframe_source(DataFrame df, size_t batch_size, ContextRef ctx = nullptr);
    StreamingPhysicalPlan filter(std::string col) const;
    StreamingPhysicalPlan select(std::vector<std::string> cols) const;
    StreamingPhysicalPlan limit(size_t n) const;
    DataStreamRef execute() const;                      // :70-133 (one operator object per node)
    RecordBatch collect() const;                        // :235-238; DataFrame sources take the fused rvl_stream_* pipeline
    std::vector<RecordBatch> collect_batches() const;   // :240-243
};

// ---------------------------------------------------------------------------------------- logical_plan/*
struct LogicalPlan {
    enum Kind { DataFrameSource, Select, Filter, Limit } kind = DataFrameSource;
    DataFrame df;
    std::vector<std::pair<std::string, DataType>> src_schema;
    std::shared_ptr<LogicalPlan> input;
    std::vector<Expr> expressions;
    Expr predicate;
    size_t n = 0;
    std::vector<std::pair<std::string, DataType>> schema() const;  // logical_plan/plan.rs:63-113
    void validate() const;                                         // :115-202
    std::string shape() const;
};
LogicalPlan optimize(LogicalPlan plan);  // optimizer.rs:7-64

class LazyFrame {  // logical_plan/builder.rs:11-114
  public:
    static LazyFrame from_dataframe(const DataFrame& df, ContextRef ctx = nullptr);
    LazyFrame select(std::vector<Expr> exprs) const;
    LazyFrame filter(Expr predicate) const;
    LazyFrame limit(size_t n) const;
    DataFrame collect() const;               // eager engine semantics on the GPU
    RecordBatch collect_streaming() const;   // streaming engine semantics on the GPU
    const LogicalPlan& logical_plan() const { return plan_; }

  private:
    LogicalPlan plan_;
    ContextRef ctx_;
};

}  // namespace rivulus
