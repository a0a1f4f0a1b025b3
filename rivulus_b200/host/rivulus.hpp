// rivulus.hpp — C++17 host layer of the B200-native filter / project / limit path.
//
// Mirrors the reference's public API for this path (same names, argument meaning and error text) and drives the GPU
// exclusively through the C ABI of include/rivulus_gpu.h — exactly what a Rust crate binding that header would do
// (the image has no rustc, so the host side is C++; INTEGRATION.md shows the Rust-side binding).
//
//   reference (under /root/reference/src)                      here
//   datatypes/series.rs      AnyValue, DataType, Series        rivulus::AnyValue, DataType, Series (columnar storage)
//   datatypes/dataframe.rs   DataFrame                         rivulus::DataFrame
//   expressions/expr.rs      Expr, BinaryOperator              rivulus::Expr, BinaryOperator
//   logical_plan/*           LogicalPlan, QueryOptimizer       rivulus::LogicalPlan, optimize()
//   logical_plan/builder.rs  LazyFrame                         rivulus::LazyFrame
//   physical_plan/planner.rs + plan.rs (eager executor)        LazyFrame::collect(): Filter [+Select] [+Limit] = ONE rvl_filter_project
//   execution/record_batch.rs RecordBatch                      rivulus::RecordBatch (device-resident)
//   execution/stream.rs      DataStream, Memory/Filter/Select  rivulus::DataStream + make_*_stream
//   physical_plan/streaming.rs StreamingPhysicalPlan, LimitStream  rivulus::StreamingPhysicalPlan, make_limit_stream
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

// Thrown for every reference `Err(..)`; what() is the reference's Display text.  `panic` marks what the reference
// reports by panicking (slice / index out of bounds, type mismatch in dataframe_to_batches).
struct Error : std::runtime_error {
    bool panic = false;
    explicit Error(const std::string& m, bool p = false) : std::runtime_error(m), panic(p) {}
};

// ---------------------------------------------------------------------------------------- datatypes/series.rs
enum class DataType : int { Int64 = 0, Float64 = 1, String = 2, Boolean = 3, Null = 4 };  // series.rs:126-133
const char* dtype_name(DataType d);                                                       // series.rs:162-172
bool dtype_is_numeric(DataType d);                                                        // series.rs:136-142
bool dtype_is_comparable_with(DataType a, DataType b);                                    // series.rs:144-159

struct AnyValue {  // series.rs:6-13
    enum Tag : uint8_t { kNull = 0, kInt64 = 1, kFloat64 = 2, kString = 3, kBoolean = 4 };
    Tag tag = kNull;
    int64_t i = 0;
    double f = 0.0;
    bool b = false;
    std::string s;
    static AnyValue Null() { return AnyValue(); }
    static AnyValue Int64(int64_t v) { AnyValue a; a.tag = kInt64; a.i = v; return a; }      // From<i64>  :31-35
    static AnyValue Float64(double v) { AnyValue a; a.tag = kFloat64; a.f = v; return a; }   // From<f64>  :37-41
    static AnyValue String(std::string v) { AnyValue a; a.tag = kString; a.s = std::move(v); return a; }  // :43-53
    static AnyValue Boolean(bool v) { AnyValue a; a.tag = kBoolean; a.b = v; return a; }     // From<bool> :55-59
    bool is_null() const { return tag == kNull; }   // :16-18
    DataType data_type() const;                     // :20-28
    std::string display() const;                    // :61-71
    std::string debug() const;                      // #[derive(Debug)]
};

// ---------------------------------------------------------------------------------------- host buffers
// Column buffers live in PAGE-LOCKED memory once they are big enough to matter (>= 64 KB), so a DataFrame is handed to the
// device by the copy engine straight from where it lies and results land where they will stay: no pageable staging copy on
// either side (SURVEY.md 8(f) rank 1 — the ingestion step in front of the hot path).  Blocks are recycled through a size-class
// pool (pinning a fresh 24 MB buffer costs milliseconds; reusing one costs nothing).  Without a usable CUDA device (CPU-only
// tests) the allocator quietly uses malloc: nothing on this path needs the GPU to hold data.
void* host_buffer_alloc(size_t bytes);
void host_buffer_free(void* p, size_t bytes);
void host_buffer_pool_trim();   // give cached pinned blocks back to the driver
template <class T>
struct HostAllocator {
    using value_type = T;
    HostAllocator() = default;
    template <class U> HostAllocator(const HostAllocator<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(host_buffer_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) { host_buffer_free(p, n * sizeof(T)); }
    // resize(n) default-initialises (no zero fill of megabytes that a copy is about to overwrite); assign(n, v) still writes v
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(std::forward<A>(a)...);
    }
    template <class U> bool operator==(const HostAllocator<U>&) const { return true; }
    template <class U> bool operator!=(const HostAllocator<U>&) const { return false; }
};
template <class T> using HostVec = std::vector<T, HostAllocator<T>>;

struct ArrayData;

// One column in Arrow layout.  dtype follows the reference's inference (series.rs:185-221): first non-null value's
// type, Null when every value is null.  Values under a null are stored as 0 / false / empty.
class Series {
  public:
    static Series make(const std::string& name, const std::vector<AnyValue>& data);  // Series::new  :185-221
    static Series empty(const std::string& name, DataType dtype);                    // Series::empty :223-229
    // columnar constructors (no reference counterpart: the reference only ingests Vec<AnyValue>); same inference rules.
    // `validity_bits`: LSB-first, 1 = valid, empty = no nulls.
    static Series from_i64(const std::string& name, std::vector<int64_t> v, std::vector<uint8_t> validity_bits = {});
    static Series from_f64(const std::string& name, std::vector<double> v, std::vector<uint8_t> validity_bits = {});
    static Series from_bool_bits(const std::string& name, std::vector<uint8_t> value_bits, size_t n, std::vector<uint8_t> validity_bits = {});
    static Series from_strings(const std::string& name, std::vector<int32_t> offsets, std::vector<uint8_t> data, std::vector<uint8_t> validity_bits = {});
    // the same over borrowed buffers (copied once, into page-locked storage); validity_bits may be NULL
    static Series from_i64(const std::string& name, const int64_t* v, size_t n, const uint8_t* validity_bits);
    static Series from_f64(const std::string& name, const double* v, size_t n, const uint8_t* validity_bits);
    static Series from_bool_bits(const std::string& name, const uint8_t* value_bits, size_t n, const uint8_t* validity_bits);
    static Series from_strings(const std::string& name, const int32_t* offsets, size_t n, const uint8_t* data, const uint8_t* validity_bits);

    const std::string& name() const { return name_; }
    size_t len() const { return len_; }
    bool is_empty() const { return len_ == 0; }
    DataType dtype() const { return dtype_; }
    AnyValue at(size_t i) const;                       // Index<usize> :273-288 (Error{panic} out of bounds)
    std::optional<AnyValue> get(size_t i) const;       // :243-245
    std::vector<AnyValue> to_values() const;
    std::string display() const;                       // "Series: numbers [{dtype}; {len}]" :267-271
    // copies share the (immutable) buffers, like an Arc clone: from_dataframe / select / rename never move column data
    Series renamed(const std::string& n) const { Series s = *this; s.name_ = n; return s; }
    size_t null_count() const;
    bool is_valid(size_t i) const { return dtype_ != DataType::Null && (b_->validity_.empty() || ((b_->validity_[i >> 3] >> (i & 7)) & 1)); }
    // A Float64-dtype Series may hold Int64 values (series.rs:210-212).  Such a mixed column keeps ONE 8-byte buffer (the f64
    // or the i64 bit pattern of each row) plus a tag bitmap (1 = the row is an AnyValue::Int64); on the device the tag travels
    // as a hidden Boolean column and the predicate kernels compare each row with its own type (rvl_predicate::tag_column).
    bool is_mixed() const { return !b_->int_tag_.empty(); }
    rvl_column tag_column(size_t offset, size_t length) const;   // Boolean column over the tag bitmap (mixed series only)
    // Build from a downloaded (values, tag) pair; `dtype` is what Series::new infers for the survivors (Null / Int64 / Float64)
    static Series from_mixed(const std::string& name, const rvl_column& values, const rvl_column& tags, DataType dtype);

    // raw buffers
    const HostVec<int64_t>& i64_values() const { return b_->i64_; }
    const HostVec<double>& f64_values() const { return b_->f64_; }
    const HostVec<uint8_t>& bool_bits() const { return b_->bits_; }
    const HostVec<int32_t>& str_offsets() const { return b_->offsets_; }
    const HostVec<uint8_t>& str_data() const { return b_->data_; }
    const HostVec<uint8_t>& validity_bits() const { return b_->validity_; }
    // rvl_column view over rows [offset, offset + length) of this Series' host buffers (borrowed).
    // flatten_nulls: numeric / boolean nulls become valid 0 / false, as dataframe_to_batches does (streaming.rs:177,188,212).
    rvl_column as_column(size_t offset, size_t length, bool flatten_nulls) const;
    // Build from a downloaded column (dtype of the device array; `dtype_if_empty` is kept when length == 0).
    static Series from_column(const std::string& name, const rvl_column& c, DataType dtype_if_empty);
    // Adopt a downloaded array's buffers (moved, not copied); `dtype` is the eager dtype of the result column
    static Series from_array(const std::string& name, ArrayData&& a, DataType dtype);

  private:
    std::string name_;
    DataType dtype_ = DataType::Null;
    size_t len_ = 0;
    struct Bufs {
        HostVec<int64_t> i64_;
        HostVec<double> f64_;
        HostVec<uint8_t> bits_;      // Boolean values, LSB-first
        HostVec<int32_t> offsets_;   // String (len + 1 entries)
        HostVec<uint8_t> data_;      // String
        HostVec<uint8_t> validity_;  // LSB-first, empty = all valid
        HostVec<uint8_t> int_tag_;   // mixed Float64 series: LSB-first bitmap, 1 = f64_[i] holds the bit pattern of an AnyValue::Int64
    };
    std::shared_ptr<Bufs> b_ = std::make_shared<Bufs>();   // written only while the Series is being built; shared by its copies
    void infer_from_validity();      // all null -> dtype Null
};

// ---------------------------------------------------------------------------------------- datatypes/dataframe.rs
class DataFrame {
  public:
    static DataFrame make(std::vector<Series> columns);  // DataFrame::new :29-56
    static DataFrame empty() { return DataFrame(); }     // :58-62
    size_t height() const { return columns_.empty() ? 0 : columns_[0].len(); }  // :64-70
    size_t width() const { return columns_.size(); }
    std::pair<size_t, size_t> shape() const { return {height(), width()}; }
    bool is_empty() const { return columns_.empty(); }   // :80-82 (no columns)
    const Series* column(const std::string& name) const; // :84-86
    std::vector<std::string> column_names() const;
    const std::vector<Series>& columns() const { return columns_; }
    DataFrame select(const std::vector<std::string>& names) const;  // :96-110
    const Series& operator[](const std::string& name) const;        // Index<&str> :136-149 (panics)
    static DataFrame unchecked(std::vector<Series> c) { DataFrame d; d.columns_ = std::move(c); return d; }

  private:
    std::vector<Series> columns_;
};

// ---------------------------------------------------------------------------------------- expressions/expr.rs
enum class BinaryOperator : int { Plus = 0, Minus, Multiply, Divide, Eq, NotEq, Lt, Gt, LtEq, GtEq, And, Or };  // :15-29 (= rvl_op)
const char* op_name(BinaryOperator op);

struct Expr {  // expr.rs:3-13
    enum Kind { Column, Literal, Binary, Alias } kind = Column;
    std::string name;
    AnyValue value;
    BinaryOperator op = BinaryOperator::Eq;
    std::shared_ptr<Expr> left, right;  // Alias: inner = left

    static Expr col(const std::string& n) { Expr e; e.kind = Column; e.name = n; return e; }       // :31-33
    static Expr lit(AnyValue v) { Expr e; e.kind = Literal; e.value = std::move(v); return e; }    // :35-37
    static Expr lit(int64_t v) { return lit(AnyValue::Int64(v)); }
    static Expr lit(int v) { return lit(AnyValue::Int64(v)); }
    static Expr lit(double v) { return lit(AnyValue::Float64(v)); }
    static Expr lit(const char* v) { return lit(AnyValue::String(v)); }
    static Expr lit(bool v) { return lit(AnyValue::Boolean(v)); }
    Expr alias(const std::string& n) const { Expr e; e.kind = Alias; e.name = n; e.left = std::make_shared<Expr>(*this); return e; }
    Expr binary(BinaryOperator o, const Expr& r) const {
        Expr e; e.kind = Binary; e.op = o; e.left = std::make_shared<Expr>(*this); e.right = std::make_shared<Expr>(r); return e;
    }
    Expr add(const Expr& o) const { return binary(BinaryOperator::Plus, o); }      // :43-121
    Expr sub(const Expr& o) const { return binary(BinaryOperator::Minus, o); }
    Expr mul(const Expr& o) const { return binary(BinaryOperator::Multiply, o); }
    Expr div(const Expr& o) const { return binary(BinaryOperator::Divide, o); }
    Expr eq(const Expr& o) const { return binary(BinaryOperator::Eq, o); }
    Expr neq(const Expr& o) const { return binary(BinaryOperator::NotEq, o); }
    Expr lt(const Expr& o) const { return binary(BinaryOperator::Lt, o); }
    Expr gt(const Expr& o) const { return binary(BinaryOperator::Gt, o); }
    Expr lte(const Expr& o) const { return binary(BinaryOperator::LtEq, o); }
    Expr gte(const Expr& o) const { return binary(BinaryOperator::GtEq, o); }
    Expr and_(const Expr& o) const { return binary(BinaryOperator::And, o); }     // :124-138
    Expr or_(const Expr& o) const { return binary(BinaryOperator::Or, o); }
    std::string debug() const;
};

// ---------------------------------------------------------------------------------------- device context (one per GPU)
class Context {
  public:
    explicit Context(int device = 0);
    ~Context();
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    rvl_ctx* handle() const { return ctx_; }
    static std::shared_ptr<Context> shared(int device = 0);  // process-wide default context per device

  private:
    rvl_ctx* ctx_ = nullptr;
};
using ContextRef = std::shared_ptr<Context>;

// ---------------------------------------------------------------------------------------- execution/schema.rs
enum class ExecType : int { Null = 0, Boolean = 1, Int64 = 2, Float64 = 3, String = 4 };  // schema.rs:1-8 (= rvl_dtype)
const char* exec_type_name(ExecType t);
struct Field {  // schema.rs:10-36
    std::string name; ExecType data_type = ExecType::Null; bool nullable = true;
    bool operator==(const Field& o) const { return name == o.name && data_type == o.data_type && nullable == o.nullable; }
};
struct Schema {  // schema.rs:38-76
    std::vector<Field> fields;
    std::optional<size_t> index_of(const std::string& n) const;
    const Field* field_by_name(const std::string& n) const;
    size_t num_fields() const { return fields.size(); }
    bool is_empty() const { return fields.empty(); }
    bool operator==(const Schema& o) const { return fields == o.fields; }
};
using SchemaRef = std::shared_ptr<Schema>;

// Host copy of one device column (what the reference's `value(i)` / `values()` / `null_bitmap()` expose).
struct ArrayData {
    ExecType dtype = ExecType::Null;
    int64_t length = 0, null_count = 0;
    HostVec<int64_t> i64; HostVec<double> f64; HostVec<uint8_t> bits;
    HostVec<int32_t> offsets; HostVec<uint8_t> data;
    HostVec<uint8_t> validity;  // empty = bitmap absent
    bool has_validity = false;
    AnyValue value(size_t i) const;  // Error{panic} "Index {} out of bounds" (primitive.rs:49 etc.)
};

// ---------------------------------------------------------------------------------------- execution/record_batch.rs
class RecordBatch {
  public:
    RecordBatch() = default;
    // try_new over HOST columns (record_batch.rs:16-58): checks field count / lengths / dtypes, then uploads
    static RecordBatch try_new(const ContextRef& ctx, SchemaRef schema, const std::vector<rvl_column>& host_columns);
    // new_unchecked (:60-66): no checks.  Columns that do not fit (schema, num_rows) cannot be uploaded as one batch; the batch then only
    // remembers what validate() has to report, as the reference's would on its mismatched arrays.
    static RecordBatch new_unchecked(const ContextRef& ctx, SchemaRef schema, const std::vector<rvl_column>& host_columns, size_t num_rows);
    void validate() const;                                                           // :348-378 (throws Error with the Err text)
    size_t memory_size() const;                                                      // :380-400
    std::optional<ArrayData> column_by_name(const std::string& name) const;          // :84-86
    static RecordBatch adopt(const ContextRef& ctx, SchemaRef schema, rvl_batch* handle);
    static RecordBatch empty(const ContextRef& ctx, SchemaRef schema);               // :402-421
    const SchemaRef& schema() const { return schema_; }
    size_t num_rows() const;                                                         // :72
    size_t num_columns() const;                                                      // :76
    bool is_empty() const { return num_rows() == 0; }
    RecordBatch slice(size_t offset, size_t length) const;                           // :92-106 (Error{panic} "Slice out of bounds")
    RecordBatch take(const std::vector<size_t>& indices) const;                      // :108-129 ("Index {} out of bounds for {} rows")
    RecordBatch select_columns(const std::vector<size_t>& indices) const;            // :180-206
    RecordBatch select_columns_by_name(const std::vector<std::string>& names) const; // :208-219
    // filter(&self, predicate: &ArrayRef) :221-243 — the predicate array is column `predicate_column` of `predicate_batch`
    RecordBatch filter(const RecordBatch& predicate_batch, size_t predicate_column) const;
    static RecordBatch concat(const std::vector<RecordBatch>& batches);              // :245-275
    RecordBatch with_schema(SchemaRef schema) const { RecordBatch r = *this; r.schema_ = std::move(schema); return r; }  // same buffers, renamed fields
    ArrayData column_data(size_t i) const;   // device -> host copy of column i (rebased to offset 0)
    int64_t column_null_count(size_t i) const;
    rvl_batch* handle() const { return h_ ? h_->b : nullptr; }
    const ContextRef& context() const { return ctx_; }

  private:
    struct Handle { rvl_batch* b = nullptr; ~Handle(); };
    ContextRef ctx_;
    SchemaRef schema_;
    std::shared_ptr<Handle> h_;
    std::string invalid_;   // new_unchecked over inconsistent columns: validate()'s message
};

// record_batch.rs:495-573.  Columns are HOST columns (borrowed until finish(), which uploads them in one batch).
class RecordBatchBuilder {
  public:
    explicit RecordBatchBuilder(SchemaRef schema, ContextRef ctx = nullptr) : schema_(std::move(schema)), ctx_(std::move(ctx)) {}   // new :501-507
    static RecordBatchBuilder with_capacity(SchemaRef schema, size_t, ContextRef ctx = nullptr) { return RecordBatchBuilder(std::move(schema), std::move(ctx)); }  // :509-516
    void add_column(const rvl_column& host_column);     // :518-546 (throws Error with the Err text)
    RecordBatch finish() const;                         // :548-558
    size_t num_columns() const { return columns_.size(); }                              // :560-562
    bool is_complete() const { return columns_.size() == schema_->fields.size(); }      // :564-566

  private:
    SchemaRef schema_;
    ContextRef ctx_;
    std::vector<rvl_column> columns_;
};

// ---------------------------------------------------------------------------------------- execution/stream.rs, streaming.rs
class DataStream {  // trait DataStream  stream.rs:25-54
  public:
    virtual ~DataStream() = default;
    virtual SchemaRef schema() const = 0;
    virtual std::optional<RecordBatch> next_batch() = 0;
    std::vector<RecordBatch> collect();     // :30-39
};
using DataStreamRef = std::unique_ptr<DataStream>;
DataStreamRef make_memory_stream(const ContextRef& ctx, SchemaRef schema, std::vector<RecordBatch> batches);  // MemoryStream::new :66-81
DataStreamRef make_filter_stream(DataStreamRef input, std::string predicate_column);         // FilterStream::new :123-128
DataStreamRef make_filter_expr_stream(DataStreamRef input, Expr predicate);                  // extension: And / Or / comparison tree
DataStreamRef make_select_stream(DataStreamRef input, std::vector<std::string> columns);     // SelectStream::new :173-194
DataStreamRef make_limit_stream(DataStreamRef input, size_t limit);                          // LimitStream::new streaming.rs:254-260
std::vector<RecordBatch> collect_all_batches(DataStream& s);                                 // streaming.rs:335-341
RecordBatch collect_stream_batches(const ContextRef& ctx, DataStream& s);                    // streaming.rs:343-352

// execution/file_stream.rs (csv_stream.cpp) — SURVEY.md 8(f) rank 3
size_t calculate_adaptive_batch_size(const Schema& schema);   // :346-369
// true = reproduce the reference's inverted validity of Int64 / Float64 columns that hold a null (:213-240, :245-272); default false
void set_csv_reference_validity(bool on);
bool csv_reference_validity();
// parse threads per reader: -1 (default) = up to 8 (half the hardware threads) for files of at least 8 MiB and none below, 0 = parse in the calling thread
void set_csv_threads(int n);
// The parser behind CsvFileStream: `batch_size` data lines at a time straight into reusable Arrow-layout host buffers.
class CsvBatchReader {
  public:
    // CsvFileStream::new :20-40; throws Error("Failed to open file: …").  delimiter: one UTF-8 encoded char, default ","
    CsvBatchReader(const std::string& path, SchemaRef schema, std::optional<size_t> batch_size, std::optional<std::string> delimiter);
    ~CsvBatchReader();
    const SchemaRef& schema() const;
    size_t batch_size() const;
    size_t read_batch();                      // read_batch :123-199: rows parsed, 0 = end of file; throws "Stream execution error: …"
    std::vector<rvl_column> columns() const;  // host views of the batch just parsed (valid until the next read_batch)
    void apply_reference_validity();          // see set_csv_reference_validity
  private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};
DataStreamRef make_csv_file_stream(const ContextRef& ctx, const std::string& path, SchemaRef schema, std::optional<size_t> batch_size,
                                   std::optional<std::string> delimiter);   // CsvFileStream as a DataStream :328-336

// streaming.rs:135-233: DataFrame -> RecordBatches of `batch_size` rows on the device
std::vector<RecordBatch> dataframe_to_batches(const ContextRef& ctx, const DataFrame& df, size_t batch_size);

struct StreamingPhysicalPlan {  // streaming.rs:28-68, 290-333
    enum Kind { MemorySource, DataFrameSource, Filter, Select, Limit, FilterExpr, CsvFileSource, HashJoin } kind = MemorySource;
    Expr predicate;   // FilterExpr (opt-in extension, see set_extensions)
    std::string csv_path; SchemaRef csv_schema; std::optional<size_t> csv_batch_size; std::optional<std::string> csv_delimiter;  // :39-44
    std::vector<RecordBatch> batches;
    DataFrame df; size_t batch_size = 0;
    std::shared_ptr<StreamingPhysicalPlan> input;
    std::string predicate_column;
    std::vector<std::string> columns;
    size_t n = 0;
    ContextRef ctx;
    static StreamingPhysicalPlan memory_source(std::vector<RecordBatch> b);
    static StreamingPhysicalPlan dataframe_source(DataFrame df, size_t batch_size, ContextRef ctx = nullptr);
    static StreamingPhysicalPlan csv_file_source(std::string path, SchemaRef schema, std::optional<size_t> batch_size,
                                                 std::optional<std::string> delimiter, ContextRef ctx = nullptr);   // :299-311
    StreamingPhysicalPlan filter(std::string col) const;
    StreamingPhysicalPlan filter_expr(Expr predicate) const;   // extension
    StreamingPhysicalPlan select(std::vector<std::string> cols) const;
    StreamingPhysicalPlan limit(size_t n) const;
    DataStreamRef execute() const;                      // :70-133 (one operator object per node, batches of `batch_size`)
    // :235-238.  The result is the concatenation of every batch, so it does not depend on the batch size: plans rooted
    // at a DataFrame source run with batches of at least kCollectBatchRows rows (fewer, larger kernel launches).
    // [Limit] -> [Select] -> [Filter] over a DataFrame or CSV source runs as one pinned, overlapped device pipeline (rvl_stream_*).
    RecordBatch collect() const;
    std::vector<RecordBatch> collect_batches() const;   // :240-243 (honours batch_size: batch boundaries are visible)
    static constexpr size_t kCollectBatchRows = 1 << 20;
};

// OPT-IN EXTENSION, off by default — the default is the reference's behaviour, error text included: both
// reference executors reject And / Or predicates (planner.rs:146-150) and the streaming planner rejects every comparison
// (streaming_planner.rs:137-168).  With set_extensions(true) a filter predicate may be a tree of And / Or over `column <op> literal`
// leaves, in collect() and in collect_streaming().  Each leaf is evaluated with the eager truth table (plan.rs:114-120 over
// series.rs:87-117) by the predicate kernels into a selection mask (rvl_predicate_mask), the masks are combined on the device
// (rvl_boolean_op = BooleanArray::{and, or}, boolean.rs:120-165) and the result drives the ordinary mask filter
// (RVL_PRED_BOOL_COLUMN = RecordBatch::filter, record_batch.rs:221-243).  A single comparison in collect_streaming() runs as the
// fused comparison operator (RVL_PRED_CMP_LITERAL) inside the rvl_stream pipeline.
void set_extensions(bool on);
bool extensions_enabled();

// ---------------------------------------------------------------------------------------- logical_plan/*
struct LogicalPlan {  // logical_plan/plan.rs:8-39
    enum Kind { DataFrameSource, Select, Filter, Limit, CsvFileSource, Join } kind = DataFrameSource;
    std::shared_ptr<LogicalPlan> right;                 // Join :32-38 (`input` is the left side); JoinType::Inner only
    std::string left_key, right_key;
    DataFrame df;
    std::vector<std::pair<std::string, DataType>> src_schema;   // DataFrameSource, CsvFileSource
    std::string csv_path; std::optional<size_t> csv_batch_size; std::optional<std::string> csv_delimiter;   // CsvFileSource :14-19
    std::shared_ptr<LogicalPlan> input;
    std::vector<Expr> expressions;
    Expr predicate;
    size_t n = 0;
    std::vector<std::pair<std::string, DataType>> schema() const;  // logical_plan/plan.rs:63-113
    void validate() const;                                         // :115-202 (throws "Logical plan error: …")
    std::string shape() const;                                     // e.g. "Limit(Select(Filter(Source)))"
    std::string describe() const;                                  // Debug-style dump: node kinds and expression trees
};
LogicalPlan optimize(LogicalPlan plan);                                   // optimizer.rs:7-64
StreamingPhysicalPlan logical_to_streaming(const LogicalPlan& plan, const ContextRef& ctx);  // streaming_planner.rs:29-168
DataFrame execute_eager(const LogicalPlan& optimized, const ContextRef& ctx);  // planner.rs:41-189 + physical_plan/plan.rs:65-173

class LazyFrame {  // logical_plan/builder.rs:11-114
  public:
    static LazyFrame from_dataframe(const DataFrame& df, ContextRef ctx = nullptr);  // :27-39
    static LazyFrame from_csv(std::string path, std::vector<std::pair<std::string, DataType>> schema, std::optional<size_t> batch_size = std::nullopt,
                              std::optional<std::string> delimiter = std::nullopt, ContextRef ctx = nullptr);   // :41-55
    LazyFrame select(std::vector<Expr> exprs) const;   // :57-64
    LazyFrame filter(Expr predicate) const;            // :66-73
    LazyFrame limit(size_t n) const;                   // :75-82
    LazyFrame inner_join(const LazyFrame& right, std::string left_key, std::string right_key) const;   // :84-94
    DataFrame collect() const;                         // :96-104   eager engine semantics, on the GPU
    RecordBatch collect_streaming() const;             // :106-113  streaming engine semantics, on the GPU
    const LogicalPlan& logical_plan() const { return plan_; }

  private:
    LogicalPlan plan_;
    ContextRef ctx_;
};

// collect() of streaming plans: fuse the operator chain into one rvl_stream pipeline (default on; off = one operator per node)
void set_stream_fusion(bool on);

// kernels launched by the default context of `device` so far (tests assert the GPU actually ran)
int64_t launch_count(int device = 0);

}  // namespace rivulus
