// rivulus_oracle.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17) of the reference's filter / project / limit path, written to be
// audited line-by-line against the Rust under /root/reference/src (citations on every item).
// It is deliberately dumb and literal: row-of-enum `AnyValue` storage for the eager engine,
// bit-at-a-time builders for the columnar engine — the same loops the reference runs.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// include, link or execute this.  The product path (rivulus_b200/) never does.
//
// Parity status: pinned against the reference's own known-answer tests (tests/test_oracle_golden.py
// ports them, with file:line per case).  The reference cannot be compiled here (no rustc/cargo in
// the image), so semantics not covered by a reference test ("parity unpinned": nulls in the eager
// predicate column, NaN, cross-type literals, dtype collapse, Select/Limit over empty input) rest
// on code reading; each such function says so.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// datatypes/series.rs
// ---------------------------------------------------------------------------------------------

// series.rs:126-133 (declaration order kept: Int64, Float64, String, Boolean, Null)
enum class DataType : int { Int64 = 0, Float64 = 1, String = 2, Boolean = 3, Null = 4 };
const char* dtype_name(DataType d);  // Display / Debug: series.rs:162-172
bool dtype_is_numeric(DataType d);                       // series.rs:136-142
bool dtype_is_comparable_with(DataType a, DataType b);   // series.rs:144-159

// series.rs:6-13
struct AnyValue {
    enum Tag : uint8_t { kNull = 0, kInt64 = 1, kFloat64 = 2, kString = 3, kBoolean = 4 };
    Tag tag = kNull;
    int64_t i = 0;
    double f = 0.0;
    bool b = false;
    std::string s;

    static AnyValue Null() { return AnyValue(); }
    static AnyValue Int64(int64_t v) { AnyValue a; a.tag = kInt64; a.i = v; return a; }
    static AnyValue Float64(double v) { AnyValue a; a.tag = kFloat64; a.f = v; return a; }
    static AnyValue String(std::string v) { AnyValue a; a.tag = kString; a.s = std::move(v); return a; }
    static AnyValue Boolean(bool v) { AnyValue a; a.tag = kBoolean; a.b = v; return a; }

    bool is_null() const { return tag == kNull; }  // series.rs:16-18
    DataType data_type() const;                    // series.rs:20-28
    std::string debug() const;                     // #[derive(Debug)]
    std::string display() const;                   // series.rs:61-71
};

// series.rs:87-98  (PartialEq)
bool any_eq(const AnyValue& a, const AnyValue& b);
// series.rs:100-117 (PartialOrd::partial_cmp). Returns nullopt for "None", else -1/0/+1.
std::optional<int> any_partial_cmp(const AnyValue& a, const AnyValue& b);

struct OracleError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct Panic : std::runtime_error {  // Rust panics (assert!/index OOB)
    using std::runtime_error::runtime_error;
};

// series.rs:119-124, 184-265
class Series {
  public:
    static Series make(const std::string& name, std::vector<AnyValue> data);  // Series::new :185-221
    static Series empty(const std::string& name, DataType dtype);             // :223-229
    const std::string& name() const { return name_; }
    size_t len() const { return data_.size(); }
    DataType dtype() const { return dtype_; }
    const std::vector<AnyValue>& data() const { return data_; }
    const AnyValue& at(size_t i) const;  // Index<usize> :273-288 (panics OOB)

  private:
    std::string name_;
    std::vector<AnyValue> data_;
    DataType dtype_ = DataType::Null;
};

// datatypes/dataframe.rs:7-110
class DataFrame {
  public:
    static DataFrame make(std::vector<Series> columns);  // DataFrame::new :29-56
    size_t height() const { return columns_.empty() ? 0 : columns_[0].len(); }  // :64-70
    size_t width() const { return columns_.size(); }
    bool is_empty() const { return columns_.empty(); }  // :80-82 (no columns)
    const Series* column(const std::string& name) const;  // :84-86
    const std::vector<Series>& columns() const { return columns_; }
    DataFrame select(const std::vector<std::string>& names) const;  // :96-110
    static DataFrame unchecked(std::vector<Series> c) { DataFrame d; d.columns_ = std::move(c); return d; }

  private:
    std::vector<Series> columns_;
};

// ---------------------------------------------------------------------------------------------
// expressions/expr.rs
// ---------------------------------------------------------------------------------------------

// expr.rs:15-29 (declaration order)
enum class BinaryOperator : int {
    Plus = 0, Minus, Multiply, Divide, Eq, NotEq, Lt, Gt, LtEq, GtEq, And, Or
};
const char* op_name(BinaryOperator op);

// expr.rs:3-13
struct Expr {
    enum Kind { Column, Literal, Binary, Alias } kind = Column;
    std::string name;  // Column name / Alias name
    AnyValue value;    // Literal
    BinaryOperator op = BinaryOperator::Eq;
    std::shared_ptr<Expr> left, right;  // Binary; Alias inner = left

    static Expr col(const std::string& n) { Expr e; e.kind = Column; e.name = n; return e; }
    static Expr lit(AnyValue v) { Expr e; e.kind = Literal; e.value = std::move(v); return e; }
    Expr alias(const std::string& n) const {
        Expr e; e.kind = Alias; e.name = n; e.left = std::make_shared<Expr>(*this); return e;
    }
    Expr binary(BinaryOperator o, const Expr& r) const {
        Expr e; e.kind = Binary; e.op = o; e.left = std::make_shared<Expr>(*this);
        e.right = std::make_shared<Expr>(r); return e;
    }
    std::string debug() const;
};

// ---------------------------------------------------------------------------------------------
// execution/schema.rs, execution/array/*
// ---------------------------------------------------------------------------------------------

// schema.rs:1-8 (declaration order: Null, Boolean, Int64, Float64, String)
enum class ExecType : int { Null = 0, Boolean = 1, Int64 = 2, Float64 = 3, String = 4 };
const char* exec_type_name(ExecType t);

struct Field {  // schema.rs:10-36
    std::string name; ExecType data_type; bool nullable;
    bool operator==(const Field& o) const { return name == o.name && data_type == o.data_type && nullable == o.nullable; }
};
struct Schema {  // schema.rs:38-76
    std::vector<Field> fields;
    std::optional<size_t> index_of(const std::string& n) const;
    bool operator==(const Schema& o) const { return fields == o.fields; }
};
using SchemaRef = std::shared_ptr<Schema>;

// bitmap.rs:3-113
struct BitMap {
    std::shared_ptr<std::vector<uint8_t>> buffer;
    size_t bit_count = 0, offset = 0;
    static BitMap zeros(size_t n);                         // BitMap::new :11-19
    static BitMap all_true(size_t n);                      // :21-38
    static BitMap from_bools(const std::vector<bool>& v);  // :44-59
    bool get_bit(size_t index) const;                      // :61-68
    size_t count(uint8_t value, size_t off, size_t len) const;  // :74-86 (raw positions)
    BitMap slice(size_t off, size_t len) const;            // :104-112
};
// bitmap.rs:115-189
struct BitmapBuilder {
    std::vector<uint8_t> buffer; size_t bit_count = 0; uint8_t current_byte = 0; size_t current_bit_pos = 0;
    void append(bool v);      // :142-155
    bool has_nulls() const;   // :157-176
    BitMap finish();          // :178-188
};

struct Array {  // array/mod.rs:10-16
    virtual ~Array() = default;
    virtual size_t len() const = 0;
    virtual ExecType data_type() const = 0;
    virtual size_t null_count() const = 0;
    virtual std::shared_ptr<Array> slice(size_t off, size_t len) const = 0;
};
using ArrayRef = std::shared_ptr<Array>;

// primitive.rs:20-122
template <typename T> struct PrimitiveArray : Array {
    std::shared_ptr<std::vector<T>> values; std::optional<BitMap> null_bitmap; size_t offset = 0, length = 0;
    static std::shared_ptr<PrimitiveArray<T>> make(std::vector<T> v, std::optional<std::vector<bool>> validity);  // :31-42
    std::optional<T> value(size_t index) const;  // :48-60
    size_t len() const override { return length; }
    ExecType data_type() const override;
    size_t null_count() const override;          // :91-105
    ArrayRef slice(size_t off, size_t len) const override;  // :107-117
};
// primitive.rs:150-198
template <typename T> struct PrimitiveArrayBuilder {
    std::vector<T> values; BitmapBuilder null_builder;
    void append_value(T v) { null_builder.append(true); values.push_back(v); }   // :170-173
    void append_null(T ph) { null_builder.append(false); values.push_back(ph); } // :175-178
    std::shared_ptr<PrimitiveArray<T>> finish();                                 // :180-197
};

// boolean.rs:9-222
struct BooleanArray : Array {
    BitMap values; std::optional<BitMap> null_bitmap; size_t offset = 0, length = 0;
    static std::shared_ptr<BooleanArray> make(const std::vector<std::optional<bool>>& v);  // :19-50
    std::optional<bool> value(size_t index) const;  // :91-103
    size_t len() const override { return length; }
    ExecType data_type() const override { return ExecType::Boolean; }
    size_t null_count() const override;             // :191-205
    ArrayRef slice(size_t off, size_t len) const override;  // :207-217
};

// string.rs:8-190
struct StringArray : Array {
    std::shared_ptr<std::vector<uint8_t>> data; std::shared_ptr<std::vector<int32_t>> offsets;
    std::optional<BitMap> null_bitmap; size_t offset = 0, length = 0;
    static std::shared_ptr<StringArray> make(const std::vector<std::optional<std::string>>& v);  // :19-58
    std::optional<std::string> value(size_t index) const;  // :80-97
    size_t len() const override { return length; }
    ExecType data_type() const override { return ExecType::String; }
    size_t null_count() const override;             // :158-172
    ArrayRef slice(size_t off, size_t len) const override;  // :174-185
};

// null.rs:5-67
struct NullArray : Array {
    size_t length = 0, offset = 0;
    explicit NullArray(size_t n) : length(n) {}
    size_t len() const override { return length; }
    ExecType data_type() const override { return ExecType::Null; }
    size_t null_count() const override { return length; }
    ArrayRef slice(size_t off, size_t len) const override;
};

// record_batch.rs:8-422
struct RecordBatch {
    SchemaRef schema; std::vector<ArrayRef> columns; size_t num_rows = 0;
    static RecordBatch try_new(SchemaRef schema, std::vector<ArrayRef> cols);   // :16-58  (throws OracleError(String))
    static RecordBatch new_unchecked(SchemaRef schema, std::vector<ArrayRef> cols, size_t num_rows);   // :60-66
    void validate() const;                                                      // :348-378 (throws OracleError(String))
    size_t memory_size() const;                                                 // :380-400
    ArrayRef column_by_name(const std::string& name) const;                     // :84-86 (nullptr = None)
    RecordBatch slice(size_t off, size_t len) const;                            // :92-106 (panics OOB)
    RecordBatch take(const std::vector<size_t>& idx) const;                     // :108-129
    RecordBatch select_columns(const std::vector<size_t>& idx) const;           // :180-206
    RecordBatch select_columns_by_name(const std::vector<std::string>& n) const;// :208-219
    RecordBatch filter(const ArrayRef& predicate) const;                        // :221-243
    static RecordBatch concat(const std::vector<RecordBatch>& batches);         // :245-275
    static RecordBatch empty(SchemaRef schema);                                 // :402-421
};
// record_batch.rs:495-573
struct RecordBatchBuilder {
    SchemaRef schema; std::vector<ArrayRef> columns;
    explicit RecordBatchBuilder(SchemaRef s) : schema(std::move(s)) {}           // new :501-507, with_capacity :509-516
    void add_column(ArrayRef column);                                            // :518-546 (throws OracleError(String))
    RecordBatch finish() const;                                                  // :548-558
    size_t num_columns() const { return columns.size(); }                        // :560-562
    bool is_complete() const { return columns.size() == schema->fields.size(); } // :564-566
};
ArrayRef take_array(const ArrayRef& a, const std::vector<size_t>& idx);         // :131-178
ArrayRef concat_arrays(const std::vector<ArrayRef>& arrays);                    // :277-342

// stream.rs:25-213, streaming.rs:246-288
struct DataStream {
    virtual ~DataStream() = default;
    virtual SchemaRef schema() const = 0;
    virtual std::optional<RecordBatch> next_batch() = 0;
};
using DataStreamRef = std::unique_ptr<DataStream>;
DataStreamRef memory_stream(SchemaRef schema, std::vector<RecordBatch> batches);   // stream.rs:66-81
DataStreamRef filter_stream(DataStreamRef in, std::string predicate_column);       // stream.rs:123-162
DataStreamRef select_stream(DataStreamRef in, std::vector<std::string> columns);   // stream.rs:173-212
DataStreamRef limit_stream(DataStreamRef in, size_t limit);                        // streaming.rs:254-287
std::vector<RecordBatch> collect_all_batches(DataStream& s);                       // streaming.rs:335-341
RecordBatch collect_stream_batches(DataStream& s);                                 // streaming.rs:343-352

// streaming.rs:135-233
std::vector<RecordBatch> dataframe_to_batches(const DataFrame& df, size_t batch_size);

// execution/file_stream.rs (csv_oracle.cpp).  SURVEY.md 8(f) rank 3.
size_t calculate_adaptive_batch_size(const Schema& schema);   // :346-369
// CsvFileStream::new :20-40 (throws OracleError("Failed to open file: …")); delimiter = one UTF-8 encoded char, default ','
DataStreamRef csv_file_stream(const std::string& path, SchemaRef schema, std::optional<size_t> batch_size, std::optional<std::string> delimiter);
// true = reproduce the reference's inverted validity of Int64 / Float64 columns holding a null (:213-240, :245-272); default false
void set_csv_reference_validity(bool on);
bool csv_reference_validity();

// ---------------------------------------------------------------------------------------------
// logical_plan/*, physical_plan/planner.rs, physical_plan/plan.rs, streaming_planner.rs
// ---------------------------------------------------------------------------------------------

struct LogicalPlan {  // logical_plan/plan.rs:8-39
    enum Kind { DataFrameSource, Select, Filter, Limit, CsvFileSource, Join } kind = DataFrameSource;
    std::shared_ptr<LogicalPlan> right;                   // Join :32-38 (`input` is the left side); JoinType::Inner only
    std::string left_key, right_key;
    DataFrame df;                                         // DataFrameSource
    std::vector<std::pair<std::string, DataType>> src_schema;   // DataFrameSource, CsvFileSource
    std::string csv_path; std::optional<size_t> csv_batch_size; std::optional<std::string> csv_delimiter;  // CsvFileSource :14-19
    std::shared_ptr<LogicalPlan> input;
    std::vector<Expr> expressions;                        // Select
    Expr predicate;                                       // Filter
    size_t n = 0;                                         // Limit

    std::vector<std::pair<std::string, DataType>> schema() const;  // :63-113
    void validate() const;                                         // :115-202 (throws LogicalPlanError text)
};
LogicalPlan optimize(LogicalPlan plan);  // optimizer.rs:7-64

DataFrame execute_eager(const LogicalPlan& optimized);       // planner.rs:41-189 + physical_plan/plan.rs:65-173

// physical_plan/streaming.rs:28-133, 235-243, 290-333
// ---------------------------------------------------------------------------------------------
// OPT-IN EXTENSION, off by default: both reference executors reject And / Or predicates
// (planner.rs:146-150) and the streaming planner rejects every comparison (streaming_planner.rs:137-168).  With
// set_extensions(true) a filter predicate may be a tree of And / Or over `column <op> literal` leaves, in collect() and in
// collect_streaming().  Each leaf is evaluated with the eager truth table (plan.rs:114-120 over series.rs:87-117: a Null row
// is Less than any literal, different types never compare) and yields true / false; And / Or combine those two-valued results.
// The streaming engine evaluates the leaves on the RecordBatch arrays as dataframe_to_batches left them (numeric / Boolean
// nulls already flattened to 0 / false, streaming.rs:177,188,212).  There is no reference behaviour to match here: this
// oracle restates the DEFINITION above so the GPU host layer can be checked against it.
// ---------------------------------------------------------------------------------------------
void set_extensions(bool on);
bool extensions_enabled();

struct StreamingPhysicalPlan {
    enum Kind { MemorySource, DataFrameSource, Filter, Select, Limit, FilterExpr, CsvFileSource, HashJoin } kind = MemorySource;
    Expr predicate;                            // FilterExpr (extension)
    std::string csv_path; SchemaRef csv_schema; std::optional<size_t> csv_batch_size; std::optional<std::string> csv_delimiter;  // CsvFileSource :39-44
    std::vector<RecordBatch> batches;          // MemorySource
    DataFrame df; size_t batch_size = 0;       // DataFrameSource
    std::shared_ptr<StreamingPhysicalPlan> input;
    std::string predicate_column;              // Filter
    std::vector<std::string> columns;          // Select
    size_t n = 0;                              // Limit
    static StreamingPhysicalPlan memory_source(std::vector<RecordBatch> b);
    static StreamingPhysicalPlan dataframe_source(DataFrame df, size_t batch_size);
    static StreamingPhysicalPlan csv_file_source(std::string path, SchemaRef schema, std::optional<size_t> batch_size,
                                                 std::optional<std::string> delimiter);   // :299-311
    StreamingPhysicalPlan filter(std::string col) const;
    StreamingPhysicalPlan filter_expr(Expr predicate) const;   // extension
    StreamingPhysicalPlan select(std::vector<std::string> cols) const;
    StreamingPhysicalPlan limit(size_t n) const;
    DataStreamRef execute() const;                    // :70-133
    RecordBatch collect() const;                      // :235-238
    std::vector<RecordBatch> collect_batches() const; // :240-243
};
StreamingPhysicalPlan logical_to_streaming(const LogicalPlan& plan);  // streaming_planner.rs:29-168

// builder.rs:26-114.  Errors are thrown as OracleError whose what() is the QueryError Display text.
struct LazyFrame {
    LogicalPlan plan;
    static LazyFrame from_dataframe(const DataFrame& df);  // :27-39
    static LazyFrame from_csv(std::string path, std::vector<std::pair<std::string, DataType>> schema, std::optional<size_t> batch_size,
                              std::optional<std::string> delimiter);   // :41-55
    LazyFrame select(std::vector<Expr> e) const;           // :57-64
    LazyFrame filter(Expr p) const;                        // :66-73
    LazyFrame limit(size_t n) const;                       // :75-82
    LazyFrame inner_join(const LazyFrame& right, std::string left_key, std::string right_key) const;   // :84-94
    DataFrame collect() const;                             // :96-104
    RecordBatch collect_streaming() const;                 // :106-113
};

// ---------------------------------------------------------------------------------------------
// Extension used to check the FUSED GPU operator: comparison predicate evaluated with the eager
// truth table (plan.rs:112-130 over series.rs:87-117) on columnar arrays, then the streaming
// engine's own filter/select/limit/concat kernels (record_batch.rs) produce the buffers.
// ---------------------------------------------------------------------------------------------
AnyValue array_value(const Array& a, size_t i);   // `value(i)` of each array type mapped to AnyValue
bool eval_cmp(const AnyValue& row, BinaryOperator op, const AnyValue& lit);  // plan.rs:114-120
RecordBatch filter_project_cmp(const RecordBatch& in, size_t pred_col, BinaryOperator op, const AnyValue& lit,
                               const std::vector<size_t>& proj, int64_t limit /* <0: none */);
RecordBatch filter_project_mask(const RecordBatch& in, size_t mask_col, const std::vector<size_t>& proj, int64_t limit);

}  // namespace orc
