// oracle_capi.cpp — TEST INFRASTRUCTURE ONLY.  extern "C" surface over rivulus_oracle.{hpp,cpp}
// so pytest (ctypes) can drive the CPU restatement, plus the timed CPU-baseline entry points
// bench.py uses (`cpu_baseline`, `--impl reference`).  Never linked into the product library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../include/rivulus_synth.h"
#include "rivulus_oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct DfBuilder { std::vector<Series> cols; };
template <typename F> int guard(F&& f) {
    try { f(); g_err.clear(); return 0; }
    catch (const Panic& p) { g_err = std::string("panic: ") + p.what(); return 2; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
struct ColExport {  // flat view of one RecordBatch column (pointers live as long as the batch handle)
    int32_t dtype; int64_t length; int64_t offset; int64_t null_count;
    const void* values; int64_t values_len;      // elements (i64/f64) or bytes (bool value bitmap)
    const uint8_t* validity; int64_t validity_len;  // bytes; NULL when bitmap absent
    const int32_t* offsets; int64_t offsets_len;
    const uint8_t* data; int64_t data_len;
};
BitMap bitmap_from_bytes(const uint8_t* bits, int64_t nbits) {
    BitMap b; b.buffer = std::make_shared<std::vector<uint8_t>>(bits, bits + (nbits + 7) / 8); b.bit_count = (size_t)nbits;
    return b;
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_free(void* p, int kind);

// ----------------------------------------------------------------------------------- DataFrame
void* orc_dfb_new() { return new DfBuilder(); }
// tags[i] = AnyValue tag (0 Null, 1 Int64, 2 Float64, 3 String, 4 Boolean); unused payload arrays may be NULL.
int orc_dfb_add_series(void* h, const char* name, int64_t n, const uint8_t* tags, const int64_t* i64, const double* f64,
                       const uint8_t* b8, const int32_t* str_off, const uint8_t* str_data) {
    return guard([&] {
        std::vector<AnyValue> d; d.reserve((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            switch (tags[i]) {
                case 0: d.push_back(AnyValue::Null()); break;
                case 1: d.push_back(AnyValue::Int64(i64[i])); break;
                case 2: d.push_back(AnyValue::Float64(f64[i])); break;
                case 3: d.push_back(AnyValue::String(std::string((const char*)str_data + str_off[i], (size_t)(str_off[i + 1] - str_off[i])))); break;
                case 4: d.push_back(AnyValue::Boolean(b8[i] != 0)); break;
                default: throw OracleError("bad tag");
            }
        }
        ((DfBuilder*)h)->cols.push_back(Series::make(name, std::move(d)));
    });
}
int orc_dfb_add_empty_series(void* h, const char* name, int dtype) {
    return guard([&] { ((DfBuilder*)h)->cols.push_back(Series::empty(name, (DataType)dtype)); });
}
int orc_dfb_finish(void* h, void** out) {
    auto* b = (DfBuilder*)h;
    int rc = guard([&] { *out = new DataFrame(DataFrame::make(std::move(b->cols))); });
    delete b;
    return rc;
}
void orc_df_free(void* df) { delete (DataFrame*)df; }
int orc_df_width(void* df) { return (int)((DataFrame*)df)->width(); }
int64_t orc_df_height(void* df) { return (int64_t)((DataFrame*)df)->height(); }
const char* orc_df_col_name(void* df, int i) { return ((DataFrame*)df)->columns()[i].name().c_str(); }
int orc_df_col_dtype(void* df, int i) { return (int)((DataFrame*)df)->columns()[i].dtype(); }
int64_t orc_df_col_len(void* df, int i) { return (int64_t)((DataFrame*)df)->columns()[i].len(); }
int64_t orc_df_col_str_bytes(void* df, int i) {
    int64_t t = 0;
    for (const auto& v : ((DataFrame*)df)->columns()[i].data()) if (v.tag == AnyValue::kString) t += (int64_t)v.s.size();
    return t;
}
// export one Series into caller arrays (each sized len; str_off sized len+1; str_data sized orc_df_col_str_bytes)
void orc_df_col_export(void* df, int i, uint8_t* tags, int64_t* i64, double* f64, uint8_t* b8, int32_t* str_off, uint8_t* str_data) {
    const auto& d = ((DataFrame*)df)->columns()[i].data();
    int32_t pos = 0;
    for (size_t r = 0; r < d.size(); ++r) {
        const auto& v = d[r];
        tags[r] = (uint8_t)v.tag; i64[r] = v.i; f64[r] = v.f; b8[r] = v.b ? 1 : 0;
        str_off[r] = pos;
        if (v.tag == AnyValue::kString) { std::memcpy(str_data + pos, v.s.data(), v.s.size()); pos += (int32_t)v.s.size(); }
    }
    str_off[d.size()] = pos;
}

// ----------------------------------------------------------------------------------- Expr / LazyFrame
void* orc_expr_col(const char* name) { return new Expr(Expr::col(name)); }
void* orc_expr_lit(int tag, int64_t i, double f, const char* s, int64_t slen, int b) {
    AnyValue v;
    switch (tag) {
        case 1: v = AnyValue::Int64(i); break;
        case 2: v = AnyValue::Float64(f); break;
        case 3: v = AnyValue::String(std::string(s, (size_t)slen)); break;
        case 4: v = AnyValue::Boolean(b != 0); break;
        default: v = AnyValue::Null();
    }
    return new Expr(Expr::lit(v));
}
void* orc_expr_binary(void* l, int op, void* r) { return new Expr(((Expr*)l)->binary((BinaryOperator)op, *(Expr*)r)); }
void* orc_expr_alias(void* e, const char* name) { return new Expr(((Expr*)e)->alias(name)); }
void orc_expr_free(void* e) { delete (Expr*)e; }

void* orc_lf_from_df(void* df) { return new LazyFrame(LazyFrame::from_dataframe(*(DataFrame*)df)); }
void* orc_lf_from_csv(const char* path, int nfields, const char** names, const int* dtypes, int64_t batch_size, const char* delimiter) {
    std::vector<std::pair<std::string, DataType>> schema;
    for (int i = 0; i < nfields; ++i) schema.emplace_back(names[i], (DataType)dtypes[i]);
    return new LazyFrame(LazyFrame::from_csv(path, std::move(schema), batch_size < 0 ? std::nullopt : std::optional<size_t>((size_t)batch_size),
                                             delimiter ? std::optional<std::string>(delimiter) : std::nullopt));
}
void orc_set_csv_reference_validity(int on) { set_csv_reference_validity(on != 0); }
int64_t orc_csv_adaptive_batch_size(int nfields, const int* exec_dtypes) {
    Schema s;
    for (int i = 0; i < nfields; ++i) s.fields.push_back(Field{"c" + std::to_string(i), (ExecType)exec_dtypes[i], true});
    return (int64_t)calculate_adaptive_batch_size(s);
}
void* orc_lf_select(void* lf, int n, void** exprs) {
    std::vector<Expr> e; for (int i = 0; i < n; ++i) e.push_back(*(Expr*)exprs[i]);
    return new LazyFrame(((LazyFrame*)lf)->select(std::move(e)));
}
void* orc_lf_filter(void* lf, void* pred) { return new LazyFrame(((LazyFrame*)lf)->filter(*(Expr*)pred)); }
void* orc_lf_inner_join(void* lf, void* right, const char* left_key, const char* right_key) {
    return new LazyFrame(((LazyFrame*)lf)->inner_join(*(LazyFrame*)right, left_key, right_key));
}
void* orc_lf_limit(void* lf, int64_t n) { return new LazyFrame(((LazyFrame*)lf)->limit((size_t)n)); }
void orc_lf_free(void* lf) { delete (LazyFrame*)lf; }
void orc_set_extensions(int on) { set_extensions(on != 0); }
int orc_dtype_is_numeric(int d) { return dtype_is_numeric((DataType)d) ? 1 : 0; }
int orc_dtype_is_comparable_with(int a, int b) { return dtype_is_comparable_with((DataType)a, (DataType)b) ? 1 : 0; }
int orc_lf_collect(void* lf, void** df_out) {
    return guard([&] { *df_out = new DataFrame(((LazyFrame*)lf)->collect()); });
}
int orc_lf_collect_streaming(void* lf, void** rb_out) {
    return guard([&] { *rb_out = new RecordBatch(((LazyFrame*)lf)->collect_streaming()); });
}
// logical_plan/plan.rs probes on the plan as built (not optimized): schema() as "name:Dtype,...", validate(), and a Debug-style dump
static std::string describe_plan(const LogicalPlan& p) {
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return "DataFrameSource";
        case LogicalPlan::CsvFileSource: return "CsvFileSource { path: \"" + p.csv_path + "\" }";
        case LogicalPlan::Join:
            return "Join { left: " + describe_plan(*p.input) + ", right: " + describe_plan(*p.right) + ", left_key: \"" + p.left_key + "\", right_key: \"" +
                   p.right_key + "\", join_type: Inner }";
        case LogicalPlan::Select: {
            std::string e;
            for (size_t i = 0; i < p.expressions.size(); ++i) e += (i ? ", " : "") + p.expressions[i].debug();
            return "Select { input: " + describe_plan(*p.input) + ", expressions: [" + e + "] }";
        }
        case LogicalPlan::Filter: return "Filter { input: " + describe_plan(*p.input) + ", predicate: " + p.predicate.debug() + " }";
        case LogicalPlan::Limit: return "Limit { input: " + describe_plan(*p.input) + ", n: " + std::to_string(p.n) + " }";
    }
    return "";
}
int orc_lf_schema(void* lf, char* buf, int cap) {
    return guard([&] {
        std::string s;
        for (const auto& pr : ((LazyFrame*)lf)->plan.schema()) s += (s.empty() ? "" : ",") + pr.first + ":" + dtype_name(pr.second);
        std::snprintf(buf, (size_t)cap, "%s", s.c_str());
    });
}
int orc_lf_validate(void* lf) { return guard([&] { ((LazyFrame*)lf)->plan.validate(); }); }
int orc_lf_describe(void* lf, char* buf, int cap) {
    return guard([&] { std::snprintf(buf, (size_t)cap, "%s", describe_plan(((LazyFrame*)lf)->plan).c_str()); });
}
// optimizer shape probe for tests: returns the optimized plan as "Filter(Select(Source))"-style text
static std::string plan_shape(const LogicalPlan& p) {
    switch (p.kind) {
        case LogicalPlan::DataFrameSource: return "Source";
        case LogicalPlan::CsvFileSource: return "CsvSource";
        case LogicalPlan::Select: return "Select(" + plan_shape(*p.input) + ")";
        case LogicalPlan::Filter: return "Filter(" + plan_shape(*p.input) + ")";
        case LogicalPlan::Limit: return "Limit(" + plan_shape(*p.input) + ")";
        case LogicalPlan::Join: return "Join(" + plan_shape(*p.input) + ", " + plan_shape(*p.right) + ")";
    }
    return "";
}
int orc_lf_plan_shape(void* lf, char* buf, int cap) {
    const std::string s = plan_shape(optimize(((LazyFrame*)lf)->plan));
    std::snprintf(buf, (size_t)cap, "%s", s.c_str());
    return (int)s.size();
}

// ----------------------------------------------------------------------------------- Arrays / RecordBatch
// validity_bits: packed LSB-first over the WHOLE buffer (nvalues bits) or NULL; (offset,length) = view.
void* orc_arr_i64(const int64_t* v, int64_t nvalues, const uint8_t* validity_bits, int64_t offset, int64_t length) {
    auto a = std::make_shared<PrimitiveArray<int64_t>>();
    a->values = std::make_shared<std::vector<int64_t>>(v, v + nvalues);
    if (validity_bits) a->null_bitmap = bitmap_from_bytes(validity_bits, nvalues);
    a->offset = (size_t)offset; a->length = (size_t)length;
    return new ArrayRef(a);
}
void* orc_arr_f64(const double* v, int64_t nvalues, const uint8_t* validity_bits, int64_t offset, int64_t length) {
    auto a = std::make_shared<PrimitiveArray<double>>();
    a->values = std::make_shared<std::vector<double>>(v, v + nvalues);
    if (validity_bits) a->null_bitmap = bitmap_from_bytes(validity_bits, nvalues);
    a->offset = (size_t)offset; a->length = (size_t)length;
    return new ArrayRef(a);
}
void* orc_arr_bool(const uint8_t* value_bits, int64_t nbits, const uint8_t* validity_bits, int64_t offset, int64_t length) {
    auto a = std::make_shared<BooleanArray>();
    a->values = bitmap_from_bytes(value_bits, nbits);
    if (validity_bits) a->null_bitmap = bitmap_from_bytes(validity_bits, nbits);
    a->offset = (size_t)offset; a->length = (size_t)length;
    return new ArrayRef(a);
}
void* orc_arr_str(const int32_t* offsets, int64_t n_strings, const uint8_t* data, int64_t data_len,
                  const uint8_t* validity_bits, int64_t offset, int64_t length) {
    auto a = std::make_shared<StringArray>();
    a->offsets = std::make_shared<std::vector<int32_t>>(offsets, offsets + n_strings + 1);
    a->data = std::make_shared<std::vector<uint8_t>>(data, data + data_len);
    if (validity_bits) a->null_bitmap = bitmap_from_bytes(validity_bits, n_strings);
    a->offset = (size_t)offset; a->length = (size_t)length;
    return new ArrayRef(a);
}
void* orc_arr_null(int64_t length) { return new ArrayRef(std::make_shared<NullArray>((size_t)length)); }
// reference constructors (exercise the builders): PrimitiveArray::new / BooleanArray::new / StringArray::new
void* orc_arr_i64_new(const int64_t* v, int64_t n, const uint8_t* valid8) {
    std::optional<std::vector<bool>> val;
    if (valid8) { val.emplace(); for (int64_t i = 0; i < n; ++i) val->push_back(valid8[i] != 0); }
    return new ArrayRef(PrimitiveArray<int64_t>::make(std::vector<int64_t>(v, v + n), val));
}
void* orc_arr_f64_new(const double* v, int64_t n, const uint8_t* valid8) {
    std::optional<std::vector<bool>> val;
    if (valid8) { val.emplace(); for (int64_t i = 0; i < n; ++i) val->push_back(valid8[i] != 0); }
    return new ArrayRef(PrimitiveArray<double>::make(std::vector<double>(v, v + n), val));
}
void* orc_arr_bool_new(const uint8_t* b8, int64_t n, const uint8_t* valid8) {
    std::vector<std::optional<bool>> v;
    for (int64_t i = 0; i < n; ++i) { if (valid8 && !valid8[i]) v.push_back(std::nullopt); else v.push_back(b8[i] != 0); }
    return new ArrayRef(BooleanArray::make(v));
}
int orc_arr_str_new(const int32_t* off, const uint8_t* data, int64_t n, const uint8_t* valid8, void** out) {
    return guard([&] {
        std::vector<std::optional<std::string>> v;
        for (int64_t i = 0; i < n; ++i) {
            if (valid8 && !valid8[i]) v.push_back(std::nullopt);
            else v.push_back(std::string((const char*)data + off[i], (size_t)(off[i + 1] - off[i])));
        }
        *out = new ArrayRef(StringArray::make(v));
    });
}
int orc_arr_slice(void* a, int64_t off, int64_t len, void** out) {
    return guard([&] { *out = new ArrayRef((*(ArrayRef*)a)->slice((size_t)off, (size_t)len)); });
}
void orc_arr_free(void* a) { delete (ArrayRef*)a; }
int64_t orc_arr_len(void* a) { return (int64_t)(*(ArrayRef*)a)->len(); }
int64_t orc_arr_null_count(void* a) { return (int64_t)(*(ArrayRef*)a)->null_count(); }

static void export_array(const Array& arr, ColExport* e) {
    std::memset(e, 0, sizeof *e);
    e->dtype = (int32_t)arr.data_type(); e->length = (int64_t)arr.len(); e->null_count = (int64_t)arr.null_count();
    auto put_validity = [&](const std::optional<BitMap>& bm) {
        if (bm) { e->validity = bm->buffer->data(); e->validity_len = (int64_t)bm->buffer->size(); }
    };
    switch (arr.data_type()) {
        case ExecType::Int64: { auto& p = dynamic_cast<const PrimitiveArray<int64_t>&>(arr);
            e->values = p.values->data(); e->values_len = (int64_t)p.values->size(); e->offset = (int64_t)p.offset; put_validity(p.null_bitmap); break; }
        case ExecType::Float64: { auto& p = dynamic_cast<const PrimitiveArray<double>&>(arr);
            e->values = p.values->data(); e->values_len = (int64_t)p.values->size(); e->offset = (int64_t)p.offset; put_validity(p.null_bitmap); break; }
        case ExecType::Boolean: { auto& p = dynamic_cast<const BooleanArray&>(arr);
            e->values = p.values.buffer->data(); e->values_len = (int64_t)p.values.buffer->size(); e->offset = (int64_t)p.offset; put_validity(p.null_bitmap); break; }
        case ExecType::String: { auto& p = dynamic_cast<const StringArray&>(arr);
            e->offsets = p.offsets->data(); e->offsets_len = (int64_t)p.offsets->size(); e->data = p.data->data(); e->data_len = (int64_t)p.data->size();
            e->offset = (int64_t)p.offset; put_validity(p.null_bitmap); break; }
        case ExecType::Null: { auto& p = dynamic_cast<const NullArray&>(arr); e->offset = (int64_t)p.offset; break; }
    }
}
void orc_arr_export(void* a, ColExport* e) { export_array(**(ArrayRef*)a, e); }

int orc_rb_new(int n, const char** names, void** arrays, void** out) {
    return guard([&] {
        auto s = std::make_shared<Schema>();
        std::vector<ArrayRef> cols;
        for (int i = 0; i < n; ++i) { cols.push_back(*(ArrayRef*)arrays[i]); s->fields.push_back(Field{names[i], cols.back()->data_type(), true}); }
        *out = new RecordBatch(RecordBatch::try_new(s, std::move(cols)));
    });
}
// try_new against an explicit schema (type-mismatch / count-mismatch tests)
int orc_rb_try_new(int nfields, const char** names, const int* dtypes, int ncols, void** arrays, void** out) {
    return guard([&] {
        auto s = std::make_shared<Schema>();
        for (int i = 0; i < nfields; ++i) s->fields.push_back(Field{names[i], (ExecType)dtypes[i], true});
        std::vector<ArrayRef> cols;
        for (int i = 0; i < ncols; ++i) cols.push_back(*(ArrayRef*)arrays[i]);
        *out = new RecordBatch(RecordBatch::try_new(s, std::move(cols)));
    });
}
int orc_rb_new_unchecked(int nfields, const char** names, const int* dtypes, int ncols, void** arrays, int64_t num_rows, void** out) {
    return guard([&] {
        auto s = std::make_shared<Schema>();
        for (int i = 0; i < nfields; ++i) s->fields.push_back(Field{names[i], (ExecType)dtypes[i], true});
        std::vector<ArrayRef> cols;
        for (int i = 0; i < ncols; ++i) cols.push_back(*(ArrayRef*)arrays[i]);
        *out = new RecordBatch(RecordBatch::new_unchecked(s, std::move(cols), (size_t)num_rows));
    });
}
int orc_rb_validate(void* rb) { return guard([&] { ((RecordBatch*)rb)->validate(); }); }
int64_t orc_rb_memory_size(void* rb) { return (int64_t)((RecordBatch*)rb)->memory_size(); }
void* orc_rbb_new(int nfields, const char** names, const int* dtypes) {
    auto s = std::make_shared<Schema>();
    for (int i = 0; i < nfields; ++i) s->fields.push_back(Field{names[i], (ExecType)dtypes[i], true});
    return new RecordBatchBuilder(s);
}
int orc_rbb_add_column(void* b, void* array) { return guard([&] { ((RecordBatchBuilder*)b)->add_column(*(ArrayRef*)array); }); }
int orc_rbb_finish(void* b, void** out) { return guard([&] { *out = new RecordBatch(((RecordBatchBuilder*)b)->finish()); }); }
int orc_rbb_num_columns(void* b) { return (int)((RecordBatchBuilder*)b)->num_columns(); }
int orc_rbb_is_complete(void* b) { return ((RecordBatchBuilder*)b)->is_complete() ? 1 : 0; }
void orc_rbb_free(void* b) { delete (RecordBatchBuilder*)b; }
void orc_rb_free(void* rb) { delete (RecordBatch*)rb; }
int64_t orc_rb_num_rows(void* rb) { return (int64_t)((RecordBatch*)rb)->num_rows; }
int orc_rb_num_columns(void* rb) { return (int)((RecordBatch*)rb)->columns.size(); }
const char* orc_rb_col_name(void* rb, int i) { return ((RecordBatch*)rb)->schema->fields[i].name.c_str(); }
void orc_rb_col_export(void* rb, int i, ColExport* e) { export_array(*((RecordBatch*)rb)->columns[i], e); }
int orc_rb_slice(void* rb, int64_t off, int64_t len, void** out) {
    return guard([&] { *out = new RecordBatch(((RecordBatch*)rb)->slice((size_t)off, (size_t)len)); });
}
int orc_rb_take(void* rb, const int64_t* idx, int64_t n, void** out) {
    return guard([&] { std::vector<size_t> v(idx, idx + n); *out = new RecordBatch(((RecordBatch*)rb)->take(v)); });
}
int orc_rb_filter(void* rb, void* mask_arr, void** out) {
    return guard([&] { *out = new RecordBatch(((RecordBatch*)rb)->filter(*(ArrayRef*)mask_arr)); });
}
int orc_rb_select(void* rb, const int32_t* idx, int n, void** out) {
    return guard([&] { std::vector<size_t> v(idx, idx + n); *out = new RecordBatch(((RecordBatch*)rb)->select_columns(v)); });
}
int orc_rb_select_by_name(void* rb, const char** names, int n, void** out) {
    return guard([&] { std::vector<std::string> v(names, names + n); *out = new RecordBatch(((RecordBatch*)rb)->select_columns_by_name(v)); });
}
int orc_rb_concat(void** rbs, int n, void** out) {
    return guard([&] { std::vector<RecordBatch> v; for (int i = 0; i < n; ++i) v.push_back(*(RecordBatch*)rbs[i]); *out = new RecordBatch(RecordBatch::concat(v)); });
}
int orc_rb_empty_like(void* rb, void** out) {
    return guard([&] { *out = new RecordBatch(RecordBatch::empty(((RecordBatch*)rb)->schema)); });
}
int orc_rb_filter_project_cmp(void* rb, int pred_col, int op, int lit_tag, int64_t li, double lf, const char* ls, int64_t lslen, int lb,
                              const int32_t* proj, int nproj, int64_t limit, void** out) {
    return guard([&] {
        AnyValue v;
        switch (lit_tag) { case 1: v = AnyValue::Int64(li); break; case 2: v = AnyValue::Float64(lf); break;
                           case 3: v = AnyValue::String(std::string(ls, (size_t)lslen)); break; case 4: v = AnyValue::Boolean(lb != 0); break; default: break; }
        std::vector<size_t> p(proj, proj + nproj);
        *out = new RecordBatch(filter_project_cmp(*(RecordBatch*)rb, (size_t)pred_col, (BinaryOperator)op, v, p, limit));
    });
}
int orc_rb_filter_project_mask(void* rb, int mask_col, const int32_t* proj, int nproj, int64_t limit, void** out) {
    return guard([&] { std::vector<size_t> p(proj, proj + nproj); *out = new RecordBatch(filter_project_mask(*(RecordBatch*)rb, (size_t)mask_col, p, limit)); });
}

// ----------------------------------------------------------------------------------- streaming plans
void* orc_sp_memory_source(void** rbs, int n) {
    std::vector<RecordBatch> v; for (int i = 0; i < n; ++i) v.push_back(*(RecordBatch*)rbs[i]);
    return new StreamingPhysicalPlan(StreamingPhysicalPlan::memory_source(std::move(v)));
}
void* orc_sp_dataframe_source(void* df, int64_t batch_size) {
    return new StreamingPhysicalPlan(StreamingPhysicalPlan::dataframe_source(*(DataFrame*)df, (size_t)batch_size));
}
void* orc_sp_csv_source(const char* path, int nfields, const char** names, const int* exec_dtypes, const int* nullable, int64_t batch_size,
                        const char* delimiter) {
    auto schema = std::make_shared<Schema>();
    for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{names[i], (ExecType)exec_dtypes[i], nullable[i] != 0});
    return new StreamingPhysicalPlan(StreamingPhysicalPlan::csv_file_source(path, schema, batch_size < 0 ? std::nullopt : std::optional<size_t>((size_t)batch_size),
                                                                            delimiter ? std::optional<std::string>(delimiter) : std::nullopt));
}
void* orc_sp_filter(void* sp, const char* col) { return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->filter(col)); }
void* orc_sp_select(void* sp, const char** names, int n) {
    return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->select(std::vector<std::string>(names, names + n)));
}
void* orc_sp_limit(void* sp, int64_t n) { return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->limit((size_t)n)); }
void orc_sp_free(void* sp) { delete (StreamingPhysicalPlan*)sp; }
int orc_sp_collect(void* sp, void** rb_out) {
    return guard([&] { *rb_out = new RecordBatch(((StreamingPhysicalPlan*)sp)->collect()); });
}
int orc_sp_collect_batches(void* sp, void** vec_out) {
    return guard([&] { *vec_out = new std::vector<RecordBatch>(((StreamingPhysicalPlan*)sp)->collect_batches()); });
}
int orc_rbv_len(void* v) { return (int)((std::vector<RecordBatch>*)v)->size(); }
void* orc_rbv_get(void* v, int i) { return new RecordBatch((*(std::vector<RecordBatch>*)v)[i]); }
void orc_rbv_free(void* v) { delete (std::vector<RecordBatch>*)v; }

// ----------------------------------------------------------------------------------- scalar semantics probes
// any_eq / partial_cmp on scalars (truth-table tests, series.rs:349-366)
static AnyValue mk(int tag, int64_t i, double f, const char* s, int b) {
    switch (tag) { case 1: return AnyValue::Int64(i); case 2: return AnyValue::Float64(f); case 3: return AnyValue::String(s ? s : "");
                   case 4: return AnyValue::Boolean(b != 0); default: return AnyValue::Null(); }
}
int orc_any_eq(int ta, int64_t ia, double fa, const char* sa, int ba, int tb, int64_t ib, double fb, const char* sb, int bb) {
    return any_eq(mk(ta, ia, fa, sa, ba), mk(tb, ib, fb, sb, bb)) ? 1 : 0;
}
// returns -1/0/1, or 2 for None
int orc_any_partial_cmp(int ta, int64_t ia, double fa, const char* sa, int ba, int tb, int64_t ib, double fb, const char* sb, int bb) {
    auto c = any_partial_cmp(mk(ta, ia, fa, sa, ba), mk(tb, ib, fb, sb, bb));
    return c ? *c : 2;
}
int orc_eval_cmp(int ta, int64_t ia, double fa, const char* sa, int ba, int op, int tb, int64_t ib, double fb, const char* sb, int bb) {
    return eval_cmp(mk(ta, ia, fa, sa, ba), (BinaryOperator)op, mk(tb, ib, fb, sb, bb)) ? 1 : 0;
}

// ----------------------------------------------------------------------------------- synthetic tables + timed CPU baselines
// Build the BASELINE config tables as AnyValue DataFrames (what the reference's eager engine consumes).
//   cols: kinds[] from rvl_synth_kind, col_ids[] = generator column ids, null_pct per column.
static Series synth_series(const char* name, int kind, uint32_t col_id, uint64_t row0, int64_t n, uint32_t null_pct) {
    std::vector<AnyValue> d; d.reserve((size_t)n);
    for (int64_t r = 0; r < n; ++r) {
        uint64_t row = row0 + (uint64_t)r;
        if (!rvl_synth_valid(RVL_SYNTH_SEED, col_id, row, null_pct)) { d.push_back(AnyValue::Null()); continue; }
        uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_id, row);
        switch (kind) {
            case RVL_SYNTH_KEY1000: case RVL_SYNTH_I64: case RVL_SYNTH_AGE100: d.push_back(AnyValue::Int64(rvl_synth_i64(u, kind))); break;
            case RVL_SYNTH_F64: d.push_back(AnyValue::Float64(rvl_synth_f64(u))); break;
            case RVL_SYNTH_BOOL: d.push_back(AnyValue::Boolean((u & 1) != 0)); break;
            case RVL_SYNTH_STR: {
                uint32_t len = rvl_synth_strlen(u); std::string s(len, 'a');
                for (uint32_t j = 0; j < len; ++j) s[j] = (char)rvl_synth_strbyte(u, j);
                d.push_back(AnyValue::String(std::move(s))); break;
            }
        }
    }
    return Series::make(name, std::move(d));
}
int orc_synth_df(int ncols, const char** names, const int* kinds, const uint32_t* col_ids, const uint32_t* null_pct,
                 uint64_t row0, int64_t n, void** out) {
    return guard([&] {
        std::vector<Series> cols;
        for (int c = 0; c < ncols; ++c) cols.push_back(synth_series(names[c], kinds[c], col_ids[c], row0, n, null_pct[c]));
        *out = new DataFrame(DataFrame::make(std::move(cols)));
    });
}

// Timed eager collect():  from_dataframe(df).filter(col pred_col <op> lit).select(proj...).collect()
// Returns seconds of the collect() call alone (the clone in from_dataframe is part of the reference's
// API cost and is inside the timed region, as in builder.rs:27-39,96-104); rows_out receives the result height.
double orc_time_eager_filter_select(void* df, const char* pred_col, int op, int lit_tag, int64_t li, double lf,
                                    const char** proj, int nproj, int64_t* rows_out) {
    double secs = -1.0;
    int rc = guard([&] {
        AnyValue v = lit_tag == 1 ? AnyValue::Int64(li) : AnyValue::Float64(lf);
        std::vector<Expr> sel; for (int i = 0; i < nproj; ++i) sel.push_back(Expr::col(proj[i]));
        auto t0 = std::chrono::steady_clock::now();
        LazyFrame lf_ = LazyFrame::from_dataframe(*(DataFrame*)df).filter(Expr::col(pred_col).binary((BinaryOperator)op, Expr::lit(v)));
        if (nproj > 0) lf_ = lf_.select(sel);
        DataFrame out = lf_.collect();
        auto t1 = std::chrono::steady_clock::now();
        *rows_out = (int64_t)out.height();
        secs = std::chrono::duration<double>(t1 - t0).count();
    });
    return rc == 0 ? secs : -1.0;
}

// Same query over `threads` independent row-range shards at once (one reference instance per host
// core — the reference itself is single-threaded).  dfs[t] are pre-built shards.
double orc_time_eager_filter_select_mt(void** dfs, int threads, const char* pred_col, int op, int lit_tag, int64_t li, double lf,
                                       const char** proj, int nproj, int64_t* rows_out) {
    std::vector<std::thread> th; std::vector<int64_t> rows((size_t)threads, 0); std::vector<double> secs((size_t)threads, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t] { secs[(size_t)t] = orc_time_eager_filter_select(dfs[t], pred_col, op, lit_tag, li, lf, proj, nproj, &rows[(size_t)t]); });
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    int64_t total = 0; for (int t = 0; t < threads; ++t) { if (secs[(size_t)t] < 0) return -1.0; total += rows[(size_t)t]; }
    *rows_out = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

// ----------------------------------------------------------------------------------- full-size property oracle
// Survivor count and order-sensitive per-column checksums (include/rivulus_synth.h: rvl_checksum_term over the
// survivor's value bits at its output rank) of  filter(pred_col <op> literal).select(proj...)  over rows
// [row0, row0+n) of a synthetic no-null table, computed straight from the generator without materialising it.
// The predicate goes through eval_cmp (plan.rs:114-120) on AnyValues, like the eager engine.
extern "C" int orc_synth_filter_checksums(int64_t n, uint64_t row0, int pred_kind, uint32_t pred_col_id, int op, int lit_tag,
                                          int64_t lit_i, double lit_f, int nproj, const int* kinds, const uint32_t* col_ids,
                                          int threads, int64_t limit, int64_t* count_out, uint64_t* checksums_out) {
    return guard([&] {
        if (threads < 1) threads = 1;
        const AnyValue lit = lit_tag == 1 ? AnyValue::Int64(lit_i) : AnyValue::Float64(lit_f);
        auto keep = [&](uint64_t row) {
            const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, pred_col_id, row);
            if (pred_kind == RVL_SYNTH_F64) return eval_cmp(AnyValue::Float64(rvl_synth_f64(u)), (BinaryOperator)op, lit);
            if (pred_kind == RVL_SYNTH_BOOL) return eval_cmp(AnyValue::Boolean((u & 1) != 0), (BinaryOperator)op, lit);
            return eval_cmp(AnyValue::Int64(rvl_synth_i64(u, pred_kind)), (BinaryOperator)op, lit);
        };
        std::vector<int64_t> counts((size_t)threads, 0);
        const int64_t per = (n + threads - 1) / threads;
        {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t)
                th.emplace_back([&, t] {
                    const int64_t b = std::min<int64_t>(n, per * t), e = std::min<int64_t>(n, per * (t + 1));
                    int64_t c = 0;
                    for (int64_t r = b; r < e; ++r) c += keep(row0 + (uint64_t)r) ? 1 : 0;
                    counts[(size_t)t] = c;
                });
            for (auto& x : th) x.join();
        }
        std::vector<int64_t> prefix((size_t)threads + 1, 0);
        for (int t = 0; t < threads; ++t) prefix[(size_t)t + 1] = prefix[(size_t)t] + counts[(size_t)t];
        const int64_t total = prefix[(size_t)threads];
        const int64_t cap = limit >= 0 ? std::min(limit, total) : total;
        std::vector<std::vector<uint64_t>> sums((size_t)threads, std::vector<uint64_t>((size_t)nproj, 0));
        {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t)
                th.emplace_back([&, t] {
                    const int64_t b = std::min<int64_t>(n, per * t), e = std::min<int64_t>(n, per * (t + 1));
                    int64_t rank = prefix[(size_t)t];
                    for (int64_t r = b; r < e && rank < cap; ++r) {
                        const uint64_t row = row0 + (uint64_t)r;
                        if (!keep(row)) continue;
                        for (int c = 0; c < nproj; ++c) {
                            const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row);
                            uint64_t bits;
                            if (kinds[c] == RVL_SYNTH_F64) { const double d = rvl_synth_f64(u); std::memcpy(&bits, &d, 8); }
                            else if (kinds[c] == RVL_SYNTH_BOOL) bits = u & 1ull;
                            else bits = (uint64_t)rvl_synth_i64(u, kinds[c]);
                            sums[(size_t)t][(size_t)c] += rvl_checksum_term(bits, (uint64_t)rank);
                        }
                        ++rank;
                    }
                });
            for (auto& x : th) x.join();
        }
        *count_out = cap;
        for (int c = 0; c < nproj; ++c) {
            uint64_t s = 0;
            for (int t = 0; t < threads; ++t) s += sums[(size_t)t][(size_t)c];
            checksums_out[c] = s;
        }
    });
}

// Same, for tables WITH nulls and string columns (BASELINE configs[2] / configs[4] shapes): every column has its own null
// percentage (rvl_synth_valid); a null predicate row goes through eval_cmp as AnyValue::Null (series.rs:105-107); null
// survivors contribute the fixed tag kNullTag; strings contribute fold(splitmix64(h ^ byte)) seeded with their length —
// the definitions rvl_batch_checksum uses on the device.  null_counts_out[c] = null survivors of projected column c.
extern "C" int orc_synth_filter_checksums_nulls(int64_t n, uint64_t row0, int pred_kind, uint32_t pred_col_id, uint32_t pred_null_pct, int op,
                                                int lit_tag, int64_t lit_i, double lit_f, int nproj, const int* kinds, const uint32_t* col_ids,
                                                const uint32_t* null_pcts, int threads, int64_t limit, int64_t* count_out,
                                                uint64_t* checksums_out, int64_t* null_counts_out, int64_t* str_bytes_out) {
    return guard([&] {
        constexpr uint64_t kNullTag = 0x6E756C6C6E756C6Cull;
        if (threads < 1) threads = 1;
        const AnyValue lit = lit_tag == 1 ? AnyValue::Int64(lit_i) : (lit_tag == 2 ? AnyValue::Float64(lit_f) : AnyValue::Null());
        auto keep = [&](uint64_t row) {
            if (!rvl_synth_valid(RVL_SYNTH_SEED, pred_col_id, row, pred_null_pct)) return eval_cmp(AnyValue::Null(), (BinaryOperator)op, lit);
            const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, pred_col_id, row);
            if (pred_kind == RVL_SYNTH_F64) return eval_cmp(AnyValue::Float64(rvl_synth_f64(u)), (BinaryOperator)op, lit);
            if (pred_kind == RVL_SYNTH_BOOL) return eval_cmp(AnyValue::Boolean((u & 1) != 0), (BinaryOperator)op, lit);
            return eval_cmp(AnyValue::Int64(rvl_synth_i64(u, pred_kind)), (BinaryOperator)op, lit);
        };
        std::vector<int64_t> counts((size_t)threads, 0);
        const int64_t per = (n + threads - 1) / threads;
        {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t)
                th.emplace_back([&, t] {
                    const int64_t b = std::min<int64_t>(n, per * t), e = std::min<int64_t>(n, per * (t + 1));
                    int64_t c = 0;
                    for (int64_t r = b; r < e; ++r) c += keep(row0 + (uint64_t)r) ? 1 : 0;
                    counts[(size_t)t] = c;
                });
            for (auto& x : th) x.join();
        }
        std::vector<int64_t> prefix((size_t)threads + 1, 0);
        for (int t = 0; t < threads; ++t) prefix[(size_t)t + 1] = prefix[(size_t)t] + counts[(size_t)t];
        const int64_t total = prefix[(size_t)threads];
        const int64_t cap = limit >= 0 ? std::min(limit, total) : total;
        std::vector<std::vector<uint64_t>> sums((size_t)threads, std::vector<uint64_t>((size_t)nproj, 0));
        std::vector<std::vector<int64_t>> nulls((size_t)threads, std::vector<int64_t>((size_t)nproj, 0)), bytes = nulls;
        {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; ++t)
                th.emplace_back([&, t] {
                    const int64_t b = std::min<int64_t>(n, per * t), e = std::min<int64_t>(n, per * (t + 1));
                    int64_t rank = prefix[(size_t)t];
                    for (int64_t r = b; r < e && rank < cap; ++r) {
                        const uint64_t row = row0 + (uint64_t)r;
                        if (!keep(row)) continue;
                        for (int c = 0; c < nproj; ++c) {
                            uint64_t bits;
                            if (!rvl_synth_valid(RVL_SYNTH_SEED, col_ids[c], row, null_pcts[c])) { bits = kNullTag; nulls[(size_t)t][(size_t)c]++; }
                            else {
                                const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row);
                                if (kinds[c] == RVL_SYNTH_F64) { const double d = rvl_synth_f64(u); std::memcpy(&bits, &d, 8); }
                                else if (kinds[c] == RVL_SYNTH_BOOL) bits = u & 1ull;
                                else if (kinds[c] == RVL_SYNTH_STR) {
                                    const uint32_t len = rvl_synth_strlen(u);
                                    uint64_t h = len;
                                    for (uint32_t j = 0; j < len; ++j) h = rvl_splitmix64(h ^ (uint64_t)rvl_synth_strbyte(u, j));
                                    bits = h; bytes[(size_t)t][(size_t)c] += len;
                                } else bits = (uint64_t)rvl_synth_i64(u, kinds[c]);
                            }
                            sums[(size_t)t][(size_t)c] += rvl_checksum_term(bits, (uint64_t)rank);
                        }
                        ++rank;
                    }
                });
            for (auto& x : th) x.join();
        }
        *count_out = cap;
        for (int c = 0; c < nproj; ++c) {
            uint64_t s = 0; int64_t nc = 0, nb = 0;
            for (int t = 0; t < threads; ++t) { s += sums[(size_t)t][(size_t)c]; nc += nulls[(size_t)t][(size_t)c]; nb += bytes[(size_t)t][(size_t)c]; }
            checksums_out[c] = s; null_counts_out[c] = nc; str_bytes_out[c] = nb;
        }
    });
}
