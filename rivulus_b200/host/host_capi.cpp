// host_capi.cpp — extern "C" surface over the C++ host layer (rivulus.hpp) so that Python (rivulus_b200/frame.py) and the
// pytest parity tests can drive LazyFrame / DataFrame / RecordBatch / StreamingPhysicalPlan.  Every query entry point
// ends in librivulus_gpu.so kernels; this file only marshals arguments.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/rivulus_synth.h"
#include "rivulus.hpp"

using namespace rivulus;

namespace {
thread_local std::string g_err;
struct DfBuilder { std::vector<Series> cols; };
struct RbHandle { RecordBatch rb; std::vector<std::unique_ptr<ArrayData>> cache; };
template <typename F> int guard(F&& f) {
    try { f(); g_err.clear(); return 0; }
    catch (const Error& e) { g_err = e.panic ? std::string("panic: ") + e.what() : std::string(e.what()); return e.panic ? 2 : 1; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
struct ColExport {  // same shape as the oracle's export struct, so one numpy decoder serves both in the tests
    int32_t dtype; int64_t length; int64_t offset; int64_t null_count;
    const void* values; int64_t values_len;
    const uint8_t* validity; int64_t validity_len;
    const int32_t* offsets; int64_t offsets_len;
    const uint8_t* data; int64_t data_len;
};
}  // namespace

extern "C" {

const char* rvh_last_error() { return g_err.c_str(); }
void rvh_set_stream_fusion(int on) { set_stream_fusion(on != 0); }
void rvh_set_extensions(int on) { set_extensions(on != 0); }
int rvh_dtype_is_numeric(int d) { return dtype_is_numeric((DataType)d) ? 1 : 0; }
int rvh_dtype_is_comparable_with(int a, int b) { return dtype_is_comparable_with((DataType)a, (DataType)b) ? 1 : 0; }
int64_t rvh_launch_count(int device) { int64_t n = -1; guard([&] { n = launch_count(device); }); return n; }

// ----------------------------------------------------------------------------------- DataFrame
void* rvh_dfb_new() { return new DfBuilder(); }
// Series::new over AnyValues: tags[i] = 0 Null, 1 Int64, 2 Float64, 3 String, 4 Boolean
int rvh_dfb_add_series(void* h, const char* name, int64_t n, const uint8_t* tags, const int64_t* i64, const double* f64,
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
                default: throw Error("bad tag");
            }
        }
        ((DfBuilder*)h)->cols.push_back(Series::make(name, d));
    });
}
int rvh_dfb_add_empty_series(void* h, const char* name, int dtype) {
    return guard([&] { ((DfBuilder*)h)->cols.push_back(Series::empty(name, (DataType)dtype)); });
}
// columnar ingestion (Arrow buffers; validity_bits LSB-first or NULL)
int rvh_dfb_add_i64(void* h, const char* name, int64_t n, const int64_t* v, const uint8_t* validity_bits) {
    return guard([&] { ((DfBuilder*)h)->cols.push_back(Series::from_i64(name, v, (size_t)n, validity_bits)); });
}
int rvh_dfb_add_f64(void* h, const char* name, int64_t n, const double* v, const uint8_t* validity_bits) {
    return guard([&] { ((DfBuilder*)h)->cols.push_back(Series::from_f64(name, v, (size_t)n, validity_bits)); });
}
int rvh_dfb_add_bool_bits(void* h, const char* name, int64_t n, const uint8_t* value_bits, const uint8_t* validity_bits) {
    return guard([&] { ((DfBuilder*)h)->cols.push_back(Series::from_bool_bits(name, value_bits, (size_t)n, validity_bits)); });
}
int rvh_dfb_add_strings(void* h, const char* name, int64_t n, const int32_t* offsets, const uint8_t* data, const uint8_t* validity_bits) {
    return guard([&] {
        ((DfBuilder*)h)->cols.push_back(Series::from_strings(name, offsets, (size_t)n, data, validity_bits));
    });
}
int rvh_dfb_finish(void* h, void** out) {
    auto* b = (DfBuilder*)h;
    int rc = guard([&] { *out = new DataFrame(DataFrame::make(std::move(b->cols))); });
    delete b;
    return rc;
}
// Synthetic DataFrame from the counter-based generator of include/rivulus_synth.h (BASELINE configs): columns are
// built directly in Arrow layout.  kinds[] = rvl_synth_kind, null_pct per column.
int rvh_synth_df(int ncols, const char** names, const int* kinds, const uint32_t* col_ids, const uint32_t* null_pct, uint64_t row0, int64_t n,
                 void** out) {
    return guard([&] {
        std::vector<Series> cols;
        for (int c = 0; c < ncols; ++c) {
            std::vector<uint8_t> validity;
            if (null_pct[c] > 0) {
                validity.assign((size_t)(n + 7) / 8, 0);
                for (int64_t r = 0; r < n; ++r)
                    if (rvl_synth_valid(RVL_SYNTH_SEED, col_ids[c], row0 + (uint64_t)r, null_pct[c])) validity[(size_t)r >> 3] |= (uint8_t)(1u << (r & 7));
            }
            auto valid = [&](int64_t r) { return validity.empty() || ((validity[(size_t)r >> 3] >> (r & 7)) & 1); };
            switch (kinds[c]) {
                case RVL_SYNTH_KEY1000: case RVL_SYNTH_I64: case RVL_SYNTH_AGE100: {
                    std::vector<int64_t> v((size_t)n);
                    for (int64_t r = 0; r < n; ++r) v[(size_t)r] = rvl_synth_i64(rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row0 + (uint64_t)r), kinds[c]);
                    cols.push_back(Series::from_i64(names[c], std::move(v), std::move(validity)));
                    break;
                }
                case RVL_SYNTH_F64: {
                    std::vector<double> v((size_t)n);
                    for (int64_t r = 0; r < n; ++r) v[(size_t)r] = rvl_synth_f64(rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row0 + (uint64_t)r));
                    cols.push_back(Series::from_f64(names[c], std::move(v), std::move(validity)));
                    break;
                }
                case RVL_SYNTH_BOOL: {
                    std::vector<uint8_t> bits((size_t)(n + 7) / 8, 0);
                    for (int64_t r = 0; r < n; ++r)
                        if (rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row0 + (uint64_t)r) & 1ull) bits[(size_t)r >> 3] |= (uint8_t)(1u << (r & 7));
                    cols.push_back(Series::from_bool_bits(names[c], std::move(bits), (size_t)n, std::move(validity)));
                    break;
                }
                case RVL_SYNTH_STR: {
                    std::vector<int32_t> off((size_t)n + 1, 0);
                    std::vector<uint8_t> data;
                    data.reserve((size_t)n * 24);
                    for (int64_t r = 0; r < n; ++r) {
                        if (valid(r)) {
                            const uint64_t u = rvl_synth_u(RVL_SYNTH_SEED, col_ids[c], row0 + (uint64_t)r);
                            const uint32_t len = rvl_synth_strlen(u);
                            for (uint32_t j = 0; j < len; ++j) data.push_back(rvl_synth_strbyte(u, j));
                        }
                        off[(size_t)r + 1] = (int32_t)data.size();
                    }
                    cols.push_back(Series::from_strings(names[c], std::move(off), std::move(data), std::move(validity)));
                    break;
                }
                default: throw Error("unknown synthetic column kind");
            }
        }
        *out = new DataFrame(DataFrame::make(std::move(cols)));
    });
}
void rvh_df_free(void* df) { delete (DataFrame*)df; }
int rvh_df_width(void* df) { return (int)((DataFrame*)df)->width(); }
int64_t rvh_df_height(void* df) { return (int64_t)((DataFrame*)df)->height(); }
const char* rvh_df_col_name(void* df, int i) { return ((DataFrame*)df)->columns()[i].name().c_str(); }
int rvh_df_col_dtype(void* df, int i) { return (int)((DataFrame*)df)->columns()[i].dtype(); }
int64_t rvh_df_col_len(void* df, int i) { return (int64_t)((DataFrame*)df)->columns()[i].len(); }
int64_t rvh_df_col_str_bytes(void* df, int i) { return (int64_t)((DataFrame*)df)->columns()[i].str_data().size(); }
// export one Series as AnyValue arrays (each sized len; str_off len+1; str_data rvh_df_col_str_bytes)
void rvh_df_col_export(void* df, int i, uint8_t* tags, int64_t* i64, double* f64, uint8_t* b8, int32_t* str_off, uint8_t* str_data) {
    const Series& s = ((DataFrame*)df)->columns()[i];
    int32_t pos = 0;
    for (size_t r = 0; r < s.len(); ++r) {
        const AnyValue v = s.at(r);
        tags[r] = (uint8_t)v.tag; i64[r] = v.i; f64[r] = v.f; b8[r] = v.b ? 1 : 0;
        str_off[r] = pos;
        if (v.tag == AnyValue::kString) { std::memcpy(str_data + pos, v.s.data(), v.s.size()); pos += (int32_t)v.s.size(); }
    }
    str_off[s.len()] = pos;
}
// raw Arrow buffers of a Series (pointers live as long as the DataFrame)
void rvh_df_col_buffers(void* df, int i, ColExport* e) {
    const Series& s = ((DataFrame*)df)->columns()[i];
    std::memset(e, 0, sizeof *e);
    e->length = (int64_t)s.len(); e->null_count = (int64_t)s.null_count();
    switch (s.dtype()) {
        case DataType::Int64: e->dtype = RVL_INT64; e->values = s.i64_values().data(); e->values_len = (int64_t)s.i64_values().size(); break;
        case DataType::Float64: e->dtype = RVL_FLOAT64; e->values = s.f64_values().data(); e->values_len = (int64_t)s.f64_values().size(); break;
        case DataType::Boolean: e->dtype = RVL_BOOLEAN; e->values = s.bool_bits().data(); e->values_len = (int64_t)s.bool_bits().size(); break;
        case DataType::String:
            e->dtype = RVL_STRING; e->offsets = s.str_offsets().data(); e->offsets_len = (int64_t)s.str_offsets().size();
            e->data = s.str_data().data(); e->data_len = (int64_t)s.str_data().size();
            break;
        case DataType::Null: e->dtype = RVL_NULL; break;
    }
    if (!s.validity_bits().empty()) { e->validity = s.validity_bits().data(); e->validity_len = (int64_t)s.validity_bits().size(); }
}

// ----------------------------------------------------------------------------------- Expr / LazyFrame
void* rvh_expr_col(const char* name) { return new Expr(Expr::col(name)); }
void* rvh_expr_lit(int tag, int64_t i, double f, const char* s, int64_t slen, int b) {
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
void* rvh_expr_binary(void* l, int op, void* r) { return new Expr(((Expr*)l)->binary((BinaryOperator)op, *(Expr*)r)); }
void* rvh_expr_alias(void* e, const char* name) { return new Expr(((Expr*)e)->alias(name)); }
void rvh_expr_free(void* e) { delete (Expr*)e; }

void* rvh_lf_from_df(void* df) { return new LazyFrame(LazyFrame::from_dataframe(*(DataFrame*)df)); }
// LazyFrame::from_csv (builder.rs:41-55).  batch_size < 0 = None; delimiter NULL = None (one UTF-8 encoded char otherwise)
void* rvh_lf_from_csv(const char* path, int nfields, const char** names, const int* dtypes, int64_t batch_size, const char* delimiter) {
    std::vector<std::pair<std::string, DataType>> schema;
    for (int i = 0; i < nfields; ++i) schema.emplace_back(names[i], (DataType)dtypes[i]);
    return new LazyFrame(LazyFrame::from_csv(path, std::move(schema), batch_size < 0 ? std::nullopt : std::optional<size_t>((size_t)batch_size),
                                             delimiter ? std::optional<std::string>(delimiter) : std::nullopt));
}
void rvh_set_csv_reference_validity(int on) { set_csv_reference_validity(on != 0); }
void rvh_set_csv_threads(int n) { set_csv_threads(n); }
int64_t rvh_csv_adaptive_batch_size(int nfields, const int* exec_dtypes) {
    Schema s;
    for (int i = 0; i < nfields; ++i) s.fields.push_back(Field{"c" + std::to_string(i), (ExecType)exec_dtypes[i], true});
    return (int64_t)calculate_adaptive_batch_size(s);
}
// CPU-only probe of the CSV parser (no device involved): parses the whole file batch by batch and renders every batch as
// "B <rows>\n" + one line per column of space-separated cells (i64 decimal, f64 as 16 hex digits of its bits, Boolean 0/1,
// String as hex bytes prefixed with 's', NULL as 'N', a column without a validity bitmap gets a leading '!').
// Returns 0 and the text in a malloc'ed buffer (*out, free with rvh_free), or 1 with rvh_last_error set.
int rvh_csv_parse_dump(const char* path, int nfields, const int* exec_dtypes, int64_t batch_size, const char* delimiter, int reference_validity,
                       char** out) {
    return guard([&] {
        auto schema = std::make_shared<Schema>();
        for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{"c" + std::to_string(i), (ExecType)exec_dtypes[i], true});
        CsvBatchReader rd(path, schema, batch_size < 0 ? std::nullopt : std::optional<size_t>((size_t)batch_size),
                          delimiter ? std::optional<std::string>(delimiter) : std::nullopt);
        std::string text;
        char tmp[40];
        while (size_t rows = rd.read_batch()) {
            if (reference_validity) rd.apply_reference_validity();
            text += "B " + std::to_string(rows) + "\n";
            for (const rvl_column& c : rd.columns()) {
                if (!c.validity) text += "! ";
                for (size_t r = 0; r < rows; ++r) {
                    const bool valid = c.dtype != RVL_NULL && (!c.validity || ((c.validity[r >> 3] >> (r & 7)) & 1));
                    if (!valid) { text += "N "; continue; }
                    switch (c.dtype) {
                        case RVL_INT64: text += std::to_string(((const int64_t*)c.values)[r]) + " "; break;
                        case RVL_FLOAT64: { uint64_t b; std::memcpy(&b, (const double*)c.values + r, 8); std::snprintf(tmp, sizeof tmp, "%016llx ", (unsigned long long)b); text += tmp; break; }
                        case RVL_BOOLEAN: text += ((((const uint8_t*)c.values)[r >> 3] >> (r & 7)) & 1) ? "1 " : "0 "; break;
                        case RVL_STRING: {
                            text += "s";
                            for (int32_t k = c.offsets[r]; k < c.offsets[r + 1]; ++k) { std::snprintf(tmp, sizeof tmp, "%02x", c.data[k]); text += tmp; }
                            text += " ";
                            break;
                        }
                        default: break;
                    }
                }
                text += "\n";
            }
        }
        *out = (char*)std::malloc(text.size() + 1);
        std::memcpy(*out, text.c_str(), text.size() + 1);
    });
}
void rvh_free(void* p) { std::free(p); }
void* rvh_lf_select(void* lf, int n, void** exprs) {
    std::vector<Expr> e; for (int i = 0; i < n; ++i) e.push_back(*(Expr*)exprs[i]);
    return new LazyFrame(((LazyFrame*)lf)->select(std::move(e)));
}
void* rvh_lf_filter(void* lf, void* pred) { return new LazyFrame(((LazyFrame*)lf)->filter(*(Expr*)pred)); }
void* rvh_lf_inner_join(void* lf, void* right, const char* left_key, const char* right_key) {
    return new LazyFrame(((LazyFrame*)lf)->inner_join(*(LazyFrame*)right, left_key, right_key));
}
void* rvh_lf_limit(void* lf, int64_t n) { return new LazyFrame(((LazyFrame*)lf)->limit((size_t)n)); }
void rvh_lf_free(void* lf) { delete (LazyFrame*)lf; }
int rvh_lf_collect(void* lf, void** df_out) {
    return guard([&] { *df_out = new DataFrame(((LazyFrame*)lf)->collect()); });
}
int rvh_lf_collect_streaming(void* lf, void** rb_out) {
    return guard([&] { *rb_out = new RbHandle{((LazyFrame*)lf)->collect_streaming(), {}}; });
}
// logical_plan/plan.rs probes on the plan as built (not optimized)
int rvh_lf_schema(void* lf, char* buf, int cap) {
    return guard([&] {
        std::string s;
        for (const auto& pr : ((LazyFrame*)lf)->logical_plan().schema()) s += (s.empty() ? "" : ",") + pr.first + ":" + dtype_name(pr.second);
        std::snprintf(buf, (size_t)cap, "%s", s.c_str());
    });
}
int rvh_lf_validate(void* lf) { return guard([&] { ((LazyFrame*)lf)->logical_plan().validate(); }); }
int rvh_lf_describe(void* lf, char* buf, int cap) {
    return guard([&] { std::snprintf(buf, (size_t)cap, "%s", ((LazyFrame*)lf)->logical_plan().describe().c_str()); });
}
int rvh_lf_plan_shape(void* lf, char* buf, int cap) {  // optimized plan, "Filter(Select(Source))"-style
    return guard([&] { std::snprintf(buf, (size_t)cap, "%s", optimize(((LazyFrame*)lf)->logical_plan()).shape().c_str()); });
}

// ----------------------------------------------------------------------------------- RecordBatch (device resident)
// try_new over host buffers: per column dtype (schema.rs order), buffers as in rvl_column
int rvh_rb_try_new(int nfields, const char** names, const int* schema_dtypes, int ncols, const rvl_column* cols, void** out) {
    return guard([&] {
        auto schema = std::make_shared<Schema>();
        for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{names[i], (ExecType)schema_dtypes[i], true});
        std::vector<rvl_column> c(cols, cols + ncols);
        *out = new RbHandle{RecordBatch::try_new(Context::shared(0), schema, c), {}};
    });
}
int rvh_rb_new_unchecked(int nfields, const char** names, const int* schema_dtypes, int ncols, const rvl_column* cols, int64_t num_rows, void** out) {
    return guard([&] {
        auto schema = std::make_shared<Schema>();
        for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{names[i], (ExecType)schema_dtypes[i], true});
        std::vector<rvl_column> c(cols, cols + ncols);
        *out = new RbHandle{RecordBatch::new_unchecked(Context::shared(0), schema, c, (size_t)num_rows), {}};
    });
}
int rvh_rb_validate(void* rb) { return guard([&] { ((RbHandle*)rb)->rb.validate(); }); }
int64_t rvh_rb_memory_size(void* rb) { int64_t n = -1; guard([&] { n = (int64_t)((RbHandle*)rb)->rb.memory_size(); }); return n; }
void* rvh_rbb_new(int nfields, const char** names, const int* schema_dtypes) {
    auto schema = std::make_shared<Schema>();
    for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{names[i], (ExecType)schema_dtypes[i], true});
    return new RecordBatchBuilder(schema);
}
int rvh_rbb_add_column(void* b, const rvl_column* col) { return guard([&] { ((RecordBatchBuilder*)b)->add_column(*col); }); }
int rvh_rbb_finish(void* b, void** out) { return guard([&] { *out = new RbHandle{((RecordBatchBuilder*)b)->finish(), {}}; }); }
int rvh_rbb_num_columns(void* b) { return (int)((RecordBatchBuilder*)b)->num_columns(); }
int rvh_rbb_is_complete(void* b) { return ((RecordBatchBuilder*)b)->is_complete() ? 1 : 0; }
void rvh_rbb_free(void* b) { delete (RecordBatchBuilder*)b; }
void rvh_rb_free(void* rb) { delete (RbHandle*)rb; }
int64_t rvh_rb_num_rows(void* rb) { int64_t n = -1; guard([&] { n = (int64_t)((RbHandle*)rb)->rb.num_rows(); }); return n; }
int rvh_rb_num_columns(void* rb) { return (int)((RbHandle*)rb)->rb.schema()->fields.size(); }
const char* rvh_rb_col_name(void* rb, int i) { return ((RbHandle*)rb)->rb.schema()->fields[(size_t)i].name.c_str(); }
int rvh_rb_col_dtype(void* rb, int i) { return (int)((RbHandle*)rb)->rb.schema()->fields[(size_t)i].data_type; }
int rvh_rb_col_export(void* rb, int i, ColExport* e) {
    return guard([&] {
        auto* h = (RbHandle*)rb;
        h->cache.push_back(std::make_unique<ArrayData>(h->rb.column_data((size_t)i)));
        const ArrayData& a = *h->cache.back();
        std::memset(e, 0, sizeof *e);
        e->dtype = (int32_t)a.dtype; e->length = a.length; e->offset = 0; e->null_count = a.null_count;
        switch (a.dtype) {
            case ExecType::Int64: e->values = a.i64.data(); e->values_len = (int64_t)a.i64.size(); break;
            case ExecType::Float64: e->values = a.f64.data(); e->values_len = (int64_t)a.f64.size(); break;
            case ExecType::Boolean: e->values = a.bits.data(); e->values_len = (int64_t)a.bits.size(); break;
            case ExecType::String:
                e->offsets = a.offsets.data(); e->offsets_len = (int64_t)a.offsets.size();
                e->data = a.data.data(); e->data_len = (int64_t)a.data.size();
                break;
            default: break;
        }
        if (a.has_validity) { e->validity = a.validity.data(); e->validity_len = (int64_t)a.validity.size(); }
    });
}
int rvh_rb_slice(void* rb, int64_t off, int64_t len, void** out) {
    return guard([&] { *out = new RbHandle{((RbHandle*)rb)->rb.slice((size_t)off, (size_t)len), {}}; });
}
int rvh_rb_take(void* rb, const int64_t* idx, int64_t n, void** out) {
    return guard([&] {
        std::vector<size_t> v; for (int64_t i = 0; i < n; ++i) v.push_back((size_t)idx[i]);
        *out = new RbHandle{((RbHandle*)rb)->rb.take(v), {}};
    });
}
int rvh_rb_select(void* rb, const int32_t* idx, int n, void** out) {
    return guard([&] {
        std::vector<size_t> v; for (int i = 0; i < n; ++i) v.push_back((size_t)idx[i]);
        *out = new RbHandle{((RbHandle*)rb)->rb.select_columns(v), {}};
    });
}
int rvh_rb_select_by_name(void* rb, const char** names, int n, void** out) {
    return guard([&] {
        std::vector<std::string> v; for (int i = 0; i < n; ++i) v.push_back(names[i]);
        *out = new RbHandle{((RbHandle*)rb)->rb.select_columns_by_name(v), {}};
    });
}
// filter(&self, predicate) where the predicate array is column `pcol` of batch `prb`
int rvh_rb_filter(void* rb, void* prb, int pcol, void** out) {
    return guard([&] { *out = new RbHandle{((RbHandle*)rb)->rb.filter(((RbHandle*)prb)->rb, (size_t)pcol), {}}; });
}
int rvh_rb_concat(void** rbs, int n, void** out) {
    return guard([&] {
        std::vector<RecordBatch> v; for (int i = 0; i < n; ++i) v.push_back(((RbHandle*)rbs[i])->rb);
        *out = new RbHandle{RecordBatch::concat(v), {}};
    });
}
int rvh_rb_empty_like(void* rb, void** out) {
    return guard([&] { auto& r = ((RbHandle*)rb)->rb; *out = new RbHandle{RecordBatch::empty(r.context(), r.schema()), {}}; });
}

// ----------------------------------------------------------------------------------- StreamingPhysicalPlan
void* rvh_sp_memory_source(void** rbs, int n) {
    std::vector<RecordBatch> v; for (int i = 0; i < n; ++i) v.push_back(((RbHandle*)rbs[i])->rb);
    return new StreamingPhysicalPlan(StreamingPhysicalPlan::memory_source(std::move(v)));
}
void* rvh_sp_dataframe_source(void* df, int64_t batch_size) {
    void* out = nullptr;
    guard([&] { out = new StreamingPhysicalPlan(StreamingPhysicalPlan::dataframe_source(*(DataFrame*)df, (size_t)batch_size)); });
    return out;
}
// StreamingPhysicalPlan::csv_file_source (streaming.rs:299-311) with the execution schema given field by field
void* rvh_sp_csv_source(const char* path, int nfields, const char** names, const int* exec_dtypes, const int* nullable, int64_t batch_size,
                        const char* delimiter) {
    auto schema = std::make_shared<Schema>();
    for (int i = 0; i < nfields; ++i) schema->fields.push_back(Field{names[i], (ExecType)exec_dtypes[i], nullable[i] != 0});
    return new StreamingPhysicalPlan(StreamingPhysicalPlan::csv_file_source(path, schema, batch_size < 0 ? std::nullopt : std::optional<size_t>((size_t)batch_size),
                                                                            delimiter ? std::optional<std::string>(delimiter) : std::nullopt));
}
void* rvh_sp_filter(void* sp, const char* col) { return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->filter(col)); }
void* rvh_sp_select(void* sp, const char** names, int n) {
    std::vector<std::string> v; for (int i = 0; i < n; ++i) v.push_back(names[i]);
    return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->select(std::move(v)));
}
void* rvh_sp_limit(void* sp, int64_t n) { return new StreamingPhysicalPlan(((StreamingPhysicalPlan*)sp)->limit((size_t)n)); }
void rvh_sp_free(void* sp) { delete (StreamingPhysicalPlan*)sp; }
int rvh_sp_collect(void* sp, void** rb_out) {
    return guard([&] { *rb_out = new RbHandle{((StreamingPhysicalPlan*)sp)->collect(), {}}; });
}
int rvh_sp_collect_batches(void* sp, void** vec_out) {
    return guard([&] { *vec_out = new std::vector<RecordBatch>(((StreamingPhysicalPlan*)sp)->collect_batches()); });
}
int rvh_rbv_len(void* v) { return (int)((std::vector<RecordBatch>*)v)->size(); }
void* rvh_rbv_get(void* v, int i) { return new RbHandle{(*(std::vector<RecordBatch>*)v)[(size_t)i], {}}; }
void rvh_rbv_free(void* v) { delete (std::vector<RecordBatch>*)v; }

}  // extern "C"
