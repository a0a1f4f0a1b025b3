// demo_main.cpp — the reference's demo queries (/root/reference/src/main.rs:47-94 filter / select / limit, :170-196 inner join,
// :233-256 CSV streaming), written against the C++ host layer exactly as
// a Rust user writes them against the reference: same type and method names, same result semantics, executed on the GPU through
// the C ABI.  Prints one line per query: column names, dtypes and values.  tests/test_host_golden.py runs it and checks the
// lines against the expected outputs derived in SURVEY.md Appendix B.
#include <cstdio>
#include <string>
#include <vector>

#include "rivulus.hpp"

using namespace rivulus;

static void show(const char* title, const DataFrame& df) {
    std::string line = std::string(title) + " rows=" + std::to_string(df.height()) + " |";
    for (const Series& s : df.columns()) {
        line += " " + s.name() + ":[";
        const std::vector<AnyValue> v = s.to_values();
        for (size_t i = 0; i < v.size(); ++i) line += (i ? "," : "") + v[i].display();
        line += "]";
    }
    std::puts(line.c_str());
}

int main(int argc, char** argv) {
    try {
        auto S = [](std::initializer_list<const char*> v) { std::vector<AnyValue> o; for (auto s : v) o.push_back(AnyValue::String(s)); return o; };
        auto I = [](std::initializer_list<int64_t> v) { std::vector<AnyValue> o; for (auto x : v) o.push_back(AnyValue::Int64(x)); return o; };
        auto F = [](std::initializer_list<double> v) { std::vector<AnyValue> o; for (auto x : v) o.push_back(AnyValue::Float64(x)); return o; };
        const DataFrame df = DataFrame::make({Series::make("name", S({"Alice", "Bob", "Charlie", "Diana", "Eve"})),
                                              Series::make("age", I({25, 30, 35, 28, 42})),
                                              Series::make("score", F({85.5, 92.0, 78.5, 94.5, 88.0}))});
        // main.rs:47-57  SELECT name, age WHERE age > 30
        show("q1", LazyFrame::from_dataframe(df).select({Expr::col("name"), Expr::col("age")}).filter(Expr::col("age").gt(Expr::lit(30))).collect());
        // main.rs:59-70  WHERE score >= 90.0 SELECT name, age AS user_age
        show("q2", LazyFrame::from_dataframe(df).filter(Expr::col("score").gte(Expr::lit(90.0)))
                       .select({Expr::col("name"), Expr::col("age").alias("user_age")}).collect());
        // main.rs:72-81  WHERE age < 40 LIMIT 2
        show("q3", LazyFrame::from_dataframe(df).filter(Expr::col("age").lt(Expr::lit(40))).limit(2).collect());
        // main.rs:83-94  no match: zero rows, names and dtypes kept
        show("q4", LazyFrame::from_dataframe(df).filter(Expr::col("age").gt(Expr::lit(100))).collect());
        // streaming engine (builder.rs:106-113): boolean-column predicate, LIMIT stops the stream
        const DataFrame dfa = DataFrame::make({Series::make("name", S({"Alice", "Bob", "Charlie"})), Series::make("age", I({25, 30, 35})),
                                               Series::make("active", {AnyValue::Boolean(true), AnyValue::Boolean(false), AnyValue::Boolean(true)})});
        const RecordBatch rb = LazyFrame::from_dataframe(dfa).filter(Expr::col("active")).select({Expr::col("name"), Expr::col("age")}).limit(1).collect_streaming();
        std::printf("q5 rows=%zu cols=%zu launches=%lld\n", rb.num_rows(), rb.schema()->fields.size(), (long long)launch_count(0));
        // main.rs:120-183  SELECT * FROM users u INNER JOIN orders o ON u.user_id = o.user_id
        const DataFrame users = DataFrame::make({Series::make("user_id", I({1, 2, 3, 4})), Series::make("name", S({"Alice", "Bob", "Charlie", "Diana"})),
                                                 Series::make("city", S({"Rome", "Milan", "Naples", "Turin"}))});
        const DataFrame orders = DataFrame::make({Series::make("order_id", I({101, 102, 103, 104, 105})), Series::make("user_id", I({1, 2, 1, 3, 2})),
                                                  Series::make("amount", F({29.99, 15.5, 45.0, 8.75, 12.99}))});
        show("q6", LazyFrame::from_dataframe(users).inner_join(LazyFrame::from_dataframe(orders), "user_id", "user_id").collect());
        // main.rs:186-196  the same join, then SELECT name, amount, city
        show("q7", LazyFrame::from_dataframe(users).inner_join(LazyFrame::from_dataframe(orders), "user_id", "user_id")
                       .select({Expr::col("name"), Expr::col("amount"), Expr::col("city")}).collect());
        // main.rs:233-256  CSV file, ';' delimiter, three of four columns, LIMIT 3, streaming engine
        if (argc > 1) {
            const RecordBatch cb = LazyFrame::from_csv(argv[1], {{"Username", DataType::String}, {"Identifier", DataType::Int64}, {"First_name", DataType::String},
                                                                  {"Last_name", DataType::String}}, 1000, std::string(";"))
                                       .select({Expr::col("Username"), Expr::col("First_name"), Expr::col("Last_name")}).limit(3).collect_streaming();
            const ArrayData first = cb.column_data(0);
            std::printf("q8 rows=%zu cols=%zu first=%s\n", cb.num_rows(), cb.schema()->fields.size(), first.value(0).display().c_str());
        }
        return 0;
    } catch (const Error& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
