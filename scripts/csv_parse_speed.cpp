// Parse rate of the host CSV reader alone (no device): g++ -O2 -std=c++17 -Irivulus_b200/host -Iinclude scripts/csv_parse_speed.cpp -Lrivulus_b200/lib -lrivulus_host -lrivulus_gpu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sys/stat.h>
#include "rivulus.hpp"
using namespace rivulus;
int main(int argc, char** argv) {
    auto schema = std::make_shared<Schema>();
    schema->fields = {Field{"id", ExecType::Int64, true}, Field{"x", ExecType::Float64, true}, Field{"s", ExecType::String, true}, Field{"flag", ExecType::Boolean, true}};
    struct stat st; stat(argv[1], &st);
    if (argc > 2) set_csv_threads(atoi(argv[2]));
    for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        CsvBatchReader rd(argv[1], schema, std::nullopt, std::nullopt);
        size_t rows = 0, batches = 0;
        while (size_t r = rd.read_batch()) { rows += r; ++batches; }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%zu rows, %zu batches, %.3f s, %.1f MB/s\n", rows, batches, s, st.st_size / 1e6 / s);
    }
}
