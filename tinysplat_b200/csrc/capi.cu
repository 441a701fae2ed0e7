// Library-level C ABI: version, last-error text, launch counter.
#include <atomic>
#include <cstdio>
#include "ts_common.cuh"

namespace ts {
static thread_local char g_err[256] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* where, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ts

extern "C" {
int ts_version(void) { return 100; /* 0.1.0 */ }
const char* ts_last_error(void) { return ts::g_err; }
int ts_rec_floats(void) { return ts::kRecFloats; }
int ts_grad_floats(void) { return ts::kGradFloats; }
int64_t ts_launch_count(void) { return (int64_t)ts::g_launches.load(std::memory_order_relaxed); }
}
