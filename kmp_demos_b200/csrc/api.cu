// Library-level state of libkmpx: version, thread-local error string, launch counter.
#include "common.cuh"

namespace kmpx {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_host_dbg = 0;   // profiling experiments (kmpx_debug_set)
}  // namespace kmpx

extern "C" int kmpx_version(void) { return KMPX_VERSION; }
extern "C" const char* kmpx_last_error(void) { return kmpx::g_err; }
extern "C" long long kmpx_launch_count(void) { return kmpx::g_launches.load(); }
