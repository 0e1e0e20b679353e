// Process-wide pieces of libikflow_b200: error string, version, launch accounting.
#include "common.h"

namespace ikf {

std::string& last_error_ref() {
  thread_local std::string msg;
  return msg;
}

std::atomic<uint64_t> g_launch_count{0};

}  // namespace ikf

extern "C" {

const char* ikf_last_error(void) { return ikf::last_error_ref().c_str(); }

const char* ikf_version(void) { return "ikflow_b200 0.1.0 (sm_100a)"; }

uint64_t ikf_launch_count(void) { return ikf::g_launch_count.load(std::memory_order_relaxed); }

}  // extern "C"
