// Library-wide pieces of the C ABI: version, error strings, launch counter.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace b200gnn {

static thread_local char g_cuda_err[256] = "";
static std::atomic<int64_t> g_launches{0};

void set_cuda_error(cudaError_t e) {
  std::strncpy(g_cuda_err, cudaGetErrorString(e), sizeof(g_cuda_err) - 1);
  g_cuda_err[sizeof(g_cuda_err) - 1] = '\0';
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace b200gnn

extern "C" int b200gnn_abi_version(void) { return B200GNN_ABI_VERSION; }

extern "C" const char* b200gnn_error_string(int code) {
  switch (code) {
    case B200GNN_OK: return "ok";
    case B200GNN_ERR_BAD_ARG: return "bad argument";
    case B200GNN_ERR_UNSUPPORTED: return "unsupported shape or layout";
    case B200GNN_ERR_CUDA: return "CUDA error (see b200gnn_last_cuda_error)";
    default: return "unknown error";
  }
}

extern "C" const char* b200gnn_last_cuda_error(void) { return b200gnn::g_cuda_err; }
extern "C" int64_t b200gnn_launch_count(void) { return b200gnn::g_launches.load(); }
extern "C" void b200gnn_reset_launch_count(void) { b200gnn::g_launches.store(0); }
