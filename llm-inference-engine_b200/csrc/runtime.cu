// runtime.cu -- error plumbing, device queries and the library workspace of libb200llm.so.
#include "common.cuh"

#include <mutex>

#ifndef B200_NO_NVTX
#include <nvtx3/nvToolsExt.h>
#endif

namespace b200 {

#ifndef B200_NO_NVTX
NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
#endif


static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return B200_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B200_ERR_CUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// One workspace per device: [tickets (kTickets x u32, zeroed once, self-resetting) | scratch].
static constexpr size_t kTickets = 16384;
struct WsSlot {
    void *base = nullptr;
    size_t bytes = 0;
    bool owned = false;
};
static WsSlot g_ws[64];
static std::mutex g_ws_mu;

bool get_workspace(Workspace *ws) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsSlot &s = g_ws[dev & 63];
    if (!s.base) {
        set_error("library workspace not set: call b200_workspace_ensure() or b200_workspace_set() first");
        return false;
    }
    ws->tickets = reinterpret_cast<unsigned int *>(s.base);
    ws->n_tickets = kTickets;
    ws->scratch = reinterpret_cast<char *>(s.base) + kTickets * sizeof(unsigned int);
    ws->scratch_bytes = s.bytes - kTickets * sizeof(unsigned int);
    return true;
}

}  // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error_string(void) { return g_err; }
int b200_abi_version(void) { return B200LLM_ABI_VERSION; }
int b200_sm_count(void) { return sm_count(); }

size_t b200_workspace_default_bytes(void) { return (size_t)64 << 20; }

int b200_workspace_set(void *ptr, size_t bytes) {
    B200_REQUIRE(ptr != nullptr && aligned16(ptr), "workspace pointer must be non-null and 16-byte aligned");
    B200_REQUIRE(bytes >= ((size_t)1 << 20), "workspace must be at least 1 MiB");
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaMemset(ptr, 0, kTickets * sizeof(unsigned int)) != cudaSuccess) return cuda_status("workspace memset");
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsSlot &s = g_ws[dev & 63];
    if (s.owned && s.base) cudaFree(s.base);
    s.base = ptr;
    s.bytes = bytes;
    s.owned = false;
    return B200_OK;
}

int b200_workspace_ensure(size_t bytes) {
    if (bytes == 0) bytes = b200_workspace_default_bytes();
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        WsSlot &s = g_ws[dev & 63];
        if (s.base && s.bytes >= bytes) return B200_OK;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace cudaMalloc(%zu) failed", bytes);
        return B200_ERR_WORKSPACE;
    }
    if (cudaMemset(p, 0, kTickets * sizeof(unsigned int)) != cudaSuccess) return cuda_status("workspace memset");
    std::lock_guard<std::mutex> lk(g_ws_mu);
    WsSlot &s = g_ws[dev & 63];
    if (s.owned && s.base) cudaFree(s.base);
    s.base = p;
    s.bytes = bytes;
    s.owned = true;
    return B200_OK;
}

}  // extern "C"
