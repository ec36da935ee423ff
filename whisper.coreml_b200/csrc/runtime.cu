// Process-global runtime state of libwhisper_b200: error log, launch counter, device, stream.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace b200 {

long g_launch_count = 0;
static std::mutex g_err_mu;
static char g_err_msg[1024] = {0};
static int g_err_count = 0;

void record_error(const char* fmt, ...) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof(g_err_msg), fmt, ap);
    va_end(ap);
    ++g_err_count;
    fprintf(stderr, "[whisper_b200] error: %s\n", g_err_msg);   // the reference NSLogs and continues (coreml.mm:54-56)
}

int take_errors(char* buf, int buf_len) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    int n = g_err_count;
    if (buf && buf_len > 0) {
        strncpy(buf, g_err_msg, buf_len - 1);
        buf[buf_len - 1] = 0;
    }
    g_err_count = 0;
    return n;
}

}  // namespace b200

// ---- per-stage CUDA-event timers -----------------------------------------------------------------
#include <vector>
#include "state.cuh"
namespace b200 {
struct Pending { int stage; cudaEvent_t e0, e1; };
static std::vector<Pending> g_pending;
static float g_stage_ms[ST_COUNT] = {0};

StageTimer::StageTimer(int st) : stage(st) {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, S().stream);
}
StageTimer::~StageTimer() {
    cudaEventRecord(e1, S().stream);
    g_pending.push_back({stage, e0, e1});
    if (g_pending.size() > 4096) stage_times(nullptr, false);
}
void stage_times(float* out_ms, bool reset) {
    for (const Pending& p : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) g_stage_ms[p.stage] += ms;
        cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
    }
    g_pending.clear();
    if (out_ms) for (int i = 0; i < ST_COUNT; ++i) out_ms[i] = g_stage_ms[i];
    if (reset) for (int i = 0; i < ST_COUNT; ++i) g_stage_ms[i] = 0.f;
}
}  // namespace b200
