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
