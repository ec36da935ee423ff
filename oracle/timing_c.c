/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of the word-timestamp kernels of the reference.
 *   oracle_dtw            <- whisper/timing.py:57-105  (dtw_cpu + backtrace, numba)
 *   oracle_median_filter  <- whisper/timing.py:19-54   (reflect pad + sliding sort, middle element)
 * Built by oracle/build.py into oracle/_build/liboracle_timing.so; never linked into the product. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x: (N, M) row-major doubles (the reference calls dtw_cpu(x.double())), cost kept in fp32
 * (timing.py:85).  Tie rule (timing.py:95-100): c0 if strictly smallest, else c1 if strictly
 * smaller than both others, else c2 (even when c2 is not the minimum).
 * out_i/out_j hold the path (length returned), ordered from (0,0) to (N-1,M-1). */
int oracle_dtw(const double* x, int N, int M, int32_t* out_i, int32_t* out_j) {
    const int W = M + 1;
    float* cost = (float*)malloc(sizeof(float) * (size_t)(N + 1) * W);
    int8_t* trace = (int8_t*)malloc((size_t)(N + 1) * W);
    for (size_t i = 0; i < (size_t)(N + 1) * W; ++i) { cost[i] = INFINITY; trace[i] = -1; }
    cost[0] = 0.f;
    for (int j = 1; j <= M; ++j)
        for (int i = 1; i <= N; ++i) {
            float c0 = cost[(i - 1) * W + j - 1], c1 = cost[(i - 1) * W + j], c2 = cost[i * W + j - 1];
            float c; int8_t t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            cost[i * W + j] = (float)(x[(size_t)(i - 1) * M + (j - 1)] + (double)c);
            trace[i * W + j] = t;
        }
    for (int j = 0; j <= M; ++j) trace[j] = 2;            /* timing.py:61-62 */
    for (int i = 0; i <= N; ++i) trace[i * W] = 1;
    int i = N, j = M, n = 0;
    while (i > 0 || j > 0) {
        out_i[n] = i - 1; out_j[n] = j - 1; ++n;
        int8_t t = trace[i * W + j];
        if (t == 0) { --i; --j; } else if (t == 1) { --i; } else { --j; }
    }
    for (int a = 0, b = n - 1; a < b; ++a, --b) {
        int32_t t = out_i[a]; out_i[a] = out_i[b]; out_i[b] = t;
        t = out_j[a]; out_j[a] = out_j[b]; out_j[b] = t;
    }
    free(cost); free(trace);
    return n;
}

static int cmp_f32(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* x, y: (rows, len) row-major fp32; width odd.  Rows with len <= width/2 are returned unchanged
 * (timing.py:22-24). */
void oracle_median_filter(const float* x, float* y, long rows, int len, int width) {
    int pad = width / 2;
    if (len <= pad) { memcpy(y, x, sizeof(float) * (size_t)rows * len); return; }
    float win[64];
    for (long r = 0; r < rows; ++r) {
        const float* xr = x + r * len; float* yr = y + r * len;
        for (int t = 0; t < len; ++t) {
            for (int k = 0; k < width; ++k) {
                int s = t + k - pad;
                if (s < 0) s = -s;                         /* reflect (no edge repeat) */
                if (s >= len) s = 2 * (len - 1) - s;
                win[k] = xr[s];
            }
            qsort(win, width, sizeof(float), cmp_f32);
            yr[t] = win[pad];
        }
    }
}
