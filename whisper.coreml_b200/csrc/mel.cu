// log-mel spectrogram (whisper/audio.py:110-157) in fp32 on the device.
//
//   pass 1 (mel_frames_kernel): 32 frames per CTA.  The windowed frame is folded around n = 200
//     (x[n] +- x[400-n]; hann(0) = 0 kills n = 0), so the 400-point real DFT becomes two 199-term
//     dot products per bin against cos/sin tables computed in fp64.  Power -> sparse triangular
//     mel filters -> log10(clamp 1e-10) -> out, plus a global max (audio.py:155 is per file).
//   pass 2 (mel_finish_kernel): max(x, gmax - 8), (x + 4) / 4.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "state.cuh"
#include "whisper_b200.h"

namespace b200 {

constexpr int N_FFT = 400, HOP = 160, N_BINS = 201, BIN_LD = 208, MEL_F = 32, MEL_THREADS = 256;
constexpr int SEG = (MEL_F - 1) * HOP + N_FFT;        // samples one CTA touches: 5360

struct MelTables {
    float* twiddle = nullptr;       // [199][2][BIN_LD]: cos(2 pi k n / 400), sin(...) for n = 1..199
    float* window = nullptr;        // [201] periodic hann, n = 0..200
    int n_mels = 0;
    int* f_start = nullptr;         // [n_mels] first bin of each filter
    int* f_len = nullptr;           // [n_mels]
    int* f_off = nullptr;           // [n_mels] offset into f_w
    float* f_w = nullptr;
    int* gmax = nullptr;            // ordered-int encoding of the running max
};
static MelTables g_mel[2];          // [0]: 80 mels, [1]: 128 mels

__device__ __forceinline__ int float_to_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void mel_tables_kernel(float* twiddle, float* window) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 199 * BIN_LD) {
        const int n = i / BIN_LD + 1, k = i % BIN_LD;
        double s = 0.0, c = 0.0;
        if (k < N_BINS) sincospi((double)((n * k) % N_FFT) / 200.0, &s, &c);
        twiddle[(size_t)(n - 1) * 2 * BIN_LD + k] = (float)c;
        twiddle[(size_t)(n - 1) * 2 * BIN_LD + BIN_LD + k] = (float)s;
    }
    if (i <= 200) window[i] = (float)(0.5 - 0.5 * cospi((double)i / 200.0));
}

__global__ void __launch_bounds__(MEL_THREADS) mel_frames_kernel(const float* __restrict__ audio, long n_samples, long n_total,
                                                                  long n_frames, MelTables t, float* __restrict__ out) {
    extern __shared__ __align__(16) float msm[];
    float* seg = msm;                                   // [SEG]
    float* E = seg + SEG;                               // [199][32]
    float* O = E + 199 * MEL_F;                         // [199][32]
    float* mid = O + 199 * MEL_F;                       // [32]  x[200] (window = 1)
    float* P = E;                                       // [32][BIN_LD + 1] power, aliases E/O after the DFT
    __shared__ float smax[MEL_THREADS / 32];
    const int tid = threadIdx.x;
    const long f0 = (long)blockIdx.x * MEL_F;
    // ---- stage the audio segment with torch.stft's center=True reflect padding (audio.py:148) ----
    const long p0 = f0 * HOP - N_FFT / 2;
    for (int i = tid; i < SEG; i += MEL_THREADS) {
        long j = p0 + i;
        if (j < 0) j = -j;
        if (j >= n_total) j = 2 * (n_total - 1) - j;
        seg[i] = (j >= 0 && j < n_samples) ? audio[j] : 0.f;      // [n_samples, n_total) is the zero padding
    }
    __syncthreads();
    for (int i = tid; i < 199 * MEL_F; i += MEL_THREADS) {
        const int n = i / MEL_F + 1, f = i % MEL_F;
        const float w = t.window[n];
        const float a = seg[f * HOP + n], b = seg[f * HOP + N_FFT - n];
        E[i] = w * a + w * b;                           // torch multiplies by the window first, then transforms
        O[i] = w * a - w * b;
    }
    if (tid < MEL_F) mid[tid] = seg[tid * HOP + 200];
    __syncthreads();
    // ---- DFT: thread = bin, 32 frames in registers ------------------------------------------------
    float re[MEL_F], im[MEL_F];
    const int k = tid;
    if (k < N_BINS) {
#pragma unroll
        for (int f = 0; f < MEL_F; ++f) { re[f] = 0.f; im[f] = 0.f; }
        const float* tw = t.twiddle + k;
#pragma unroll 2
        for (int n = 0; n < 199; ++n) {
            const float c = __ldg(tw + (size_t)n * 2 * BIN_LD), s = __ldg(tw + (size_t)n * 2 * BIN_LD + BIN_LD);
            const float4* e4 = reinterpret_cast<const float4*>(E + n * MEL_F);
            const float4* o4 = reinterpret_cast<const float4*>(O + n * MEL_F);
#pragma unroll
            for (int q = 0; q < MEL_F / 4; ++q) {
                const float4 e = e4[q], o = o4[q];
                re[4 * q] = fmaf(e.x, c, re[4 * q]); re[4 * q + 1] = fmaf(e.y, c, re[4 * q + 1]);
                re[4 * q + 2] = fmaf(e.z, c, re[4 * q + 2]); re[4 * q + 3] = fmaf(e.w, c, re[4 * q + 3]);
                im[4 * q] = fmaf(o.x, s, im[4 * q]); im[4 * q + 1] = fmaf(o.y, s, im[4 * q + 1]);
                im[4 * q + 2] = fmaf(o.z, s, im[4 * q + 2]); im[4 * q + 3] = fmaf(o.w, s, im[4 * q + 3]);
            }
        }
        const float sgn = (k & 1) ? -1.f : 1.f;
#pragma unroll
        for (int f = 0; f < MEL_F; ++f) re[f] += sgn * mid[f];
    }
    __syncthreads();                                    // everyone is done reading E/O
    if (k < N_BINS) {
#pragma unroll
        for (int f = 0; f < MEL_F; ++f) P[f * (BIN_LD + 1) + k] = re[f] * re[f] + im[f] * im[f];
    }
    __syncthreads();
    // ---- mel filters + log10 ------------------------------------------------------------------------
    float lmax = -INFINITY;
    for (int i = tid; i < t.n_mels * MEL_F; i += MEL_THREADS) {
        const int m = i / MEL_F, f = i % MEL_F;
        if (f0 + f >= n_frames) continue;
        const int ks = t.f_start[m], len = t.f_len[m];
        const float* w = t.f_w + t.f_off[m];
        const float* p = P + f * (BIN_LD + 1) + ks;
        float acc = 0.f;
        for (int j = 0; j < len; ++j) acc = fmaf(w[j], p[j], acc);
        const float v = log10f(fmaxf(acc, 1e-10f));
        out[(size_t)m * n_frames + f0 + f] = v;
        lmax = fmaxf(lmax, v);
    }
    lmax = warp_max(lmax);
    if ((tid & 31) == 0) smax[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        float m = smax[0];
        for (int i = 1; i < MEL_THREADS / 32; ++i) m = fmaxf(m, smax[i]);
        atomicMax(t.gmax, float_to_ordered(m));
    }
}

__global__ void mel_finish_kernel(float* __restrict__ x, long n, const int* __restrict__ gmax) {
    const float floor_v = ordered_to_float(*gmax) - 8.0f;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = (fmaxf(x[i], floor_v) + 4.0f) / 4.0f;
}
__global__ void mel_reset_kernel(int* gmax) { *gmax = float_to_ordered(-INFINITY); }

// librosa.filters.mel(sr=16000, n_fft=400, n_mels, htk=False, norm="slaney") - the recipe that produced
// whisper/assets/mel_filters.npz (audio.py:97-101) - evaluated in fp64, stored as fp32 band by band.
static double hz_to_mel(double f) { return f >= 1000.0 ? 15.0 + log(f / 1000.0) / (log(6.4) / 27.0) : f / (200.0 / 3); }
static double mel_to_hz(double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : m * (200.0 / 3); }

static bool init_tables(int n_mels) {
    MelTables& t = g_mel[n_mels == 80 ? 0 : 1];
    if (t.twiddle && t.n_mels == n_mels) return true;
    t.n_mels = n_mels;
    std::vector<double> pts(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(8000.0);
    for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<int> start(n_mels), len(n_mels), off(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
        const double enorm = 2.0 / (pts[m + 2] - pts[m]);
        int ks = -1, ke = -1;
        std::vector<float> row(N_BINS);
        for (int k = 0; k < N_BINS; ++k) {
            const double f = 8000.0 * k / 200.0;
            const double lower = (f - pts[m]) / (pts[m + 1] - pts[m]), upper = (pts[m + 2] - f) / (pts[m + 2] - pts[m + 1]);
            const double v = fmax(0.0, fmin(lower, upper)) * enorm;
            row[k] = (float)v;
            if (row[k] != 0.f) { if (ks < 0) ks = k; ke = k; }
        }
        if (ks < 0) { ks = 0; ke = -1; }
        start[m] = ks; len[m] = ke - ks + 1; off[m] = (int)w.size();
        for (int k = ks; k <= ke; ++k) w.push_back(row[k]);
    }
    bool ok = true;
    auto up = [&](auto** dst, const auto& v) {
        typedef typename std::remove_reference<decltype(**dst)>::type T;
        if (cudaMalloc((void**)dst, v.size() * sizeof(T) + 16) != cudaSuccess) { ok = false; return; }
        cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    };
    up(&t.f_start, start); up(&t.f_len, len); up(&t.f_off, off); up(&t.f_w, w);
    ok &= cudaMalloc((void**)&t.twiddle, (size_t)199 * 2 * BIN_LD * sizeof(float)) == cudaSuccess;
    ok &= cudaMalloc((void**)&t.window, 201 * sizeof(float)) == cudaSuccess;
    ok &= cudaMalloc((void**)&t.gmax, sizeof(int)) == cudaSuccess;
    if (!ok) { record_error("log-mel: table allocation failed"); t.twiddle = nullptr; return false; }
    mel_tables_kernel<<<cdiv(199 * BIN_LD, 256), 256>>>(t.twiddle, t.window);
    B200_LAUNCH_CHECK();
    B200_CHECK(cudaDeviceSynchronize());                                // one-time: the frames kernels run on the library stream
    return true;
}

long log_mel_device(const float* d_audio, long n_samples, long padding, int n_mels, float* d_out, cudaStream_t st) {
    if (n_mels != 80 && n_mels != 128) { record_error("log-mel: n_mels must be 80 or 128, got %d", n_mels); return 0; }
    const long n_total = n_samples + padding;
    if (n_total < N_FFT / 2 + 1) { record_error("log-mel: input shorter than the reflect padding"); return 0; }
    if (!init_tables(n_mels)) return 0;
    const MelTables& t = g_mel[n_mels == 80 ? 0 : 1];
    const long n_frames = n_total / HOP;                // 1 + n/hop frames, last one dropped (audio.py:149)
    if (n_frames <= 0) return 0;
    const size_t smem = (size_t)(SEG + 2 * 199 * MEL_F + MEL_F) * sizeof(float);
    static bool attr = false;
    if (!attr) { B200_CHECK(cudaFuncSetAttribute(mel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    mel_reset_kernel<<<1, 1, 0, st>>>(t.gmax);
    B200_LAUNCH_CHECK();
    mel_frames_kernel<<<(unsigned)((n_frames + MEL_F - 1) / MEL_F), MEL_THREADS, smem, st>>>(d_audio, n_samples, n_total, n_frames, t, d_out);
    B200_LAUNCH_CHECK();
    const long n = n_frames * n_mels;
    mel_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_out, n, t.gmax);
    B200_LAUNCH_CHECK();
    return n_frames;
}

}  // namespace b200

using namespace b200;

// d_audio may have been produced on the caller's stream (torch's default stream): wait for it once, then run on the library
// stream under the mel stage timer; the result is complete when the call returns.
extern "C" long logMelSpectrogramDev(const float* d_audio, long n_samples, long padding, int n_mels, float* d_out_mel) {
    use_device();
    B200_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
    cudaStream_t st = S().stream;
    long f;
    {
        StageTimer t(ST_MEL);
        f = log_mel_device(d_audio, n_samples, padding, n_mels, d_out_mel, st);
    }
    B200_CHECK(cudaStreamSynchronize(st));
    return f;
}

// host buffers in and out: staged through two grow-only device buffers (no allocation per call)
extern "C" long logMelSpectrogram(const float* audio, long n_samples, long padding, int n_mels, float* out_mel) {
    static float *d_a = nullptr, *d_o = nullptr;
    static size_t cap_a = 0, cap_o = 0;
    use_device();
    cudaStream_t st = S().stream;
    const long n_frames = (n_samples + padding) / HOP;
    const size_t need_a = (size_t)(n_samples > 0 ? n_samples : 1), need_o = (size_t)(n_frames > 0 ? n_frames : 1) * n_mels;
    if (need_a > cap_a) { if (d_a) cudaFree(d_a); d_a = nullptr; cap_a = 0; if (cudaMalloc((void**)&d_a, need_a * sizeof(float)) == cudaSuccess) cap_a = need_a; }
    if (need_o > cap_o) { if (d_o) cudaFree(d_o); d_o = nullptr; cap_o = 0; if (cudaMalloc((void**)&d_o, need_o * sizeof(float)) == cudaSuccess) cap_o = need_o; }
    if (cap_a < need_a || cap_o < need_o) { record_error("logMelSpectrogram: device allocation failed"); return 0; }
    B200_CHECK(cudaMemcpyAsync(d_a, audio, (size_t)n_samples * sizeof(float), cudaMemcpyHostToDevice, st));
    long f;
    {
        StageTimer t(ST_MEL);
        f = log_mel_device(d_a, n_samples, padding, n_mels, d_o, st);
    }
    B200_CHECK(cudaMemcpyAsync(out_mel, d_o, (size_t)f * n_mels * sizeof(float), cudaMemcpyDeviceToHost, st));
    B200_CHECK(cudaStreamSynchronize(st));
    return f;
}
