// Audio ingest on the device (the step in front of log_mel_spectrogram): PCM -> mono fp32 and rational resampling
// to 16 kHz.  The reference shells out to ffmpeg (whisper/audio.py:45-62: `-ac 1 -ar 16000 -f s16le`, then int16 / 32768); ffmpeg
// is not part of the reference's sources, so what is restated here is a published polyphase design instead - the Kaiser(5.0)
// windowed-sinc low-pass of scipy.signal.resample_poly (cut-off 1 / max(up, down), half length 10 * max(up, down), unit DC gain
// times `up`) - and parity is checked against a numpy restatement of that design (tests/test_mel_gpu.py), NOT against ffmpeg's
// own resampler.
//
//   y[n] = sum_k x[k] * h[n * down - k * up + half],   |n * down - k * up| <= half      (~20 * max(up, down) / up taps per output)
//
// HBM bound: one thread per output sample; a warp's outputs read overlapping input windows (L1 / L2 serve the reuse) and the
// filter taps with stride `up` (the table is a few tens of KB and stays in L1).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "whisper_b200.h"

namespace b200 {

__global__ void __launch_bounds__(256) pcm16_to_mono_kernel(const short* __restrict__ pcm, long n_frames, int channels, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames) return;
    float s = 0.f;
    for (int c = 0; c < channels; ++c) s += (float)pcm[i * channels + c];
    out[i] = s / (32768.f * channels);                  // audio.py:62 scaling; channels averaged like ffmpeg's default stereo down-mix
}

__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long n_in, const float* __restrict__ h, int half, int up, int down,
                                                       float* __restrict__ y, long n_out) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_out) return;
    const long t = n * down;
    long k0 = (t - half + up - 1) / up;                 // ceil((t - half) / up), t - half may be negative
    if (t - half < 0) k0 = 0;
    long k1 = (t + half) / up;
    if (k1 > n_in - 1) k1 = n_in - 1;
    float acc = 0.f;
    for (long k = k0; k <= k1; ++k) acc = fmaf(x[k], __ldg(h + (t - k * up + half)), acc);
    y[n] = acc;
}

static double bessel_i0(double v) {                     // power series, converges quickly for |v| <= 5
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 64; ++k) { term *= (v / (2.0 * k)) * (v / (2.0 * k)); sum += term; if (term < 1e-18 * sum) break; }
    return sum;
}

static long gcd_l(long a, long b) { while (b) { const long t = a % b; a = b; b = t; } return a; }

}  // namespace b200

using namespace b200;

extern "C" long b200ResampleDev(const float* d_in, long n_in, int sr_in, float* d_out, long out_capacity) {
    const int sr_out = 16000;
    if (n_in <= 0 || sr_in <= 0) return 0;
    const long g = gcd_l(sr_in, sr_out);
    const int up = (int)(sr_out / g), down = (int)(sr_in / g);
    const long n_out = (n_in * up + down - 1) / down;
    if (!d_out) return n_out;                           // size query
    if (out_capacity < n_out) { record_error("b200ResampleDev: output holds %ld samples, %ld needed", out_capacity, n_out); return 0; }
    if (up == 1 && down == 1) { B200_CHECK(cudaMemcpyAsync(d_out, d_in, (size_t)n_in * sizeof(float), cudaMemcpyDeviceToDevice, 0)); B200_CHECK(cudaStreamSynchronize(0)); return n_out; }
    const int max_rate = up > down ? up : down, half = 10 * max_rate;
    std::vector<double> hd(2 * (size_t)half + 1);
    const double fc = 1.0 / max_rate, i0b = bessel_i0(5.0);
    double sum = 0.0;
    for (int j = -half; j <= half; ++j) {
        const double a = M_PI * fc * j, sinc = j == 0 ? 1.0 : sin(a) / a, r = (double)j / half;
        const double w = bessel_i0(5.0 * sqrt(r * r < 1.0 ? 1.0 - r * r : 0.0)) / i0b;
        hd[j + half] = fc * sinc * w;
        sum += hd[j + half];
    }
    std::vector<float> hf(hd.size());
    for (size_t j = 0; j < hd.size(); ++j) hf[j] = (float)(hd[j] / sum * up);
    float* d_h = nullptr;
    if (cudaMalloc((void**)&d_h, hf.size() * sizeof(float)) != cudaSuccess) { record_error("b200ResampleDev: device allocation failed"); return 0; }
    B200_CHECK(cudaMemcpyAsync(d_h, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice, 0));
    resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, 0>>>(d_in, n_in, d_h, half, up, down, d_out, n_out);
    B200_LAUNCH_CHECK();
    B200_CHECK(cudaStreamSynchronize(0));
    cudaFree(d_h);
    return n_out;
}

extern "C" long b200Pcm16ToMonoDev(const short* d_pcm, long n_frames, int channels, float* d_out) {
    if (n_frames <= 0 || channels < 1) return 0;
    pcm16_to_mono_kernel<<<(unsigned)((n_frames + 255) / 256), 256, 0, 0>>>(d_pcm, n_frames, channels, d_out);
    B200_LAUNCH_CHECK();
    B200_CHECK(cudaStreamSynchronize(0));
    return n_frames;
}
