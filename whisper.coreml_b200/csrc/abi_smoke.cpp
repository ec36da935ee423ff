// ABI smoke / latency driver: the analogue of coreml/coremlTest.cpp:26-103 for libwhisper_b200.so.
//
//     build/abi_smoke <model folder> <n_audio_layer> <n_text_layer> <n_state> <n_mels> <n_vocab> [beam slots = 5] [n_alignment_head = 0]
//
// Plain C++ against include/whisper_b200.h Part 1 only (no CUDA, no torch): loads the four sub-models from the .b2w files written
// by export.py, calls every reference entry point with host buffers a few times, prints milliseconds per call, closes everything
// and does it all again (the reference's "Run 0 / Run 1": loads must be idempotent and close -> load must work).  Any error the
// library recorded fails the run (exit code 1), so tools/run_gpu_tests.sh can use it as a check.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "whisper_b200.h"

static int check(const char* where) {
    char buf[1024];
    const int n = b200LastError(buf, sizeof(buf));
    if (n) std::fprintf(stderr, "abi_smoke: %d error(s) in %s: %s\n", n, where, buf);
    return n;
}

template <class F>
static double ms_of(F f) {
    const auto t0 = std::chrono::steady_clock::now();
    f();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s <folder> <n_audio_layer> <n_text_layer> <n_state> <n_mels> <n_vocab> [beam slots] [n_alignment_head]\n", argv[0]);
        return 2;
    }
    const std::string folder = argv[1];
    const int n_audio_layer = std::atoi(argv[2]), n_text_layer = std::atoi(argv[3]), n_state = std::atoi(argv[4]);
    const int n_mels = std::atoi(argv[5]), n_vocab = std::atoi(argv[6]);
    const int bs = argc > 7 ? std::atoi(argv[7]) : 5, n_align = argc > 8 ? std::atoi(argv[8]) : 0;
    const int n_head = n_state / 64, max_n_ctx = 256;
    int failures = 0;

    std::vector<float> mel((size_t)n_mels * 3000);
    for (size_t i = 0; i < mel.size(); ++i) mel[i] = 0.25f * std::sin(0.001f * (float)i);            // any bounded signal
    std::vector<float> x256((size_t)max_n_ctx * n_state, 0.f), mask256((size_t)max_n_ctx * max_n_ctx), out256(x256.size());
    std::vector<float> chw((size_t)(n_align > 0 ? n_align : 1) * max_n_ctx * 1500);
    const int n_prompt = 3;
    for (int r = 0; r < n_prompt; ++r)
        for (int c = 0; c < n_state; ++c) x256[(size_t)r * n_state + c] = 0.02f * std::cos(0.37f * (float)(r * n_state + c));
    for (int r = 0; r < max_n_ctx; ++r)                                                               // whisper/decoder.py:212-213
        for (int c = 0; c < max_n_ctx; ++c) mask256[(size_t)r * max_n_ctx + c] = (c <= r && c < n_prompt) ? 0.f : -INFINITY;
    std::vector<float> x1((size_t)bs * n_state), mask1(450, 0.f), logits((size_t)bs * n_vocab);
    for (size_t i = 0; i < x1.size(); ++i) x1[i] = 0.02f * std::sin(0.11f * (float)i);
    std::vector<int> perm(bs);

    for (int run = 0; run < 2; ++run) {
        std::printf("///////// run %d\n", run);
        std::printf("load       %8.2f ms\n", ms_of([&] {
                        loadEncoder(folder.c_str(), n_audio_layer, n_state, n_mels);
                        loadCrossKV((folder + "/CrossKV.b2w").c_str(), n_text_layer, n_state);
                        loadDecoder256((folder + "/Decoder.b2w").c_str(), n_text_layer, n_state, n_head, n_align, bs);
                        loadDecoder1((folder + "/Decoder.b2w").c_str(), n_text_layer, n_state, n_head, n_vocab);
                        loadEncoder(folder.c_str(), n_audio_layer, n_state, n_mels);                  // idempotent (coreml.mm:43-45)
                    }));
        failures += check("load");
        for (int i = 0; i < 3; ++i) std::printf("encoder    %8.3f ms\n", ms_of([&] { encoderPredict(mel.data()); }));
        for (int i = 0; i < 3; ++i) std::printf("crossKV    %8.3f ms\n", ms_of([&] { crossKVPredict(); }));
        for (int b = 0; b < bs; ++b)
            std::printf("decoder256 %8.3f ms (beam slot %d)\n", ms_of([&] { decoder256Predict(x256.data(), mask256.data(), out256.data(), n_align > 0 ? chw.data() : nullptr, b); }), b);
        failures += check("encoder / crossKV / decoder256");
        int text_offset = n_prompt;
        for (int i = 0; i < 5; ++i, ++text_offset) {
            for (int c = 0; c < 449; ++c) mask1[c] = (c < text_offset || c == 448) ? 0.f : -INFINITY;  // whisper/decoder.py:242-245
            mask1[449] = -INFINITY;                                                                    // the bs == 1 column (:246-248)
            std::printf("decoder1   %8.3f ms (text_offset %d)\n", ms_of([&] { decoder1Predict(x1.data(), mask1.data(), text_offset, logits.data()); }), text_offset);
            for (int b = 0; b < bs; ++b) perm[b] = (b + 1) % bs;
            rearrange_mkv(perm.data(), text_offset + 1);                                               // after text_offset was incremented (decoding.py:182,392)
        }
        failures += check("decoder1 / rearrange_mkv");
        int bad = 0;
        for (size_t i = 0; i < logits.size(); ++i) bad += !std::isfinite(logits[i]);
        for (size_t i = 0; i < (size_t)n_prompt * n_state; ++i) bad += !std::isfinite(out256[i]);
        if (bad) { std::fprintf(stderr, "abi_smoke: %d non-finite outputs\n", bad); ++failures; }
        std::printf("logits[0][0..3] = %.5f %.5f %.5f %.5f   kernel launches so far %ld\n", logits[0], logits[1], logits[2], logits[3], b200KernelLaunchCount());
        closeDecoder1(); closeDecoder256(); closeCrossKV(); closeEncoder();
        failures += check("close");
    }
    std::printf(failures ? "abi_smoke: FAILED\n" : "abi_smoke: ok\n");
    return failures ? 1 : 0;
}
