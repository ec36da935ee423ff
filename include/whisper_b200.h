/* libwhisper_b200.so - C ABI of the B200-native Whisper backend.
 *
 * Part 1 is, prototype for prototype, the plugin surface of wangchou/whisper.coreml
 * (reference coreml/coreml.h:5-31, bodies in coreml/coreml.mm).  A reference user switches by
 * loading this library instead of ./coreml/<model>/coreml.so (whisper/coreml.py:21); see
 * INTEGRATION.md for the ctypes stub.  All pointers in Part 1 are HOST buffers owned by the
 * caller (whisper/coreml.py:58-59,137-140); every call is synchronous: outputs are complete in
 * host memory on return.  State (Xa, CK/CV, the 448-slot self-attention KV cache) is
 * process-global, exactly like the reference's MLMultiArray globals (coreml.mm:18-23).
 *
 * Part 2 holds additive entry points (device-resident fast path, log-mel, word-timestamp
 * kernels, error reporting).  Nothing in Part 1 changes meaning when Part 2 is used.
 *
 * Weights: `load*` receive a path, like the reference.  For this backend the path names a
 * `.b2w` container written by whisper.coreml_b200/export.py (the analogue of convert_*.py +
 * convert_coreml.sh): loadEncoder gets the model FOLDER (reads <folder>/Encoder.b2w),
 * the other three get the file itself (CrossKV.b2w / Decoder.b2w).
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#if __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 - reference plugin ABI (coreml/coreml.h)
 * ---------------------------------------------------------------------------------------- */

/* coreml.h:5  / coreml.mm:35-65.  Idempotent.  Allocates Xa (1,1500,n_state). */
void loadEncoder(const char* modelFolderPath, int n_layer, int n_state, int n_mels);
/* coreml.h:6  / coreml.mm:102-114 */
void closeEncoder();
/* coreml.h:7  / coreml.mm:67-100.  melSegment: (1, n_mels, 3000) fp32.  Result stays on device (Xa). */
void encoderPredict(float* melSegment);

/* coreml.h:9  / coreml.mm:122-146.  Allocates CK/CV for n_layer decoder layers. */
void loadCrossKV(const char* modelPath, int n_layer, int n_state);
/* coreml.h:10 / coreml.mm:171-181 */
void closeCrossKV();
/* coreml.h:11 / coreml.mm:148-169.  Xa -> CK, CV (device resident). */
void crossKVPredict();

/* coreml.h:13 / coreml.mm:201-243.  beam_size fixes the number of KV-cache slots (coreml.mm:230). */
void loadDecoder256(const char* modelPath, int n_layer, int n_state, int n_head, int n_alignment_head, int beam_size);
/* coreml.h:14 / coreml.mm:329-352 */
void closeDecoder256();
/* coreml.h:15-21 / coreml.mm:279-327.  Prefill of one beam; its 256 K/V rows are copied into
 * slot `beam_idx` of the 448-row cache (coreml.mm:313-326). */
void decoder256Predict(
    float* x,                       /* (1, 256, n_state)  embedded + zero-padded tokens      */
    float* qk_mask,                 /* (256, 256)         additive mask                      */
    float* out_x,                   /* (1, 256, n_state)  ln(x)                              */
    float* out_cross_head_weights,  /* (n_alignment_head, 256, 1500) raw (pre-softmax) QK    */
    int beam_idx);

/* coreml.h:23 / coreml.mm:369-402 */
void loadDecoder1(const char* modelPath, int n_layer, int n_state, int n_head, int n_vocab);
/* coreml.h:24 / coreml.mm:446-459 */
void closeDecoder1();
/* coreml.h:25 / coreml.mm:251-277.  cache[:, b, :text_offset] = cache[:, indices[b], :text_offset]. */
void rearrange_mkv(int* indices, int text_offset);
/* coreml.h:26-31 / coreml.mm:404-444.  One beam-batched token step; the new K/V row is written
 * at `text_offset` of every layer x beam (coreml.mm:432-443). */
void decoder1Predict(
    float* x,         /* (bs, 1, n_state)                                                    */
    float* qk_mask,   /* (1, 449); (1, 450) when bs == 1 (whisper/decoder.py:247-248)        */
    int text_offset,
    float* out_x);    /* (bs, 1, n_vocab) logits (coreml.mm:398,430; the header comment in   */
                      /*  the reference says n_state, the implementation writes n_vocab)     */

/* ------------------------------------------------------------------------------------------
 * Part 2 - additive entry points
 * ---------------------------------------------------------------------------------------- */

/* Errors: Part 1 keeps the reference's `void` signatures (coreml.mm logs and continues);
 * failures are also recorded here.  Returns the number of errors since the last call and
 * copies the most recent message.  Every message is also written to stderr as it happens; a
 * bounded wait of the persistent step kernel that gives up (a protocol bug, never a normal
 * condition) additionally logs which wait, the stage every CTA had reached and the words the
 * long waits were polling (csrc/decoder_batch.cu: db_fault_word). */
int  b200LastError(char* buf, int buf_len);
/* Select the CUDA device (default 0) before any load*.  One process drives one GPU. */
void b200SetDevice(int device);
/* Which alignment heads decoder256 reports: n pairs (layer, head), in CHW row order
 * (whisper/decoder.py:306-308).  Default: all heads of the last n_layer/2 layers
 * (whisper/model.py:55-58), truncated to n_alignment_head. */
void b200SetAlignmentHeads(const int* layer_head_pairs, int n);
/* Token ids used by the device-side logit filters / beam search (whisper/decoding.py:450-532):
 * suppress[] is SuppressTokens' list; blank[] is tokenizer.encode(" "). */
void b200SetDecodeSpec(int sot, int eot, int no_timestamps, int timestamp_begin, int no_speech,
                       const int* suppress, int n_suppress, const int* blank, int n_blank);

/* whisper/audio.py:110-157.  audio: n_samples fp32 (16 kHz) on the HOST; `padding` zeros are
 * appended; out: (n_mels, (n_samples+padding)/160) fp32 on the HOST.  Returns frames written. */
long logMelSpectrogram(const float* audio, long n_samples, long padding, int n_mels, float* out_mel);
/* Same, device pointers; the result may stay on the GPU to feed encoderPredictWindows. */
long logMelSpectrogramDev(const float* d_audio, long n_samples, long padding, int n_mels, float* d_out_mel);

/* Audio ingest on the device (the step in front of log_mel_spectrogram; the reference shells out to ffmpeg, whisper/audio.py:45-62).
 * b200Pcm16ToMonoDev: interleaved int16 PCM (DEVICE) -> mono fp32 in [-1, 1) (channels averaged, / 32768 as audio.py:62).
 * b200ResampleDev: rational resampling of a DEVICE fp32 waveform to 16 kHz with the Kaiser(5.0) windowed-sinc polyphase design of
 * scipy.signal.resample_poly; d_out == NULL returns the output length ceil(n_in * 16000 / sr_in) without computing. */
long b200Pcm16ToMonoDev(const short* d_pcm, long n_frames, int channels, float* d_out);
long b200ResampleDev(const float* d_in, long n_in, int sr_in, float* d_out, long out_capacity);
/* FLAC container decode on the HOST (the reference leaves every container to ffmpeg, whisper/audio.py:45-62).  b200FlacInfo reads
 * STREAMINFO (returns 0, or -1 if `data` is not a FLAC stream); md5_16 receives the encoder's MD5 of the decoded PCM, which is what
 * the parity test checks the decoder against.  b200FlacDecode writes interleaved int32 samples (right-justified, bits_per_sample
 * significant bits) and returns the samples per channel, or -1. */
int  b200FlacInfo(const unsigned char* data, long n_bytes, int* sample_rate, int* channels, int* bits_per_sample, long* total_samples,
                  unsigned char* md5_16);
long b200FlacDecode(const unsigned char* data, long n_bytes, int* out_interleaved, long cap_samples_per_channel);

/* Batched encoder + crossKV over `n_windows` independent 30-s windows (the reference loops
 * windows in Python, whisper/transcribe.py:276-306).  d_mel: DEVICE (n_mels, total_frames) fp32
 * log-mel of the whole file; window w reads frames [seeks[w], seeks[w]+3000) zero-padded past
 * `total_frames` (pad_or_trim, whisper/audio.py:65-88).  Results stay on the device per window. */
void encoderPredictWindows(const float* d_mel, long total_frames, const int* seeks, int n_windows);
/* Same with the file's content length: frames at or past `content_frames` read as ZEROS, which is what the reference feeds the
 * encoder for a partial last window - it slices mel[:, seek : seek + min(3000, content_frames - seek)] out of the mel that
 * carries 30 s of trailing padding and zero-pads the slice (whisper/transcribe.py:143, 286-290). */
void encoderPredictWindowsContent(const float* d_mel, long total_frames, long content_frames, const int* seeks, int n_windows);
void crossKVPredictWindows(int n_windows);
/* Make window w the current one for decoder256Predict/decoder1Predict/b200DecodeWindow. */
void b200SelectWindow(int w);

/* Device-resident DecodingTask.run at temperature 0 for the current window
 * (whisper/decoding.py:707-816): prefill of `n_initial` tokens, then up to `sample_len` steps of
 * decoder1 + logit filters + greedy/beam update, all on the GPU, one host sync at the end.
 * beam_size <= 0 -> GreedyDecoder (:299-325, the reference's beam_size=None), else BeamSearchDecoder
 * (:328-431) with that many beams (<= the beam_size given to loadDecoder256), patience 1.
 * Outputs (HOST):
 *   out_tokens     (n_cand, 449) int32  candidate sequences incl. initial tokens, EOT padded
 *   out_lengths    (n_cand)      int32  tokens before the first EOT after sample_begin
 *   out_sum_logprobs (n_cand)    float
 *   out_no_speech  (1)           float  softmax(logits[sot_index])[no_speech]
 * n_cand = max(beam_size, 1); unused rows have length -1.  Ranking (MaximumLikelihoodRanker, :217-240) is
 * left to the caller.  Returns the number of sampling steps executed (the prefill step included). */
int b200DecodeWindow(const int* initial_tokens, int n_initial, int beam_size, int sample_len,
                     int without_timestamps, int max_initial_timestamp_index,
                     int* out_tokens, int* out_lengths, float* out_sum_logprobs, float* out_no_speech);
/* The same for n_windows encoded windows (ids as for b200SelectWindow).  Independent windows are decoded CONCURRENTLY on up to
 * 8 decode lanes (B200_DECODE_LANES), each lane with its own KV cache, decode states and stream and its share of the SMs; the
 * windows of a lane (more than one from nine windows up, or with fewer lanes) advance together in ONE batched step kernel, so a
 * lane streams the decoder weights once per step for all of them.  The token loop is a latency-bound chain of dependent stages:
 * lanes overlap each other's stalls, batching saves weight traffic.
 * Outputs are the b200DecodeWindow outputs per window, window-major: out_tokens (n_windows, n_cand, 449), out_lengths and
 * out_sum_logprobs (n_windows, n_cand), out_no_speech (n_windows), out_steps (n_windows, may be NULL).
 * Returns the total number of sampling steps.  This is what transcribe() calls (whisper/transcribe.py:276-306 loops windows). */
int b200DecodeWindows(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size,
                      int sample_len, int without_timestamps, int max_initial_timestamp_index, int* out_tokens,
                      int* out_lengths, float* out_sum_logprobs, float* out_no_speech, int* out_steps);
/* The same with sampling at temperature > 0 (the reference's fallback decodes, whisper/transcribe.py:188-228): beam_size <= 0,
 * n_group = best_of independent samples per window (GreedyDecoder over n_group rows, whisper/decoding.py:299-325, 761), every
 * token drawn from Categorical(logits / temperature) after the logit filters (:307-310) with a counter-based generator
 * (Philox4x32-10 keyed by `seed`; window, row, step and token index form the counter, so a decode is reproducible and
 * independent of batching); sum_logprobs accumulate log_softmax(logits) of the drawn tokens (:311-313).  n_cand = n_group.
 * temperature == 0 is b200DecodeWindows. */
int b200DecodeWindowsEx(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int n_group,
                        float temperature, unsigned long long seed, int sample_len, int without_timestamps,
                        int max_initial_timestamp_index, int* out_tokens, int* out_lengths, float* out_sum_logprobs,
                        float* out_no_speech, int* out_steps);

/* decoder1 with on-device filters + log-softmax + top-(bs+1) instead of returning full logits:
 * tokens_hist (bs, n_hist) int32 HOST = whole context so far (last column is fed to the
 * decoder), out_logprob/out_token (bs, bs+1). */
void decoder1StepFused(const int* tokens_hist, int n_hist, int sample_begin, int text_offset,
                       int without_timestamps, int max_initial_timestamp_index,
                       float* out_logprob, int* out_token);

/* whisper/timing.py:19-54: median filter of odd `width` along the last axis with reflect padding.
 * x, y: (rows, len) fp32 HOST. */
void medianFilter(const float* x, float* y, long rows, int len, int width);
/* whisper/timing.py:57-105 (dtw_cpu tie rule + backtrace).  x: (N, M) fp32 HOST cost matrix;
 * out_i/out_j: capacity N+M.  Returns the path length. */
int dtw(const float* x, int N, int M, int* out_i, int* out_j);
/* whisper/timing.py:194-204 on the device for the CHW of the last decoder256Predict / alignment
 * pass: softmax over frames[:num_frames/2] -> z-score over tokens -> median(width) -> mean over
 * heads -> rows [n_skip, n_tokens-1) -> DTW(-matrix).  tokens = [*sot_sequence, no_timestamps, *text, eot]
 * (timing.py:176-183), n_skip = len(sot_sequence).  Runs one decoder256 pass for the current window (it
 * overwrites KV slot 0).  Returns the path length (out_i/out_j capacity n_tokens + num_frames/2);
 * optional HOST outputs: out_matrix (n_tokens-1-n_skip, num_frames/2) and out_text_token_probs
 * (n_tokens-n_skip-2): softmax over [:eot] at each text token (timing.py:187-190). */
int b200AlignTokens(const int* tokens, int n_tokens, int n_skip, int num_frames, int medfilt_width,
                    int* out_i, int* out_j, float* out_matrix, float* out_text_token_probs);

/* Per-stage device time (ms, CUDA events) accumulated since the last reset - the analogue of
 * whisper/coreml.py:9-13,247-263.  stages: 0 mel, 1 encoder, 2 crossKV, 3 decoder256, 4 decoder1,
 * 5 sampling, 6 align. */
void b200GetStageTimes(float* out_ms7, int reset);
/* Number of kernel launches issued by this library since process start (bench.py gpu_launches). */
long b200KernelLaunchCount();

/* Test hooks (used by tests/ only): C[M,N] = A[M,K] * B[N,K]^T (+bias)(gelu) through the
 * tcgen05 GEMM and through the SIMT checker; device pointers, bf16 inputs. */
void b200TestGemm(const void* dA, const void* dB, const float* dBias, void* dC, int M, int N, int K,
                  int out_fp32, int gelu, int use_simt);
/* Force the GEMM tile configuration of every following tcgen05 GEMM: 0 automatic, 1 = 128x128, 2 = 128x256,
 * 3 = CTA pair (cta_group::2) 256x256. */
void b200TestGemmTile(int sel);
/* Average device time in ms of `iters` back-to-back tcgen05 GEMMs; mode bits: 1 bias, 2 GELU, 4 fp32 output + fp32 residual. */
float b200TestGemmTime(const void* dA, const void* dB, void* dC, int M, int N, int K, int mode, int iters);
/* clock64 marks of CTA 0 of one tcgen05 GEMM launch, out[tile * 16 + k]: epilogue warp 2: 0 entered, 1 accumulator ready,
 * 2+3c / 3+3c / 4+3c loads of chunk c issued / landed / chunk done, 14 accumulator released; 15 = MMA thread: tile issued. */
int b200TestGemmTimeline(const void* dA, const void* dB, void* dC, int M, int N, int K, int mode, unsigned long long* out, int cap_tiles);
/* State read-back as fp32 HOST arrays in the reference's layouts: Xa (1500, d) of window w;
 * CK (Ld,H,64,1500) / CV (Ld,H,1500,64) of window w; logical KV cache rows (2Ld, bs, n_rows, d). */
void b200TestGetXa(float* out, int w);
void b200TestGetCrossKV(float* out_ck, float* out_cv, int w);
void b200TestGetKV(float* out, int n_rows);
/* softmax(Q K^T) V per head over a fused DEVICE [batch][n_tok][3*heads*64] bf16 QKV buffer ->
 * [batch][n_tok][heads*64] bf16: the tcgen05 flash-attention kernel, or the SIMT checker. */
void b200TestAttention(const void* dQKV, void* dO, int n_tok, int heads, int batch, int use_simt);
/* Average device time in ms of `iters` back-to-back launches of the tcgen05 flash-attention kernel. */
float b200TestAttentionTime(const void* dQKV, void* dO, int n_tok, int heads, int batch, int iters);
/* clock64 marks of CTA (0,0,0) of one flash-attention launch, out[block * 16 + k]: k = 0-4 softmax warp (enter, S ready,
 * S loaded, P buffer free, P published), 5-7 MMA thread (loop top, next QK issued, P seen); returns the key-block count. */
int b200TestAttentionTimeline(const void* dQKV, void* dO, int n_tok, int heads, int batch, unsigned long long* out, int cap_blocks);
/* Stage timeline of the persistent decoder step kernel: enable=1 starts recording CTA 0's %globaltimer after every
 * grid barrier of the following steps; enable=0 copies up to `cap` timestamps (ns) to `out` and returns the count. */
int b200TestStepTimeline(int enable, unsigned long long* out, int cap);

#if __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
