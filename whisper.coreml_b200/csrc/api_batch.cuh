// Batched device-resident decode loop (api_batch.cu): interface towards the C ABI files.
#pragma once
#include "decoder_batch.cuh"
#include "sampling.cuh"

namespace b200 {

void batch_set_model(const DbModel& m);       // both decoders loaded: descriptor -> constant memory
void batch_free();                            // decoder closed
void batch_clear_graphs();                    // captured step graphs hold the decode spec by value
bool batch_available();                       // batched step kernel usable for the loaded model (B200_STEP_IMPL unset)?
int batch_max_windows(int nb);
bool run_step_batch_abi(int nb, int text_offset, const float* d_mask, const float* d_x_in);
int decode_windows_batch(const int* windows, int n_windows, const int* initial_tokens, int n_initial, int beam_size, int n_group,
                         float temperature, unsigned long long seed, int sample_len, int without_timestamps, int max_initial_timestamp_index, int* out_tokens, int* out_lengths,
                         float* out_sum_logprobs, float* out_no_speech, int* out_steps);
int batch_timeline(int enable, unsigned long long* out, int cap_ctas);
// decoder1StepFused: one token step of the process-global cache from token histories + logit filters + top-(bs + 1) per beam
void step_fused_abi(const int* tokens_hist, int n_hist, int sample_begin, int text_offset, int without_timestamps,
                    int max_initial_timestamp_index, float* out_logprob, int* out_token);
DecodeSpec decode_spec();                     // api_decode.cu: the spec set by b200SetDecodeSpec

}  // namespace b200
