#!/bin/bash
# round-2 evidence: launch lists of the bench command and one `ncu --set full` capture per hand-written kernel (run under gpurun,
# AFTER the same commands have exited 0 without ncu).  tools/ncu_summaries.py r2 turns the outputs into profiles/.
cd /root/repo; mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --cpu-baseline 0 --long-clip 0 --word-timestamps-pass 0"
$B --sample-len 24 > gpurun_out/plain_r2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r2.log; exit 1; }
$B --sample-len 24 --word-timestamps > gpurun_out/plain_r2_words.log 2>&1 || { echo "plain word run failed"; tail -5 gpurun_out/plain_r2_words.log; exit 1; }
echo "$B --sample-len 224" > gpurun_out/launches_r2.cmd
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv $B --sample-len 224 > gpurun_out/ncu_r2_l.log 2>&1
echo "$B --sample-len 24 --word-timestamps" > gpurun_out/launches_r2_words.cmd
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2_words.csv $B --sample-len 24 --word-timestamps > gpurun_out/ncu_r2_lw.log 2>&1
cap() {  # name, regex, skip, count, extra bench args
  echo "-k regex:$2 -s $3 -c $4, $B $5" > gpurun_out/prof_r2_$1.cmd
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_r2_$1 $B $5 > gpurun_out/ncu_r2_$1.log 2>&1
}
cap decoder_batch decoder_batch_kernel 40 1 "--sample-len 24"
cap sample_update sample_update_batch 20 1 "--sample-len 24"
cap mel_frames mel_frames_kernel 1 1 "--sample-len 2"
cap layernorm layernorm_kernel 40 2 "--sample-len 2"
cap flash_attn flash_attn 36 1 "--sample-len 2"
cap gemm_tcgen05 gemm_tcgen05 146 4 "--sample-len 2"
cap attention_simt attention_simt 2 2 "--sample-len 8 --word-timestamps"
cap median median_kernel 0 1 "--sample-len 8 --word-timestamps"
cap dtw dtw 0 1 "--sample-len 8 --word-timestamps"
cap align align_ 0 3 "--sample-len 8 --word-timestamps"
ls -la gpurun_out/prof_r2_*.ncu-rep | awk '{print $5, $9}'
