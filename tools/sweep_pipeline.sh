#!/bin/bash
# Sweeps of the noise-pipeline knobs on one B200 (profiles/r02b_noise_pipeline_sweeps.txt is the concatenation of the
# runs made while tuning): device time of the C5 job (bench.py's timed region) and every interval between model calls.
#   SONAR_BENCH_ITEMS=k        k of the 8 video latents (1 = the per-GPU share of the 8-GPU split)
#   SONAR_B200_NOISE_PIPELINE  0: batched look-ahead
#   SONAR_B200_PIPELINE_{STEP,FILL,FFT}_CTAS, _STEP_SPLIT, _STEP_B_CTAS, _CHUNK, _MIN_NUMEL: see samplers.py
export SONAR_BENCH_NO_CLOCKS=1
run() { env "$@" python tools/c5_job_probe.py 2>&1 | tail -1 | cut -c1-260; }
for items in 8 4 2 1; do
  export SONAR_BENCH_ITEMS=$items
  run SONAR_B200_NOISE_PIPELINE=0
  run SONAR_B200_PIPELINE_MIN_NUMEL=0
done
export SONAR_BENCH_ITEMS=8
for s in 0.5 0.6 0.7 1.0; do run SONAR_B200_PIPELINE_STEP_SPLIT=$s; done
run SONAR_B200_PIPELINE_FFT_CTAS=0 SONAR_B200_PIPELINE_STEP_SPLIT=1.0
for p in 528 4224 9504; do PLANES=$p python tools/fft_co_probe.py 2>&1 | sed -n 1,2p; done
