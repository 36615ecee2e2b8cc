#!/bin/bash
# A/B: pre-change library vs the experiment build (ASR_LSTM_OPT: bit 0 = early probe generation, bit 1 = relaxed.gpu)
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw --format=csv
echo "== OLD lib =="
ASR_B200_LIB=$PWD/asr-study_b200/libasr_b200_old.so REPS=6 python profiles/prof_lstm_phases.py 2>&1 | grep " ms"
for opt in 0 1 2 3; do
  echo "== NEW lib ASR_LSTM_OPT=$opt =="
  ASR_LSTM_OPT=$opt REPS=6 python profiles/prof_lstm_phases.py 2>&1 | grep " ms"
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw --format=csv
