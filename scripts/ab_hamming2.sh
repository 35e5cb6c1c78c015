#!/bin/bash
# decomposition of the packed Hamming pass with the bring-up switches (one box)
export ONLY_TC=1
for cfg in "16 0 512" "128 0 512" "16 0 1" "128 0 1" "16 0 2" "16 0 16" "16 0 17" "16 0 18" "16 2 0" "128 2 1"; do
  set -- $cfg
  echo "== HAM_LDW=$1 HAM_PAIR=$2 DEBUG=$3"
  YAEL_B200_HAM_LDW=$1 YAEL_B200_HAM_PAIR=$2 YAEL_B200_TF32_DEBUG=$3 timeout 300 python scripts/prof_hamming.py 10000 10000000 8 100 2>&1 | grep -E "e4m3 pass|clock|issuer|epilogue|Error|error"
done
