#!/bin/bash
export ONLY_TC=1 YAEL_B200_HAM_SLOTS=1 YAEL_B200_HAM_J2=16
for cfg in "0 0" "0 1" "0 3" "0 17" "0 19" "2 1" "2 3" "0 513" "0 512"; do
  set -- $cfg
  echo "== PAIR=$1 DEBUG=$2"
  YAEL_B200_HAM_PAIR=$1 YAEL_B200_TF32_DEBUG=$2 timeout 300 python scripts/prof_hamming.py 10000 10000000 8 100 2>&1 | grep -E "e4m3|issuer|epilogue"
done
