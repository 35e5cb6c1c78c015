#!/bin/bash
# A/B of the Hamming tensor pass inside ONE box: packed (3 rows per accumulator, integer epilogue) vs
# one row per accumulator with the constant-norm MAX-tree epilogue; 1-SM vs cta_group::2
export ONLY_TC=1
for nb in 10000000 1250000; do
for cfg in "3 16 0" "1 128 0" "1 128 2"; do
  set -- $cfg
  echo "== nb=$nb HAM_SLOTS=$1 HAM_LDW=$2 HAM_PAIR=$3"
  YAEL_B200_HAM_SLOTS=$1 YAEL_B200_HAM_LDW=$2 YAEL_B200_HAM_PAIR=$3 YAEL_B200_TF32_DEBUG=${DBG:-0} timeout 300 python scripts/prof_hamming.py 10000 $nb 8 100 2>&1 | grep -E "engine 1.*rep 2|phase|agree|Error|error|issuer|epilogue"
done; done
