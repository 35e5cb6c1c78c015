#!/bin/bash
export ONLY_TC=1
for nb in 10000000 1250000; do
for j2 in 32 24 20 16 12; do
  echo "== nb=$nb J2=$j2"
  YAEL_B200_HAM_J2=$j2 YAEL_B200_HAM_SLOTS=1 YAEL_B200_HAM_PAIR=0 timeout 300 python scripts/prof_hamming.py 10000 $nb 8 100 2>&1 | grep -E "engine 1.*rep 2|e4m3|certify"
done; done
