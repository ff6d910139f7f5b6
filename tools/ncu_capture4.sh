#!/bin/bash
# Round-2 ncu captures, fourth set: the CTA-pair (cta_group::2) conv kernels inside a config-2 step.   tools/ncu_capture4.sh <outdir>
set -u
OUT=${1:-gpurun_out/ncu}
mkdir -p "$OUT"
C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
cap() {  # name cmd regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s "$4" -c "$5" -f -o "$OUT/$1" $2 > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
  ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.csv" 2>/dev/null
}
cap conv_fused_pair "$C2" 'tc_conv3x3_c64_pair_kernel<\(bool\)1>' 8 1
cap conv_c64_pair "$C2" 'tc_conv3x3_c64_pair_kernel<\(bool\)0>' 16 2
cap conv_halo_pair "$C2" 'tc_conv3x3_halo_pair_kernel<\(int\)128' 48 3
ls -la "$OUT"
