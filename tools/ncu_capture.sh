#!/bin/bash
# Round-2 ncu captures of the hot kernels of the bench step (run under gpurun on ONE GPU):
#   tools/ncu_capture.sh <outdir>
# writes <outdir>/<name>.ncu-rep (--set full, source import) + the launch list of one steady-state step.
set -u
OUT=${1:-gpurun_out/ncu}
mkdir -p "$OUT"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "$OUT/$1" $CMD > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
}
cap attn 'tc_attn_kernel' 20 1
cap gemm 'tc_gemm_tf32_kernel' 60 2
cap conv_fused 'tc_conv3x3_c64_kernel<\(bool\)1>' 6 1
cap conv_generic 'tc_conv3x3_kernel<\(int\)128, \(int\)3, \(bool\)0>' 30 2
cap conv_c64 'tc_conv3x3_c64_kernel<\(bool\)0>' 10 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 260 --csv --log-file "$OUT/launches_step.csv" $CMD > "$OUT/launches.log" 2>&1
echo "launch list rc=$?"
ls -la "$OUT"
