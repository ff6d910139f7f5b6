#!/bin/bash
# Round-2 ncu captures, second set: the LightGlue linears at full batch (layer 0 of a warm step), the streaming similarity
# kernel on BASELINE config 5, and the launch list of one whole steady-state step of config 2.
#   tools/ncu_capture2.sh <outdir>
set -u
OUT=${1:-gpurun_out/ncu}
mkdir -p "$OUT"
C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
C5="python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
cap() {  # name cmd regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s "$4" -c "$5" -f -o "$OUT/$1" $2 > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
}
# one warm step = 54 EpiStore launches: skip two steps, take out_proj / ffn0 / ffn3 of layer 0 (all 64 pairs active)
cap gemm_store "$C2" 'tc_gemm_tf32_kernel.*EpiStore' 108 3
cap gemm_qkv "$C2" 'tc_gemm_tf32_kernel.*EpiQKVRotary' 18 1
cap simreduce "$C5" 'tc_simreduce_kernel' 4 2
cap nms "$C2" 'nms_kernel' 2 1
ls -la "$OUT"
