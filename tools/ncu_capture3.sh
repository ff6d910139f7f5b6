#!/bin/bash
# Round-2 ncu captures, third set (after the split-fp16 rewrites): attention, the QKV / FFN linears, NMS on config 2 and the
# backbone conv + QKV linear of LoFTR (config 3).   tools/ncu_capture3.sh <outdir>
set -u
OUT=${1:-gpurun_out/ncu}
mkdir -p "$OUT"
C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
C3="python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
cap() {  # name cmd regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s "$4" -c "$5" -f -o "$OUT/$1" $2 > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
}
cap attn "$C2" 'tc_attn_kernel' 36 2
cap gemm_store "$C2" 'tc_gemm_f16_kernel.*EpiStore' 108 3
cap gemm_qkv "$C2" 'tc_gemm_f16_kernel.*EpiQKVRotary' 18 1
cap nms "$C2" 'nms_fast_kernel' 2 1
cap conv_fused "$C2" 'tc_conv3x3_c64_kernel<\(bool\)1>' 8 1
cap conv_generic "$C2" 'tc_conv3x3_kernel<\(int\)128, \(int\)3, \(bool\)0>' 36 2
ls -la "$OUT"
