#!/bin/bash
# Round-2 ncu captures, final set: the hot kernels of config 2 (SuperPoint + LightGlue) and config 3 (LoFTR) as shipped.
#   tools/ncu_capture5.sh <outdir>     (writes <name>.csv = `--page raw --csv` of every capture; the .ncu-rep files are removed)
set -u
OUT=${1:-gpurun_out/ncu}
mkdir -p "$OUT"
C2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
C3="python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --f1-pairs 0"
cap() {  # name cmd regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s "$4" -c "$5" -f -o "$OUT/$1" $2 > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
  ncu -i "$OUT/$1.ncu-rep" --page raw --csv > "$OUT/$1.csv" 2>/dev/null
  rm -f "$OUT/$1.ncu-rep"
}
cap conv_fused_pair "$C2" 'tc_conv3x3_c64_pair_kernel<\(bool\)1>' 8 1
cap conv_c64_pair "$C2" 'tc_conv3x3_c64_pair_kernel<\(bool\)0>' 16 2
cap conv_halo_pair "$C2" 'tc_conv3x3_halo_pair_kernel<\(int\)128' 48 6
cap attn "$C2" 'tc_attn_kernel' 36 2
cap gemm_store "$C2" 'tc_gemm_f16_kernel.*EpiStore' 108 3
cap gemm_qkv "$C2" 'tc_gemm_f16_kernel.*EpiQKVRotary' 18 1
cap loftr_conv_halo_pair "$C3" 'tc_conv3x3_halo_pair_kernel' 168 3
cap loftr_gemm "$C3" 'tc_gemm_f16_kernel' 288 2
ls -la "$OUT"
