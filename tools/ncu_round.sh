#!/bin/bash
# ncu captures of one GPU visit (never a bench number): per-kernel `--set full` reports at the bench batch and the launch list
# of the bench command.  usage: bash tools/ncu_round.sh [batch]   -> gpurun_out/prof_*.ncu-rep, gpurun_out/launches.csv
B=${1:-256}
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1 \
    python tools/profile_step.py --batch $B --steps 3 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap attention attention_kernel 14 1          # layer 2 of step 1
cap ffn1 'gemm_pair_kernel<.int.1' 14 1          # FFN1 (bias + erf-GELU epilogue)
cap gemm_epi0 'gemm_pair_kernel<.int.0' 39 3     # step 1, layer 1: QKV (K=768,N=2304), out-proj (768x768), FFN2 (K=3072)
cap round_tc round_tc_kernel 1 1
cap posterior posterior_step 1 1
cap layernorm layernorm_kernel 30 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
