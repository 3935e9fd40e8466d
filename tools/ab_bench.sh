#!/bin/bash
# A/B on ONE box: bench.py (no e2e / cpu legs) for each library variant / env setting given as "name:lib:ENV=VAL,..."
# usage: bash tools/ab_bench.sh new:musediffusion_b200/libmusediff_b200.so: old:musediffusion_b200/libmusediff_old.so: ...
mkdir -p gpurun_out
cp musediffusion_b200/libmusediff_b200.so /tmp/lib_current.so
for spec in "$@"; do
  name="${spec%%:*}"; rest="${spec#*:}"; lib="${rest%%:*}"; envs="${rest#*:}"
  cp "$lib" /tmp/lib_variant.so; cp /tmp/lib_variant.so musediffusion_b200/libmusediff_b200.so
  envcmd=$(echo "$envs" | tr ',' ' ')
  env $envcmd timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<PY
import json,sys
d=json.load(open("gpurun_out/ab_%s.json"%sys.argv[1]))
k=d["kernels"]
def ms(sub): return sum(v["ms"] for n,v in k.items() if sub in n)
print("%-10s step %.2f ms | att %.2f qkv %.2f out %.2f ffn1 %.2f ffn2 %.2f ln %.2f | clk %s" % (sys.argv[1], d["ms_per_step"], ms("attention"), ms("x2304x768"), ms("x768x768 epi=0"), ms("epi=1"), ms("x768x3072"), ms("layernorm"), d["clocks"]["sm_mhz"]))
PY
done
cp /tmp/lib_current.so musediffusion_b200/libmusediff_b200.so
