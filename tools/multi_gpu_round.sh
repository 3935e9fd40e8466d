#!/bin/bash
# One 8-GPU visit: BASELINE.json configs[2] (generation, 128 sequences per GPU x 8) and configs[4] (scaled denoiser, 2048 sequences
# over 2 / 4 / 8 GPUs, strong scaling).  Each run is launched exactly as the driver launches bench.py.
mkdir -p gpurun_out
run() {  # tag nproc args...
  tag=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" \
    > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag rc=$?"
  python - $tag <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step %.1f"%d["ms_per_step"], "value %.4f"%d["value"], "e2e", (d.get("e2e") or {}).get("value"), "step TFLOP/s/GPU %.0f"%d["step_tflops"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "no line:", e); print(open("gpurun_out/bench_%s.err"%sys.argv[1]).read()[-1500:])
PY
}
run gen_8gpu 8 --mode generation --batch 128 --steps 10 --warmup 3 --no-cpu-baseline --no-breakdown
run scaled_2gpu 2 --scaled --batch 1024 --scaling strong --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown
run scaled_4gpu 4 --scaled --batch 512 --scaling strong --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown
run scaled_8gpu 8 --scaled --batch 256 --scaling strong --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown
