#!/bin/bash
# One GPU visit: parity tests, smoke, bench line, ncu launch list, ncu full captures.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout 900 python bench.py --steps ${BENCH_STEPS:-8} --warmup 3 ${BENCH_FLAGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ "${NCU_LIST:-0}" == "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-breakdown > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
fi
for k in ${NCU_FULL}; do
  kk=$(echo $k | tr -c "A-Za-z0-9_\n" "_")
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_$kk python tools/profile_step.py --batch 64 --steps 2 > gpurun_out/ncu_$kk.log 2>&1; echo "ncu full $k rc=$?"
done
