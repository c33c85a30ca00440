#!/bin/bash
# One GPU call, many answers: every step has its own timeout and log under gpurun_out/, a failing step does not stop the rest.
# usage (under gpurun): bash profiles/gpu_batch.sh [steps...]   steps: att gemmtests shapes bf16model suite bench_bf16 bench_fp32 bench_x3 synth launches launches_fp32 launches_synth
mkdir -p gpurun_out
steps="$@"
[ -z "$steps" ] && steps="att gemmtests shapes bf16model suite bench_bf16"
for s in $steps; do
  echo "=== $s ==="
  case $s in
    att)        timeout 300 python -m pytest tests/test_bf16_gpu.py -q -m gpu -k attention 2>&1 | tail -25 | tee gpurun_out/att.log ;;
    gemmtests)  timeout 400 python -m pytest tests/test_bf16_gpu.py -q -m gpu -k "not attention" 2>&1 | tail -25 | tee gpurun_out/gemmtests.log ;;
    shapes)     timeout 200 python profiles/gemm_shapes.py bf16 2>&1 | tail -22 | tee gpurun_out/shapes_bf16.log ;;
    shapes_x3)  timeout 200 python profiles/gemm_shapes.py tf32x3 2>&1 | tail -22 | tee gpurun_out/shapes_tf32x3.log ;;
    bf16model)  timeout 400 python -m pytest tests/test_bf16_model_gpu.py -q -m gpu -s 2>&1 | grep -E "bf16|passed|failed|Error|error|assert" | tail -40 | tee gpurun_out/bf16model.log ;;
    suite)      timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 --deselect tests/test_bf16_gpu.py --deselect tests/test_bf16_model_gpu.py 2>&1 | tail -40 | tee gpurun_out/suite.log ;;
    bench_bf16) timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 2500 gpurun_out/bench_bf16.json; tail -3 gpurun_out/bench_bf16.err ;;
    bench_fp32) timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 2500 gpurun_out/bench_fp32.json; tail -3 gpurun_out/bench_fp32.err ;;
    bench_x3)   timeout 600 python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_x3.json 2> gpurun_out/bench_x3.err; tail -c 2500 gpurun_out/bench_x3.json; tail -3 gpurun_out/bench_x3.err ;;
    bench_full) timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 6000 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err ;;
    synth)      for w in synth_c1 synth_c4 synth_c5 mas_c2 mas_c5; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$w.json")); print("$w", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
                done ;;
    launches)   python profiles/one_step.py bf16 train_c2 > gpurun_out/one_step_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_bf16.csv python profiles/one_step.py bf16 train_c2 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log ;;
    launches_fp32) python profiles/one_step.py tf32x3 train_c2 > gpurun_out/one_step_plain_fp32.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_fp32.csv python profiles/one_step.py tf32x3 train_c2 > gpurun_out/ncu_fp32.log 2>&1; tail -2 gpurun_out/ncu_fp32.log ;;
    launches_synth) python profiles/one_step.py bf16 synth_c1 > gpurun_out/one_step_plain_synth.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_synth_bf16.csv python profiles/one_step.py bf16 synth_c1 > gpurun_out/ncu_synth.log 2>&1; tail -2 gpurun_out/ncu_synth.log ;;
    *) echo "unknown step $s" ;;
  esac
done
