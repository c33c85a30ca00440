mkdir -p gpurun_out
VARIANT=all timeout 300 python profiles/debug_graph_vs_eager.py 2>&1 | grep -v Warning | grep -E "VARIANT|step|graph |eager " | head -14
timeout 400 python -m pytest tests/test_bf16_model_gpu.py tests/test_bf16_gpu.py -q -m gpu 2>&1 | tail -8
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py -q -m gpu -s -k "c2_training or adamw or gst or graph" 2>&1 | grep -E "C2 shape|passed|failed|Error|assert" | head
python profiles/gemm_one.py panel > gpurun_out/gemm_one_panel.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_panel -s 2 -c 2 -o gpurun_out/r2_panel -f python profiles/gemm_one.py panel > gpurun_out/ncu_panel.log 2>&1
tail -2 gpurun_out/ncu_panel.log
python profiles/gemm_one.py attn > gpurun_out/gemm_one_attn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 3 -o gpurun_out/r2_attn -f python profiles/gemm_one.py attn > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_bf16_b.json 2> gpurun_out/bench_bf16_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_bf16_b.json')); print('bf16 train ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['kernel'], d['roofline']['frac']); print(d['roofline']['shares'])"
