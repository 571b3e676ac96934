mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_pipeline.py tests/test_gpu_head_metrics.py -q -m gpu -x ) > gpurun_out/u16_tests.log 2>&1
tail -n 8 gpurun_out/u16_tests.log
timeout 300 python tools/kbench_fused.py 2>&1 | head -16 > gpurun_out/u16_kbench.txt
cat gpurun_out/u16_kbench.txt
( time timeout 600 python bench.py ) > gpurun_out/u16_bench.json 2> gpurun_out/u16_bench.err
tail -5 gpurun_out/u16_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/u16_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['pass_latency_us'])
for k,v in d['configs'].items(): print(k, v.get('us_per_pass'), v.get('patches_per_s'))
for k,v in d['roofline_kernels'].items():
    if 'argmax' in k or 'paint' in k: print(k, v['us'], v['frac'])
P
