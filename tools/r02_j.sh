mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_pipeline.py tests/test_zz_gpu_boundary.py -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1
tail -4 gpurun_out/r02j_pytest.log
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k e2e 96:20
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k u8 96:20
PROBE_FRAMES=96 timeout 150 python tools/pipe_probe.py synthmod4k value 64:26
) > gpurun_out/r02j_probe.txt 2>&1
cat gpurun_out/r02j_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02j_bench8k.json 2> gpurun_out/r02j_bench8k.err
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench8k.json'))
print('8k value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', d['roofline']['frac'], d['roofline']['per_kernel_ms'], 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
print(d['entropy']); print(d['clocks'])"
tail -3 gpurun_out/r02j_bench8k.err
