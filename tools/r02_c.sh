mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1
tail -5 gpurun_out/r02c_pytest.log
( 
timeout 120 python tools/pipe_probe.py synth4k value 8:8 16:8 32:10 48:12 --phases
timeout 120 python tools/pipe_probe.py synth4k value 32:10 --no-affinity
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 120 python tools/pipe_probe.py synth4k value 32:10
timeout 200 python tools/pipe_probe.py synth8k value 16:8 32:10 48:12 --phases
PROBE_HF=4 timeout 120 python tools/pipe_probe.py synth8k value 32:10
PROBE_HF=32 timeout 120 python tools/pipe_probe.py synth8k value 32:10
timeout 200 python tools/pipe_probe.py synth8k e2e 32:10 --phases
timeout 200 python tools/pipe_probe.py synth8k e2e 32:10 --no-affinity
timeout 200 python tools/pipe_probe.py synth8k u8 32:10
) > gpurun_out/r02c_probe.txt 2>&1
cat gpurun_out/r02c_probe.txt
python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline'])"
