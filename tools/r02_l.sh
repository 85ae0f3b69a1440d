mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_boundary.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1
tail -6 gpurun_out/r02l_pytest.log
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20
PROBE_HF=64 PROBE_FRAMES=192 timeout 150 python tools/pipe_probe.py synth8k value 96:20 96:40 128:56
PROBE_HF=128 PROBE_FRAMES=192 timeout 150 python tools/pipe_probe.py synth8k value 96:40 128:56
PROBE_HF=32 PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20
PROBE_HF=8 PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20
nvidia-smi --query-gpu=memory.used --format=csv
) > gpurun_out/r02l_probe.txt 2>&1
cat gpurun_out/r02l_probe.txt
