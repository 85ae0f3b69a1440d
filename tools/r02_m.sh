mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_boundary.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1
tail -3 gpurun_out/r02m_pytest.log
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20 --phases
PROBE_HF=128 PROBE_FRAMES=192 timeout 150 python tools/pipe_probe.py synth8k value 96:40 --phases
PROBE_HF=128 PROBE_FRAMES=192 timeout 150 python tools/pipe_probe.py synth8k value 48:40 64:40 --phases
) > gpurun_out/r02m_probe.txt 2>&1
cat gpurun_out/r02m_probe.txt
