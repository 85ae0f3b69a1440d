mkdir -p gpurun_out
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20 96:26
PROBE_BATCH=10 PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:22
export PROBE_HF=128 PROBE_FRAMES=192
timeout 100 python tools/pipe_probe.py synth8k value 96:26
PROBE_BATCH=10 timeout 100 python tools/pipe_probe.py synth8k value 96:22 128:22
JXLB_HF_LANE_STRIDE=4 timeout 100 python tools/pipe_probe.py synth8k value 96:26
JXLB_HF_LANE_STRIDE=4 timeout 100 python tools/pipe_probe.py synth8k value 96:26 --phases
) > gpurun_out/r02p_probe.txt 2>&1
cat gpurun_out/r02p_probe.txt
