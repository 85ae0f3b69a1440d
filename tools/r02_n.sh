mkdir -p gpurun_out
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:40 128:48
export PROBE_HF=128 PROBE_FRAMES=192
JXLB_DEBUG_SKIP=1 timeout 100 python tools/pipe_probe.py synth8k value 96:40
JXLB_DEBUG_SKIP=6 timeout 100 python tools/pipe_probe.py synth8k value 96:40
JXLB_DEBUG_SKIP=7 timeout 100 python tools/pipe_probe.py synth8k value 96:40 160:40
JXLB_DEBUG_SKIP=7 timeout 100 python tools/pipe_probe.py synth8k value 96:40 --phases
) > gpurun_out/r02n_probe.txt 2>&1
cat gpurun_out/r02n_probe.txt
