mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1
tail -3 gpurun_out/r02o_pytest.log
JXLB_HF_LANE_STRIDE=4 timeout 600 python -m pytest tests/test_zz_gpu_schedules.py -m gpu -x -q > gpurun_out/r02o_pytest2.log 2>&1
tail -3 gpurun_out/r02o_pytest2.log
for S in 1 2 4 8; do
JXLB_HF_LANES=128 JXLB_HF_LANE_STRIDE=$S timeout 120 python - bench_data/synth_7680x4320_d1.0_s1.jxl 2>&1 <<'PY' | grep -o "decode_hf': [0-9.]*" | sed "s/^/stride $S /"
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import quick_time as q
q.latency(sys.argv[1], 3)
PY
done
(
PROBE_FRAMES=192 timeout 100 python tools/pipe_probe.py synth8k value 96:20
export PROBE_HF=128 PROBE_FRAMES=192
timeout 100 python tools/pipe_probe.py synth8k value 96:26
JXLB_HF_LANE_STRIDE=2 timeout 100 python tools/pipe_probe.py synth8k value 96:26
JXLB_HF_LANE_STRIDE=4 timeout 100 python tools/pipe_probe.py synth8k value 96:26 128:26
JXLB_HF_LANE_STRIDE=8 timeout 100 python tools/pipe_probe.py synth8k value 96:26
) > gpurun_out/r02o_probe.txt 2>&1
cat gpurun_out/r02o_probe.txt
