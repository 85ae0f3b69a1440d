mkdir -p gpurun_out
(
timeout 100 python tools/pipe_probe.py synth4k value 8:8 32:10 48:12 --trace
timeout 100 python tools/pipe_probe.py synth8k value 32:10 --trace
PROBE_FRAMES=48 timeout 100 python tools/pipe_probe.py synth8k hostin 32:10 --phases
PROBE_FRAMES=48 timeout 100 python tools/pipe_probe.py synth8k slotout 32:10 --phases
PROBE_FRAMES=48 timeout 100 python tools/pipe_probe.py synth8k slotout 4:4
JXLB_NO_TMA=1 timeout 60 python tools/quick_time.py 2>&1 | head -3
) > gpurun_out/r02d_probe.txt 2>&1
cat gpurun_out/r02d_probe.txt
for L in 16; do
  JXLB_HF_LANES=$L timeout 120 python - bench_data/synth_7680x4320_d1.0_s1.jxl > gpurun_out/r02d_solo_$L.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import quick_time as q
q.latency(sys.argv[1], 4)
PY
cat gpurun_out/r02d_solo_$L.txt
done
timeout 600 python -m pytest tests/test_zz_gpu_pipeline.py tests/test_zz_gpu_schedules.py -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1
tail -5 gpurun_out/r02d_pytest.log
