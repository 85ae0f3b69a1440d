mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1
tail -6 gpurun_out/r02f_pytest.log
JXLB_HF_LANES=16 timeout 120 python - bench_data/synth_7680x4320_d1.0_s1.jxl > gpurun_out/r02f_solo.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import quick_time as q
q.latency(sys.argv[1], 4)
PY
cat gpurun_out/r02f_solo.txt
(
timeout 100 python tools/pipe_probe.py synth4k value 32:10 64:12 --trace
timeout 150 python tools/pipe_probe.py synth8k value 32:10 48:12 64:14 --phases
PROBE_FRAMES=64 timeout 100 python tools/pipe_probe.py synth8k e2e 48:12 --phases
PROBE_FRAMES=64 timeout 100 python tools/pipe_probe.py synth8k u8 48:12
) > gpurun_out/r02f_probe.txt 2>&1
cat gpurun_out/r02f_probe.txt
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02f_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02f_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02f_launches_8k.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
out=[(r[ki][:60], r[vi]) for r in rows[hdr+1:] if len(r)>vi]
for k,v in out[len(out)//2:]: print(k, v)
PY
