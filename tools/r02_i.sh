mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1
tail -6 gpurun_out/r02i_pytest.log
(
PROBE_FRAMES=128 timeout 150 python tools/pipe_probe.py synth8k value 32:12 64:16 96:20 --phases
PROBE_FRAMES=128 timeout 100 python tools/pipe_probe.py synth4k value 64:16 128:20
PROBE_FRAMES=96 timeout 100 python tools/pipe_probe.py synth8k e2e 64:16
PROBE_FRAMES=96 timeout 100 python tools/pipe_probe.py synth8k u8 64:16
PROBE_FRAMES=64 timeout 150 python tools/pipe_probe.py synthmod4k value 64:20 --phases
) > gpurun_out/r02i_probe.txt 2>&1
cat gpurun_out/r02i_probe.txt
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02i_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02i_launches_8k.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
out=[(r[ki][:50], float(r[vi])) for r in rows[hdr+1:] if len(r)>vi]
out=out[len(out)//2:]
for k,v in out: print("  %-52s %.3f ms"%(k,v/1e6))
PY
