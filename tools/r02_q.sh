mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1
tail -4 gpurun_out/r02q_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02q_ncu.log 2>&1
F3=bench_data/synth_7680x4320_d2.0_s1epfiters3.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches_8k_d2.csv python tools/decode_once.py $F3 2 > gpurun_out/r02q_ncu3.log 2>&1
python - <<'PY'
import csv, collections
for name in ("8k","8k_d2"):
    rows=list(csv.reader(open('gpurun_out/r02q_launches_%s.csv'%name)))
    hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
    out=[(r[ki][:50], float(r[vi])) for r in rows[hdr+1:] if len(r)>vi]
    out=out[len(out)//2:]
    acc=collections.OrderedDict()
    for k,v in out:
        a=acc.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    print(name)
    for k,(n,v) in acc.items(): print("  %-52s x%-4d %.3f ms"%(k,n,v/1e6))
PY
(
export PROBE_FRAMES=480
timeout 100 python tools/pipe_probe.py synth8k value 96:26
PROBE_HF=128 timeout 100 python tools/pipe_probe.py synth8k value 96:26
PROBE_HF=128 JXLB_HF_LANE_STRIDE=8 timeout 100 python tools/pipe_probe.py synth8k value 96:26
) > gpurun_out/r02q_probe.txt 2>&1
cat gpurun_out/r02q_probe.txt
