mkdir -p gpurun_out
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02e_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02e_launches_8k.csv')))
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i; break
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
out=[]
for r in rows[hdr+1:]:
    if len(r)>vi: out.append((r[ki][:70], r[vi], r[ui]))
half=len(out)//2
for k,v,u in out[half:]: print(k, v, u)
PY
