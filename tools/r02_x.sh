# call X: two-thread 64-point passes in the 64-sample kernel, medium kernel at 5 CTAs/SM: parity, launch list, whole-job A/B of the
# L2 fetch granularity, the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1
tail -4 gpurun_out/r02x_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'idct|filter|classify' -c 40 --csv --log-file gpurun_out/r02x_launches.csv python tools/decode_once.py $F 2 > gpurun_out/r02x_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02x_launches.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ii=h.index('ID')
recs=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    recs.setdefault(r[ii],{'k':r[ki][:46]})[r[mi]]=float(r[vi].replace(',',''))
ids=list(recs); ids=ids[len(ids)//2:]
for i in ids:
    d=recs[i]
    print("  %-48s %.3f ms  read %.0f MB  write %.0f MB"%(d['k'], d.get('gpu__time_duration.sum',0)/1e6, d.get('dram__bytes_read.sum',0)/1e6, d.get('dram__bytes_write.sum',0)/1e6))
PY
(
export PROBE_FRAMES=480 PROBE_HF=128
timeout 100 python tools/pipe_probe.py synth8k value 96:26
JXLB_L2_FETCH=32 timeout 100 python tools/pipe_probe.py synth8k value 96:26
timeout 100 python tools/pipe_probe.py synth8k value 96:26
JXLB_L2_FETCH=32 timeout 100 python tools/pipe_probe.py synth8k value 96:26
) > gpurun_out/r02x_probe.txt 2>&1
cat gpurun_out/r02x_probe.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02x_bench8k.json 2> gpurun_out/r02x_bench8k.err
python -c "
import json; d=json.load(open('gpurun_out/r02x_bench8k.json'))
print('8k value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', d['roofline']['frac'], d['roofline']['per_kernel_ms'])
print(d['clocks']); print(d['cpu_baseline'])"
tail -3 gpurun_out/r02x_bench8k.err
