mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py tests/test_zz_gpu_schedules.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1
tail -4 gpurun_out/r02r_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02r_ncu.log 2>&1
python - <<'PY'
import csv, collections
for name in ("8k",):
    rows=list(csv.reader(open('gpurun_out/r02r_launches_%s.csv'%name)))
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
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02r_bench8k.json 2> gpurun_out/r02r_bench8k.err
python -c "
import json; d=json.load(open('gpurun_out/r02r_bench8k.json'))
print('8k value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', d['roofline']['frac'], d['roofline']['per_kernel_ms'])
print(d['entropy']); print(d['clocks'])"
tail -3 gpurun_out/r02r_bench8k.err
