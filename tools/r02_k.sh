mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1
tail -4 gpurun_out/r02k_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02k_ncu.log 2>&1
F3=bench_data/synth_7680x4320_d2.0_s1epfiters3.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches_8k_d2.csv python tools/decode_once.py $F3 2 > gpurun_out/r02k_ncu3.log 2>&1
F2=bench_data/synth_3840x2160_d1.0_s1modular.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02k_launches_mod4k.csv python tools/decode_once.py $F2 2 > gpurun_out/r02k_ncu2.log 2>&1
python - <<'PY'
import csv, collections
for name in ("8k","8k_d2","mod4k"):
    rows=list(csv.reader(open('gpurun_out/r02k_launches_%s.csv'%name)))
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
for W in synth8k_d2 synthmod4k synth4k; do
  timeout 600 python bench.py --steps 4 --warmup 2 --workload $W > gpurun_out/r02k_bench_$W.json 2> gpurun_out/r02k_bench_$W.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02k_bench_$W.json'))
print('$W', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', d['roofline'] and (round(d['roofline']['frac'],4), d['roofline']['per_kernel_ms']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1))"
done
