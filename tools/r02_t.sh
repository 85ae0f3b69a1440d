# call T: state of HEAD after the re-entry — parity tests, launch list, bench line, full ncu captures of the pixel chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1
tail -4 gpurun_out/r02t_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_launches_8k.csv python tools/decode_once.py $F 2 > gpurun_out/r02t_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02t_launches_8k.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
out=[(r[ki][:50], float(r[vi])) for r in rows[hdr+1:] if len(r)>vi]
out=out[len(out)//2:]
acc=collections.OrderedDict()
for k,v in out:
    a=acc.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,v) in acc.items(): print("  %-52s x%-4d %.3f ms"%(k,n,v/1e6))
PY
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02t_bench8k.json 2> gpurun_out/r02t_bench8k.err
python -c "
import json; d=json.load(open('gpurun_out/r02t_bench8k.json'))
print('8k value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', d['roofline']['frac'], d['roofline']['per_kernel_ms'])
print(d['entropy']); print(d['clocks']); print(d['cpu_baseline'])"
tail -3 gpurun_out/r02t_bench8k.err
cap() { name=$1; kern=$2; skip=$3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f \
      -o gpurun_out/r02t_full_$name python tools/decode_once.py $F 2 > gpurun_out/r02t_full_$name.log 2>&1; }
cap filter fused_filter_kernel 1
cap idct_medium idct_medium_kernel 1
cap idct_small idct_small_kernel 1
cap idct_large idct_large_kernel 2
ls -la gpurun_out/*.ncu-rep
