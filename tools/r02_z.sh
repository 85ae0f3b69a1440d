# call Z: rolled loops (code size) against the unrolled build, DRAM bytes with the pipeline's L2 fetch granularity, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1
tail -4 gpurun_out/r02z_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
run() { name=$1; shift
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'idct|filter|classify' -c 40 --csv --log-file gpurun_out/r02z_launches_$name.csv python tools/decode_once.py $F 2 > gpurun_out/r02z_ncu_$name.log 2>&1
  python - $name <<'PY'
import csv, collections, sys
name=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/r02z_launches_%s.csv'%name)))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ii=h.index('ID')
recs=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    recs.setdefault(r[ii],{'k':r[ki][:46]})[r[mi]]=float(r[vi].replace(',',''))
ids=list(recs); ids=ids[len(ids)//2:]
print(name)
for i in ids:
    d=recs[i]
    print("  %-48s %.3f ms  read %.0f MB  write %.0f MB"%(d['k'], d.get('gpu__time_duration.sum',0)/1e6, d.get('dram__bytes_read.sum',0)/1e6, d.get('dram__bytes_write.sum',0)/1e6))
PY
}
run rolled JXLB_L2_FETCH=32
run unrolled JXLB_L2_FETCH=32 JXLB_LIB=$PWD/jxl_oxide_b200/_variants/libjxlb200_unrolled.so
show() { python -c "
import json,sys; d=json.load(open('$1'))
print('$2 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', round(d['roofline']['frac'],4), d['roofline']['per_kernel_ms'], 'busy', d['clocks'].get('gpu_busy_pct_mean'), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],1))"; }
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02z_bench_synth8k.json 2> gpurun_out/r02z_bench_synth8k.err; show gpurun_out/r02z_bench_synth8k.json synth8k; tail -2 gpurun_out/r02z_bench_synth8k.err
for w in synth4k synth8k_d2 synthmod4k; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02z_bench_$w.json 2> gpurun_out/r02z_bench_$w.err; show gpurun_out/r02z_bench_$w.json $w; tail -2 gpurun_out/r02z_bench_$w.err
done
