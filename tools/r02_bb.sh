# call BB: the final defaults (medium kernel unrolled, no L2 prefetch, strip kernel with rolled channel loops): parity, launch list with
# the pipeline's L2 fetch granularity (-> profiles/r02_traffic.json), the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02bb_pytest.log 2>&1
tail -4 gpurun_out/r02bb_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
for g in 32 128; do
JXLB_L2_FETCH=$g timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'idct|filter|classify' -c 40 --csv --log-file gpurun_out/r02bb_launches_l2f$g.csv python tools/decode_once.py $F 2 > gpurun_out/r02bb_ncu_$g.log 2>&1
python - $g <<'PY'
import csv, collections, sys, json
g=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/r02bb_launches_l2f%s.csv'%g)))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ii=h.index('ID')
recs=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    recs.setdefault(r[ii],{'k':r[ki][:46]})[r[mi]]=float(r[vi].replace(',',''))
ids=list(recs); ids=ids[len(ids)//2:]
tot=0; ms=0; per={}
print('L2 fetch', g)
for i in ids:
    d=recs[i]
    b=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0)
    tot+=b; ms+=d.get('gpu__time_duration.sum',0)/1e6
    per[d['k']]={'ms':round(d.get('gpu__time_duration.sum',0)/1e6,4),'dram_read_mb':round(d.get('dram__bytes_read.sum',0)/1e6,1),'dram_write_mb':round(d.get('dram__bytes_write.sum',0)/1e6,1)}
    print("  %-48s %.3f ms  read %.0f MB  write %.0f MB"%(d['k'], d.get('gpu__time_duration.sum',0)/1e6, d.get('dram__bytes_read.sum',0)/1e6, d.get('dram__bytes_write.sum',0)/1e6))
print('  chain %.3f ms, %.0f MB'%(ms,tot/1e6))
json.dump({'workload':'synth8k','l2_fetch_granularity':int(g),'chain_dram_bytes_per_frame':tot,'chain_ms_ncu':ms,'kernels':per,
           'how':'ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, second decode of synth_7680x4320_d1.0_s1.jxl, tools/r02_bb.sh'}, open('gpurun_out/r02bb_traffic_l2f%s.json'%g,'w'), indent=1)
PY
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02bb_bench_synth8k.json 2> gpurun_out/r02bb_bench_synth8k.err
python -c "
import json; d=json.load(open('gpurun_out/r02bb_bench_synth8k.json'))
print('synth8k value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', round(d['roofline']['frac'],4), d['roofline']['per_kernel_ms'], 'busy', d['clocks'].get('gpu_busy_pct_mean'), 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],1))"
tail -2 gpurun_out/r02bb_bench_synth8k.err
