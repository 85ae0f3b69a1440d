# call Y: the whole GPU suite on the final kernels, L2 prefetch A/B, full ncu captures of the pixel-chain kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1
tail -4 gpurun_out/r02y_pytest.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
run() { name=$1; shift
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'idct|filter|classify' -c 40 --csv --log-file gpurun_out/r02y_launches_$name.csv python tools/decode_once.py $F 2 > gpurun_out/r02y_ncu_$name.log 2>&1
  python - $name <<'PY'
import csv, collections, sys
name=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/r02y_launches_%s.csv'%name)))
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
run prefetch A=1
run noprefetch JXLB_NO_L2_PREFETCH=1
cap() { name=$1; kern=$2; skip=$3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f \
      -o gpurun_out/r02y_full_$name python tools/decode_once.py $F 2 > gpurun_out/r02y_full_$name.log 2>&1; }
cap strip strip_filter_kernel 1
cap medium idct_medium_deq_kernel 1
cap small idct_small_kernel 1
cap large64 idct_large64_deq_kernel 1
ls -la gpurun_out/r02y*.ncu-rep
