# call W: occupancy variants of the small / medium transform kernels, L2 fetch granularity (time + DRAM bytes)
mkdir -p gpurun_out
F=bench_data/synth_7680x4320_d1.0_s1.jxl
run() { name=$1; shift
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'idct|filter|classify' -c 40 --csv --log-file gpurun_out/r02w_launches_$name.csv python tools/decode_once.py $F 2 > gpurun_out/r02w_ncu_$name.log 2>&1
  python - $name <<'PY'
import csv, collections, sys
name=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/r02w_launches_%s.csv'%name)))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ii=h.index('ID')
recs=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    recs.setdefault(r[ii],{'k':r[ki][:46]})[r[mi]]=float(r[vi].replace(',',''))
ids=list(recs)
ids=ids[len(ids)//2:]
print(name)
for i in ids:
    d=recs[i]
    print("  %-48s %.3f ms  read %.0f MB  write %.0f MB"%(d['k'], d.get('gpu__time_duration.sum',0)/1e6, d.get('dram__bytes_read.sum',0)/1e6, d.get('dram__bytes_write.sum',0)/1e6))
PY
}
run default A=1
run l2f32 JXLB_L2_FETCH=32
run l2f128 JXLB_L2_FETCH=128
for v in smb4 smb5 mmb5 mmb6; do run $v JXLB_LIB=$PWD/jxl_oxide_b200/_variants/libjxlb200_$v.so; done
