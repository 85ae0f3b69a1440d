# call V: shape-templated medium transform kernel + shared-memory staged 64-sample kernel: parity, memcheck, per-kernel times,
# pipeline probes with more frames in flight
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1
tail -4 gpurun_out/r02v_pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/decode_once.py bench_data/synth_2000x1500_d1.0_s3.jxl 1 > gpurun_out/r02v_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r02v_memcheck.log
F=bench_data/synth_7680x4320_d1.0_s1.jxl
run() { name=$1; shift
  env "$@" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02v_launches_$name.csv python tools/decode_once.py $F 2 > gpurun_out/r02v_ncu_$name.log 2>&1
  python - $name <<'PY'
import csv, collections, sys
name=sys.argv[1]
rows=list(csv.reader(open('gpurun_out/r02v_launches_%s.csv'%name)))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
out=[(r[ki][:50], float(r[vi])) for r in rows[hdr+1:] if len(r)>vi]
out=out[len(out)//2:]
acc=collections.OrderedDict()
for k,v in out:
    a=acc.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
print(name)
for k,(n,v) in acc.items():
    if 'idct' in k or 'filter' in k or 'classify' in k: print("  %-52s x%-4d %.3f ms"%(k,n,v/1e6))
PY
}
run default A=1
run generic JXLB_MEDIUM_GENERIC=1 JXLB_LARGE_GENERIC=1
(
export PROBE_FRAMES=480 PROBE_HF=128
timeout 200 python tools/pipe_probe.py synth8k value 96:26 144:26 192:26 144:40
timeout 100 python tools/pipe_probe.py synth8k value 96:26 --phases
) > gpurun_out/r02v_probe.txt 2>&1
cat gpurun_out/r02v_probe.txt
