# call CC: three back-to-back bench runs of one build (reproducibility of `value`, e2e, e2e_u8)
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02cc_bench_run$i.json 2> gpurun_out/r02cc_bench_run$i.err
  python -c "
import json; d=json.load(open('gpurun_out/r02cc_bench_run$i.json'))
print('run $i value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'roof', round(d['roofline']['frac'],4), 'busy', d['clocks'].get('gpu_busy_pct_mean'), 'traffic', d['roofline']['traffic'])"
done
