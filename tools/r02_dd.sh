# call DD: EPF step 0 in strip form (JXLB_STRIP3=1) and the re-laid-out one- / two-step kernels: parity, d2.0 launch lists, d2.0 bench
mkdir -p gpurun_out
export JXLB_STRIP3=1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_pipeline.py tests/test_zz_gpu_schedules.py -m gpu -x -q -k "not hf_lanes and not schedule" > gpurun_out/r02dd_pytest.log 2>&1
tail -3 gpurun_out/r02dd_pytest.log
F=bench_data/synth_7680x4320_d2.0_s1epfiters3.jxl
run() { name=$1; shift
  env "$@" timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'filter' -c 8 --csv --log-file gpurun_out/r02dd_launches_$name.csv python tools/decode_once.py $F 2 > gpurun_out/r02dd_ncu_$name.log 2>&1
  grep -o '"[^"]*filter_kernel[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","[^"]*","[^"]*"' gpurun_out/r02dd_launches_$name.csv | sed 's/(unnamed.*"gpu__time/ gpu__time/' | tail -4
}
run strip3 JXLB_STRIP3=1
run general JXLB_NO_STRIP=1
timeout 200 python bench.py --steps 3 --warmup 3 --workload synth8k_d2 --no-cpu-baseline > gpurun_out/r02dd_bench_d2.json 2> gpurun_out/r02dd_bench_d2.err
python -c "
import json; d=json.load(open('gpurun_out/r02dd_bench_d2.json'))
print('d2 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],4), d['roofline']['per_kernel_ms'])"
