# call EE (diagnostic, the round's last GPU seconds): where does the step-0 strip kernel of tools/wip/strip_step0.patch fault?
mkdir -p gpurun_out
export JXLB_LIB=$PWD/jxl_oxide_b200/_variants/libjxlb200_strip3wip.so JXLB_STRIP3=1
timeout 60 compute-sanitizer --tool memcheck --print-limit 5 python tools/decode_once.py bench_data/synth_1000x600_d1.0_s7epfiters3.jxl 1 > gpurun_out/r02ee_memcheck.log 2>&1
grep -m12 "Invalid\|Illegal\|illegal\|at 0x\|by thread\|in \|ERROR SUMMARY\|Error" gpurun_out/r02ee_memcheck.log | cut -c1-220
timeout 60 python -m pytest tests/test_zz_gpu_schedules.py -m gpu -x -q -k "epf_iteration" > gpurun_out/r02ee_pytest.log 2>&1
tail -3 gpurun_out/r02ee_pytest.log
