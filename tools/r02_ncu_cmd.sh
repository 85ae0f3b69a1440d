# Second GPU call of round 2 (one box, ~6 min): full ncu captures of the HF coefficient kernel under each schedule on
# the 8K bench frame, plus the launch list of one bench step (default schedule; the auto probe spawns children, which ncu would follow).
#   gpurun --timeout 900 -- 'bash tools/r02_ncu_cmd.sh'
# Read back here with:  ncu -i gpurun_out/r02_hf_<schedule>.ncu-rep --page raw --csv | ...   (B200_PROFILING.md)
mkdir -p gpurun_out
python -c "import bench; bench.synth_frame(7680, 4320, 1)" > /dev/null
F=bench_data/synth_7680x4320_d1.0_s1.jxl
cap() { name=$1; lanes=$2; kern=$3
  JXLB_HF_LANES=$lanes timeout 280 ncu --set full --clock-control none --import-source on -k "regex:$kern" -c 1 -f \
      -o gpurun_out/r02_hf_$name python tools/decode_once.py $F 1 > gpurun_out/r02_hf_$name.log 2>&1; }
cap default 0 decode_hf_fast_kernel
cap warps16 16 decode_hf_fast_kernel
cap lanes64 64 decode_hf_lanes_kernel
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --hf-lanes 0 --steps 1 --warmup 1 --contexts 4 --frames-per-step 4 > gpurun_out/r02_launches_default.log 2>&1
ls -la gpurun_out | tail -12
