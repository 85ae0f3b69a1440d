# One `ncu --set full` capture per hot kernel on the 8K bench frame (one frame alone on the GPU, second decode of the process).
mkdir -p gpurun_out
F=bench_data/synth_7680x4320_d1.0_s1.jxl
cap() { name=$1; kern=$2; skip=$3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f \
      -o gpurun_out/r02_full_$name python tools/decode_once.py $F 2 > gpurun_out/r02_full_$name.log 2>&1; }
cap idct_medium idct_medium_kernel 1
cap filter fused_filter_kernel 1
cap idct_small idct_small_kernel 1
cap decode_hf decode_hf_warp_kernel 1
cap modular_lf modular_stream_kernel 2
ls -la gpurun_out/*.ncu-rep
for n in idct_medium filter idct_small decode_hf modular_lf; do
  ncu -i gpurun_out/r02_full_$n.ncu-rep --page raw --csv > gpurun_out/r02_full_$n.raw.csv 2>/dev/null
done
