# Round-2 GPU call B: pipeline parity tests, then pipeline throughput under a few configurations.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_pipeline.py -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
tail -5 gpurun_out/r02b_pytest.log
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err; }
B="python bench.py --no-cpu-baseline --steps 3 --warmup 1"
run w4_c40 $B --hf-lanes 0
run w16_c40 $B --hf-lanes 16
run w16_c64_h16 $B --hf-lanes 16 --contexts 64 --heavy-frames 16
run w8_c40 $B --hf-lanes 8
run w16_c24_h6 $B --hf-lanes 16 --contexts 24 --heavy-frames 6
run d2_w16 $B --hf-lanes 16 --workload synth8k_d2
run k4_w16 $B --hf-lanes 16 --workload synth4k
python - <<'PY'
import json
for n in ("w4_c40","w16_c40","w16_c64_h16","w8_c40","w16_c24_h6","d2_w16","k4_w16"):
    try:
        d=json.load(open("gpurun_out/r02b_%s.json"%n))
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_u8"]["value"]), "ms/step", round(d["ms_per_step"]),
              "roof", d["roofline"]["frac"] if d.get("roofline") else None, "solo", d["kernel_ms_per_frame_solo"])
    except Exception as e:
        print(n, "ERR", e); print(open("gpurun_out/r02b_%s.err"%n).read()[-600:])
PY
