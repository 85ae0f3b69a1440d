# First GPU call of round 2 (one box, ~12 min): parity of the schedules written blind at the end of round 1, then an
# A/B of the thread-per-stream HF kernel against the default on the same box.
#   gpurun --timeout 1500 -- 'bash tools/r02_lanes_cmd.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_schedules.py -m gpu -q > gpurun_out/r02_zz_pytest.log 2>&1
tail -3 gpurun_out/r02_zz_pytest.log
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err; }
B="python bench.py --steps 3 --warmup 3"
run A_default $B --hf-lanes 0
run B_lanes32 JXLB_HF_LANES=32 $B
run C_lanes64 JXLB_HF_LANES=64 $B
run D_lanes128 JXLB_HF_LANES=128 $B
run D2_warps16 JXLB_HF_LANES=16 $B
run D3_warps8 JXLB_HF_LANES=8 $B
run E_lanes64_pipe JXLB_HF_LANES=64 $B --pipeline-steps
run F_lanes64_c48 JXLB_HF_LANES=64 $B --contexts 48 --frames-per-step 48
run G_default_c48 $B --hf-lanes 0 --contexts 48 --frames-per-step 48
run H_auto $B
python - <<PY
import json
for n in ("A_default","B_lanes32","C_lanes64","D_lanes128","D2_warps16","D3_warps8","E_lanes64_pipe","F_lanes64_c48","G_default_c48","H_auto"):
    try:
        d=json.load(open("gpurun_out/r02_%s.json"%n))
        k=d["kernel_ms_per_step"]; s=d["kernel_ms_per_frame_solo"]
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"]),
              "decode_hf/step", k.get("decode_hf"), "solo", s.get("decode_hf"), "hf_block_ctx solo", s.get("hf_block_ctx"))
    except Exception as e: print(n, "ERR", e)
PY
