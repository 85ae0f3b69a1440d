mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 3 --warmup 2 "$@" > gpurun_out/sw_$name.json 2>/dev/null; }
run A_barrier
run B_pipe --pipeline-steps
run C_pipe_stag60 --pipeline-steps --stagger-ms 60
run D_pipe_c32 --pipeline-steps --contexts 32 --frames-per-step 64
run E_pipe_f96 --pipeline-steps --frames-per-step 96
python - <<PY
import json
for n in ("A_barrier","B_pipe","C_pipe_stag60","D_pipe_c32","E_pipe_f96"):
    try:
        d=json.load(open("gpurun_out/sw_%s.json"%n)); print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"]))
    except Exception as e: print(n, "ERR", e)
PY
