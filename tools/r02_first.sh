# Round-2 first GPU call: state of the existing schedules + host link + timeline under load.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc >> gpurun_out/r02a_smi.txt; numactl -H >> gpurun_out/r02a_smi.txt 2>&1
python - > gpurun_out/r02a_pcie.txt 2>&1 <<'PY'
import torch, time
x = torch.empty(400*1024*1024, dtype=torch.uint8, device='cuda')
h = torch.empty(400*1024*1024, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: h.copy_(x, non_blocking=True)), ("h2d", lambda: x.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(name, "GB/s", 0.4194304 / dt)
PY
python -c "import bench; bench.synth_frame(7680, 4320, 1)" > /dev/null
F=bench_data/synth_7680x4320_d1.0_s1.jxl
for L in 0 8 16; do
  JXLB_HF_LANES=$L timeout 120 python - $F > gpurun_out/r02a_solo_$L.txt 2>&1 <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import quick_time as q
q.latency(sys.argv[1], 3)
PY
done
run() { name=$1; shift; timeout 300 env "$@" > gpurun_out/r02a_$name.json 2> gpurun_out/r02a_$name.err; }
B="python bench.py --steps 3 --warmup 3"
run b_default $B --hf-lanes 0
run b_w16 $B --hf-lanes 16
run b_w16_c48 $B --hf-lanes 16 --contexts 48 --frames-per-step 48
run b_w16_pipe $B --hf-lanes 16 --pipeline-steps
JXLB_HF_LANES=16 timeout 200 python tools/timeline.py $F 24 > gpurun_out/r02a_timeline24.txt 2>&1
python - <<'PY'
import json
for n in ("b_default","b_w16","b_w16_c48","b_w16_pipe"):
    try:
        d=json.load(open("gpurun_out/r02a_%s.json"%n))
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_u8"]["value"]), "ms/step", round(d["ms_per_step"]),
              "solo", d["kernel_ms_per_frame_solo"])
    except Exception as e: print(n, "ERR", e)
PY
cat gpurun_out/r02a_pcie.txt gpurun_out/r02a_solo_*.txt; head -30 gpurun_out/r02a_timeline24.txt
