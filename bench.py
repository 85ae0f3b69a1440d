#!/usr/bin/env python
"""bench.py — megapixels/s decoded (JPEG XL VarDCT d1.0) on N B200s; CPU baseline beside it.

A "step" is one pass of the decode hot path over one batch of independent frames per GPU
(weak scaling: every rank decodes its own batch; there is no data-path collective — groups and
frames are independent). `value` is timed with the encoded frames already
resident in HBM and the decoded planes left in HBM; `e2e` goes through the public API with host
bytes in and planar f32 pixels copied back to pinned host memory every step.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference      # CPU arm: the oracle (port of jxl-oxide's generic path)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# Several decoder contexts (CUDA streams) run concurrently; with the default 8 hardware work queues
# streams alias and a long entropy kernel delays other streams' launches.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------------------------
# workloads
def synth_frame(w, h, seed, distance=1.0, extra=()):
    """Synthetic encoded frame from tools/synth_enc.cc (built on demand; cached under bench_data/)."""
    tool = os.path.join(ROOT, "tools", "_build_synth_enc")
    src = os.path.join(ROOT, "tools", "synth_enc.cc")
    host = os.path.join(ROOT, "jxl_oxide_b200", "csrc", "host")
    if not os.path.exists(tool) or os.path.getmtime(tool) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", tool, src] +
                              [os.path.join(host, f) for f in ("entropy.cc", "frame_syntax.cc", "modular_syntax.cc", "headers.cc")])
    os.makedirs(os.path.join(ROOT, "bench_data"), exist_ok=True)
    tag = "".join(extra).replace("-", "")
    path = os.path.join(ROOT, "bench_data", f"synth_{w}x{h}_d{distance}_s{seed}{tag}.jxl")
    if not os.path.exists(path):
        subprocess.check_call([tool, "--width", str(w), "--height", str(h), "--seed", str(seed), "--distance", str(distance),
                               "-o", path] + list(extra), stderr=subprocess.DEVNULL)
    with open(path, "rb") as f:
        return f.read()


def load_workload(name, nframes=16):
    """Returns (description, list of encoded frames (bytes) for ONE step on ONE GPU, (w, h) per frame)."""
    if name == "mosaic8k":
        with open(os.path.join(GOLDEN, "benchmark-data", "starrail.d1-e6.jxl"), "rb") as f:
            tile = f.read()
        desc = ("8K-equivalent (33.18 MP/step/GPU): 3x3 mosaic of a real libjxl VarDCT d1.0 2560x1440 frame "
                "(starrail.d1-e6.jxl, Gaborish + EPF), decoded as 9 independent frames")
        return desc, [tile] * 9, (2560, 1440)
    if name in ("synth8k", "synth4k"):
        w, h = (7680, 4320) if name == "synth8k" else (3840, 2160)
        frames = []
        for seed in (1, 2, 3, 4):
            frames.append(synth_frame(w, h, seed))
        frames = [frames[i % len(frames)] for i in range(max(1, nframes))]
        desc = (f"{w}x{h} VarDCT d1.0 synthetic encoded frames (tools/synth_enc.cc seeds 1-4, ~0.93 bit/px, "
                "libjxl-like: WP-coded LF, mixed varblocks 8x8..64x64, Gaborish + EPF 2 iters), "
                f"{len(frames)} independent frames per step")
        return desc, frames, (w, h)
    if name.startswith("file:"):
        with open(name[5:], "rb") as f:
            data = f.read()
        import oracle_lib
        img = oracle_lib.OracleImage(data, output_colour=2, threads=os.cpu_count())
        return f"file {name[5:]}", [data], (img.width, img.height)
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region, from ONE long-running `nvidia-smi -lms 200` process
    (the profiling recipe's clocks line) - starting a new nvidia-smi per sample re-initialises NVML each time and
    perturbs the run it is supposed to observe."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, uuid=None):
        # `uuid` ("GPU-...") identifies the CUDA device whatever CUDA_VISIBLE_DEVICES remaps; else the index is used
        self.gpu = uuid or gpu_index
        self.index = gpu_index
        self.proc = None
        self.lines = []
        self.reader = None

    def _start_nvml(self):
        """In-process NVML (what nvidia-smi itself reads): one init, then cheap polls every 200 ms."""
        import pynvml
        pynvml.nvmlInit()
        try:
            h = (pynvml.nvmlDeviceGetHandleByUUID(self.gpu) if isinstance(self.gpu, str)
                 else pynvml.nvmlDeviceGetHandleByIndex(self.gpu))
        except Exception:
            if not isinstance(self.gpu, str):
                raise
            self.gpu = self.index  # the UUID did not resolve: fall back to the CUDA ordinal
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # raises when unsupported
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        self.nvml_samples, self.nvml_reasons, self.nvml_max = [], set(), None
        self._halt = threading.Event()

        def poll():
            while not self._halt.is_set():
                try:
                    self.nvml_samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    for n, b in bits.items():
                        if mask & b:
                            self.nvml_reasons.add(n)
                except Exception:
                    pass
                self._halt.wait(0.2)
        self.nvml_thread = threading.Thread(target=poll, daemon=True)
        self.nvml_thread.start()

    def start(self):
        self.nvml_thread = None
        try:
            self._start_nvml()
            return
        except Exception:
            self.nvml_thread = None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["stdbuf", "-oL", "nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            try:
                for line in self.proc.stdout:
                    self.lines.append(line)
            except Exception:
                pass
        self.reader = threading.Thread(target=pump, daemon=True)
        self.reader.start()

    def stop(self):
        if self.nvml_thread is not None:
            self._halt.set()
            self.nvml_thread.join(timeout=5)
            if self.nvml_samples:
                return {"sm_mhz": float(np.median(self.nvml_samples)), "sm_max_mhz": self.nvml_max,
                        "reasons": sorted(self.nvml_reasons), "samples": len(self.nvml_samples),
                        "how": "NVML polled every 200 ms during the timed region"}
        if self.proc is not None:
            try:
                self.proc.terminate()  # the exact process started above
                self.proc.wait(timeout=5)
            except Exception:
                pass
            if self.reader is not None:
                self.reader.join(timeout=5)
        mode = "nvidia-smi -lms 200 during the timed region"
        if not self.lines:  # nothing arrived through the pipe: one query right after the region instead
            mode = "single nvidia-smi query right after the timed region"
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.lines = [subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                             capture_output=True, text=True, timeout=10).stdout]
            except Exception:
                self.lines = []
        samples, reasons, max_mhz = [], set(), None
        for line in self.lines:
            out = [x.strip() for x in line.strip().split(",")]
            try:
                samples.append(float(out[0]))
                max_mhz = float(out[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(self.NAMES, out[2:]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        med = float(np.median(samples)) if samples else None
        return {"sm_mhz": med, "sm_max_mhz": max_mhz, "reasons": sorted(reasons), "samples": len(samples), "how": mode}


def algorithmic_bytes(kernel, w, h, stream_bytes):
    """Algorithmic HBM bytes of ONE frame for a kernel family (DESIGN.md 'Kernels')."""
    px = w * h
    lf = ((w + 7) // 8) * ((h + 7) // 8)
    table = {
        "modular_decode": stream_bytes * 0.12 + lf * 3 * 4 + lf * 4 + (px / 4096) * 8,  # LF + HfMetadata streams
        "decode_hf": stream_bytes * 0.88 + px * 12,       # HF sections read, 3 x i32 coefficients written
        "build_block_info": lf * 4 * 4,
        "hf_block_ctx": lf * 6 * 4,                        # type, multiplier, 3 quantised LF read, 1 word written
        "hf_dequant_cfl": px * 24,
        "hf_transform": px * 24 + lf * 12,
        "filters_fused": px * 24,                          # Gaborish + EPF + colour in one pass
        "gaborish": px * 24,                               # 3 launches x 8 B/px
        "epf_step": px * 24,
        "xyb_to_rgb": px * 24,
        "lf_dequant": lf * 24, "lf_cfl": lf * 24, "lf_smooth": lf * 24,
    }
    return table.get(kernel)


KERNELS = ["modular_decode", "build_block_info", "hf_block_ctx", "decode_hf", "lf_dequant", "lf_cfl", "lf_smooth", "hf_dequant_cfl",
           "hf_transform", "filters_fused", "gaborish", "epf_step", "xyb_to_rgb", "copy_rect", "squeeze_inverse", "rct_inverse",
           "int_to_float", "modular_xyb", "palette_inverse_simple"]


# streams per CTA tried against the default (4, one warp each): 16 one-warp streams; 64 / 128 one-thread streams
HF_CANDIDATES = (16, 64, 128)


def run_probe(args):
    """Child process of choose_hf_schedule(): one JSON line on stdout. Runs in its own process so that a fault in the
    candidate kernel cannot poison the CUDA context of the measuring process."""
    import torch
    import jxl_oxide_b200 as J
    dev = args.probe_device
    torch.cuda.set_device(dev)
    nctx = max(1, args.contexts)
    _, frames, (w, h) = load_workload(args.workload, nctx)
    # bit pattern of one decoded frame under this schedule (the parent compares it with the default kernel's)
    import hashlib
    d0 = J.Decoder(dev)
    d0.set_hf_streams_per_cta(args.probe_lanes)
    d0.decode(frames[0])
    digest = hashlib.sha256(d0.frame_planar(0).tobytes()).hexdigest()
    d0.close()
    decs = [J.Decoder(dev) for _ in range(nctx)]
    for i, d in enumerate(decs):
        d.set_hf_streams_per_cta(args.probe_lanes)
        d.preload(0, frames[i % len(frames)])

    def one_step():
        errs = []

        def work(d):
            try:
                d.decode_slot(0)
                d.sync()
                d.release_frames()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(d,)) for d in decs]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
    for _ in range(2):
        one_step()
    torch.cuda.synchronize()
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"probe": "speed", "lanes": args.probe_lanes, "value": w * h * nctx / dt / 1e6, "ms_per_step": dt * 1e3,
                      "sha256": digest}))


def choose_hf_schedule(args, device):
    """HF coefficient schedule for the timed run: (streams per CTA, report). An explicit --hf-lanes / JXLB_HF_LANES wins.
    Otherwise one child process per schedule (the default first, then every candidate) decodes a frame of the workload,
    reports the SHA-256 of its f32 planes and times a short lock-step run (all contexts, one frame each); a candidate is
    kept only if its planes are bit-identical to the default kernel's and it is at least 3 % faster. A candidate whose
    child fails or times out is dropped; if the default's child fails the default kernel is used without a probe."""
    env_knob = os.environ.get("JXLB_HF_LANES")
    if env_knob is not None or args.hf_lanes != "auto":
        n = int(env_knob if env_knob is not None else args.hf_lanes)
        return (0 if n <= 0 else (8 if n <= 8 else (16 if n <= 16 else (32 if n <= 32 else (64 if n <= 64 else 128))))), {"mode": "explicit"}
    report = {"mode": "auto", "rule": "bit-identical to the default kernel and >= 3 % faster in a lock-step probe"}

    def child(extra, timeout):
        env = dict(os.environ)
        env.pop("JXLB_HF_LANES", None)
        fake = os.environ.get("JXLB_BENCH_FAKE_PROBE")  # host-logic test hook: canned child outputs, no GPU
        if fake:
            key = extra[3]
            return json.loads(fake)[key]
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--contexts",
                            str(args.contexts), "--probe-device", str(device)] + extra,
                           capture_output=True, text=True, timeout=timeout, env=env)
        if p.returncode != 0:
            raise RuntimeError(f"probe {extra} exited with {p.returncode}: {p.stderr[-300:]}")
        return json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    try:
        ref = child(["--probe", "speed", "--probe-lanes", "0"], 300)
        speeds, same = {"0": ref["value"]}, {}
        for n in HF_CANDIDATES:
            try:
                r = child(["--probe", "speed", "--probe-lanes", str(n)], 300)
            except Exception as e:  # noqa: BLE001  (this candidate is out; the others and the default are unaffected)
                same[str(n)] = f"failed: {type(e).__name__}: {e}"[:200]
                continue
            same[str(n)] = r["sha256"] == ref["sha256"]
            if same[str(n)] is True:
                speeds[str(n)] = r["value"]
        report["identical_to_default_kernel"] = same
        report["probe_mp_s"] = speeds
        cands = [n for n in HF_CANDIDATES if str(n) in speeds]
        best = max(cands, key=lambda n: speeds[str(n)]) if cands else 0
        chosen = best if best and speeds[str(best)] >= 1.03 * speeds["0"] else 0
    except Exception as e:  # noqa: BLE001
        report["fallback"] = f"{type(e).__name__}: {e}"[:300]
        chosen = 0
    report["chosen"] = chosen
    return chosen, report


def run_ours(args, rank, world, local_rank):
    import torch
    import jxl_oxide_b200 as J
    from jxl_oxide_b200 import build as jb
    if not os.path.exists(J.LIB_PATH):
        jb.build()
    torch.cuda.set_device(local_rank)
    # rank 0 picks the HF schedule before it touches its GPU; the other ranks adopt it after the process group is up
    hf_lanes, hf_report = choose_hf_schedule(args, local_rank) if rank == 0 else (0, None)
    desc, frames, (w, h) = load_workload(args.workload, args.frames_per_step)
    px_per_frame = w * h
    nthreads = max(1, min(args.contexts, len(frames)))
    decs = [J.Decoder(local_rank) for _ in range(nthreads)]
    shares = [list(range(i, len(frames), nthreads)) for i in range(nthreads)]
    # encoded frames resident in HBM ("inputs already resident"): one preloaded slot per frame
    for d, idxs in zip(decs, shares):
        for k in idxs:
            d.preload(k, frames[k])
    pinned = [torch.empty((3, h, w), dtype=torch.float32).pin_memory() for _ in range(nthreads)]
    pinned_np = [p.numpy() for p in pinned]
    decs[0].decode_slot(shares[0][0])  # shape of the packed 8-bit image (orientation and alpha are the stream's)
    u8_shape = decs[0].frame_to_buffer(0, np.uint8).shape
    decs[0].release_frames()
    pinned_u8 = [torch.empty(u8_shape, dtype=torch.uint8).pin_memory() for _ in range(nthreads)]
    pinned_u8_np = [p.numpy() for p in pinned_u8]

    def step(e2e, nsteps=1):
        """`nsteps` passes over the batch. Every context walks its share of the batch `nsteps` times; contexts are
        joined only at the end, so consecutive steps pipeline (no barrier between steps, one on each side of the
        timed region) when --pipeline-steps asks for it; by default one step per call."""
        errs = []

        def work(d, idxs, out, out_u8, delay):
            try:
                if delay > 0.0:
                    time.sleep(delay)  # de-phase the contexts (inside the timed region)
                for _ in range(nsteps):
                    for k in idxs:
                        if e2e == "u8":
                            d.decode(frames[k])                          # host bytes in
                            d.frame_to_buffer(0, np.uint8, out=out_u8)   # interleaved 8-bit RGB out, packed on the device
                        elif e2e:
                            d.decode(frames[k])          # host bytes in
                            d.frame_to_host(0, out)      # planar f32 out (pinned host)
                        else:
                            d.decode_slot(k)
                            d.sync()
                        d.release_frames()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(d, idxs, o, o8, (i % args.stagger_groups) * args.stagger_ms / 1e3))
              for i, (d, idxs, o, o8) in enumerate(zip(decs, shares, pinned_np, pinned_u8_np))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        knob = torch.tensor([hf_lanes], dtype=torch.int32, device="cuda")
        dist.broadcast(knob, src=0)
        hf_lanes = int(knob.item())
    for d in decs:
        d.set_hf_streams_per_cta(hf_lanes)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if args.pipeline_steps:
            step(e2e, steps)
        else:
            for _ in range(steps):
                step(e2e)
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(False)
    launches0 = sum(d.launch_count() for d in decs)
    try:
        dev_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    sampler.start()
    ms = timed(False, args.steps)
    clocks = sampler.stop()
    launches = sum(d.launch_count() for d in decs) - launches0
    step(True)
    ms_e2e = timed(True, args.steps)
    # the same end-to-end call with the output an 8-bit image (ImageStream::write_to_buffer::<u8>): 3 B/px cross the
    # host link instead of the 12 B/px of f32 planes that bound `e2e`
    step("u8")
    ms_e2e_u8 = timed("u8", args.steps)

    # per-kernel device time (CUDA events on the launching stream), one extra profiled step
    for d in decs:
        d.set_profile(True)
        d.profile_reset()
    step(False)
    prof = {}
    for k in KERNELS:
        n = sum(d.profile(k)[0] for d in decs)
        t = sum(d.profile(k)[1] for d in decs)
        if n:
            prof[k] = {"launches": n, "ms": t}
    # the same kernels with one frame alone on the GPU (no queueing behind other streams' kernels)
    decs[0].profile_reset()
    solo_reps = 2
    for _ in range(solo_reps):
        decs[0].decode_slot(shares[0][0])
        decs[0].sync()
        decs[0].release_frames()
    solo = {k: decs[0].profile(k)[1] / solo_reps for k in KERNELS if decs[0].profile(k)[0]}
    for d in decs:
        d.set_profile(False)

    total_px = px_per_frame * len(frames) * world
    gather = None
    if args.gather != "none":
        # BASELINE config #5's delivery: every frame packed on its GPU (interleaved u8 / u16, 3-6 B/px instead of 12 B/px
        # of f32 planes) and gathered to rank 0 over NCCL (jxl_oxide_b200.sharding.gather_frames); timed like `value`.
        from jxl_oxide_b200 import sharding
        gdt = np.uint8 if args.gather == "u8" else np.uint16
        packed = [None] * len(frames)

        def gather_step():
            errs = []

            def work(d, idxs):
                try:
                    for k in idxs:
                        d.decode_slot(k)
                        packed[k] = d.frame_to_torch(0, gdt, out=packed[k])
                        d.release_frames()
                except Exception as e:  # noqa: BLE001
                    errs.append(e)
            ts = [threading.Thread(target=work, args=(d, idxs)) for d, idxs in zip(decs, shares)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errs:
                raise errs[0]
            return sharding.gather_frames(packed, len(frames) * world, dst=0)

        gather_step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            got = gather_step()
        torch.cuda.synchronize()
        g1.record()
        g1.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1)], device="cuda")
        if dist is not None:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_g = float(tg.item())
        nbytes = int(packed[0].numel() * packed[0].element_size())
        gather = {"value": total_px / (ms_g / args.steps / 1e3) / 1e6, "unit": "MP/s", "ms_per_step": ms_g / args.steps,
                  "format": f"{args.gather} interleaved RGB, packed on the device",
                  "bytes_to_rank0_per_step": nbytes * len(frames) * (world - 1),
                  "collective": "torch.distributed gather (nccl), one per round of world_size frames" if world > 1 else "none (1 rank)",
                  "frames_at_rank0": (len([g for g in got if g is not None]) if got is not None else 0) if rank == 0 else None}
    value = total_px / (ms / args.steps / 1e3) / 1e6
    e2e_value = total_px / (ms_e2e / args.steps / 1e3) / 1e6
    e2e_u8_value = total_px / (ms_e2e_u8 / args.steps / 1e3) / 1e6
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    dom = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    roofline = None
    if dom:
        stream_bytes = float(np.mean([len(f) for f in frames]))
        per_launch_frames = len(frames) / max(1, prof[dom]["launches"])
        ab = algorithmic_bytes(dom, w, h, stream_bytes)
        avg_ms = prof[dom]["ms"] / prof[dom]["launches"]
        achieved = (ab * per_launch_frames) / (avg_ms / 1e3) / 1e9 if ab else None
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("workload") == args.workload:
                traffic = tj.get(dom)
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s",
                    "avg_launch_ms": avg_ms,
                    "measured": "CUDA events around each launch during one step with all contexts running",
                    "note": "entropy decode is latency-bound (serial ANS/context chain per stream), see DESIGN.md"}
    # the HBM-bound pixel pipeline, reported beside the dominant kernel
    pipe = ["hf_dequant_cfl", "hf_transform", "filters_fused", "gaborish", "epf_step", "xyb_to_rgb"]
    pipe_ms = sum(prof[k]["ms"] for k in pipe if k in prof)
    pipe_bytes = sum((algorithmic_bytes(k, w, h, 0) or 0) * (prof[k]["launches"] / (3 if k == "gaborish" else 1))
                     for k in pipe if k in prof)
    pipeline = None
    if pipe_ms > 0:
        ach = pipe_bytes / (pipe_ms / 1e3) / 1e9
        pipeline = {"kernels": [k for k in pipe if k in prof], "ms_per_step": pipe_ms, "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "bytes": "sum of each kernel's own algorithmic bytes (24 B/px each)"}
        solo_ms = sum(solo.get(k, 0.0) for k in pipe)
        if solo_ms > 0:
            fused_bytes = px_per_frame * 24.3  # coefficients in -> RGB out, fully fused chain
            pipeline["solo"] = {"ms_per_frame": solo_ms, "achieved_vs_fused_chain_bytes": fused_bytes / (solo_ms / 1e3) / 1e9,
                                "frac_of_peak": fused_bytes / (solo_ms / 1e3) / 1e9 / peak,
                                "note": "one frame alone on the GPU; 24.3 B/px algorithmic bytes of the fully fused chain"}
    cpu = cpu_baseline(args, frames, px_per_frame)
    line = {
        "metric": "Megapixels/s decoded (8K VarDCT d1.0)", "value": value, "unit": "MP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic" if "synth" in args.workload else "real-file mosaic",
        "config": {"workload": desc, "frames_per_step_per_gpu": len(frames), "decoder_contexts_per_gpu": nthreads,
                   "cache": "inputs+planes per step (>= 33 MP x 24 B) exceed L2 (126 MB); no explicit L2 flush",
                   "step_barrier": "before and after the K timed steps (steps pipeline across decoder contexts)"
                                   if args.pipeline_steps else "after every step",
                   "stagger_ms": args.stagger_ms,
                   "hf_streams_per_cta": hf_lanes, "hf_schedule": hf_report},
        "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": int(sum(len(f) for f in frames)),
                "d2h_bytes_per_step": int(px_per_frame * 12 * len(frames)), "ms_per_step": ms_e2e / args.steps},
        "e2e_u8": {"value": e2e_u8_value, "unit": "MP/s", "h2d_bytes_per_step": int(sum(len(f) for f in frames)),
                   "d2h_bytes_per_step": int(np.prod(u8_shape)) * len(frames), "ms_per_step": ms_e2e_u8 / args.steps,
                   "note": "same call path as e2e, output = interleaved 8-bit RGB packed on the device "
                           "(jxlb_frame_write_to_buffer); reported beside e2e, not instead of it"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "pipeline_roofline": pipeline,
        "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in prof.items()},
        "kernel_ms_per_frame_solo": {k: round(v, 3) for k, v in solo.items()}, "cpu_baseline": cpu,
    }
    if gather is not None:
        line["gather"] = gather
    print(json.dumps(line))


def cpu_throughput(frames, px_per_frame, steps, warm=1):
    """The CPU restatement at its best on this host: frame-level parallelism (what the reference's CLI does across
    keyframes, jxl-oxide-cli/src/decode.rs:293-301) on top of the per-frame thread pool. The restatement's intra-frame
    scaling flattens after a few threads (measured: 1.5x at 8 threads), so the cores are split into P concurrent
    frames x T threads each. One step = P frames of the workload, each decoded once (bytes -> planar f32)."""
    import oracle_lib
    oracle_lib.build()
    cores = os.cpu_count() or 1
    t_per_frame = 4 if cores >= 4 else cores
    par = max(1, cores // t_per_frame)
    try:  # keep P concurrent decodes (~64 B of planes per pixel each) within half of the free host memory
        import psutil
        par = max(1, min(par, int(psutil.virtual_memory().available * 0.5 // (px_per_frame * 64))))
    except Exception:
        par = min(par, 16)
    sample = [frames[i % len(frames)] for i in range(par)]

    def one_step():
        errs = []

        def work(f):
            try:
                oracle_lib.OracleImage(f, threads=t_per_frame).close()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=work, args=(f,)) for f in sample]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
    for _ in range(warm):
        one_step()
    t0 = time.time()
    for _ in range(steps):
        one_step()
    dt = (time.time() - t0) / steps
    value = px_per_frame * par / dt / 1e6
    return value, dt, {"value": value, "unit": "MP/s", "cores": par * t_per_frame, "kind": "port",
                       "sample": f"{par} frame(s) of the step's workload decoded concurrently, {t_per_frame} threads each "
                                 f"({cores} host cores), decode chain only (bytes -> planar f32); CPU restatement of "
                                 "jxl-oxide's generic path (not jxl-oxide itself: no Rust toolchain)"}


def cpu_baseline(args, frames, px_per_frame, steps=1):
    return cpu_throughput(frames, px_per_frame, steps)[2]


def run_reference(args, rank, world):
    if rank != 0:
        return
    desc, frames, (w, h) = load_workload(args.workload, args.frames_per_step)
    value, dt, cpu = cpu_throughput(frames, w * h, max(1, args.steps), warm=max(1, min(args.warmup, 1)))
    print(json.dumps({
        "impl": "reference", "metric": "Megapixels/s decoded (8K VarDCT d1.0)", "value": value, "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic" if "synth" in args.workload else "real-file mosaic",
        "config": {"workload": desc, "note": "CPU restatement of jxl-oxide's generic render path, all host cores: "
                                              "concurrent frames x per-frame thread pool"},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synth8k", help="synth8k | synth4k | mosaic8k | file:PATH")
    ap.add_argument("--contexts", type=int, default=24, help="decoder contexts (CUDA streams) per GPU")
    ap.add_argument("--frames-per-step", type=int, default=48, help="independent frames decoded per step per GPU")
    ap.add_argument("--cpu-sample-frames", type=int, default=1)
    ap.add_argument("--stagger-ms", type=float, default=0.0, help="start offset between context groups within a step")
    ap.add_argument("--stagger-groups", type=int, default=4)
    ap.add_argument("--gather", default="none", choices=["none", "u8", "u16"],
                    help="also time decode + device-side packing + NCCL gather of every frame to rank 0 (BASELINE config #5)")
    ap.add_argument("--hf-lanes", default="auto", choices=["auto", "0", "8", "16", "32", "64", "128"],
                    help="HF coefficient schedule = streams per CTA: 0 (= 4) / 8 / 16 one warp per stream, 32 / 64 / 128 one "
                         "thread per stream; "
                         "auto = probe in child processes (parity against the default kernel, then a short A/B) and keep the "
                         "faster one. JXLB_HF_LANES in the environment overrides.")
    ap.add_argument("--probe", default=None, choices=["speed"], help=argparse.SUPPRESS)
    ap.add_argument("--probe-lanes", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--probe-device", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--pipeline-steps", action="store_true",
                    help="run the K timed steps back to back, contexts joined only at the end (default: joined after every "
                         "step - measured faster, profiles/r01_progress.md)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.probe:
        run_probe(args)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
