"""Ad-hoc timing helper (not part of the bench contract): single-context latency and a
concurrency sweep (N decoder contexts = N CUDA streams decoding independent frames)."""
import sys
import threading
import time

sys.path.insert(0, '.')
import jxl_oxide_b200 as J  # noqa: E402


def latency(path, n=5):
    data = open(path, 'rb').read()
    d = J.Decoder(0)
    for _ in range(2):
        d.decode(data)
        d.sync()
        d.release_frames()
    d.set_profile(True)
    d.profile_reset()
    t = time.time()
    for _ in range(n):
        d.decode(data)
        d.sync()
        d.release_frames()
    dt = (time.time() - t) / n
    info = {k: round(d.profile(k)[1] / n, 2) for k in ("modular_decode", "build_block_info", "decode_hf", "hf_dequant_cfl",
                                                          "hf_transform", "filters_fused", "gaborish", "epf_step", "xyb_to_rgb")}
    print('%s: %.2f ms/frame; kernel ms/frame: %s' % (path.split('/')[-1], dt * 1e3, info))
    host = {k: round(d.profile('host:' + k)[1] / n, 2) for k in ("lf_global", "alloc", "lf_coeff", "mlf", "hf_metadata", "hf_global",
                                                                  "pass_groups", "inverse_transforms", "render_vardct", "filters",
                                                                  "filters_colour")}
    print('   wall ms per host phase (device waits included): %s' % host)
    d.close()


def sweep(path, px, counts=(1, 4, 9, 18, 36, 72)):
    data = open(path, 'rb').read()
    for n in counts:
        decs = [J.Decoder(0) for _ in range(n)]
        for d in decs:
            d.preload(0, data)

        def work(d, reps):
            for _ in range(reps):
                d.decode_slot(0)
                d.sync()
                d.release_frames()
        for reps in (1, 3):
            ts = [threading.Thread(target=work, args=(d, reps)) for d in decs]
            t = time.time()
            for th in ts:
                th.start()
            for th in ts:
                th.join()
            dt = time.time() - t
        print('contexts=%3d: %.1f ms per round, %.1f MP/s' % (n, dt / 3 * 1e3, px * n * 3 / dt / 1e6))
        for d in decs:
            d.close()


if __name__ == '__main__':
    latency('tests/golden/benchmark-data/starrail.d1-e6.jxl')
    latency('tests/golden/benchmark-data/minecraft.d0-e6.jxl')
    sweep('tests/golden/benchmark-data/starrail.d1-e6.jxl', 2560 * 1440)
