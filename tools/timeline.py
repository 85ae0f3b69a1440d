"""Timeline of one concurrent round: N decoder contexts each decode one frame with profiling on.

Prints, per context, the kernel / host-phase intervals (ms since a common origin) and a summary of
how long each context's stream had a kernel running vs. was idle. Usage:
    python tools/timeline.py FILE N [--full]
"""
import sys
import threading
import time

sys.path.insert(0, '.')
import jxl_oxide_b200 as J  # noqa: E402


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    full = '--full' in sys.argv
    data = open(path, 'rb').read()
    decs = [J.Decoder(0) for _ in range(n)]
    for d in decs:
        d.preload(0, data)

    def work(d, reps):
        for _ in range(reps):
            d.decode_slot(0)
            d.sync()
            d.release_frames()

    def round_(reps):
        ts = [threading.Thread(target=work, args=(d, reps)) for d in decs]
        t = time.time()
        for th in ts:
            th.start()
        for th in ts:
            th.join()
        return (time.time() - t) / reps

    round_(2)
    print('unprofiled round: %.1f ms' % (round_(2) * 1e3))
    mode = 2 if '--trace' in sys.argv else 1
    for d in decs:
        d._L.jxlb_set_profile(d._h, mode)
        d.profile_reset()
    print('profiled round (mode %d): %.1f ms' % (mode, round_(1) * 1e3))
    if mode == 2:
        tls = [d.timeline() for d in decs]
        t_min = min(t0 for tl in tls for (_, t0, _) in tl)
        for i, tl in enumerate(tls):
            parts = []
            for k in range(0, len(tl), 2):
                (hn, h0, h1), (dn, d0, d1) = tl[k], tl[k + 1]
                parts.append('launch %7.1f  dev %7.1f..%7.1f  return %7.1f |' % (h0 - t_min, d0 - t_min, d1 - t_min, h1 - t_min))
            print('ctx %2d: %s' % (i, ' '.join(parts)))
        return
    tls = [d.timeline() for d in decs]
    t_min = min(t0 for tl in tls for (_, t0, _) in tl)
    t_max = max(t1 for tl in tls for (_, _, t1) in tl)
    print('span %.1f ms' % (t_max - t_min))
    for i, tl in enumerate(tls):
        kern = sorted([(t0 - t_min, t1 - t_min, nm) for (nm, t0, t1) in tl if not nm.startswith('host:')])
        host = sorted([(t0 - t_min, t1 - t_min, nm) for (nm, t0, t1) in tl if nm.startswith('host:')])
        busy = sum(b - a for a, b, _ in kern)
        first, last = kern[0][0], kern[-1][1]
        print('ctx %2d: first kernel at %7.1f, last ends %7.1f, kernel-busy %6.1f ms, stream idle inside %6.1f ms' %
              (i, first, last, busy, (last - first) - busy))
        if full or i == 0:
            for a, b, nm in kern:
                print('      K %-18s %8.2f -> %8.2f  (%7.2f)' % (nm, a, b, b - a))
            for a, b, nm in host:
                print('      H %-18s %8.2f -> %8.2f  (%7.2f)' % (nm, a, b, b - a))


if __name__ == '__main__':
    main()
