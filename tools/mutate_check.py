"""Robustness check: decode bit-flipped / truncated variants of valid streams; every outcome must be a
decode or a JxlError value (run under `timeout`, optionally under compute-sanitizer)."""
import random
import sys

sys.path.insert(0, '.')
import jxl_oxide_b200 as J  # noqa: E402


def main():
    files = sys.argv[1:-1]
    n = int(sys.argv[-1])
    d = J.Decoder(0)
    rng = random.Random(1234)
    ok = err = 0
    for path in files:
        data = bytearray(open(path, 'rb').read())
        for i in range(n):
            m = bytearray(data)
            kind = i % 3
            if kind == 0:  # flip a few bits past the headers
                for _ in range(1 + i % 4):
                    pos = rng.randrange(min(40, len(m) // 4), len(m))
                    m[pos] ^= 1 << rng.randrange(8)
            elif kind == 1:  # truncate
                m = m[: rng.randrange(len(m) // 8, len(m))]
            else:  # overwrite a run with random bytes
                pos = rng.randrange(len(m) // 3, len(m) - 8)
                for k in range(8):
                    m[pos + k] = rng.randrange(256)
            try:
                d.decode(bytes(m))
                d.sync()
                d.release_frames()
                ok += 1
            except J.JxlError as e:
                err += 1
        d.decode(bytes(data))  # still usable, and the pristine stream still decodes
        d.release_frames()
    print('mutations decoded: %d, clean errors: %d' % (ok, err))


if __name__ == '__main__':
    main()
