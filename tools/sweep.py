"""Concurrency sweep: N decoder contexts each decoding the same frame once per round."""
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
import quick_time as q

if __name__ == '__main__':
    path = sys.argv[1]
    w, h = int(sys.argv[2]), int(sys.argv[3])
    counts = tuple(int(c) for c in sys.argv[4].split(','))
    q.sweep(path, w * h, counts=counts)
