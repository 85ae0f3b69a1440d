import sys
sys.path.insert(0, '.')
import jxl_oxide_b200 as J
d = J.Decoder(0)
data = open(sys.argv[1], 'rb').read()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    d.decode(data); d.sync(); d.release_frames()
