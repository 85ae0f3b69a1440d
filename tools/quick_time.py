import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
import jxl_oxide_b200 as J
data = open('tests/golden/benchmark-data/starrail.d1-e6.jxl','rb').read()
d = J.Decoder(0)
for i in range(3):
    d.decode(data); d.sync(); d.release_frames()
t=time.time(); n=5
for i in range(n):
    d.decode(data); d.sync(); d.release_frames()
dt=(time.time()-t)/n
print('starrail 2560x1440 decode (resident output): %.2f ms  -> %.1f MP/s'%(dt*1e3, 2560*1440/dt/1e6))
data2 = open('tests/golden/benchmark-data/minecraft.d0-e6.jxl','rb').read()
for i in range(2):
    d.decode(data2); d.sync(); d.release_frames()
t=time.time()
for i in range(n):
    d.decode(data2); d.sync(); d.release_frames()
dt=(time.time()-t)/n
print('minecraft lossless 2560x1440: %.2f ms -> %.1f MP/s'%(dt*1e3, 2560*1440/dt/1e6))
