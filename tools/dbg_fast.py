import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from svo_pro_universal_b200 import capi, synth
from oracle import orc
ctx = capi.Context(0)
img = synth.make_image(7, blur=2)
p = capi.Pyramid(ctx, 1, 752, 480, 5); p.upload(img); p.build()
sm, nm = capi.fast_level_maps(ctx, p, 0, 0, 10, 10)
xy = orc.fast_detect(img, 10, 10); sc = orc.fast_score10(img, xy, 10)
exp = np.zeros_like(sm); exp[xy[:,1], xy[:,0]] = sc
bad = np.argwhere(sm != exp)
print('bad px', len(bad), 'first', bad[:5])
print('bad x mod 64 hist', np.bincount(bad[:,1] % 64, minlength=64))
print('bad y mod 16 hist', np.bincount(bad[:,0] % 16, minlength=16))
y, x = bad[0]
print('at', x, y, 'gpu', sm[y, x], 'exp', exp[y, x])
print(img[y-3:y+4, x-3:x+4])
