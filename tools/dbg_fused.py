import sys, itertools
sys.path.insert(0, '/root/repo')
import numpy as np
import dplug_b200 as b2
from tests import oracle
oracle.lib()
from tests.test_gpu_mip_fused import _rand, _pair, _fill_stale, _oracle_chain
rng = np.random.default_rng(1000 + 3)
w, h = 301, 167
src = _rand(rng, 0, w, h)
for trial in range(3):
  for q in [(2,0,0),(2,1,0),(2,1,1)]:
    g, o = _pair(b2, oracle, 0, 3, w, h)
    _fill_stale(rng, g, o, 0)
    stale1 = g.download(1).copy()
    g.upload(0, src); o.upload(0, src)
    g.generateLevels(1, list(q), (0,0,w,h)); _oracle_chain(o, 1, list(q), (0,0,w,h))
    for level in (1,2,3):
        a, b = g.download(level), o.download(level)
        bad = np.argwhere((a != b).any(axis=-1))
        if len(bad):
            y, x = bad[0]
            print(trial, q, 'level', level, 'nbad', len(bad), 'first', (y, x), 'gpu', a[y, x], 'ref', b[y, x], 'stale', stale1[y, x] if level == 1 else None)
            if level == 1:
                print('src quad', src[2*y:2*y+2, 2*x:2*x+2].reshape(-1, 4))
print('done')
