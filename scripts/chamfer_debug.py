import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import chamfer as OC
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.utils import synthetic
a = torch.zeros(2, 300, 3);  b = torch.zeros(2, 700, 3)
cases = [('zeros', a.clone(), b.clone())]
a2, b2 = synthetic.point_clouds(2, 300, 700, seed=9, duplicates=True)
a2[0, 5] = float('nan');  b2[1, 100:400] = float('nan');  cases.append(('nan-some', a2, b2))
a3, b3 = synthetic.point_clouds(1, 64, 600, seed=10)
a3[:] = float('nan');  cases.append(('nan-all', a3, b3))
a4, b4 = synthetic.point_clouds(1, 40, 300, seed=11)
a4[0, :20] = 3e19;  b4[0, :10] = -3e19;  cases.append(('inf', a4, b4))
for knob in (1, -1):
    lib.set_knob('chamfer_sym', knob)
    for name, x, y in cases:
        ref = OC.chamfer_c(x.numpy(), y.numpy())       # d1, d2, i1, i2
        got = [t.cpu().numpy() for t in ops.chamfer_forward(x.cuda(), y.cuda())]
        msg = []
        for nm, g, r in zip(('d1', 'd2', 'i1', 'i2'), got, ref):
            bad = ~((g == r) | (np.isnan(g) & np.isnan(r))) if g.dtype.kind == 'f' else (g != r)
            if bad.any():
                w = np.argwhere(bad)[:4]
                msg.append('%s: %d bad, first %s got %s ref %s' % (nm, bad.sum(), w.tolist(), [g[tuple(i)] for i in w], [r[tuple(i)] for i in w]))
        print('knob %2d %-9s %s' % (knob, name, '; '.join(msg) or 'ok'), flush=True)
