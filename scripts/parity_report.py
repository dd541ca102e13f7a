"""Prints the measured parity errors (CUDA path vs oracle) for every precision / cost-volume mode at the small
test config and at the smoke config, so the tolerances written in the tests carry a known margin."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import models as O
from stereo_3d_reconstruction_b200 import models as M
from stereo_3d_reconstruction_b200.utils import synthetic
from tests.common import small_cfg
from config import cfg as dcfg

def smoke_cfg(prec):
    c = dcfg.clone()
    c.CONST.IMG_H = c.CONST.IMG_W = 64
    c.NETWORK.MAX_DISP = 8
    c.NETWORK.REC_CHANNELS = [16, 32, 32, 64, 64]
    c.NETWORK.LATENT_HW = 2
    c.NETWORK.DEC_CHANNELS = [32, 32, 16, 16, 8]
    c.NETWORK.PRECISION = prec
    return c

for label, mk in (('small', lambda p, cv: small_cfg(NETWORK__PRECISION=p, NETWORK__COST_VOLUME=cv)),
                  ('smoke', lambda p, cv: smoke_cfg(p))):
    for cv in (('concat', 'corr') if label == 'small' else ('concat',)):
        for prec in ('fp32', 'tf32x3', 'tf32', 'bf16'):
            for seed in (0, 1, 2):
                cfg = mk(prec, cv)
                oracle = O.make_model('Stereo2Voxel', cfg, seed=seed)
                model = M.build_model('Stereo2Voxel', cfg); model.load_state_dict(oracle.state_dict()); model.cuda().pack()
                l, r, _ = synthetic.stereo_pair(3, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 2 * cfg.NETWORK.MAX_DISP, seed=seed)
                with torch.no_grad():
                    rdl, rdr, rv = oracle(l, r)
                    dl, dr, v = model(l.cuda(), r.cuda())
                ed = max((dl.cpu() - rdl).abs().max().item(), (dr.cpu() - rdr).abs().max().item()) / rdl.abs().max().item()
                ev = (v.cpu() - rv).abs().max().item()
                print('%-6s %-7s %-5s seed %d: disp rel err %.3e   voxel abs err %.3e   (vox mean abs err %.2e)' %
                      (label, cv, prec, seed, ed, ev, (v.cpu() - rv).abs().mean().item()))
