import copy

import torch

from config import cfg as default_cfg


def small_cfg(**kw):
    """A reduced network (same structure, fewer channels, 64x64 input) the CPU oracle runs in well under a second."""
    c = default_cfg.clone()
    c.CONST.IMG_H = c.CONST.IMG_W = 64
    c.CONST.N_POINTS = 128
    c.NETWORK.MAX_DISP = 8
    c.NETWORK.FEAT_CHANNELS = 16
    c.NETWORK.ENC_CHANNELS = [16, 32]
    c.NETWORK.AGG_CHANNELS = 32
    c.NETWORK.REC_CHANNELS = [16, 32, 32, 64, 64]
    c.NETWORK.LATENT_HW = 2
    c.NETWORK.DEC_CHANNELS = [32, 32, 16, 16, 8]
    c.NETWORK.POINT_FC = 64
    for k, v in kw.items():
        sec, key = k.split('__')
        c[sec][key] = v
    return c
