# -*- coding: utf-8 -*-
"""ORACLE -- test infrastructure, not product code.

Stock ``torch.nn`` CPU restatement of the Stereo2Voxel / Stereo2Point inference
path named by BASELINE.json's ``north_star``.

PARITY UNPINNED.  The reference's model code lives on the upstream
``Stereo2Voxel`` / ``Stereo2Point`` branches (/root/reference/README.md:5,56,62)
which are not mounted and cannot be fetched; /root/reference holds README.md and
requirements.txt only, with no tests, fixtures or golden vectors.  Every layer
below is therefore a [SPEC] decision of this repo (SURVEY.md section 8(a) rows
E,V,A,S,X,D,F,P,M), frozen in DESIGN.md.  Results compared with this module are
"parity vs north_star restatement", never "vs upstream".

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product package
(``stereo_3d_reconstruction_b200``) must never import it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def conv_bn2d(cin, cout, k, s, p):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p, bias=False), nn.BatchNorm2d(cout))


def conv_bn3d(cin, cout):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm3d(cout))


class FeatureEncoder(nn.Module):
    """Row E: siamese 2D conv encoder, stride 4, shared by left and right image."""

    def __init__(self, cfg):
        super().__init__()
        c1, c2 = cfg.NETWORK.ENC_CHANNELS
        cf = cfg.NETWORK.FEAT_CHANNELS
        self.conv0 = conv_bn2d(3, c1, 3, 2, 1)      # H/2
        self.conv1 = conv_bn2d(c1, c1, 3, 1, 1)
        self.conv2 = conv_bn2d(c1, c2, 3, 2, 1)     # H/4
        self.conv3 = conv_bn2d(c2, c2, 3, 1, 1)     # residual block: x + bn(conv(relu(bn(conv(x)))))
        self.conv4 = conv_bn2d(c2, c2, 3, 1, 1)
        self.conv5 = nn.Conv2d(c2, cf, 3, 1, 1, bias=True)   # features, no BN / activation

    def forward(self, x):
        x = F.relu(self.conv0(x))
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        y = F.relu(self.conv3(x))
        x = x + self.conv4(y)
        return self.conv5(x)


def build_concat_volume(ref, tgt, max_disp, direction):
    """Row V (concat).  ref/tgt: [B,C,h,w].

    direction=-1 (left reference):  vol[b,:C,d,y,x]=ref[b,:,y,x], vol[b,C:,d,y,x]=tgt[b,:,y,x-d]
    direction=+1 (right reference): vol[b,C:,d,y,x]=tgt[b,:,y,x+d]
    Out-of-image target samples are zero.
    """
    B, C, h, w = ref.shape
    vol = ref.new_zeros(B, 2 * C, max_disp, h, w)
    for d in range(max_disp):
        vol[:, :C, d] = ref
        if d == 0:
            vol[:, C:, d] = tgt
        elif d < w:
            if direction < 0:
                vol[:, C:, d, :, d:] = tgt[:, :, :, :w - d]
            else:
                vol[:, C:, d, :, :w - d] = tgt[:, :, :, d:]
    return vol


def build_corr_volume(ref, tgt, max_disp, direction):
    """Row V (corr): cost[b,d,y,x] = (1/C) sum_c ref[b,c,y,x] * tgt[b,c,y,x -/+ d]."""
    B, C, h, w = ref.shape
    cost = ref.new_zeros(B, max_disp, h, w)
    for d in range(max_disp):
        if d == 0:
            cost[:, d] = (ref * tgt).mean(1)
        elif d < w:
            if direction < 0:
                cost[:, d, :, d:] = (ref[:, :, :, d:] * tgt[:, :, :, :w - d]).mean(1)
            else:
                cost[:, d, :, :w - d] = (ref[:, :, :, :w - d] * tgt[:, :, :, d:]).mean(1)
    return cost


def soft_argmin(cost):
    """Row S: p = softmax_d(-cost); disp = sum_d d * p[d].  cost: [B,D,h,w] -> [B,h,w]."""
    D = cost.shape[1]
    p = F.softmax(-cost, dim=1)
    d = torch.arange(D, dtype=cost.dtype, device=cost.device).view(1, D, 1, 1)
    return (p * d).sum(1)


def soft_argmax(score):
    """Correlation variant: higher correlation = better match, so p = softmax_d(+score)."""
    return soft_argmin(-score)


class CostAggregation(nn.Module):
    """Row A: 3x3x3 Conv3d stack over the concat volume -> 1-channel cost."""

    def __init__(self, cfg):
        super().__init__()
        cin = 2 * cfg.NETWORK.FEAT_CHANNELS
        a = cfg.NETWORK.AGG_CHANNELS
        self.dres0a = conv_bn3d(cin, a)
        self.dres0b = conv_bn3d(a, a)
        self.dres1a = conv_bn3d(a, a)
        self.dres1b = conv_bn3d(a, a)        # residual: x + bn(conv(relu(bn(conv(x)))))
        self.cls_a = conv_bn3d(a, a)
        self.cls_b = nn.Conv3d(a, 1, 3, 1, 1, bias=False)

    def forward(self, vol):
        x = F.relu(self.dres0a(vol))
        x = F.relu(self.dres0b(x))
        y = F.relu(self.dres1a(x))
        x = x + self.dres1b(y)
        x = F.relu(self.cls_a(x))
        return self.cls_b(x).squeeze(1)     # [B,D,h,w]


class DispNet(nn.Module):
    """Rows E+V+A+S: bidirectional disparity at full resolution, in input pixels."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.encoder = FeatureEncoder(cfg)
        if cfg.NETWORK.COST_VOLUME == 'concat':
            self.aggregation = CostAggregation(cfg)

    def forward(self, left, right):
        B, _, H, W = left.shape
        D = self.cfg.NETWORK.MAX_DISP
        feats = self.encoder(torch.cat([left, right], 0))
        fl, fr = feats[:B], feats[B:]
        if self.cfg.NETWORK.COST_VOLUME == 'concat':
            vol = torch.cat([build_concat_volume(fl, fr, D, -1),
                             build_concat_volume(fr, fl, D, +1)], 0)
            disp_q = soft_argmin(self.aggregation(vol))            # [2B,h,w], 1/4-res units
        else:
            cost = torch.cat([build_corr_volume(fl, fr, D, -1),
                              build_corr_volume(fr, fl, D, +1)], 0)
            disp_q = soft_argmax(cost)
        disp = F.interpolate(disp_q.unsqueeze(1) * 4.0, size=(H, W), mode='bilinear',
                             align_corners=False)
        return disp[:B], disp[B:], disp_q                            # [B,1,H,W] x2


class RGBDEncoder(nn.Module):
    """Row X: per-view (image, disparity) -> latent [C_last, L, L]."""

    def __init__(self, cfg):
        super().__init__()
        ch = [4] + list(cfg.NETWORK.REC_CHANNELS)
        self.layers = nn.ModuleList([conv_bn2d(ch[i], ch[i + 1], 3, 2, 1) for i in range(len(ch) - 1)])
        self.latent_hw = cfg.NETWORK.LATENT_HW

    def forward(self, rgbd):
        x = rgbd
        for l in self.layers:
            x = F.relu(l(x))
        return F.adaptive_avg_pool2d(x, self.latent_hw)


class VoxelDecoder(nn.Module):
    """Row D: ConvTranspose3d(k4,s2,p1) stack 2^3 -> 32^3, then 1x1x1 + sigmoid."""

    def __init__(self, cfg):
        super().__init__()
        ch = list(cfg.NETWORK.DEC_CHANNELS)
        self.ch0 = ch[0]
        self.layers = nn.ModuleList([
            nn.Sequential(nn.ConvTranspose3d(ch[i], ch[i + 1], 4, 2, 1, bias=False),
                          nn.BatchNorm3d(ch[i + 1])) for i in range(len(ch) - 1)])
        self.out = nn.ConvTranspose3d(ch[-1], 1, 1, bias=False)

    def forward(self, latent):
        x = latent.reshape(latent.shape[0], self.ch0, 2, 2, 2)
        for l in self.layers:
            x = F.relu(l(x))
        raw = x                                         # [N, 8, 32,32,32]
        vol = torch.sigmoid(self.out(x))                # [N, 1, 32,32,32]
        return raw, vol


class Merger(nn.Module):
    """Row F: context-aware fusion.  score per voxel per view -> softmax over views."""

    def __init__(self, cfg):
        super().__init__()
        ch = list(cfg.NETWORK.MERGER_CHANNELS)
        self.leaky = cfg.NETWORK.LEAKY_VALUE
        self.layers = nn.ModuleList([conv_bn3d(ch[i], ch[i + 1]) for i in range(len(ch) - 1)])

    def forward(self, raw, vol, n_views):
        # raw: [V*B, 8, 32^3], vol: [V*B, 1, 32^3]; views are the outer index.
        x = torch.cat([raw, vol], 1)
        for l in self.layers:
            x = F.leaky_relu(l(x), self.leaky)
        VB = x.shape[0]
        B = VB // n_views
        score = x.view(n_views, B, *x.shape[2:])
        w = F.softmax(score, dim=0)
        fused = (w * vol.view(n_views, B, *vol.shape[2:])).sum(0)
        return torch.clamp(fused, 0.0, 1.0)             # [B,32,32,32]


class Stereo2Voxel(nn.Module):
    """forward(left, right) -> (disp_left [B,1,H,W], disp_right [B,1,H,W], voxels [B,32,32,32])."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.dispnet = DispNet(cfg)
        self.rgbd_encoder = RGBDEncoder(cfg)
        self.decoder = VoxelDecoder(cfg)
        self.merger = Merger(cfg)

    def forward(self, left, right):
        dl, dr, _ = self.dispnet(left, right)
        scale = 1.0 / (4.0 * self.cfg.NETWORK.MAX_DISP)
        rgbd = torch.cat([torch.cat([left, dl * scale], 1), torch.cat([right, dr * scale], 1)], 0)
        latent = self.rgbd_encoder(rgbd)
        raw, vol = self.decoder(latent)
        voxels = self.merger(raw, vol, 2)
        return dl, dr, voxels


class PointDecoder(nn.Module):
    """Row P: (left latent, right latent) -> N_POINTS x 3 in [-0.5, 0.5]^3."""

    def __init__(self, cfg):
        super().__init__()
        c = cfg.NETWORK.REC_CHANNELS[-1]
        L = cfg.NETWORK.LATENT_HW
        self.n_points = cfg.CONST.N_POINTS
        self.conv = conv_bn2d(2 * c, 2 * c, 3, 2, 1)             # L -> L/2
        self.fc1 = nn.Linear(2 * c * (L // 2) * (L // 2), cfg.NETWORK.POINT_FC)
        self.fc2 = nn.Linear(cfg.NETWORK.POINT_FC, self.n_points * 3)

    def forward(self, lat_l, lat_r):
        x = F.relu(self.conv(torch.cat([lat_l, lat_r], 1)))
        x = F.relu(self.fc1(x.flatten(1)))
        x = torch.tanh(self.fc2(x)) * 0.5
        return x.view(-1, self.n_points, 3)


class Stereo2Point(nn.Module):
    """forward(left, right) -> (disp_left, disp_right, points [B,N_POINTS,3])."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.dispnet = DispNet(cfg)
        self.rgbd_encoder = RGBDEncoder(cfg)
        self.point_decoder = PointDecoder(cfg)

    def forward(self, left, right):
        B = left.shape[0]
        dl, dr, _ = self.dispnet(left, right)
        scale = 1.0 / (4.0 * self.cfg.NETWORK.MAX_DISP)
        rgbd = torch.cat([torch.cat([left, dl * scale], 1), torch.cat([right, dr * scale], 1)], 0)
        latent = self.rgbd_encoder(rgbd)
        pts = self.point_decoder(latent[:B], latent[B:])
        return dl, dr, pts


def init_weights(model, seed=0):
    """Synthetic weights [SPEC] (pretrained weights are unavailable offline): He-normal conv /
    linear weights from a seeded generator so activations stay O(1) through the stack, output
    layers scaled so costs / logits / scores have a non-degenerate spread, and randomised
    eval-mode BN statistics (mean~N(0,0.1), var~U(0.5,1.5), gamma~U(0.8,1.2), beta~N(0,0.1))
    so that BN folding is actually exercised (SURVEY.md 8(d) 'synthetic data')."""
    g = torch.Generator().manual_seed(seed + 12345)
    gains = {'decoder.out': 3.0, 'merger.layers.4.0': 4.0, 'point_decoder.fc2': 2.0}
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear, nn.ConvTranspose3d)):
                w = m.weight
                if isinstance(m, nn.ConvTranspose3d):
                    taps = 1
                    for k, s in zip(m.kernel_size, m.stride):
                        taps *= max(k // s, 1)
                    fan_in = w.shape[0] * taps
                else:
                    fan_in = w[0].numel()
                std = (2.0 / fan_in) ** 0.5 * gains.get(name, 1.0)
                w.copy_(torch.randn(w.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                n = m.num_features
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
                m.weight.copy_(torch.rand(n, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
    return model.eval()


def make_model(name, cfg, seed=0):
    torch.manual_seed(seed)
    m = {'Stereo2Voxel': Stereo2Voxel, 'Stereo2Point': Stereo2Point}[name](cfg)
    return init_weights(m, seed)


def iou_counts(voxels, gt, thresholds):
    """Row M: per sample, per threshold integer (intersection, union) counts.
    voxels [B,32,32,32] float, gt [B,32,32,32] {0,1}.  Returns int64 [B,T,2]."""
    out = []
    g = gt > 0.5
    for t in thresholds:
        p = voxels >= t
        inter = (p & g).flatten(1).sum(1)
        union = (p | g).flatten(1).sum(1)
        out.append(torch.stack([inter, union], 1))
    return torch.stack(out, 1).to(torch.int64)
