"""Developer diagnostic (GPU box): s3d_conv_igemm vs s3d_conv_direct on identical inputs, one
config per subprocess so a trap / hang in one config cannot poison the others.

  python scripts/dev_check_igemm.py            # run all configs, summary to gpurun_out/igemm_check.txt
  python scripts/dev_check_igemm.py NAME       # run one config in this process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (kind, dtype, N, D, H, W, Cin, Cout, extra)
    'pw_c64_bf16':      ('pointwise', 'bf16', 2, 1, 8, 8, 64, 64, {}),
    'pw_c32_bf16':      ('pointwise', 'bf16', 2, 1, 8, 8, 32, 32, {}),
    'pw_c16_bf16':      ('pointwise', 'bf16', 2, 1, 8, 8, 16, 16, {}),
    'pw_c128_n256':     ('pointwise', 'bf16', 4, 1, 8, 8, 128, 512, {}),
    'conv2d_s1_bf16':   ('conv2d', 'bf16', 2, 1, 16, 16, 64, 64, {'stride': 1}),
    'conv2d_s1_odd':    ('conv2d', 'bf16', 3, 1, 35, 35, 32, 48, {'stride': 1}),
    'conv2d_s2_bf16':   ('conv2d', 'bf16', 2, 1, 16, 16, 16, 32, {'stride': 2}),
    'conv2d_s2_odd':    ('conv2d', 'bf16', 2, 1, 37, 37, 16, 32, {'stride': 2}),
    'conv3d_bf16':      ('conv3d', 'bf16', 2, 8, 8, 8, 64, 64, {}),
    'conv3d_res_relu':  ('conv3d', 'bf16', 1, 4, 16, 16, 64, 64, {'residual': True, 'act': 'relu'}),
    'conv3d_cout1_f32': ('conv3d', 'bf16', 2, 8, 8, 8, 64, 1, {'plane_out': True}),
    'deconv_bf16':      ('deconv', 'bf16', 4, 2, 2, 2, 64, 32, {}),
    'deconv_16cube':    ('deconv', 'bf16', 2, 16, 16, 16, 32, 8, {'act': 'relu'}),
    'linear_map':       ('linear', 'bf16', 5, 1, 4, 4, 32, 80, {}),
    'pw_c32_tf32':      ('pointwise', 'tf32', 2, 1, 8, 8, 32, 32, {}),
    'conv2d_s1_tf32':   ('conv2d', 'tf32', 2, 1, 16, 16, 32, 64, {'stride': 1}),
    'conv3d_tf32':      ('conv3d', 'tf32', 1, 4, 8, 8, 16, 32, {}),
    'conv3d_big_bf16':  ('conv3d', 'bf16', 4, 32, 64, 64, 64, 64, {'act': 'relu', 'time': True}),
    'conv3d_big_nohalo': ('conv3d', 'bf16', 4, 32, 64, 64, 64, 64, {'act': 'relu', 'time': True, 'env': {'S3D_NO_HALO': '1'}}),
    'conv3d_big_tf32':  ('conv3d', 'tf32', 4, 32, 64, 64, 32, 64, {'act': 'relu', 'time': True}),
    'conv3d_x32_bf16':  ('conv3d', 'bf16', 32, 32, 64, 64, 64, 64, {'act': 'relu', 'time': True}),
    'conv3d_bf16_bo1':  ('conv3d', 'bf16', 2, 8, 8, 8, 64, 64, {'env': {'S3D_HALO_BASE_OFFSET': '1'}}),
    'conv2d_s1_bo1':    ('conv2d', 'bf16', 2, 1, 16, 16, 64, 64, {'stride': 1, 'env': {'S3D_HALO_BASE_OFFSET': '1'}}),
    'conv3d_odd':       ('conv3d', 'bf16', 2, 5, 19, 13, 64, 48, {'act': 'relu'}),
    'conv3d_odd_bo1':   ('conv3d', 'bf16', 2, 5, 19, 13, 64, 48, {'act': 'relu', 'env': {'S3D_HALO_BASE_OFFSET': '1'}}),
    'conv2d_c128':      ('conv2d', 'bf16', 3, 1, 40, 24, 128, 256, {'stride': 1}),
    'conv3d_c16':       ('conv3d', 'bf16', 2, 8, 8, 8, 16, 16, {'act': 'relu'}),
    'conv3d_c32_odd':   ('conv3d', 'bf16', 1, 5, 37, 11, 32, 32, {}),
    'conv2d_big':       ('conv2d', 'bf16', 128, 1, 64, 64, 64, 64, {'stride': 1, 'act': 'relu', 'time': True}),
}
ONLY = os.environ.get('S3D_DEV_ONLY')


def run_one(name):
    import torch
    import torch.nn as nn
    from stereo_3d_reconstruction_b200 import lib
    from stereo_3d_reconstruction_b200.layers import PackedConv
    kind, dt, N, D, H, W, Cin, Cout, ex = CONFIGS[name]
    torch.manual_seed(0)
    dev = 'cuda'
    code = lib.DTYPE_BF16 if dt == 'bf16' else lib.DTYPE_F32
    tdt = torch.bfloat16 if dt == 'bf16' else torch.float32
    act = {'relu': lib.ACT_RELU, None: lib.ACT_NONE}[ex.get('act')]
    if kind == 'pointwise':
        w = torch.randn(Cout, Cin) * (1.0 / Cin ** 0.5)
        pc = PackedConv.from_pointwise(w, torch.randn(Cout) * 0.1, None, act, code, dev)
    elif kind == 'conv2d':
        conv = nn.Conv2d(Cin, Cout, 3, ex['stride'], 1, bias=True)
        pc = PackedConv.from_conv(conv, None, act, code, dev)
    elif kind == 'conv3d':
        conv = nn.Conv3d(Cin, Cout, 3, 1, 1, bias=not ex.get('plane_out'))
        pc = PackedConv.from_conv(conv, None, act, code, dev)
    elif kind == 'deconv':
        dc = nn.ConvTranspose3d(Cin, Cout, 4, 2, 1, bias=False)
        pc = PackedConv.from_deconv_k4s2p1(dc, None, act, code, dev)
    elif kind == 'linear':
        fc = nn.Linear(Cin * H * W, Cout)
        pc = PackedConv.from_linear_over_map(fc, Cin, H, W, act, code, dev)
    x = torch.randn(N, D, H, W, pc.cin_pad, device=dev).to(tdt)
    kw = {}
    outs = []
    for engine in ('direct', 'igemm'):
        if ex.get('plane_out'):
            oD, oH, oW = pc.out_grid(D, H, W)
            out = torch.full((N, oD, oH, oW), -7.0, dtype=torch.float32, device=dev)
            pc(x, out=out, out_view=(0, (oD * oH * oW, oH * oW, oW, 1)), cout_store=1, engine=engine)
        else:
            res = None
            if ex.get('residual'):
                torch.manual_seed(1)
                oD, oH, oW = pc.out_grid(D, H, W)
                res = torch.randn(N, oD, oH, oW, pc.cout_pad, device=dev).to(tdt)
            out = pc(x, residual=res, engine=engine)
        torch.cuda.synchronize()
        outs.append(out.float())
    ref, got = outs
    err = (ref - got).abs()
    scale = ref.abs().max().item() + 1e-6
    tol = (2e-2 if dt == 'bf16' else 5e-3) * scale
    bad = (err > tol)
    ok = not bad.any().item() and torch.isfinite(got).all().item()
    print('%-18s %s max_err=%.4g (scale %.3g) bad=%.4f%% shape=%s tile=%s bn=%d' % (
        name, 'OK  ' if ok else 'FAIL', err.max().item(), scale, 100.0 * bad.float().mean().item(), tuple(got.shape),
        None, pc.bn))
    if not ok and got.dim() == 5:
        e2 = bad.float()
        print('   bad by channel octet:', [round(v, 3) for v in e2.mean(dim=(0, 1, 2, 3)).view(-1, 8).mean(1).tolist()][:16])
        print('   bad by x%8:', [round(e2[:, :, :, i::8].mean().item(), 3) for i in range(min(8, e2.shape[3]))])
        print('   bad by y:', [round(e2[:, :, i].mean().item(), 3) for i in range(min(16, e2.shape[2]))])
        print('   bad by z:', [round(e2[:, i].mean().item(), 3) for i in range(min(16, e2.shape[1]))])
        print('   bad by n:', [round(e2[i].mean().item(), 3) for i in range(min(8, e2.shape[0]))])
        print('   sample ref', ref.flatten()[:8].tolist())
        print('   sample got', got.flatten()[:8].tolist())
    if ex.get('time') and ok:
        for _ in range(3):
            pc(x, out=out, engine='igemm')
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            pc(x, out=out, engine='igemm')
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        fl = pc.flops(N, D, H, W)
        print('   time %.3f ms  %.1f TFLOP/s (algorithmic)' % (ms, fl / ms / 1e9))
    return ok


def main():
    if len(sys.argv) > 1:
        ok = run_one(sys.argv[1])
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    lines = []
    for name in CONFIGS:
        if ONLY and not any(name.startswith(o) for o in ONLY.split(',')):
            continue
        t = time.time()
        env = dict(os.environ)
        env.update(CONFIGS[name][8].get('env', {}))
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=240, env=env)
            txt = r.stdout.strip() or ''
            if r.returncode != 0 and 'FAIL' not in txt:
                txt += '\n   rc=%d stderr tail: %s' % (r.returncode, r.stderr.strip()[-1500:])
        except subprocess.TimeoutExpired:
            txt = '%-18s TIMEOUT' % name
        lines.append(txt + '   [%.0fs]' % (time.time() - t))
        print(lines[-1], flush=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'igemm_check.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
