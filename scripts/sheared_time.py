"""Cost volume + first aggregation layer at the benchmark shape (B = 64, 64 x 64, D = 32, C = 32 -> 64): reference-once
kernel against the sheared form (map convolutions + streaming pass), CUDA events, L2 flushed between launches."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import bench
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.layers import PackedConv
B, C, D, h, w = 64, 32, 32, 64, 64
torch.manual_seed(0)
conv = nn.Conv3d(2 * C, 64, 3, 1, 1)
pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
featp = torch.zeros(2 * B, 1, h, w + 2 * D, C, dtype=torch.bfloat16, device='cuda')
featp[:, :, :, D:D + w] = torch.randn(2 * B, 1, h, w, C, device='cuda').to(torch.bfloat16)
out = torch.empty(2 * B, D, h, w, 64, dtype=torch.bfloat16, device='cuda')
bufs = {'maps_l': torch.empty(B, 1, h, w + 4, 384, device='cuda'), 'maps_r': torch.empty(B, 1, h, w + 4, 384, device='cuda'),
        'edge_l': torch.empty(B, 1, h, D, 256, device='cuda'), 'edge_r': torch.empty(B, 1, h, D, 256, device='cuda')}
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device='cuda')
print('reference-once kernel   %.4f ms' % bench._events_ms(lambda: ops.conv_concat_volume(pc, featp, B, D, D, out=out, ref_once=True), 10, flush=flush))
print('sheared form (5 launches) %.4f ms' % bench._events_ms(lambda: ops.conv_concat_volume_sheared(pc, featp, B, D, D, out=out, bufs=bufs), 10, flush=flush))
g = pc.gonce_convs(C, D, w, D)
L = lib.load()
st = lambda: torch.cuda.current_stream().cuda_stream
def mc(name, key):
    off, ntx, ow = g[name + '_geom']
    lib.check(L.s3d_map_conv(featp[:B].data_ptr(), g[name].weight.data_ptr(), bufs[key].data_ptr(), B, h, w + 2 * D, ow, off, ntx, g[name].cout_pad, lib.DTYPE_BF16, st()), 'mc')
print('  map conv, left images   %.4f ms (generic engine %.4f)' % (bench._events_ms(lambda: mc('left', 'maps_l'), 10, flush=flush),
      bench._events_ms(lambda: g['left'](featp[:B], out=bufs['maps_l']), 10, flush=flush)))
print('  edge conv, left images  %.4f ms (generic engine %.4f)' % (bench._events_ms(lambda: mc('edge_left', 'edge_l'), 10, flush=flush),
      bench._events_ms(lambda: g['edge_left'](featp[:B], out=bufs['edge_l']), 10, flush=flush)))
def asm():
    lib.check(L.s3d_concat_gonce_assemble(bufs['maps_l'].data_ptr(), bufs['maps_r'].data_ptr(), bufs['edge_l'].data_ptr(), bufs['edge_r'].data_ptr(),
                                          pc.bias.data_ptr(), out.data_ptr(), B, D, h, w, w + 4, lib.DTYPE_BF16, torch.cuda.current_stream().cuda_stream), 'asm')
print('  streaming pass          %.4f ms' % bench._events_ms(asm, 10, flush=flush))
