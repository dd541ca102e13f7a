"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: this library's kernels of the LAST full forward
pass(es), grouped by name.  usage: launch_summary.py launches.csv [n_forwards]   (window: the n forwards after the first, warm-up, one)"""
import collections, csv, re, sys
path = sys.argv[1]
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
ours = re.compile(r'conv_|conv2d_|cls_fused|fuse_views|upsample_disp|depth_to_space|pool_|concat_volume|soft_argmin|corr_|chamfer|pack_image|split_|partials_|gonce_')
L = []
for r in rows:
    n = r[ki]
    if not ours.search(n):
        continue
    n = re.sub(r'\(CUtensorMap_st.*|\(const .*|\(.*\)$', '', n).replace('void ', '').replace('s3d::', '').replace('(anonymous namespace)::', '').replace('unnamed>::', '')
    t = float(r[vi].replace(',', '')) / (1e6 if r[ui] == 'ns' else 1e3 if r[ui] == 'us' else 1)
    L.append((n, t))
L = [(n.lstrip('<'), t) for n, t in L]
# one forward = the launches from one feature-encoder first layer (conv_first_kernel<.., 3, ..>) to the next
first = re.compile(r'conv_first_kernel<\d+, 3,|conv_first_tc_kernel<3,')      # the feature encoder's first layer (3 input channels)
starts = [i for i, (n, _) in enumerate(L) if first.match(n) and (i == 0 or not first.match(L[i - 1][0]))]
per = starts[1] - starts[0] if len(starts) > 1 else len(L)
win = L[starts[1]:starts[1] + nf * per] if len(starts) > nf else L       # skip the first (warm-up) forward
tot = sum(t for _, t in win)
agg = collections.OrderedDict()
for n, t in win:
    c = agg.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += t
print('%-78s %5s %10s %7s' % ('kernel', 'count', 'total ms', 'share'))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-78s %5d %10.3f %6.1f%%' % (n[:78], c, t, 100 * t / tot))
print('%-78s %5d %10.3f' % ('TOTAL (%d forwards of %d launches)' % (nf, per), len(win), tot))
a = [t for n, t in win if re.search(r'conv_scatter_kernel<0, 1, (128|256), 64|conv_scatter_concat|conv_scatter_rm|conv_scatter_cls', n) and t > 1.0 or re.search(r'map_conv|gonce_assemble', n)]   # (the 2-D encoder layers share the kernel)
if a:
    print('\n# aggregation layers (fused volume + dres0a, dres0b, dres1a, dres1b, cls_a): %d launches, mean %.3f ms, %.1f%% of the window'
          % (len(a), sum(a) / len(a), 100 * sum(a) / tot))
print('\n# one forward, in launch order (ms):')
for i, (n, t) in enumerate(win[-per:]):
    print('%2d  %-70s %8.3f' % (i, n[:70], t))
