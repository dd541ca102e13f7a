"""The C-ABI library loads on a CPU-only box and exports every symbol include/s3d.h declares."""
import ctypes
import os
import re

import pytest

from stereo_3d_reconstruction_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 's3d.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(s3d_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 14
    L = lib.load()
    for n in names:
        assert hasattr(L, n), 'libs3d_b200.so does not export %s' % n
        assert n in lib.SIGNATURES, 'lib.py has no ctypes signature for %s' % n
    assert set(lib.SIGNATURES) == set(names)


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of S3dConvParams as compiled by gcc == the ctypes mirror."""
    c = tmp_path / 'sz.c'
    c.write_text('#include "%s"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu",'
                 'sizeof(S3dConvParams),offsetof(S3dConvParams,dz),offsetof(S3dConvParams,osN),'
                 'offsetof(S3dConvParams,act_param),offsetof(S3dConvParams,bn));return 0;}\n'
                 % os.path.join(ROOT, 'include', 's3d.h'))
    exe = tmp_path / 'sz'
    import subprocess
    subprocess.check_call(['gcc', str(c), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    P = lib.S3dConvParams
    assert got == [ctypes.sizeof(P), P.dz.offset, P.osN.offset, P.act_param.offset, P.bn.offset]


def test_version_and_no_cpu_fallback():
    import torch
    L = lib.load()
    assert b'sm_100a' in L.s3d_version()
    if not torch.cuda.is_available():
        assert L.s3d_device_check(0) < 0            # no device -> error code, not a silent CPU path
        from stereo_3d_reconstruction_b200 import ops, models
        from tests.common import small_cfg
        with pytest.raises(lib.S3dError):
            ops.chamfer_forward(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
        with pytest.raises(lib.S3dError):
            models.build_model('Stereo2Voxel', small_cfg(), seed=0).pack()


def test_chamfer_extension_entry_point():
    """The reference-shaped entry (README.md:62-65): top-level `extensions.chamfer_dist`, built by its setup.py as a thin
    torch C++ extension over the C ABI (no kernel in it: it links libs3d_b200.so)."""
    import subprocess
    import torch
    from extensions.chamfer_dist import ChamferDistance, backend
    ext_dir = os.path.join(ROOT, 'extensions', 'chamfer_dist')
    assert os.path.exists(os.path.join(ext_dir, 'setup.py'))
    assert backend() == 'torch_extension', 'run __graft_entry__.build() (python setup.py build_ext --inplace)'
    so = [f for f in os.listdir(ext_dir) if f.startswith('chamfer') and f.endswith('.so')][0]
    needed = subprocess.check_output(['readelf', '-d', os.path.join(ext_dir, so)]).decode()
    assert 'libs3d_b200.so' in needed                     # the shim links the C-ABI library
    if not torch.cuda.is_available():
        with pytest.raises((RuntimeError, lib.S3dError)):   # CPU tensors: loud failure, no CPU path
            ChamferDistance()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    x = torch.zeros(1, 4, 3, requires_grad=True)
    with pytest.raises(NotImplementedError):              # never silently gradient-less
        ChamferDistance()(x, torch.zeros(1, 4, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'stereo_3d_reconstruction_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), os.path.join(dp, f)
    assert not re.search(r'^\s*(from|import)\s+oracle\b', open(os.path.join(ROOT, 'runner.py')).read(), flags=re.M)
