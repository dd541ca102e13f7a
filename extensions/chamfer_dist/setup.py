# -*- coding: utf-8 -*-
"""Build of the chamfer_dist extension, same command as the reference (README.md:64-65):

    cd extensions/chamfer_dist && python setup.py install --user        # or: python setup.py build_ext --inplace

It (1) makes sure libs3d_b200.so -- the sm_100a kernels behind the C ABI of include/s3d.h -- is built (nvcc, no torch), and
(2) compiles the thin torch C++ shim chamfer_cuda.cpp (g++ only: no kernel lives in it) and links it against that library.
"""
import os
import subprocess

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, 'stereo_3d_reconstruction_b200')

subprocess.check_call(['make', '-s', '-j8', '-C', os.path.join(PKG, 'csrc')])

setup(
    name='chamfer',
    version='2.0.0',
    ext_modules=[
        CUDAExtension(
            'chamfer', [os.path.join(os.path.relpath(HERE), 'chamfer_cuda.cpp') if os.getcwd() != HERE else 'chamfer_cuda.cpp'],
            include_dirs=[os.path.join(ROOT, 'include')],
            library_dirs=[PKG],
            libraries=[':libs3d_b200.so'],
            extra_compile_args={'cxx': ['-O2']},
            extra_link_args=['-Wl,-rpath,' + PKG, '-Wl,-rpath,$ORIGIN/../../stereo_3d_reconstruction_b200']),
    ],
    cmdclass={'build_ext': BuildExtension.with_options(use_ninja=False)})
