import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def knobs():
    """A/B switches (lib.set_knob) for the duration of one test; everything is reset afterwards."""
    from stereo_3d_reconstruction_b200 import lib
    touched = []

    def set_(name, value):
        touched.append(name)
        lib.set_knob(name, value)

    yield set_
    for name in touched:
        lib.set_knob(name, 0)
