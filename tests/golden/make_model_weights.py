"""Export the tensors of the shipped dafx22 checkpoint that the control-rate graph needs
 and of the v2 checkpoint (maestro-v2.gin) into tests/golden/{dafx22,v2}_weights.npz,
keyed by their checkpoint keys (optimizer slots dropped), so that the
model tests run where /root/reference does not exist (the GPU box).  The reverb embedding is cut
to its first two instruments (2 x 24000 floats instead of 10 x 24000).

usage (in the build container): python tests/golden/make_model_weights.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ddsp_piano_b200.checkpoint import Checkpoint  # noqa: E402

MODELS = {'dafx22': '/root/reference/ddsp_piano/model_weights/dafx22/ckpt-0',
          'v2': '/root/reference/ddsp_piano/model_weights/v2/ckpt-225000'}
SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'


def export(name, prefix):
    ck = Checkpoint(prefix)
    out = {}
    for key in ck.keys():
        if not key.endswith(SUFFIX) or '.OPTIMIZER_SLOT' in key or not key.startswith('model/'):
            continue
        if ck.entries[key]['dtype'] != 1:          # float32 variables only
            continue
        x = ck.tensor(key)
        if 'reverb_dict' in key:
            x = x[:2]
        out[key] = x
    path = os.path.join(ROOT, 'tests', 'golden', f'{name}_weights.npz')
    np.savez_compressed(path, **out)
    print(f'{len(out)} tensors, {sum(v.size for v in out.values())} floats -> {path} '
          f'({os.path.getsize(path) / 1e6:.2f} MB)')
    for k, v in sorted(out.items()):
        print(f'  {k[len("model/"):-len(SUFFIX)]:70s} {list(v.shape)}')


if __name__ == '__main__':
    for model_name, model_prefix in MODELS.items():
        export(model_name, model_prefix)
