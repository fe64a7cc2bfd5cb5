"""Driver of tests/test_gpu_reference_python_on_shim.py (run in its own process: it registers the
REFERENCE's python package as `torchsparse`).  The reference's own python layer (torchsparse/nn/functional
glue, spnn modules, core/models/utils.py, ELKBlock / ELKEncoder of linkencoder.py -- staged untouched
under baseline/_ref/py by oracle/build_ref.py::stage_python) runs with ONE substitution: the module
`torchsparse.backend` is link_b200.backend (the ctypes binding of liblinkb200).  Its outputs are
compared with link_b200's own fused modules on the same weights and inputs.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import link_b200.backend as shim  # noqa: E402
from link_b200.utils.synthetic import random_voxels  # noqa: E402

stage = os.path.join(ROOT, 'baseline', '_ref', 'py')
sys.path.insert(0, stage)
sys.modules['torchsparse.backend'] = shim            # the drop-in: the only substitution
import torchsparse  # noqa: E402  (the reference's package)
torchsparse.backend = shim
assert os.path.realpath(torchsparse.__file__).startswith(os.path.realpath(stage)), torchsparse.__file__
from core.models.semantic_kitti import linkencoder as ref_models  # noqa: E402

import link_b200  # noqa: E402
from link_b200.elk import ELKBlock  # noqa: E402
from link_b200.linkencoder import ELKEncoder  # noqa: E402

dev = torch.device('cuda:0')
out = {}
with torch.no_grad():
    for name, (C, groups, op, s, r) in {'block_cos_3x7': (64, 2, 'cos', 7, 3), 'block_cosx_2x3': (16, 1, 'cos_x', 3, 2)}.items():
        coords = torch.from_numpy(random_voxels(6000, 48, seed=C)).to(dev)
        torch.manual_seed(C)
        ref_blk = ref_models.ELKBlock(C, C, groups=groups, baseop=op).to(dev).eval()
        ours = ELKBlock(C, C, groups=groups, baseop=op).to(dev).eval()
        ours.load_state_dict(ref_blk.state_dict(), strict=True)
        feats = torch.randn(coords.shape[0], C, device=dev)
        y_ref = ref_blk(torchsparse.SparseTensor(feats.clone(), coords, 1), s, r).F
        y_ours = ours(link_b200.SparseTensor(feats.clone(), coords, 1), s, r).F
        out[name] = {'max_abs_diff': float((y_ref - y_ours).abs().max()), 'ref_abs_max': float(y_ref.abs().max())}
    # the whole encoder: the reference's model class on the shim vs link_b200's
    coords = torch.from_numpy(random_voxels(9000, 64, seed=7, batch=2)).to(dev)
    torch.manual_seed(0)
    kw = dict(num_classes=19, cr=0.5, baseop='cos', r=3, s=7, groups=2)
    ref_enc = ref_models.ELKEncoder(**kw).to(dev).eval()
    ours_enc = ELKEncoder(**kw).to(dev).eval()
    missing = ours_enc.load_state_dict(ref_enc.state_dict(), strict=False)
    feats = torch.randn(coords.shape[0], 4, device=dev)
    y_ref = ref_enc(torchsparse.SparseTensor(feats.clone(), coords, 1))
    y_ours = ours_enc(link_b200.SparseTensor(feats.clone(), coords, 1))
    out['encoder_cos_3x7'] = {'max_abs_diff': float((y_ref - y_ours).abs().max()), 'ref_abs_max': float(y_ref.abs().max()),
                              'missing_keys': list(missing.missing_keys), 'unexpected_keys': list(missing.unexpected_keys)}
print(json.dumps(out))
