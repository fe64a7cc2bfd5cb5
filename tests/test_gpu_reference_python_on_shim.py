"""Backend-only drop-in, end to end: the REFERENCE's own python layer and model classes, staged untouched
under baseline/_ref/py, run on link_b200.backend registered as `torchsparse.backend` and reproduce the
outputs of link_b200's fused modules (tests/ref_on_shim_driver.py, own process)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_elkblock_and_encoder_on_link_b200_backend():
    if not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'py', 'torchsparse')):
        pytest.skip('baseline/_ref/py not staged (oracle/build_ref.py::stage_python needs /root/reference at build time)')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_on_shim_driver.py')], capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith('{')][-1])
    # block outputs are LayerNorm'd, unit scale: same tolerance as the golden-fixture tests (atol 4e-5 + rtol 1e-4)
    for k in ('block_cos_3x7', 'block_cosx_2x3'):
        assert res[k]['max_abs_diff'] <= 4e-5 + 1e-4 * res[k]['ref_abs_max'], (k, res[k])
    e = res['encoder_cos_3x7']
    assert not e['unexpected_keys'] and all(k.endswith('num_batches_tracked') for k in e['missing_keys']), e
    assert e['max_abs_diff'] <= 1e-4 + 1e-3 * e['ref_abs_max'], e
