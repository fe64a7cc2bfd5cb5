"""Import the UNMODIFIED reference python layer (torchsparse + seg models) from
/root/reference with oracle/_ref/backend.so standing in for `torchsparse.backend`.

TEST INFRASTRUCTURE ONLY -- works only in the build container (the GPU box has no
/root/reference).  Used by tests/golden/make_golden.py to generate fixtures and by
the oracle self-check.
"""
import os
import sys
import warnings

REF_ROOT = '/root/reference'
TS_DIR = os.path.join(REF_ROOT, 'segmentation', 'torchsparse-u')
SEG_DIR = os.path.join(REF_ROOT, 'segmentation')


def available() -> bool:
    from . import build_ref
    return os.path.isdir(TS_DIR) and os.path.exists(build_ref.OUT)


def import_reference():
    """Returns (torchsparse, seg_utils_module, linkencoder_module)."""
    from . import build_ref
    warnings.filterwarnings('ignore', category=FutureWarning)
    backend = build_ref.load()
    sys.modules['torchsparse.backend'] = backend
    for p in (SEG_DIR, TS_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torchsparse
    torchsparse.backend = backend
    import core.models.utils as seg_utils
    import core.models.semantic_kitti.linkencoder as linkencoder
    return torchsparse, seg_utils, linkencoder
