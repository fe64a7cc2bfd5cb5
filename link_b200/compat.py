"""`torchsparse` import shim: makes `import torchsparse`, `torchsparse.nn as spnn`,
`torchsparse.nn.functional as F`, `torchsparse.nn.utils`, `torchsparse.utils.{collate,quantize}`
resolve to link_b200, so the reference's segmentation/core and detection/det3d code imports
unchanged (call sites: linkencoder.py:4-6, utils.py:2-5, ts_elk.py:3,6,62-63)."""
import importlib
import sys

_MAP = {
    'torchsparse': 'link_b200',
    'torchsparse.tensor': 'link_b200.tensor',
    'torchsparse.backend': 'link_b200.backend',
    'torchsparse.operators': 'link_b200.operators',
    'torchsparse.nn': 'link_b200.nn',
    'torchsparse.nn.functional': 'link_b200.nn.functional',
    'torchsparse.nn.modules': 'link_b200.nn.modules',
    'torchsparse.nn.utils': 'link_b200.nn.utils',
    'torchsparse.utils': 'link_b200.utils',
    'torchsparse.utils.collate': 'link_b200.utils.collate',
    'torchsparse.utils.quantize': 'link_b200.utils.quantize',
}


def install() -> None:
    for alias, real in _MAP.items():
        sys.modules[alias] = importlib.import_module(real)


def uninstall() -> None:
    for alias in _MAP:
        sys.modules.pop(alias, None)
