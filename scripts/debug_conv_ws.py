"""Debug aid: where does the tensor-core conv differ from a float64 contraction? (rows mod 128, columns)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from link_b200 import SparseTensor
from link_b200.nn.functional import conv as conv_mod
from link_b200.utils.synthetic import random_voxels
dev = torch.device('cuda:0')
n, c = int(os.environ.get('DBG_N', '3000')), int(os.environ.get('DBG_C', '64'))
coords = torch.from_numpy(random_voxels(n, 24, seed=1)).to(dev)
st = SparseTensor(torch.zeros(len(coords), c, device=dev), coords, 1)
km = conv_mod.build_kernel_map(st, (3, 3, 3), (1, 1, 1), (1, 1, 1), want_plan=False)
K, rows = km.nbr.shape
x = torch.randn(rows, c, device=dev)
w = torch.randn(K, c, c, device=dev) / 16
rel = km.nbr.cpu().numpy()
want = np.zeros((rows, c))
xn, wn = x.double().cpu().numpy(), w.double().cpu().numpy()
for k in range(K):
    hit = rel[k] >= 0
    want[hit] += xn[rel[k][hit]] @ wn[k]
got = conv_mod._conv_fwd(x, w, km.nbr, rows).double().cpu().numpy()
err = np.abs(got - want)
bad = err > 1e-3
print('rows', rows, 'bad elements', int(bad.sum()), 'of', bad.size, 'max err', err.max())
print('bad rows:', int(bad.any(1).sum()), ' bad cols:', np.nonzero(bad.any(0))[0].tolist())
br = np.nonzero(bad.any(1))[0]
print('bad row mod 128 histogram (first 20 rows):', br[:20].tolist(), ' mod128 unique:', sorted(set((br % 128).tolist()))[:40])
print('bad tiles:', sorted(set((br // 128).tolist()))[:40])
# single-offset probe: only offset k active
for k in (0, 13, 26):
    nb1 = torch.full_like(km.nbr, -1); nb1[k] = km.nbr[k]
    g1 = conv_mod._conv_fwd(x, w, nb1, rows).double().cpu().numpy()
    w1 = np.zeros((rows, c)); hit = rel[k] >= 0; w1[hit] = xn[rel[k][hit]] @ wn[k]
    e1 = np.abs(g1 - w1)
    print(f'offset {k} alone: max err {e1.max():.3e}, bad cols {np.nonzero((e1 > 1e-3).any(0))[0].tolist()[:70]}')
# identity weights: out[:, c] = x[nbr, c] -> shows which input channels land where
wi = torch.zeros(K, c, c, device=dev); wi[13] = torch.eye(c, device=dev)
nb1 = torch.full_like(km.nbr, -1); nb1[13] = km.nbr[13]
gi = conv_mod._conv_fwd(x, wi, nb1, rows).cpu().numpy()
xs = x.cpu().numpy()
row = 5
print('identity probe row 5: out vs x', np.round(gi[row, :8], 3).tolist(), np.round(xs[row, :8], 3).tolist())
match = [[int(np.argmin(np.abs(xs[row] - gi[row, j]))) if abs(gi[row, j]) > 1e-6 else -1 for j in range(c)]]
print('out col j holds x col:', match[0])
