"""Kernel table (torch.profiler) of the detection backbone's inference forward (config 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scripts.train_profile import det_inputs
from link_b200.scn import SpMiddleResNetFHDELKv3
dev = torch.device('cuda:0')
torch.manual_seed(0)
feats, idx = det_inputs(dev)
net = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8).to(dev).eval()
def fwd():
    with torch.no_grad():
        return net(feats, idx, 1, [1440, 1440, 40])[0]
for _ in range(3): fwd()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): fwd()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=28, max_name_column_width=60))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): fwd()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
