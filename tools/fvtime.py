import sys, os, torch
sys.path.insert(0, '/root/repo')
from dpdist_b200 import dpdist_util
g = torch.Generator().manual_seed(5)
for n in (2048, 16384):
    big = (torch.rand((n, 64, 3), generator=g) * 1.6 - 0.8).cuda()
    for _ in range(3): dpdist_util.get_3dmfv_tf(big, n_gaussians=512, sigma=0.125, flatten=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dpdist_util.get_3dmfv_tf(big, n_gaussians=512, sigma=0.125, flatten=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("stagger", os.environ.get("DPD_FV_STAGGER"), "clouds", n, "us", ms * 1e3, "GB/s", n * 41728 / ms / 1e6)
