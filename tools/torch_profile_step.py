"""Which ATen ops (and from where) still run inside one eager train step."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
wl = bench.ProductWorkload(0, 1, torch.device("cuda", 0))
for _ in range(2):
    wl.step(False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    wl.step(False)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=6).table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=40, max_src_column_width=90))
