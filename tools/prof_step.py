"""Profiling driver: full train steps of the bench workload (see bench.py), for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/prof_step.py 1 2"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.cuda.set_device(0)
torch.backends.cudnn.benchmark = True
wl = bench.ProductWorkload(0, 1, torch.device("cuda", 0))
if os.environ.get("AGR_GRAPH", "0") == "1":
    wl.capture()
for _ in range(warm):
    wl.step(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    loss = wl.step(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
from animatablegaussians_b200 import rasterizer
print("loss", float(loss), "capacity hints", rasterizer._capacity_hint)
