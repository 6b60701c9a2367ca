import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0"); stream = torch.cuda.Stream()
which = sys.argv[1:] or ["walkers", "windows"]
for w in which:
    if w == "walkers":
        print(bench.run_walkers(dev, stream, 0, 1, None)); torch.cuda.synchronize(); print("walkers ok")
    if w == "windows":
        print(bench.run_windows(dev, stream, 0, 1, None)); torch.cuda.synchronize(); print("windows ok")
    if w == "weak":
        print(bench.run_windows(dev, stream, 0, 1, None, per_gpu=True)); torch.cuda.synchronize(); print("weak ok")
