"""One fit step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python tools/step_launches.py [--mode global|local] [--T 300] [--M 1000000]"""
import argparse, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="global")
ap.add_argument("--T", type=int, default=300)
ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--scene-order", default="kd")
args = ap.parse_args()
fpv = importlib.import_module("4dcapture-fpv_b200")
dev = torch.device("cuda:0")
prob = fpv.FitProblem(T=args.T, M=0 if args.mode == "local" else args.M, device=dev, seed=1235, front_end=True, mode=args.mode,
                      idx_dtype=torch.int32, scene_order=args.scene_order)
for _ in range(4):
    prob.step(update=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(args.steps):
    prob.step(update=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
