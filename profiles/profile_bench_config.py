"""Drive a few steady-state iterations of one BASELINE configuration (bench.py's own workload builder) for ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cN.csv \
        python profiles/profile_bench_config.py --config N --iters 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--backend", default="tc")
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--timed", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
case = bench.build_case(args.config)
runner = bench.EngineRunner(args.config, case, dev, args.backend, 0)
runner.eng.set_option("use_graph", args.graph)
runner.warm(args.iters)
if args.timed:
    ms = runner.timed(args.timed)
    print(f"config {args.config}: {args.timed / ms * 1e3:.1f} it/s, {ms / args.timed:.3f} ms/iteration")
print("history", runner.eng.history().tolist()[-3:], "launches/iter", runner.eng.launches_per_iteration())
