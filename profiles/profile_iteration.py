"""Drive a few steady-state iterations of BASELINE config 2 for ncu (launch list / full capture).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_iteration.py --iters 3
"""
import argparse
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from breaching_b200 import get_attack_config, synthetic  # noqa: E402
from breaching_b200.engine import Engine  # noqa: E402
from breaching_b200.schedule import lr_table  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--backend", default="tc")
ap.add_argument("--graph", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda:0")
model, loss_fn, payload, shared, true = synthetic.make_case("resnet18", "imagenet", batch=1, seed=233)
cfg = get_attack_config("invertinggradients")
meta = payload[0]["metadata"]
eng = Engine(copy.deepcopy(model).to(dev).eval(), (1, 3, 224, 224), cfg, dev, backend=args.backend)
eng.set_option("use_graph", args.graph)
eng.load_model()
eng.load_targets([g.to(dev) for g in shared[0]["gradients"]], true["labels"].to(dev), mean=meta.mean, std=meta.std)
eng.begin_trial(torch.randn(1, 3, 224, 224, device=dev), lr_table(0.1, "step-lr", 0, 24000, 64))
eng.run(args.iters)
eng.sync()
print("history", eng.history().tolist(), "launches/iter", eng.launches_per_iteration())
