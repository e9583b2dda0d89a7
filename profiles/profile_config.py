"""Drive a few iterations of another BASELINE configuration for ncu.  usage: profile_config.py {config3|config4} [iters]"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from breaching_b200 import get_attack_config, synthetic  # noqa: E402
from breaching_b200.engine import Engine  # noqa: E402
from breaching_b200.schedule import lr_table  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "config3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
timed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
local = None
if which == "config3":
    model, loss_fn, payload, shared, true = synthetic.make_case("resnet50", "imagenet", batch=8, seed=17, user_buffers=True)
    cfg, shape = get_attack_config("seethroughgradients"), (8, 3, 224, 224)
    m = copy.deepcopy(model)
    for buf, src in zip(m.buffers(), shared[0]["buffers"]):
        buf.data.copy_(src)
    labels, total = true["labels"], 8
else:  # config 4: modern hyper-parameters, FedAvg 4 steps x 1 image, ResNet-18 224^2 (features / DI priors off)
    model, loss_fn, payload, shared, true = synthetic.make_fedavg_case("resnet18", "imagenet", num_data_points=4, steps=4,
                                                                       data_per_step=1, lr=1e-3, seed=233)
    cfg, shape, m = get_attack_config("modern", {"regularization.features.scale": 0.0}), (1, 3, 224, 224), model
    local = shared[0]["metadata"]["local_hyperparams"]
    labels, total = local["labels"][0], 4
meta = payload[0]["metadata"]
eng = Engine(copy.deepcopy(m).to(dev).eval(), shape, cfg, dev)
eng.set_option("use_graph", 1 if timed else 0)
eng.load_model()
eng.load_targets([g.to(dev) for g in shared[0]["gradients"]], labels.to(dev), mean=meta.mean, std=meta.std)
if local is not None:
    eng.set_local_steps(total, local["steps"], local["lr"], local["labels"])
opt = cfg.optim
eng.begin_trial(torch.randn(total, *shape[1:], device=dev), lr_table(opt.step_size, opt.step_size_decay, opt.warmup, opt.max_iterations))
eng.run(iters)
eng.sync()
if timed:
    ms = eng.run_timed(timed)
    print(f"{which}: {timed / ms * 1e3:.1f} it/s, {ms / timed:.3f} ms/iteration, {eng.launches_per_iteration()} launches/iteration")
print("history", eng.history().tolist()[-3:], "launches/iter", eng.launches_per_iteration())
