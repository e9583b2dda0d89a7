"""Steady-state iterations/s of the engine for other BASELINE configurations (single GPU, synthetic data).
usage: bench_config.py {config2|config3|config1} [steps] [backend]"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from breaching_b200 import get_attack_config, synthetic  # noqa: E402
from breaching_b200.engine import Engine  # noqa: E402
from breaching_b200.schedule import lr_table  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "config3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
backend = sys.argv[3] if len(sys.argv) > 3 else "tc"
dev = torch.device("cuda:0")
if which == "config3":
    model, loss_fn, payload, shared, true = synthetic.make_case("resnet50", "imagenet", batch=8, seed=17, user_buffers=True)
    cfg, shape, gflop = get_attack_config("seethroughgradients"), (8, 3, 224, 224), 454.1
    m = copy.deepcopy(model)
    for buf, src in zip(m.buffers(), shared[0]["buffers"]):
        buf.data.copy_(src)
elif which == "config1":
    model, loss_fn, payload, shared, true = synthetic.make_case("convnet", "cifar", batch=1, seed=233)
    cfg, shape, gflop, m = get_attack_config("invertinggradients"), (1, 3, 32, 32), 2 * 7 * 1.86, model
else:
    model, loss_fn, payload, shared, true = synthetic.make_case("resnet18", "imagenet", batch=1, seed=233)
    cfg, shape, gflop, m = get_attack_config("invertinggradients"), (1, 3, 224, 224), 24.92, model
meta = payload[0]["metadata"]
eng = Engine(copy.deepcopy(m).to(dev).eval(), shape, cfg, dev, backend=backend)
eng.load_model()
eng.load_targets([g.to(dev) for g in shared[0]["gradients"]], true["labels"].to(dev), mean=meta.mean, std=meta.std)
opt = cfg.optim
eng.begin_trial(torch.randn(*shape, device=dev), lr_table(opt.step_size, opt.step_size_decay, opt.warmup, opt.max_iterations))
eng.run(10)
eng.sync()
ms = eng.run_timed(steps)
print(f"{which} backend {backend}: {steps / ms * 1e3:.1f} it/s, {ms / steps:.3f} ms/iteration, {gflop / (ms / steps) :.1f} TFLOP/s algorithmic, "
      f"{eng.launches_per_iteration()} launches/iteration, objective {eng.history()[-1].item():.5f}")
