"""Config 3 (see-through-gradients, ResNet-50, batch 8): how far is d(objective)/dx from a float64 evaluation when the first / last
K conv layers of the tensor-core back end run in fp32 (engine option precise_first / precise_last), and what does it cost?
Prints one line per setting: rel. l2 error vs float64, objective, ms per iteration."""
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from breaching_b200 import get_attack_config, synthetic  # noqa: E402
from breaching_b200.engine import Engine  # noqa: E402
from breaching_b200.schedule import lr_table  # noqa: E402
from oracle import restate  # noqa: E402

DEV = torch.device("cuda:0")


def relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


model, loss_fn, payload, shared, true = synthetic.make_case("resnet50", "imagenet", batch=8, seed=17, user_buffers=True)
cfg = get_attack_config("seethroughgradients")
meta = payload[0]["metadata"]
dm, ds = torch.tensor(meta.mean)[None, :, None, None], torch.tensor(meta.std)[None, :, None, None]
labels = restate.recover_labels(cfg.label_strategy, shared, 8)
m = copy.deepcopy(model)
for buf, src in zip(m.buffers(), shared[0]["buffers"]):
    buf.data.copy_(src)
m.eval()
x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(2))
t0 = time.time()
o64 = restate.TrialOracle(copy.deepcopy(m).double(), loss_fn, cfg, [g.double() for g in shared[0]["gradients"]], labels, dm.double(), ds.double(),
                          dtype=torch.double)
phi64, _, raw64, _ = o64.closure_gradient(x.double(), 0, 0.0)
print(f"float64 oracle: objective {float(phi64):.6f} ({time.time() - t0:.0f} s)", flush=True)
settings = [("simt", 0, 0), ("tc", 0, 0), ("tc", 1, 0), ("tc", 4, 0), ("tc", 11, 0), ("tc", 24, 0), ("tc", 0, 4), ("tc", 0, 11), ("tc", 0, 24),
            ("tc", 11, 11), ("tc", 54, 0)]
if len(sys.argv) > 1:
    settings = [tuple([s.split(",")[0], int(s.split(",")[1]), int(s.split(",")[2])]) for s in sys.argv[1:]]
opt = cfg.optim
table = lr_table(opt.step_size, opt.step_size_decay, opt.warmup, opt.max_iterations)
for backend, first, last in settings:
    eng = Engine(copy.deepcopy(m).to(DEV).eval(), (8, 3, 224, 224), cfg, DEV, backend=backend)
    eng.set_option("precise_first", first)
    eng.set_option("precise_last", last)
    eng.load_model()
    eng.load_targets([g.to(DEV) for g in shared[0]["gradients"]], labels.to(DEV), mean=meta.mean, std=meta.std)
    val, grad = eng.objective_and_gradient(x.to(DEV))
    eng.begin_trial(x.to(DEV), table)
    eng.run(5)
    eng.sync()
    ms = eng.run_timed(20) / 20
    print(f"{backend} precise_first={first} precise_last={last}: rel-l2 vs float64 {relerr(grad, raw64):.3e}, objective {val:.6f}, {ms:.3f} ms/it", flush=True)
    eng.close()
