"""Which term limits fp32 agreement on config 3 (ResNet-50, batch 8, 224^2)?  engine (SIMT fp32) vs CPU oracle in fp32 and fp64."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from breaching_b200 import get_attack_config, synthetic
from breaching_b200.engine import Engine
from oracle import restate

dev = torch.device("cuda:0")
model, loss_fn, payload, shared, true = synthetic.make_case("resnet50", "imagenet", batch=8, seed=17, user_buffers=True)
meta = payload[0]["metadata"]
m = copy.deepcopy(model)
for buf, src in zip(m.buffers(), shared[0]["buffers"]):
    buf.data.copy_(src)
m.eval()
labels = true["labels"]
x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(2))
dm, ds = torch.tensor(meta.mean)[None, :, None, None], torch.tensor(meta.std)[None, :, None, None]
variants = {
    "euclid only": {"regularization": None},
    "euclid+tv+norm": {"regularization.deep_inversion.scale": 0.0},
    "DI only (euclid scale 1e-12)": {"objective.scale": 1e-12, "regularization.total_variation.scale": 0.0, "regularization.norm.scale": 0.0},
    "full": {},
}
for name, ov in variants.items():
    cfg = get_attack_config("seethroughgradients", ov)
    eng = Engine(copy.deepcopy(m).to(dev).eval(), (8, 3, 224, 224), cfg, dev, backend="simt")
    eng.load_model()
    eng.load_targets([g.to(dev) for g in shared[0]["gradients"]], labels.to(dev), mean=meta.mean, std=meta.std)
    val, grad = eng.objective_and_gradient(x.to(dev))
    grad = grad.cpu().double()
    eng.close()
    o32 = restate.TrialOracle(copy.deepcopy(m), loss_fn, cfg, shared[0]["gradients"], labels, dm, ds)
    phi32, _, raw32, _ = o32.closure_gradient(x, 0, 0.0)
    o32.close()
    m64 = copy.deepcopy(m).double()
    o64 = restate.TrialOracle(m64, loss_fn, cfg, [g.double() for g in shared[0]["gradients"]], labels, dm.double(), ds.double(), dtype=torch.double)
    phi64, _, raw64, _ = o64.closure_gradient(x.double(), 0, 0.0)
    o64.close()
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    print(f"{name:32s} phi eng {val:.6e} cpu32 {float(phi32):.6e} cpu64 {float(phi64):.6e} | grad rel: eng-vs-64 {rel(grad, raw64):.2e}  "
          f"cpu32-vs-64 {rel(raw32, raw64):.2e}  eng-vs-cpu32 {rel(grad, raw32):.2e}  |grad| {raw64.norm().item():.3e}", flush=True)
