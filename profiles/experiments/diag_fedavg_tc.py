"""Localise the difference between the SIMT and tcgen05 back ends on the FedAvg ResNet-18 fixture."""
import copy
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import torch  # noqa: E402

from breaching_b200.engine import Engine  # noqa: E402
from helpers import case_from_fixture, cfg_from_fixture, load_golden  # noqa: E402

DEV = torch.device("cuda:0")
fx = load_golden("trial_fedavg_resnet18.pt")


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


def engine(backend, **opts):
    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    local = shared[0]["metadata"]["local_hyperparams"]
    meta = payload[0]["metadata"]
    shape = (local["data_per_step"], *fx["x0"].shape[1:])
    eng = Engine(copy.deepcopy(model).to(DEV).eval(), shape, cfg, DEV, backend=backend)
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.load_model()
    eng.load_targets([g.to(DEV) for g in shared[0]["gradients"]], local["labels"][0], mean=meta.mean, std=meta.std)
    eng.set_local_steps(fx["x0"].shape[0], local["steps"], local["lr"], local["labels"])
    return eng


ref = engine("simt")
v0, g0 = ref.objective_and_gradient(fx["x0"].to(DEV))
print("simt: val", v0, "fixture", fx["objective0"], "grad rel", rel(g0, fx["raw_grad0"]))
nparams = len(ref.prog.params)
for opts in [dict(), dict(overlap_wgrad=0), dict(pdl=0), dict(overlap_wgrad=0, pdl=0)]:
    eng = engine("tc", **opts)
    v1, g1 = eng.objective_and_gradient(fx["x0"].to(DEV))
    v2, g2 = eng.objective_and_gradient(fx["x0"].to(DEV))
    print("tc", opts, "val", v1, "grad rel vs fixture", rel(g1, fx["raw_grad0"]), "vs simt", rel(g1, g0), "repeat identical", torch.equal(g1, g2),
          "per step", [round(rel(g1[k], g0[k]), 4) for k in range(g1.shape[0])])
    if not opts:
        for which in ("G", "v"):
            worst = sorted(((rel(eng.debug_param(which, i), ref.debug_param(which, i)), i) for i in range(nparams)), reverse=True)[:8]
            print("  arena", which, "worst params (rel, index, shape):", [(round(r, 4), i, tuple(ref.prog.params[i].shape)) for r, i in worst])
    eng.close()
