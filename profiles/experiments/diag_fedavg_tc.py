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

# ---- what does the reference's own GPU path (cuDNN TF32 convolutions, torch default) do on this case? -------------------
from oracle import restate  # noqa: E402
from breaching_b200 import synthetic  # noqa: E402
from breaching_b200 import get_attack_config  # noqa: E402


def torch_oracle(tf32, fixture=fx):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    model, loss_fn, payload, shared, true = case_from_fixture(fixture)
    cfg = cfg_from_fixture(fixture)
    meta = payload[0]["metadata"]
    dm = torch.tensor(meta.mean, device=DEV)[None, :, None, None]
    ds = torch.tensor(meta.std, device=DEV)[None, :, None, None]
    local = copy.deepcopy(shared[0]["metadata"]["local_hyperparams"])
    local["labels"] = [l.to(DEV) for l in local["labels"]]
    orc = restate.TrialOracle(copy.deepcopy(model).to(DEV).eval(), loss_fn, cfg, [g.to(DEV) for g in shared[0]["gradients"]],
                              torch.cat(local["labels"]), dm, ds, local_hyperparams=local)
    phi, _, raw, terms = orc.closure_gradient(fixture["x0"].to(DEV), 0, 0.0)
    orc.close()
    return float(phi), raw


for tf32 in (False, True):
    phi, raw = torch_oracle(tf32)
    print(f"torch eager on the GPU, allow_tf32={tf32}: val {phi} grad rel vs fp32 CPU fixture {rel(raw, fx['raw_grad0'])}",
          "per step", [round(rel(raw[k], fx['raw_grad0'][k]), 4) for k in range(raw.shape[0])])

# ---- growth with the number of local steps / step size (engine simt vs tc) ---------------------------------------------
for steps, lr in [(1, 0.01), (2, 0.01), (4, 0.01), (4, 0.001), (4, 0.1)]:
    model, loss_fn, payload, shared, true = synthetic.make_fedavg_case("resnet18", "imagenet", num_data_points=steps, steps=steps,
                                                                       data_per_step=1, lr=lr, seed=6, bn_random=True, image_size=64, classes=10)
    cfg = get_attack_config("modern", {"regularization.features.scale": 0.0})
    local = shared[0]["metadata"]["local_hyperparams"]
    meta = payload[0]["metadata"]
    x0 = torch.randn(steps, 3, 64, 64, generator=torch.Generator().manual_seed(1)).to(DEV)
    res = {}
    for backend in ("simt", "tc"):
        eng = Engine(copy.deepcopy(model).to(DEV).eval(), (1, 3, 64, 64), cfg, DEV, backend=backend)
        eng.load_model()
        eng.load_targets([g.to(DEV) for g in shared[0]["gradients"]], local["labels"][0], mean=meta.mean, std=meta.std)
        eng.set_local_steps(steps, local["steps"], local["lr"], local["labels"])
        res[backend] = eng.objective_and_gradient(x0)
        eng.close()
    print(f"steps={steps} lr={lr}: val simt {res['simt'][0]:.6f} tc {res['tc'][0]:.6f}; grad rel tc vs simt {rel(res['tc'][1], res['simt'][1]):.4f}",
          "per step", [round(rel(res['tc'][1][k], res['simt'][1][k]), 4) for k in range(steps)])
