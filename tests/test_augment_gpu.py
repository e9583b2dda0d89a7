"""Candidate augmentations on the engine (reference attacks/auxiliaries/augmentations.py wired into the closure at
optimization_based_attack.py:149-153): the device view pipeline and its transposed pull-back against the reference modules restated
with explicit draws (oracle.restate.augment_candidate), stand-alone and inside a closure evaluation."""
import copy
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from breaching_b200 import engine as E  # noqa: E402
from breaching_b200 import get_attack_config, synthetic  # noqa: E402
from breaching_b200.attacks import augment  # noqa: E402
from oracle import restate  # noqa: E402

DEV = torch.device("cuda:0")


def _relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


CASES = [
    dict(steps=[(1, 4)], offsets=[(3, -2)]),
    dict(steps=[(2, 0.5)], offsets=[(1, 0)]),
    dict(steps=[(1, 8), (2, 0.5), (1, 3)], offsets=[(-5, 7), (1, 0), (2, 2)]),
    dict(steps=[], offsets=[], continuous_shift=6.0, circular=True, uniforms=([0.13, 0.81], [0.66, 0.05])),
    dict(steps=[], offsets=[], continuous_shift=5.0, circular=False, uniforms=([0.9, 0.2], [0.4, 0.75])),
    dict(steps=[(1, 4), (2, 0.5)], offsets=[(2, -1), (1, 0)], continuous_shift=20.0, circular=True, uniforms=([0.31, 0.5], [0.97, 0.02]), colour=True),
    dict(steps=[], offsets=[], colour=True),
]


@pytest.mark.parametrize("case", CASES)
def test_view_and_pullback_equal_the_reference_modules(case):
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 18, 18, generator=gen)
    g = torch.randn(2, 3, 18, 18, generator=gen)
    mean = std = scale = shift = None
    if case.get("colour"):
        mean = (torch.rand(2, 3, 1, 1, generator=gen) - 0.5) * 0.4
        std = ((torch.rand(2, 3, 1, 1, generator=gen) - 0.5) * 0.6).exp()
        scale, shift = (1 / std).view(2, 3), (-mean / std).view(2, 3)
    kw = dict(steps=case["steps"], offsets=case["offsets"], continuous_shift=case.get("continuous_shift"), circular=case.get("circular", True),
              uniforms=case.get("uniforms"))
    xr = x.clone().double().requires_grad_(True)
    want = restate.augment_candidate(xr, colour_mean=None if mean is None else mean.double(), colour_std=None if std is None else std.double(), **kw)
    (gwant,) = torch.autograd.grad((want * g.double()).sum(), xr)
    got = E.augment_view(x.to(DEV), colour_scale=scale, colour_shift=shift, **kw)
    assert (got.cpu().double() - want.detach()).abs().max().item() < 2e-5, (got.cpu().double() - want.detach()).abs().max().item()
    pulled = E.augment_view(g.to(DEV), colour_scale=scale, colour_shift=shift, transpose=True, **kw)
    assert _relerr(pulled, gwant) < 1e-5, _relerr(pulled, gwant)


@pytest.mark.parametrize("differentiable", [True, False])
def test_closure_with_augmentations_matches_the_oracle(differentiable):
    """One closure evaluation with the augmentation block of a config (discrete shift, flip, continuous shift, colour jitter): the
    engine draws on the device; the oracle applies the reference modules with the same draws (read back) and differentiates through
    them with autograd.  Non-differentiable mode: the candidate itself is replaced by its view (the reference assigns candidate.data)."""
    model, loss_fn, payload, shared, true = synthetic.make_case("convnet-tiny", "cifar", batch=2, seed=8, bn_random=True)
    over = {"augmentations": {"discrete_shift": {"lim": 5}, "flip": {"p": 0.5}, "continuous_shift": {"shift": 6, "padding": "circular"},
                              "colorjitter": {"mean": 0.1, "std": 0.3}}, "differentiable_augmentations": differentiable,
            "objective.task_regularization": 0.2}
    cfg = get_attack_config("invertinggradients", over)
    meta = payload[0]["metadata"]
    setup = dict(device=DEV, dtype=torch.float)
    torch.manual_seed(5)
    plan = augment.build_plan(cfg, 2, 3, setup)
    assert plan.steps == [(1, 5.0), (2, 0.5)] and plan.continuous_shift == 6.0 and plan.circular and plan.differentiable == differentiable
    eng = E.Engine(copy.deepcopy(model).to(DEV).eval(), (2, 3, 32, 32), cfg, DEV, backend="simt")
    eng.load_model()
    eng.load_targets([g.to(DEV) for g in shared[0]["gradients"]], true["labels"].to(DEV), mean=meta.mean, std=meta.std)
    eng.set_augmentations(plan)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    val, grad = eng.objective_and_gradient(x.to(DEV))
    o1, o2, sx, sy = eng.last_augmentation()
    offsets = [(o1[0], o2[0]), (o1[1], 0)]
    assert -5 <= o1[0] < 5 and -5 <= o2[0] < 5 and o1[1] in (0, 1) and all(0 <= u < 1 for u in sx + sy)
    std = (1 / plan.colour_scale).cpu().view(2, 3, 1, 1)
    mean = (-plan.colour_shift / plan.colour_scale).cpu().view(2, 3, 1, 1)
    dm, ds = torch.tensor(meta.mean)[None, :, None, None], torch.tensor(meta.std)[None, :, None, None]
    orc = restate.TrialOracle(model.eval(), loss_fn, cfg, shared[0]["gradients"], true["labels"], dm, ds)
    xr = x.clone().requires_grad_(True)
    xa = restate.augment_candidate(xr, steps=plan.steps, offsets=offsets, continuous_shift=6.0, circular=True, uniforms=(sx, sy), colour_mean=mean,
                                   colour_std=std)
    if differentiable:
        total, terms = orc.objective_terms(xa)
        (gref,) = torch.autograd.grad(total, xr)
    else:
        xa = xa.detach().requires_grad_(True)
        total, terms = orc.objective_terms(xa)
        (gref,) = torch.autograd.grad(total, xa)
        assert (eng.candidate().cpu() - xa.detach()).abs().max().item() < 2e-5     # the candidate became its view
    assert math.isclose(val, float(total), rel_tol=2e-4), (val, float(total), terms, eng.last_terms())
    assert _relerr(grad, gref) < 2e-3, _relerr(grad, gref)
    # a second evaluation at a later iteration draws again (Philox keyed by the iteration counter)
    from breaching_b200.schedule import lr_table

    eng.begin_trial(x.to(DEV), lr_table(0.1, "step-lr", 0, 100, 8))
    eng.run(3)
    eng.sync()
    assert len(eng.history()) == 3 and all(math.isfinite(v) for v in eng.history().tolist())
    eng.close()
    orc.close()


def test_multiscale_preset_with_its_augmentations_runs():
    """multiscale_ghiasi.yaml as shipped (continuous_shift 224 circular + colorjitter, differentiable): two short stages."""
    from breaching_b200.attacks import prepare_attack

    model, loss_fn, payload, shared, true = synthetic.make_case("resnet18", "imagenet", batch=1, seed=4, bn_random=True, image_size=64, classes=10)
    cfg = get_attack_config("multiscale_ghiasi", {"num_stages": 2, "scale_pyramid": "log", "optim.max_iterations": 5, "optim.callback": 5})
    torch.manual_seed(1)
    rec, stats = prepare_attack(model, loss_fn, cfg, dict(device=DEV, dtype=torch.float)).reconstruct(payload, copy.deepcopy(shared), {})
    assert rec["data"].shape == (1, 3, 64, 64) and torch.isfinite(rec["data"]).all() and len(stats["Trial_0_Val"]) == 10
    with pytest.raises(NotImplementedError):
        prepare_attack(model, loss_fn, get_attack_config("invertinggradients", {"augmentations": {"median": {}}, "optim.max_iterations": 2}),
                       dict(device=DEV, dtype=torch.float)).reconstruct(payload, copy.deepcopy(shared), {})
