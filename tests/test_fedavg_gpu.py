"""FedAvg / multi-step local updates (reference ``GradientLoss._grad_fn_multi_step``, objectives.py:48-72) on the engine:
K forward/backward passes at W_0..W_{K-1}, matching of W_K - W_0, adjoint carried back over the steps with
Hessian-vector products (dual-source tangent wgrads) -- against the fixtures produced by the unmodified reference."""
import copy
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from breaching_b200.attacks import prepare_attack  # noqa: E402
from breaching_b200.engine import Engine  # noqa: E402
from breaching_b200.schedule import lr_table  # noqa: E402
from helpers import FEDAVG_FIXTURES, case_from_fixture, cfg_from_fixture, load_golden, oracle_for_fixture  # noqa: E402

DEV = torch.device("cuda:0")


def _relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


def _engine(fx, backend):
    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    local = shared[0]["metadata"]["local_hyperparams"]
    meta = payload[0]["metadata"]
    shape = (local["data_per_step"], *fx["x0"].shape[1:])
    eng = Engine(copy.deepcopy(model).to(DEV).eval(), shape, cfg, DEV, backend=backend)
    eng.load_model()
    eng.load_targets([g.to(DEV) for g in shared[0]["gradients"]], local["labels"][0], mean=meta.mean, std=meta.std)
    eng.set_local_steps(fx["x0"].shape[0], local["steps"], local["lr"], local["labels"])
    return eng, cfg


def _reference_tf32_deviation(fx, with_score=False):
    """rel. l2 distance between the reference algorithm run in eager PyTorch on the GPU with TF32 convolutions (the
    reference's default GPU numerics) and the fp32 fixture (optionally also its score of the fixture's best candidate)."""
    from oracle import restate

    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    meta = payload[0]["metadata"]
    dm = torch.tensor(meta.mean, device=DEV)[None, :, None, None]
    ds = torch.tensor(meta.std, device=DEV)[None, :, None, None]
    local = copy.deepcopy(shared[0]["metadata"]["local_hyperparams"])
    local["labels"] = [lab.to(DEV) for lab in local["labels"]]
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        orc = restate.TrialOracle(copy.deepcopy(model).to(DEV).eval(), loss_fn, cfg, [g.to(DEV) for g in shared[0]["gradients"]],
                                  torch.cat(local["labels"]), dm, ds, local_hyperparams=local)
        _, _, raw, _ = orc.closure_gradient(fx["x0"].to(DEV), 0, 0.0)
        score = orc.score(fx["best"].to(DEV), fx["scoring"]) if with_score else None
        orc.close()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return (_relerr(raw, fx["raw_grad0"]), score) if with_score else _relerr(raw, fx["raw_grad0"])


@pytest.mark.parametrize("backend", ["simt", "tc"])
@pytest.mark.parametrize("name", FEDAVG_FIXTURES)
def test_fedavg_closure_matches_reference_fixture(name, backend):
    fx = load_golden(f"trial_{name}.pt")
    eng, cfg = _engine(fx, backend)
    val, grad = eng.objective_and_gradient(fx["x0"].to(DEV))
    # the matched quantity W_K - W_0 is a difference of nearly equal fp32 vectors in the reference (the engine accumulates it
    # directly); that cancellation noise, amplified by the cosine objective, bounds the agreement with the fp32 fixture
    tol_v, tol_g = (2e-3, 1e-2) if backend == "simt" else (2e-2, 5e-2)
    assert math.isclose(val, fx["objective0"], rel_tol=tol_v, abs_tol=1e-6), (val, fx["objective0"], eng.last_terms())
    # loss of the last local step (evaluated at W_{K-1}, i.e. after K-1 TF32 / fp32 updates)
    assert math.isclose(eng.last_terms()["task_loss"], fx["task_loss0"], rel_tol=1e-3 if backend == "simt" else 1e-2)
    rel = _relerr(grad, fx["raw_grad0"])
    if backend == "tc":
        # Hessian-vector products through K local steps are badly conditioned under TF32: the reference's own GPU path
        # (eager PyTorch, cuDNN TF32 convolutions = torch's default) is 27 % away from its fp32 CPU result on the
        # ResNet-18 fixture, step by step in the same pattern as the TF32 engine (profiles/experiments/diag_fedavg_tc.py).
        # The TF32 back end is therefore held to the reference's TF32 deviation, the fp32 back end to the fp32 fixture.
        ref_dev, ref_score = _reference_tf32_deviation(fx, with_score=True)
        tol_g = max(tol_g, 1.5 * ref_dev)
    assert rel < tol_g, rel
    score = eng.score(fx["best"].to(DEV), fx["scoring"])
    tol_s = 2e-2 * abs(fx["score"]) if backend == "simt" else max(5e-2 * abs(fx["score"]), 2.5 * abs(ref_score - fx["score"]))
    # (the score of a converged candidate is a small difference of two nearly equal updates: under TF32 the reference's own GPU run moves it
    # by ten per cent on the narrow ConvNet fixture (measured: reference TF32 -8.9 %, engine -15.6 % -- its 32-channel layers run on the
    # 128 x 32 tensor-core tiles, cuDNN keeps some of them in fp32); the bound is 2.5x the reference's own TF32 shift)
    assert abs(score - fx["score"]) <= tol_s + 1e-5, (score, fx["score"], tol_s)
    eng.close()


@pytest.mark.parametrize("name", FEDAVG_FIXTURES)
def test_fedavg_trajectory_matches_reference_fixture(name):
    fx = load_golden(f"trial_{name}.pt")
    eng, cfg = _engine(fx, "simt")
    opt = cfg.optim
    eng.begin_trial(fx["x0"].to(DEV), lr_table(opt.step_size, opt.step_size_decay, opt.warmup, opt.max_iterations))
    eng.run(fx["iters"])
    eng.sync()
    hist = eng.history().tolist()
    assert len(hist) == fx["iters"]
    for a, b in zip(hist, fx["history"]):
        assert math.isclose(a, b, rel_tol=2e-3, abs_tol=1e-5), (hist, fx["history"])
    assert (eng.candidate().cpu() - fx["candidate_final"]).abs().mean().item() < 5e-3  # soft sign: smooth trajectory
    eng.close()


def test_fedavg_through_the_attacker_api():
    """BASELINE config 4 shape of the call: modern hyper-parameters, FedAvg user with shared local hyper-parameters."""
    from breaching_b200 import get_attack_config, synthetic

    model, loss_fn, payload, shared, true = synthetic.make_fedavg_case("resnet18", "imagenet", num_data_points=4, steps=4,
                                                                       data_per_step=1, lr=1e-3, seed=3, image_size=64, classes=10)
    cfg = get_attack_config("modern", {"regularization.features.scale": 0.0, "optim.max_iterations": 12, "optim.callback": 6,
                                       "optim.warmup": 2})
    attacker = prepare_attack(model, loss_fn, cfg, dict(device=DEV, dtype=torch.float))
    rec, stats = attacker.reconstruct(payload, copy.deepcopy(shared), {}, dryrun=False)
    assert rec["data"].shape == (4, 3, 64, 64) and len(stats["Trial_0_Val"]) == 12
    assert math.isfinite(stats["opt_value"]) and torch.isfinite(rec["data"]).all()
    cfg_bad = get_attack_config("modern", {"optim.max_iterations": 4})  # features prior + FedAvg: the reference crashes, we refuse
    with pytest.raises(Exception):
        prepare_attack(model, loss_fn, cfg_bad, dict(device=DEV, dtype=torch.float)).reconstruct(payload, copy.deepcopy(shared), {})
