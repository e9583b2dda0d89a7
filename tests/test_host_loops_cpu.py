"""The host-driven loops of the attackers (joint data + label optimisation incl. the text / TAG path, L-BFGS, multi-query)
verified on the CPU: the engine is replaced by a stand-in that evaluates the closure with the oracle (autograd), so what is
tested is exactly the product's Python loop code -- optimiser updates, gradient post-processing, softmax chain of the label
leaf, row reshapes of token models, best-so-far, history -- against the trajectories of the unmodified reference."""
import math

import pytest
import torch

from breaching_b200.attacks import lbfgs
from breaching_b200.attacks.joint_attack import OptimizationJointAttacker
from breaching_b200.schedule import lr_table
from helpers import (JOINT_FIXTURES, LBFGS_FIXTURES, MULTI_QUERY_FIXTURES, cfg_from_fixture, joint_oracle_for_fixture, load_golden,
                     multi_query_oracle_for_fixture, oracle_for_fixture)


class _OracleEngine:
    """The calls the host loops make on ``breaching_b200.engine.Engine``, answered by the CPU oracle."""

    def __init__(self, oracle, joint=False):
        self.oracle, self.joint, self.q, self.gq = oracle, joint, None, None

    def load_soft_labels(self, q):
        self.q = None if q is None else q.detach().clone()

    def objective_and_gradient(self, x):
        from torch.nn.attention import SDPBackend, sdpa_kernel

        if not self.joint:
            phi, _, raw, _ = self.oracle.closure_gradient(x.reshape(self.shape), 0, 0.0)
            return float(phi), raw.reshape(x.shape)
        xs = x.detach().clone().reshape(self.shape).requires_grad_(True)
        q = self.q.reshape(self.label_shape).clone().requires_grad_(True)
        self.oracle.labels = q
        with sdpa_kernel(SDPBackend.MATH):
            total, _ = self.oracle.objective_terms(xs)
            gx, gq = torch.autograd.grad(total, [xs, q])
        self.gq = gq
        return float(total), gx.reshape(x.shape)

    def label_gradient(self, shape):
        return self.gq.reshape(shape)


def _bare_joint_attacker(cfg, dm, ds):
    att = object.__new__(OptimizationJointAttacker)   # the constructor insists on a CUDA device; the loop code does not need one
    att.cfg, att.dm, att.ds = cfg, dm, ds
    att.setup = dict(device=torch.device("cpu"), dtype=torch.float)
    return att


@pytest.mark.parametrize("name", JOINT_FIXTURES)
def test_joint_loop_reproduces_reference_trajectory(name):
    fx = load_golden(f"trial_{name}.pt")
    orc, cfg = joint_oracle_for_fixture(fx)
    eng = _OracleEngine(orc, joint=True)
    eng.shape, eng.label_shape = tuple(fx["x0"].shape), tuple(fx["l0"].shape)
    att = _bare_joint_attacker(cfg, orc.dm, orc.ds)
    from collections import defaultdict

    stats = defaultdict(list)
    best, best_l = att._run_joint_trial(eng, fx["x0"], fx["l0"], stats, 0, iterations=fx["iters"])
    is_lbfgs = str(cfg.optim.optimizer).lower() == "l-bfgs"
    tol, atol = (5e-2, 1e-4 * abs(fx["history"][0])) if is_lbfgs else (5e-4, 1e-6)
    assert len(stats["Trial_0_Val"]) == len(fx["history"])
    for a, b in zip(stats["Trial_0_Val"], fx["history"]):
        assert math.isclose(a, b, rel_tol=tol, abs_tol=atol), (stats["Trial_0_Val"], fx["history"])
    if not is_lbfgs:
        x_final, l_final = att._last_joint_state
        assert (x_final - fx["candidate_final"]).abs().mean().item() < 2e-3
        assert (l_final - fx["labels_final"]).abs().mean().item() < 2e-3
    orc.close()


@pytest.mark.parametrize("name", LBFGS_FIXTURES + MULTI_QUERY_FIXTURES)
def test_single_leaf_host_loop_reproduces_reference_trajectory(name):
    """attacks/lbfgs.run_trial: the L-BFGS trials and the multi-query loop (first-order optimisers on the host)."""
    fx = load_golden(f"trial_{name}.pt")
    if name in MULTI_QUERY_FIXTURES:
        orc, cfg, _ = multi_query_oracle_for_fixture(fx)
    else:
        orc, cfg, _ = oracle_for_fixture(fx)
    eng = _OracleEngine(orc)
    eng.shape = tuple(fx["x0"].shape)
    opt = cfg.optim
    table = lr_table(opt.step_size, opt.step_size_decay, opt.warmup, int(opt.max_iterations))
    lo, hi = -orc.dm / orc.ds, (1 - orc.dm) / orc.ds
    best, hist = lbfgs.run_trial(eng, fx["x0"], cfg, table, lo, hi, iterations=fx["iters"])
    tol = 3e-2 if name in LBFGS_FIXTURES else 5e-4
    for a, b in zip(hist, fx["history"]):
        assert math.isclose(a, b, rel_tol=tol, abs_tol=1e-6), (hist, fx["history"])
    assert (best - fx["best"]).abs().mean().item() < (3e-2 if name in LBFGS_FIXTURES else 2e-3)
    orc.close()
