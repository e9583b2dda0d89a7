"""Shared helpers for the parity tests."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from breaching_b200 import config as bcfg  # noqa: E402
from breaching_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def cfg_from_fixture(fx):
    return bcfg.get_attack_config(fx["attack"], dict(fx["overrides"]))


def case_from_fixture(fx):
    if "seq_len" in fx["case"]:
        model, loss_fn, payload, shared, true = synthetic.make_text_case(**fx["case"])
    elif "steps" in fx["case"]:
        model, loss_fn, payload, shared, true = synthetic.make_fedavg_case(**fx["case"])
    elif "queries" in fx["case"]:
        model, loss_fn, payload, shared, true = synthetic.make_multi_query_case(**fx["case"])
    else:
        model, loss_fn, payload, shared, true = synthetic.make_case(**fx["case"])
    checksum = float(sum(p.double().sum() for p in model.parameters()))
    assert abs(checksum - fx["weight_checksum"]) <= 1e-6 * max(1.0, abs(fx["weight_checksum"])), \
        "synthetic case differs from the one the fixture was generated with"
    return model, loss_fn, payload, shared, true


def oracle_for_fixture(fx):
    """TrialOracle (CPU restatement) set up exactly like the reference attacker was for this fixture."""
    from oracle import restate

    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    shared = copy.deepcopy(shared)
    m = copy.deepcopy(model)
    if shared[0]["buffers"] is not None:
        for buf, src in zip(m.buffers(), shared[0]["buffers"]):
            buf.data.copy_(src)
    m.eval()
    if shared[0]["buffers"] is None and payload[0]["buffers"] is None:  # base_attack.py:192-197: no buffers anywhere -> train mode
        m.train()
        for mod in m.modules():
            if hasattr(mod, "track_running_stats"):
                mod.track_running_stats = False
    meta = payload[0]["metadata"]
    dm = torch.tensor(meta.mean)[None, :, None, None]
    ds = torch.tensor(meta.std)[None, :, None, None]
    labels = restate.recover_labels(cfg.label_strategy, shared, shared[0]["metadata"]["num_data_points"])
    return restate.TrialOracle(m, loss_fn, cfg, shared[0]["gradients"], labels, dm, ds,
                               local_hyperparams=shared[0]["metadata"]["local_hyperparams"]), cfg, labels


TRIAL_FIXTURES = ["ig_convnet", "ig_resnet18", "stg_resnet18", "modern_convnet", "tag_clip_convnet", "l1_sgd_convnet"]
FEDAVG_FIXTURES = ["fedavg_convnet", "fedavg_resnet18"]
LBFGS_FIXTURES = ["lbfgs_convnet", "lbfgs_wei_convnet", "lbfgs_cosine_convnet"]
JOINT_FIXTURES = ["joint_dlg_convnet", "joint_adam_convnet", "joint_tag_transformer"]
MULTI_QUERY_FIXTURES = ["multiquery_convnet"]
TRAIN_BN_FIXTURES = ["trainbn_convnet", "trainbn_resnet18"]


def joint_oracle_for_fixture(fx):
    """JointTrialOracle (CPU restatement of OptimizationJointAttacker) for a joint-optimisation fixture."""
    from oracle import restate

    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    m = copy.deepcopy(model).eval()
    meta = payload[0]["metadata"]
    grads = list(shared[0]["gradients"])
    if getattr(meta, "modality", "vision") == "text":
        # base_attack.py:76-128 ("run-embedding"): optimise in embedding space -- drop the token-embedding gradient and bypass
        # the embedding layer; no input normalisation
        names = [n for n, _ in m.named_parameters()]
        grads.pop(names.index("encoder.weight"))
        m.encoder = torch.nn.Identity()
        dm, ds = torch.tensor(0.0), torch.tensor(1.0)
    else:
        dm = torch.tensor(meta.mean)[None, :, None, None]
        ds = torch.tensor(meta.std)[None, :, None, None]
    return restate.JointTrialOracle(m, loss_fn, cfg, grads, None, dm, ds), cfg


def multi_query_oracle_for_fixture(fx):
    """MultiQueryOracle: one TrialOracle per (model, update) pair, regularisers only on the first."""
    from oracle import restate

    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    no_priors = copy.deepcopy(cfg)
    if no_priors.get("regularization") is not None:
        for key in no_priors["regularization"].keys():
            no_priors["regularization"][key]["scale"] = 0.0
    meta = payload[0]["metadata"]
    dm = torch.tensor(meta.mean)[None, :, None, None]
    ds = torch.tensor(meta.std)[None, :, None, None]
    labels = restate.recover_labels(cfg.label_strategy, copy.deepcopy(shared), shared[0]["metadata"]["num_data_points"])
    oracles = []
    for i, (pl, sh) in enumerate(zip(payload, shared)):
        m = copy.deepcopy(model)
        with torch.no_grad():
            for p, src in zip(m.parameters(), pl["parameters"]):
                p.copy_(src)
            for b, src in zip(m.buffers(), pl["buffers"]):
                b.copy_(src)
        m.eval()
        oracles.append(restate.TrialOracle(m, loss_fn, cfg if i == 0 else no_priors, sh["gradients"], labels, dm, ds))
    return restate.MultiQueryOracle(oracles), cfg, labels
