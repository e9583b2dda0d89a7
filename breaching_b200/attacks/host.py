"""Host-side prologue of an optimisation attack (runs once per ``reconstruct`` call; stays PyTorch).

Counterpart of the reference's ``_BaseAttacker`` (``attacks/base_attack.py``): preprocessing constants, rebuilding
the attacked model from the server payload, casting the shared update, label recovery, candidate initialisation
and gradient normalisation.  None of this is on the per-iteration hot path.
"""
import copy
import logging
import math

import torch

log = logging.getLogger(__name__)


def preprocessing_constants(metadata, setup):
    """base_attack.py:51-57."""
    if hasattr(metadata, "mean"):
        dm = torch.as_tensor(metadata.mean, **setup)[None, :, None, None]
        ds = torch.as_tensor(metadata.std, **setup)[None, :, None, None]
    else:
        dm, ds = torch.tensor(0, **setup), torch.tensor(1, **setup)
    return dm, ds


def construct_models(model_template, server_payload, shared_data, setup):
    """base_attack.py:169-212 (without the JIT options, which only wrap the eager model)."""
    models = []
    for idx, payload in enumerate(server_payload):
        new_model = copy.deepcopy(model_template)
        new_model.to(**setup)
        parameters = payload["parameters"]
        if shared_data[idx]["buffers"] is not None:
            buffers = shared_data[idx]["buffers"]
            new_model.eval()
        elif payload["buffers"] is not None:
            buffers = payload["buffers"]
            new_model.eval()
        else:
            new_model.train()
            for module in new_model.modules():
                if hasattr(module, "track_running_stats"):
                    module.reset_parameters()
                    module.track_running_stats = False
            buffers = []
        with torch.no_grad():
            for param, server_state in zip(new_model.parameters(), parameters):
                param.copy_(server_state.to(**setup))
            for buffer, server_state in zip(new_model.buffers(), buffers):
                buffer.copy_(server_state.to(**setup))
        models.append(new_model)
    return models


def cast_shared_data(shared_data, dtype):
    """base_attack.py:214-220 (mutates the inner dicts exactly like the reference)."""
    for data in shared_data:
        data["gradients"] = [g.to(dtype=dtype) for g in data["gradients"]]
        if data["buffers"] is not None:
            data["buffers"] = [b.to(dtype=dtype) for b in data["buffers"]]
    return shared_data


def normalize_gradients(shared_data, fudge_factor=1e-6):
    """base_attack.py:298-303."""
    for data in shared_data:
        grad_norm = torch.stack([g.pow(2).sum() for g in data["gradients"]]).sum().sqrt()
        torch._foreach_div_(data["gradients"], max(grad_norm, fudge_factor))
    return shared_data


# ---- label recovery (base_attack.py:305-475) ----------------------------------------------------------------------------
# One small function per strategy; every one maps the last-layer gradients of all queries to a (possibly short) label vector.
# Integer logic on the reference's own arithmetic (same reductions in the same order): results are bit-exact
# (tests/golden/labels.pt, 25 synthetic cases x 5 strategies).
def _head_weight_grads(user_data):
    return [d["gradients"][-2] for d in user_data]


def _labels_idlg(user_data, n, classes, device):          # :319-327 -- most negative row sum of the head weight gradient
    per_query = [g.sum(dim=-1).argmin(dim=-1).detach() for g in _head_weight_grads(user_data)]
    return torch.stack(per_query).unique()


def _labels_analytic(user_data, n, classes, device):      # :328-335 -- classes whose bias gradient is negative
    per_query = [(d["gradients"][-1] < 0).nonzero() for d in user_data]
    return torch.stack(per_query).unique()[:n]


def _labels_yin(user_data, n, classes, device):           # :336-345 -- smallest row minima, summed over the queries
    score = 0
    for g in _head_weight_grads(user_data):
        score = score + g.min(dim=-1)[0]
    return score.argsort()[:n]


def _labels_wainakh_simple(user_data, n, classes, device):   # :346-407
    queries = len(user_data)
    impact = 0
    for g in _head_weight_grads(user_data):
        row_sums = g.sum(dim=1)
        negative_mass = torch.where(row_sums < 0, row_sums, torch.zeros_like(row_sums)).sum()
        impact = impact + negative_mass * (1 + 1 / classes) / n / queries
    residual = torch.stack([g.sum(dim=1) for g in _head_weight_grads(user_data)]).mean(dim=0)
    found = []
    cursor = 0
    for cursor in range(classes):                          # stage 1: every class with negative mass, once
        if residual[cursor] < 0:
            found.append(torch.as_tensor(cursor, device=device))
            residual[cursor] -= impact
    while len(found) < n:                                  # stage 2: the reference keeps decrementing the *last stage-1 cursor*
        found.append(torch.as_tensor(residual.argmin(), device=device))   # (:405, reproduced), so argmin never moves
        residual[cursor] -= impact
    return torch.stack(found)


def _labels_bias_corrected(user_data, n, classes, device):   # :409-425
    residual = torch.stack([d["gradients"][-1] for d in user_data]).mean(dim=0)
    negative = (residual < 0).nonzero()                     # [k, 1], ascending class order
    found = list(negative.squeeze(dim=-1))                  # stage 1: one label per class with a negative mean bias gradient
    impact = residual[negative].sum() / n
    residual[negative] = residual[negative] - impact
    for _ in range(max(n - len(found), 0)):                 # stage 2: repeated classes, most negative residual first
        pick = residual.argmin()
        found.append(pick)
        residual[pick] -= impact
    return torch.stack(found)


def _labels_random(user_data, n, classes, device):        # :459-461
    return torch.randint(0, classes, (n,), device=device)


_LABEL_STRATEGIES = {
    "iDLG": _labels_idlg, "analytic": _labels_analytic, "yin": _labels_yin, "wainakh-simple": _labels_wainakh_simple,
    "bias-corrected": _labels_bias_corrected, "random": _labels_random,
}


def recover_labels(strategy, user_data, setup, data_shape=None):
    """base_attack.py:305-475 for the strategies that need no model queries."""
    if strategy is None:
        return None
    n = user_data[0]["metadata"]["num_data_points"]
    classes = user_data[0]["gradients"][-1].shape[0]
    device = setup["device"]
    if strategy == "wainakh-whitebox":
        raise NotImplementedError("label_strategy=wainakh-whitebox (model queries per class) is not implemented")
    if strategy == "exhaustive":
        raise ValueError(f"Exhaustive label searching not implemented. A naive strategy would attack {classes ** n} label vectors.")
    if strategy not in _LABEL_STRATEGIES:
        raise ValueError(f"Invalid label recovery strategy {strategy} given.")
    labels = _LABEL_STRATEGIES[strategy](user_data, n, classes, device)
    if len(labels) < n:                                     # :468-471: fill up with random classes
        labels = torch.cat([labels, torch.randint(0, classes, (n - len(labels),), device=device)])
    labels = labels.sort()[0]                               # :473
    log.info(f"Recovered labels {labels.tolist()} through strategy {strategy}.")
    return labels


# ---- candidate initialisation (base_attack.py:222-285) --------------------------------------------------------------------
# Draws come from torch's global generator on ``setup['device']`` in the reference's order, so trial k of a restarted attack
# starts from the reference's trial-k candidate.
_PLAIN_INITS = {
    "randn": lambda shape, setup: torch.randn(shape, **setup),
    "randn-trunc": lambda shape, setup: (torch.randn(shape, **setup) * 0.1).clamp(-0.1, 0.1),
    "rand": lambda shape, setup: (torch.rand(shape, **setup) * 2) - 1.0,
    "zeros": lambda shape, setup: torch.zeros(shape, **setup),
}
_COLOUR_WORDS = ("red", "green", "blue", "dark", "light")


def _colour_init(init_type, shape, dm, ds, setup):          # :236-246
    if "light" in init_type:
        candidate = torch.ones(shape, **setup)
    else:
        candidate = torch.zeros(shape, **setup)
        channel = 0 if "red" in init_type else (1 if "green" in init_type else 2)   # "dark" falls through to blue like the reference
        candidate[:, channel, :, :] = 1
    return (candidate - dm) / ds if "-true" in init_type else candidate


def _pattern_init(init_type, shape, setup):                 # :247-281: one random k x k tile repeated over the image
    width = int("".join(ch for ch in init_type if ch.isdigit()))
    wants_uniform = "rand" in init_type and ("randn" not in init_type if "patterned" in init_type else True)
    tile_shape = [shape[0], 3, width, width]
    tile = (torch.rand(tile_shape, **setup) * 2) - 1 if wants_uniform else torch.randn(tile_shape, **setup)
    reps = (1, 1, int(math.ceil(shape[2] / width)), int(math.ceil(shape[3] / width)))
    return torch.tile(tile, reps)[:, :, : shape[2], : shape[3]].contiguous().clone()


def initialize_data(init_type, data_shape, dm, ds, setup):
    if init_type in _PLAIN_INITS:
        return _PLAIN_INITS[init_type](data_shape, setup)
    if any(word in init_type for word in _COLOUR_WORDS):
        return _colour_init(init_type, data_shape, dm, ds, setup)
    if "patterned" in init_type or "wei" in init_type:
        return _pattern_init(init_type, data_shape, setup)
    raise ValueError(f"Unknown initialization scheme {init_type} given.")


def measured_features(shared_data, labels):
    """regularizers.py:31-43: per-label rows of the de-biased last-layer weight gradient."""
    out = []
    for user_data in shared_data:
        weights, bias = user_data["gradients"][-2], user_data["gradients"][-1]
        debiased = weights / bias[:, None]
        rows = [debiased[label] if bias[label] != 0 else torch.zeros_like(debiased[0]) for label in labels]
        out.append(torch.stack(rows))
    return out


# ---- text models: optimise in embedding space (base_attack.py:76-167) ---------------------------------------------------
EMBEDDING_LAYER_NAMES = ("encoder.weight", "word_embeddings.weight", "transformer.wte")   # base_attack.py:15


def prepare_for_text_data(rec_models, shared_data, text_strategy="run-embedding"):
    """base_attack.py:76-128 ("run-embedding"): the token embedding cannot be optimised through, so the candidate lives in
    embedding space -- the embedding layer of every attacked model is replaced by ``Identity`` and its gradient entry is
    removed from the shared update.  Returns ``(embeddings, token_embedding_dim)`` with ``embeddings[i] = dict(weight, grads)``;
    ``shared_data[i]["gradients"]`` is edited in place exactly like the reference does."""
    if text_strategy == "no-preprocessing":
        return [], None
    if text_strategy != "run-embedding":
        raise ValueError(f"Invalid text strategy {text_strategy} given.")
    embeddings = []
    for model, data in zip(rec_models, shared_data):
        name_to_idx = dict(zip([n for n, _ in model.named_parameters()], range(len(data["gradients"]))))
        position = None
        for name in EMBEDDING_LAYER_NAMES:      # the last matching name wins, as in the reference's nested loop
            for key in name_to_idx:
                if name in key:
                    position = name_to_idx[key]
        if position is None:
            raise ValueError("no token-embedding layer found in the model")
        weight = list(model.parameters())[position]
        embeddings.append(dict(weight=weight, grads=data["gradients"].pop(position)))

        def replace(module):
            for child_name, child in module.named_children():
                if isinstance(child, torch.nn.Embedding):
                    if child.weight is weight:
                        setattr(module, child_name, torch.nn.Identity())
                else:
                    replace(child)

        replace(model)
    return embeddings, embeddings[0]["weight"].shape[1]


def _max_similarity(recovered, true, subset=None):
    """base_attack.py:126-133 -- note the *squared* norms in the denominator (reference behaviour).  On the GPU this is
    one engine kernel (``bre_token_match``); host tensors (the CPU tests of the epilogue) take the formula below."""
    if recovered.is_cuda:
        from ..engine import token_match

        return token_match(recovered, true, subset)
    if subset is not None:
        true = true[subset, :]
    centred = recovered - recovered.mean(dim=-1, keepdim=True)
    vocab = true - true.mean(dim=-1, keepdim=True)
    scores = centred @ vocab.T
    scores = scores / centred.square().sum(dim=-1, keepdim=True) / vocab.square().sum(dim=-1)
    return scores.argmax(dim=1)


def postprocess_text_data(reconstructed, embedding_weight, token_recovery):
    """base_attack.py:123-167: map the reconstructed embeddings back to token ids."""
    if token_recovery == "from-embedding":
        rec = reconstructed["data"]
        base_shape = rec.shape[0:2]
        tokens = _max_similarity(rec.reshape(-1, rec.shape[-1]), embedding_weight).view(*base_shape)
    elif token_recovery == "from-labels":
        tokens = reconstructed["labels"]
    elif token_recovery == "from-limited-embedding":
        rec = reconstructed["data"]
        base_shape = rec.shape[0:2]
        active = reconstructed["labels"].unique()
        matches = _max_similarity(rec.reshape(-1, rec.shape[-1]), embedding_weight, subset=active)
        tokens = active[matches].view(*base_shape)
    else:
        raise ValueError(f"Invalid token recovery {token_recovery} given.")
    reconstructed["data"] = tokens
    return reconstructed
