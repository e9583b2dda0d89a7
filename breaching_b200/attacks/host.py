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


def recover_labels(strategy, user_data, setup, data_shape=None):
    """base_attack.py:305-475 for the strategies that need no model queries.  Integer logic: bit-exact."""
    num_data_points = user_data[0]["metadata"]["num_data_points"]
    num_classes = user_data[0]["gradients"][-1].shape[0]
    num_queries = len(user_data)
    device = setup["device"]

    if strategy is None:
        return None
    if strategy == "iDLG":
        label_list = [torch.argmin(torch.sum(d["gradients"][-2], dim=-1), dim=-1).detach() for d in user_data]
        labels = torch.stack(label_list).unique()
    elif strategy == "analytic":
        label_list = [(d["gradients"][-1] < 0).nonzero() for d in user_data]
        labels = torch.stack(label_list).unique()[:num_data_points]
    elif strategy == "yin":
        total_min_vals = 0
        for d in user_data:
            total_min_vals += d["gradients"][-2].min(dim=-1)[0]
        labels = total_min_vals.argsort()[:num_data_points]
    elif strategy == "wainakh-simple":
        m_impact = 0
        for d in user_data:
            g_i = d["gradients"][-2].sum(dim=1)
            m_query = torch.where(g_i < 0, g_i, torch.zeros_like(g_i)).sum() * (1 + 1 / num_classes) / num_data_points
            m_impact += m_query / num_queries
        label_list = []
        g_i = torch.stack([d["gradients"][-2].sum(dim=1) for d in user_data]).mean(dim=0)
        idx = 0
        for idx in range(num_classes):  # stage 1
            if g_i[idx] < 0:
                label_list.append(torch.as_tensor(idx, device=device))
                g_i[idx] -= m_impact
        while len(label_list) < num_data_points:  # stage 2 (decrements g_i[idx] like the reference, :405)
            selected_idx = g_i.argmin()
            label_list.append(torch.as_tensor(selected_idx, device=device))
            g_i[idx] -= m_impact
        labels = torch.stack(label_list)
    elif strategy == "wainakh-whitebox":
        raise NotImplementedError("label_strategy=wainakh-whitebox (model queries per class) is not implemented")
    elif strategy == "bias-corrected":
        bias_per_query = [d["gradients"][-1] for d in user_data]
        label_list = []
        average_bias = torch.stack(bias_per_query).mean(dim=0)
        valid_classes = (average_bias < 0).nonzero()
        label_list += [*valid_classes.squeeze(dim=-1)]
        m_impact = average_bias[valid_classes].sum() / num_data_points
        average_bias[valid_classes] = average_bias[valid_classes] - m_impact
        while len(label_list) < num_data_points:
            selected_idx = average_bias.argmin()
            label_list.append(selected_idx)
            average_bias[selected_idx] -= m_impact
        labels = torch.stack(label_list)
    elif strategy == "random":
        labels = torch.randint(0, num_classes, (num_data_points,), device=device)
    elif strategy == "exhaustive":
        raise ValueError(
            f"Exhaustive label searching not implemented. A naive strategy would attack "
            f"{num_classes ** num_data_points} label vectors."
        )
    else:
        raise ValueError(f"Invalid label recovery strategy {strategy} given.")

    if len(labels) < num_data_points:
        labels = torch.cat([labels, torch.randint(0, num_classes, (num_data_points - len(labels),), device=device)])
    labels = labels.sort()[0]
    log.info(f"Recovered labels {labels.tolist()} through strategy {strategy}.")
    return labels


def initialize_data(init_type, data_shape, dm, ds, setup):
    """base_attack.py:222-285: candidate initialisation, drawn from torch's global generator on ``setup['device']``
    so that the draw order matches the reference trial by trial."""
    if init_type == "randn":
        candidate = torch.randn(data_shape, **setup)
    elif init_type == "randn-trunc":
        candidate = (torch.randn(data_shape, **setup) * 0.1).clamp(-0.1, 0.1)
    elif init_type == "rand":
        candidate = (torch.rand(data_shape, **setup) * 2) - 1.0
    elif init_type == "zeros":
        candidate = torch.zeros(data_shape, **setup)
    elif any(c in init_type for c in ["red", "green", "blue", "dark", "light"]):
        candidate = torch.zeros(data_shape, **setup)
        if "light" in init_type:
            candidate = torch.ones(data_shape, **setup)
        else:
            nonzero_channel = 0 if "red" in init_type else 1 if "green" in init_type else 2
            candidate[:, nonzero_channel, :, :] = 1
        if "-true" in init_type:
            candidate = (candidate - dm) / ds
    elif "patterned" in init_type or "wei" in init_type:
        pattern_width = int("".join(filter(str.isdigit, init_type)))
        if "patterned" in init_type:
            uniform = ("rand" in init_type) and ("randn" not in init_type)
        else:
            uniform = "rand" in init_type
        if uniform:
            seed = (torch.rand([data_shape[0], 3, pattern_width, pattern_width], **setup) * 2) - 1
        else:
            seed = torch.randn([data_shape[0], 3, pattern_width, pattern_width], **setup)
        x_factor = int(math.ceil(data_shape[2] / pattern_width))
        y_factor = int(math.ceil(data_shape[3] / pattern_width))
        candidate = torch.tile(seed, (1, 1, x_factor, y_factor))[:, :, : data_shape[2], : data_shape[3]].contiguous().clone()
    else:
        raise ValueError(f"Unknown initialization scheme {init_type} given.")
    return candidate


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


def _max_similarity(recovered, true):
    """base_attack.py:126-133 -- note the *squared* norms in the denominator (reference behaviour)."""
    recovered = recovered - recovered.mean(dim=-1, keepdim=True)
    true = true - true.mean(dim=-1, keepdim=True)
    cosim = recovered.matmul(true.T) / recovered.pow(2).sum(dim=-1)[:, None] / true.pow(2).sum(dim=-1)[None, :]
    return cosim.argmax(dim=1)


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
        matches = _max_similarity(rec.reshape(-1, rec.shape[-1]), embedding_weight[active, :])
        tokens = active[matches].view(*base_shape)
    else:
        raise ValueError(f"Invalid token recovery {token_recovery} given.")
    reconstructed["data"] = tokens
    return reconstructed
