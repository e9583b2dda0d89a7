"""Attack hyper-parameter presets and an OmegaConf-like attribute dictionary.

The reference composes these with hydra from ``breaching/config/attack/*.yaml``; hydra and
omegaconf are not available here (and are not needed by the hot path), so the presets are
restated as plain Python data.  The attacker only needs attribute + item access and
``keys()`` / ``items()`` (reference ``attacks/optimization_based_attack.py:33-38``), which
:class:`AttackConfig` provides; real OmegaConf ``DictConfig`` objects are accepted as well.

Values follow (reference file:line):
  * defaults ............ ``config/attack/_default_optimization_attack.yaml:1-46``
  * invertinggradients .. ``config/attack/invertinggradients.yaml:4-33``
  * modern .............. ``config/attack/modern.yaml:4-36``
  * seethroughgradients . ``config/attack/seethroughgradients.yaml:4-36``
  * clsattack / legacy / sanitycheck / tag / wei / beyondinfering / deepleakage: same folder.
"""
import copy


class AttackConfig(dict):
    """dict with attribute access (``cfg.optim.step_size``) mirroring OmegaConf's DictConfig surface."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value

    def __deepcopy__(self, memo):
        return AttackConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(value):
    if isinstance(value, dict):
        return AttackConfig({k: _wrap(v) for k, v in value.items()})
    if isinstance(value, (list, tuple)):
        return [_wrap(v) for v in value]
    return value


def _merge(base, new):
    for key, val in new.items():
        if isinstance(val, dict) and isinstance(base.get(key), dict):
            _merge(base[key], val)
        else:
            base[key] = copy.deepcopy(val)
    return base


_DEFAULT = dict(
    type="default",
    attack_type="optimization",
    label_strategy="bias-corrected",
    text_strategy="run-embedding",
    token_recovery="from-labels",
    objective=dict(type="euclidean", scale=1.0, task_regularization=0.0),
    restarts=dict(num_trials=1, scoring="euclidean"),
    init="randn",
    normalize_gradients=False,
    optim=dict(
        optimizer="Adam",
        signed=None,
        step_size=1.0,
        boxed=False,
        max_iterations=400,
        step_size_decay=None,
        langevin_noise=0.0,
        warmup=0,
        grad_clip=None,
        callback=100,
    ),
    augmentations=None,
    differentiable_augmentations=False,
    regularization=None,
    impl=dict(dtype="float", mixed_precision=False, JIT=None),
)

_PRESETS = {
    "_default_optimization_attack": {},
    "invertinggradients": dict(
        type="invertinggradients",
        objective=dict(type="cosine-similarity", scale=1.0),
        restarts=dict(num_trials=1, scoring="cosine-similarity"),
        optim=dict(
            optimizer="adam", signed="hard", step_size=0.1, boxed=True, max_iterations=24_000,
            step_size_decay="step-lr", callback=1000,
        ),
        regularization=dict(total_variation=dict(scale=0.2, inner_exp=1, outer_exp=1)),
    ),
    "modern": dict(
        type="invertinggradients",
        objective=dict(type="cosine-similarity", scale=1.0),
        init="patterned-4",
        restarts=dict(num_trials=1, scoring="cosine-similarity"),
        optim=dict(
            optimizer="adam", signed="soft", step_size=0.1, boxed=True, max_iterations=24_000,
            step_size_decay="cosine-decay", warmup=50, callback=1000,
        ),
        regularization=dict(
            total_variation=dict(scale=0.1, inner_exp=2, outer_exp=0.5, double_opponents=True),
            features=dict(scale=0.1),
            deep_inversion=dict(scale=0.0),
        ),
    ),
    "clsattack": dict(
        type="invertinggradients",
        objective=dict(type="cosine-similarity", scale=1.0),
        init="patterned-4-randn",
        restarts=dict(num_trials=1, scoring="cosine-similarity"),
        optim=dict(
            optimizer="adam", signed="soft", step_size=0.1, boxed=True, max_iterations=24_000,
            step_size_decay="cosine-decay", warmup=50, callback=1000,
        ),
        regularization=dict(
            total_variation=dict(scale=0.2, inner_exp=2, outer_exp=0.5, double_opponents=True),
            features=dict(scale=0.0),
            deep_inversion=dict(scale=0.0),
        ),
    ),
    "legacy": dict(
        type="invertinggradients",
        objective=dict(type="cosine-similarity", scale=1.0),
        init="zeros",
        restarts=dict(num_trials=1, scoring="cosine-similarity"),
        optim=dict(
            optimizer="adam", signed="soft", step_size=0.1, boxed=True, max_iterations=24_000,
            step_size_decay="cosine-decay", callback=1000,
        ),
        regularization=dict(
            total_variation=dict(scale=0.1, inner_exp=2, outer_exp=0.5, double_opponents=True),
            features=dict(scale=0.1),
            deep_inversion=dict(scale=0.00005),
        ),
    ),
    "seethroughgradients": dict(
        type="see-through-gradients",
        label_strategy="yin",
        objective=dict(type="euclidean", scale=1e-4),
        restarts=dict(num_trials=1, scoring="euclidean"),
        optim=dict(
            optimizer="adam", signed=False, step_size=0.1, boxed=True, max_iterations=20_000,
            step_size_decay="cosine-decay", langevin_noise=0.01, warmup=50, callback=1000,
        ),
        regularization=dict(
            total_variation=dict(scale=1e-4, inner_exp=1, outer_exp=1),
            norm=dict(scale=1e-6, pnorm=2),
            deep_inversion=dict(scale=0.1),
        ),
    ),
    "sanitycheck": dict(
        type="sanitycheck",
        objective=dict(type="cosine-similarity", scale=1.0),
        optim=dict(
            optimizer="adam", signed=None, step_size=1, boxed=True, max_iterations=1,
            step_size_decay="none", callback=0,
        ),
    ),
    "tag": dict(
        type="tag",
        attack_type="joint-optimization",
        label_strategy="None",
        token_recovery="from-embedding",
        init="randn-trunc",
        objective=dict(type="tag-euclidean", scale=1.0, task_regularization=0.0, tag_scale=0.1, scale_scheme="linear"),
        optim=dict(
            optimizer="bert-adam", step_size=0.05, boxed=False, max_iterations=1000, grad_clip=1.0,
            warmup=50, step_size_decay="linear", callback=100,
        ),
    ),
    "deepleakage": dict(
        type="deep-leakage",
        attack_type="joint-optimization",
        label_strategy="None",
        token_recovery="from-embedding",
        optim=dict(optimizer="L-BFGS", step_size=1.0, boxed=False, max_iterations=1200, callback=100),
    ),
    "beyondinfering": dict(
        type="beyond-infering",
        optim=dict(optimizer="L-BFGS", step_size=1.0, boxed=True, max_iterations=400),
        regularization=dict(total_variation=dict(scale=0.2352, inner_exp=2, outer_exp=1.25)),
    ),
    "multiscale_ghiasi": dict(          # config/attack/multiscale_ghiasi.yaml (defaults: invertinggradients, then itself)
        _inherits="invertinggradients",
        attack_type="multiscale",
        type="multiscale-invertinggradients",
        num_stages=7,
        augmentations=dict(continuous_shift=dict(shift=224, padding="circular"), colorjitter=dict(none=None)),
        resize="focus",
        scale_pyramid="linear",
        optim=dict(optimizer="adam-safe", max_iterations=2000),
        differentiable_augmentations=True,
        update_augmentations="None",
    ),
    "wei": dict(
        type="beyond-infering",
        objective=dict(type="euclidean", scale=1.0, task_regularization=1.0),
        init="patterned-16",
        optim=dict(optimizer="L-BFGS", step_size=1.0, boxed=True, max_iterations=300),
    ),
}


def _set_dotted(cfg, dotted, value):
    node = cfg
    parts = dotted.split(".")
    for part in parts[:-1]:
        if node.get(part) is None:
            node[part] = AttackConfig()
        node = node[part]
    node[parts[-1]] = _wrap(value)


def _parse_override_value(text):
    low = text.strip()
    if low in ("null", "None", "~", ""):
        return None
    if low in ("true", "True"):
        return True
    if low in ("false", "False"):
        return False
    try:
        return int(low.replace("_", ""))
    except ValueError:
        pass
    try:
        return float(low)
    except ValueError:
        return low.strip("'\"")


def get_attack_config(attack="invertinggradients", overrides=()):
    """Counterpart of reference ``breaching.get_attack_config`` (``breaching/__init__.py:24-29``).

    ``overrides`` accepts hydra-style strings (``"optim.max_iterations=100"``) or a dict of dotted keys.
    """
    if attack not in _PRESETS:
        raise ValueError(f"Unknown attack configuration {attack}. Known: {sorted(_PRESETS)}")
    chain, name = [], attack
    while name is not None:                       # hydra `defaults:` lists: base presets first
        chain.append(name)
        name = _PRESETS[name].get("_inherits")
    merged = copy.deepcopy(_DEFAULT)
    for name in reversed(chain):
        _merge(merged, {k: v for k, v in _PRESETS[name].items() if k != "_inherits"})
    cfg = _wrap(merged)
    if isinstance(overrides, dict):
        for key, val in overrides.items():
            _set_dotted(cfg, key, val)
    else:
        for entry in overrides:
            key, _, val = entry.partition("=")
            key = key.lstrip("+")
            if key.startswith("attack."):
                key = key[len("attack."):]
            _set_dotted(cfg, key, _parse_override_value(val))
    return cfg


def cfg_get(node, key, default=None):
    """Attribute-or-item lookup that works for AttackConfig, OmegaConf DictConfig and plain dicts."""
    if node is None:
        return default
    try:
        if isinstance(node, dict):
            return node.get(key, default)
        if key in node:
            return node[key]
        return default
    except Exception:
        return getattr(node, key, default)
