"""CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
``/root/reference/breaching``).  The numerical substrate is PyTorch on the CPU: the
reference itself has no arithmetic of its own on this path -- model forward, the
double backward, Adam and the LR schedulers all live in the third-party dependency
``torch`` (reference pins ``pytorch=1.10.1`` in ``environment.yml:18``; this image has
torch 2.11.0, which is what both the reference and this oracle run on here).  What is
restated is the reference's *algorithm*: the order of operations of one trial, the
objective / regulariser formulas, gradient post-processing, the optimiser update rule
(restated explicitly, not through ``torch.optim``), the LR tables, box projection, the
best-so-far bookkeeping and the label-recovery heuristics.

Pinned against the live reference by ``tests/test_oracle_vs_reference.py`` and against
the committed fixtures in ``tests/golden/`` (made by ``tests/golden/make_golden.py``).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# LR tables  (attacks/auxiliaries/common.py:5-40 optimizer_lookup, :74-162 GradualWarmupScheduler)
# --------------------------------------------------------------------------------------
def lr_table(step_size, scheduler, warmup, max_iterations, n=None):
    """LR used by optimiser step ``it`` (0-based) for ``it < n``.

    Restated closed forms of what stepping the reference's scheduler objects produces:
      * ``step-lr``: MultiStepLR with *float* milestones ``T//2.667, T//1.6, T//1.142`` (common.py:21-24);
        torch keeps them in a Counter keyed by float, ``last_epoch`` (int) hits ``float(m)`` when equal.
      * ``cosine-decay``: CosineAnnealingLR(T_max=T, eta_min=0) (common.py:25-26).
      * ``linear``: LambdaLR ``max(0, (T - step)/max(1, T))`` (common.py:27-32).
      * anything else: constant (common.py:33-34).
      * ``warmup > 0``: GradualWarmupScheduler(multiplier=1): lr = base * e/warmup for e <= warmup
        (so the very first step runs at lr 0), then the wrapped scheduler, whose clock starts
        at 0 when e == warmup + 1 (common.py:100-117, :131-146).
    """
    T = int(max_iterations)
    n = T if n is None else int(n)
    base = float(step_size)

    def after(e):  # lr of the wrapped scheduler after ``e`` of its own steps
        if scheduler == "step-lr":
            milestones = [T // 2.667, T // 1.6, T // 1.142]
            k = sum(1 for m in milestones if e >= m)
            return base * (0.1 ** k)
        if scheduler == "cosine-decay":
            return base * (1 + math.cos(math.pi * e / T)) / 2 if T > 0 else base
        if scheduler == "linear":
            return base * max(0.0, float(T - e) / float(max(1, T)))
        return base

    table = []
    for it in range(n):
        if warmup and warmup > 0:
            if it <= warmup:
                table.append(base * (float(it) / warmup))
            else:
                # After warm-up the wrapped scheduler is stepped once per iteration; its
                # ``last_epoch`` is (it - warmup - 1) when optimiser step ``it`` runs.
                table.append(after(it - warmup - 1))
        else:
            table.append(after(it))
    return table


def lr_table_by_stepping(step_size, scheduler, warmup, max_iterations, n, optimizer_lookup):
    """LR table obtained by literally stepping scheduler objects built by ``optimizer_lookup``
    (pass the reference's ``attacks.auxiliaries.common.optimizer_lookup``)."""
    p = torch.zeros(1, requires_grad=True)
    opt, sched = optimizer_lookup([p], "adam", step_size, scheduler=scheduler, warmup=warmup,
                                  max_iterations=max_iterations)
    out = []
    for _ in range(n):
        out.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return out


# --------------------------------------------------------------------------------------
# objectives (attacks/auxiliaries/objectives.py)
# --------------------------------------------------------------------------------------
def matching_objective(kind, G, g, scale=1.0, tag_scale=0.1, scale_scheme="linear", fudge=1e-7, mask_value=1e-6):
    """Gradient-matching scalar for lists of tensors ``G`` (reconstructed, differentiable) and ``g`` (data)."""
    zero = G[0].new_zeros(1)
    if kind == "euclidean":  # objectives.py:91-95
        acc = zero.clone()
        for a, b in zip(G, g):
            acc = acc + (a - b).pow(2).sum()
        return 0.5 * acc * scale
    if kind == "l1":  # objectives.py:160-164
        acc = zero.clone()
        for a, b in zip(G, g):
            acc = acc + (a - b).abs().sum()
        return 0.5 * acc * scale
    if kind == "tag-euclidean":  # objectives.py:113-141
        L = len(G)
        if scale_scheme == "linear":
            w = torch.arange(L, 0, -1, dtype=G[0].dtype) / L
        elif scale_scheme == "exp":
            w = torch.arange(L, 0, -1, dtype=G[0].dtype).softmax(dim=0)
            w = w / w[0]
        else:
            w = G[0].new_ones(L)
        acc = zero.clone()
        for a, b, wl in zip(G, g, w):
            acc = acc + (a - b).pow(2).sum() + tag_scale * wl * (a - b).abs().sum()
        return 0.5 * acc * scale
    if kind in ("cosine-similarity", "angular", "fast-cosine-similarity", "masked-cosine-similarity"):
        sp, rn, dn = zero.clone(), zero.clone(), zero.clone()
        for a, b in zip(G, g):
            if kind == "masked-cosine-similarity":  # objectives.py:234-244
                m = b.abs() > mask_value
                sp = sp + (a * b * m).sum()
                rn = rn + (a * m).pow(2).sum()
                dn = dn + (b * m).pow(2).sum()
            elif kind == "fast-cosine-similarity":  # objectives.py:261-273 (norms detached)
                sp = sp + (a * b).sum()
                rn = rn + a.detach().pow(2).sum()
                dn = dn + b.detach().pow(2).sum()
            else:  # objectives.py:185-196
                sp = sp + (a * b).sum()
                rn = rn + a.pow(2).sum()
                dn = dn + b.pow(2).sum()
        cos_obj = 1 - sp / (rn.sqrt() * dn.sqrt())
        if kind == "angular":  # objectives.py:210-214
            cosine = 1 - cos_obj
            angle = torch.acos(cosine.clamp(min=-1 + fudge, max=1 - fudge))
            return angle / math.pi * scale
        return cos_obj * scale
    raise ValueError(f"Unknown objective type {kind} given.")


# --------------------------------------------------------------------------------------
# regularisers (attacks/auxiliaries/regularizers.py, deepinversion.py)
# --------------------------------------------------------------------------------------
def total_variation(x, scale=0.1, inner_exp=1, outer_exp=1, double_opponents=False, eps=1e-8):
    """regularizers.py:130-147.  The reference's grouped 3x3 convolution with kernels
    ``[[0,0,0],[0,-1,1],[0,0,0]]`` (and its transpose), zero padding 1, is a forward difference
    ``x[.., j+1] - x[.., j]`` along W (resp. H) with the *out-of-range neighbour read as 0*."""
    if double_opponents:
        x = torch.cat([x, x[:, 0:1] - x[:, 1:2], x[:, 0:1] - x[:, 2:3], x[:, 1:2] - x[:, 2:3]], dim=1)
    dh = F.pad(x, (0, 0, 0, 1))[:, :, 1:, :] - x  # transposed kernel comes first in the weight (channel 0::2)
    dw = F.pad(x, (0, 1, 0, 0))[:, :, :, 1:] - x
    sq_h = (dh.abs() + eps).pow(inner_exp)
    sq_w = (dw.abs() + eps).pow(inner_exp)
    return (sq_h + sq_w).pow(outer_exp).mean() * scale


def norm_regularization(x, scale=0.1, pnorm=2.0):
    """regularizers.py:197-198."""
    return 1 / pnorm * x.pow(pnorm).mean() * scale


def orthogonality_regularization(x):
    """regularizers.py:169-178 (note: the reference ignores ``scale`` here)."""
    if x.shape[0] == 1:
        return 0
    B = x.shape[0]
    prod = (x.unsqueeze(0) * x.unsqueeze(1)).pow(2).view(B, B, -1).mean(dim=2)
    idx = torch.arange(0, B)
    prod[idx, idx] = 0
    return prod.sum()


class _BNInputStats:
    """deepinversion.py:93-103 restated: distance of batch statistics of the BN *input* to the running stats."""

    def __init__(self, module):
        self.value = None
        self.handle = module.register_forward_hook(self._hook)

    def _hook(self, module, inputs, output):
        z = inputs[0]
        mean = z.mean([0, 2, 3])
        var = z.permute(1, 0, 2, 3).reshape(z.shape[1], -1).var(1, unbiased=False)
        self.value = torch.norm(module.running_var.data - var, 2) + torch.norm(module.running_mean.data - mean, 2)


class _LinearInput:
    """regularizers.py:8-20 restated."""

    def __init__(self, module):
        self.features = None
        self.handle = module.register_forward_hook(self._hook)

    def _hook(self, module, inputs, output):
        self.features = inputs[0]


# --------------------------------------------------------------------------------------
# label recovery (attacks/base_attack.py:305-475)
# --------------------------------------------------------------------------------------
def recover_labels(strategy, shared_data, num_data_points, generator=None):
    num_classes = shared_data[0]["gradients"][-1].shape[0]
    if strategy is None:
        return None
    if strategy == "iDLG":  # :320-328
        lst = [torch.argmin(torch.sum(d["gradients"][-2], dim=-1), dim=-1).detach() for d in shared_data]
        labels = torch.stack(lst).unique()
    elif strategy == "analytic":  # :329-335
        lst = [(d["gradients"][-1] < 0).nonzero() for d in shared_data]
        labels = torch.stack(lst).unique()[:num_data_points]
    elif strategy == "yin":  # :336-345
        total = 0
        for d in shared_data:
            total = total + d["gradients"][-2].min(dim=-1)[0]
        labels = total.argsort()[:num_data_points]
    elif strategy == "bias-corrected":  # :409-425
        avg = torch.stack([d["gradients"][-1] for d in shared_data]).mean(dim=0).clone()
        valid = (avg < 0).nonzero()
        lst = [*valid.squeeze(dim=-1)]
        m_impact = avg[valid].sum() / num_data_points
        avg[valid] = avg[valid] - m_impact
        while len(lst) < num_data_points:
            sel = avg.argmin()
            lst.append(sel)
            avg[sel] -= m_impact
        labels = torch.stack(lst)
    elif strategy == "wainakh-simple":  # :347-358, :389-407 (incl. the stage-2 ``g_i[idx]`` quirk)
        m_impact = 0
        for d in shared_data:
            g_i = d["gradients"][-2].sum(dim=1)
            m_query = torch.where(g_i < 0, g_i, torch.zeros_like(g_i)).sum() * (1 + 1 / num_classes) / num_data_points
            m_impact = m_impact + m_query / len(shared_data)
        g_i = torch.stack([d["gradients"][-2].sum(dim=1) for d in shared_data]).mean(dim=0)
        lst = []
        idx = 0
        for idx in range(num_classes):
            if g_i[idx] < 0:
                lst.append(torch.as_tensor(idx))
                g_i[idx] -= m_impact
        while len(lst) < num_data_points:
            sel = g_i.argmin()
            lst.append(torch.as_tensor(sel))
            g_i[idx] -= m_impact
        labels = torch.stack(lst)
    elif strategy == "random":  # :452-454
        labels = torch.randint(0, num_classes, (num_data_points,), generator=generator)
    else:
        raise ValueError(f"Invalid label recovery strategy {strategy} given.")
    if len(labels) < num_data_points:  # :466-470
        labels = torch.cat([labels, torch.randint(0, num_classes, (num_data_points - len(labels),), generator=generator)])
    return labels.sort()[0]  # :473


# --------------------------------------------------------------------------------------
# candidate initialisation (attacks/base_attack.py:222-285)
# --------------------------------------------------------------------------------------
def initialize_data(init_type, shape, dm, ds, dtype=torch.float32):
    if init_type == "randn":
        cand = torch.randn(shape, dtype=dtype)
    elif init_type == "randn-trunc":
        cand = (torch.randn(shape, dtype=dtype) * 0.1).clamp(-0.1, 0.1)
    elif init_type == "rand":
        cand = (torch.rand(shape, dtype=dtype) * 2) - 1.0
    elif init_type == "zeros":
        cand = torch.zeros(shape, dtype=dtype)
    elif any(c in init_type for c in ["red", "green", "blue", "dark", "light"]):
        cand = torch.zeros(shape, dtype=dtype)
        if "light" in init_type:
            cand = torch.ones(shape, dtype=dtype)
        else:
            ch = 0 if "red" in init_type else 1 if "green" in init_type else 2
            cand[:, ch, :, :] = 1
        if "-true" in init_type:
            cand = (cand - dm) / ds
    elif "patterned" in init_type or "wei" in init_type:
        width = int("".join(filter(str.isdigit, init_type)))
        uniform = ("rand" in init_type) and not ("randn" in init_type and "patterned" in init_type)
        if "wei" in init_type:
            uniform = "rand" in init_type
        if uniform:
            seed = (torch.rand([shape[0], 3, width, width], dtype=dtype) * 2) - 1
        else:
            seed = torch.randn([shape[0], 3, width, width], dtype=dtype)
        fx = int(math.ceil(shape[2] / width))
        fy = int(math.ceil(shape[3] / width))
        cand = torch.tile(seed, (1, 1, fx, fy))[:, :, : shape[2], : shape[3]].contiguous().clone()
    else:
        raise ValueError(f"Unknown initialization scheme {init_type} given.")
    return cand


# --------------------------------------------------------------------------------------
# one trial (attacks/optimization_based_attack.py:90-189)
# --------------------------------------------------------------------------------------
_OPTIMS = {
    # name -> (kind, beta1, beta2, eps, weight_decay, momentum, nesterov)   (common.py:6-17)
    "adam": ("adam", 0.9, 0.999, 1e-8, 0.0, 0.0, False),
    "adam-safe": ("adam", 0.5, 0.99, 1e-4, 0.0, 0.0, False),
    "bert-adam": ("adamw", 0.9, 0.999, 1e-6, 0.01, 0.0, False),
    "momgd": ("sgd", 0.0, 0.0, 0.0, 0.0, 0.9, True),
    "gd": ("sgd", 0.0, 0.0, 0.0, 0.0, 0.0, False),
}


def cfg_get(node, key, default=None):
    if node is None:
        return default
    try:
        val = node[key]
    except (KeyError, TypeError, AttributeError):
        return default
    return val


def active_regularizers(cfg):
    """optimization_based_attack.py:33-38: every configured regulariser with ``scale > 0``."""
    out = {}
    reg = cfg_get(cfg, "regularization")
    if reg is None:
        return out
    for key in reg.keys():
        if reg[key]["scale"] > 0:
            out[key] = dict(reg[key])
    return out


class TrialOracle:
    """Runs the reference algorithm for one model / one payload (the BASELINE configs all use one).

    ``model`` is an ``nn.Module`` on the CPU already in the state the reference attacker would have put it in
    (``base_attack.py:169-212``: parameters and buffers loaded, ``.eval()`` when buffers are known).
    """

    def __init__(self, model, loss_fn, cfg, gradients, labels, dm, ds, dtype=torch.float32, local_hyperparams=None):
        self.model, self.loss_fn, self.cfg = model, loss_fn, cfg
        self.local = local_hyperparams
        self.g = [t.detach().to(dtype) for t in gradients]
        self.labels = labels
        self.dm, self.ds = dm, ds
        self.dtype = dtype
        self.regs = active_regularizers(cfg)
        self._bn_hooks, self._feat_hook, self._measured = [], None, None
        if "deep_inversion" in self.regs:  # regularizers.py:214-220
            for m in model.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    self._bn_hooks.append(_BNInputStats(m))
        if "features" in self.regs:  # regularizers.py:31-51
            w, b = self.g[-2], self.g[-1]
            deb = w / b[:, None]
            rows = [deb[l] if b[l] != 0 else torch.zeros_like(deb[0]) for l in labels]
            self._measured = torch.stack(rows)
            for m in model.modules():
                if isinstance(m, torch.nn.Linear):
                    last = m
            self._feat_hook = _LinearInput(last)

    def close(self):
        for h in self._bn_hooks:
            h.handle.remove()
        if self._feat_hook is not None:
            self._feat_hook.handle.remove()

    # objectives.py:26-46
    def param_gradient(self, x, create_graph):
        if self.local is not None:
            return self._multi_step_difference(x, create_graph)
        self.model.zero_grad()
        task_loss = self.loss_fn(self.model(x), self.labels)
        G = torch.autograd.grad(task_loss, list(self.model.parameters()), create_graph=create_graph)
        return G, task_loss

    # objectives.py:48-72 (FedAvg): K local SGD steps kept in the graph; the "gradient" is W_K - W_0.  The reference
    # deep-copies the module through its vendored make_functional_with_buffers; torch.func.functional_call evaluates
    # the same module with substituted parameters.
    def _multi_step_difference(self, x, create_graph):
        names = [n for n, _ in self.model.named_parameters()]
        params = [p.detach().clone().requires_grad_(True) for p in self.model.parameters()]
        initial = [p.clone() for p in params]
        buffers = dict(self.model.named_buffers())
        seen, task_loss = 0, None
        for i in range(self.local["steps"]):
            data = x[seen: seen + self.local["data_per_step"]]
            seen = (seen + self.local["data_per_step"]) % x.shape[0]
            out = torch.func.functional_call(self.model, ({n: p for n, p in zip(names, params)}, buffers), (data,))
            task_loss = self.loss_fn(out, self.local["labels"][i])
            grads = torch.autograd.grad(task_loss, params, create_graph=create_graph)
            params = [p - self.local["lr"] * g for p, g in zip(params, grads)]
        return [pl - p0 for pl, p0 in zip(params, initial)], task_loss

    def objective_terms(self, x):
        o = self.cfg["objective"]
        G, task_loss = self.param_gradient(x, create_graph=True)
        kw = {k: o[k] for k in ("tag_scale", "scale_scheme") if k in o}
        match = matching_objective(o["type"], G, self.g, scale=cfg_get(o, "scale", 1.0), **kw)
        total = match
        treg = cfg_get(o, "task_regularization", 0.0) or 0.0
        if treg != 0:
            total = total + treg * task_loss
        terms = {"match": float(match.detach()), "task_loss": float(task_loss.detach())}
        for key, r in self.regs.items():
            if key == "total_variation":
                val = total_variation(x, **{k: v for k, v in r.items()})
            elif key == "norm":
                val = norm_regularization(x, **r)
            elif key == "orthogonality":
                val = orthogonality_regularization(x)
            elif key == "deep_inversion":  # regularizers.py:222-227
                mult = cfg_get(r, "first_bn_multiplier", 10)
                val = r["scale"] * sum(h.value * (mult if i == 0 else 1.0) for i, h in enumerate(self._bn_hooks))
            elif key == "features":  # regularizers.py:53-57
                val = (self._feat_hook.features - self._measured).pow(2).mean() * r["scale"]
            else:
                raise ValueError(key)
            terms[key] = float(val.detach()) if torch.is_tensor(val) else float(val)
            total = total + val
        return total, terms

    def closure_gradient(self, x, iteration, lr, noise=None):
        """optimization_based_attack.py:146-187 -> (objective, processed candidate gradient, raw gradient)."""
        opt = self.cfg["optim"]
        x = x.detach().clone().requires_grad_(True)
        total, terms = self.objective_terms(x)
        (grad,) = torch.autograd.grad(total, x)
        raw = grad.clone()
        if (cfg_get(opt, "langevin_noise", 0.0) or 0.0) > 0:  # :167-170
            noise_map = torch.randn_like(grad) if noise is None else noise
            grad = grad + opt["langevin_noise"] * lr * noise_map
        clip = cfg_get(opt, "grad_clip")
        if clip is not None:  # :171-174
            norm = grad.norm()
            if norm > clip:
                grad = grad * (clip / (norm + 1e-6))
        signed = cfg_get(opt, "signed")
        if signed is not None:  # :175-184
            if signed == "soft":
                s = 1 - iteration / opt["max_iterations"]
                grad = (grad * s).tanh() / s
            elif signed == "hard":
                grad = grad.sign()
        return total.detach(), grad, raw, terms

    def run(self, x0, iterations=None, dryrun=False, record=False, noises=None):
        """optimization_based_attack.py:90-143 with the optimiser written out (torch.optim.Adam/AdamW/SGD formulas)."""
        opt = self.cfg["optim"]
        T = opt["max_iterations"]
        n = T if iterations is None else min(T, iterations)
        lrs = lr_table(opt["step_size"], cfg_get(opt, "step_size_decay"), cfg_get(opt, "warmup", 0) or 0, T, n)
        if opt["optimizer"].lower() == "l-bfgs":
            return self._run_lbfgs(x0, n, lrs, dryrun, record)
        kind, b1, b2, eps, wd, mom, nesterov = _OPTIMS[opt["optimizer"].lower()]
        x = x0.detach().clone().to(self.dtype)
        m = torch.zeros_like(x)
        v = torch.zeros_like(x)
        best = x.clone()
        fmin = float("inf")
        hist, trace = [], []
        lo, hi = -self.dm / self.ds, (1 - self.dm) / self.ds
        for it in range(n):
            lr = lrs[it]
            noise = None if noises is None else noises[it]
            phi, g, raw, terms = self.closure_gradient(x, it, lr, noise)
            t = it + 1
            if kind in ("adam", "adamw"):
                if kind == "adamw":
                    x = x * (1 - lr * wd)
                m = b1 * m + (1 - b1) * g
                v = b2 * v + (1 - b2) * g * g
                bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
                denom = v.sqrt() / math.sqrt(bc2) + eps
                x = x - (lr / bc1) * (m / denom)
            else:  # SGD (torch.optim.SGD semantics: buf = g at first step)
                d = g
                if mom != 0:
                    m = g.clone() if it == 0 else mom * m + g
                    d = g + mom * m if nesterov else m
                x = x - lr * d
            if cfg_get(opt, "boxed", False):  # :117-118
                x = torch.max(torch.min(x, hi), lo)
            phi_f = float(phi)
            if phi_f < fmin:  # :119-121  (objective before the step, candidate after it)
                fmin = phi_f
                best = x.clone()
            if record:
                trace.append(dict(objective=phi_f, raw_grad=raw, grad=g, candidate=x.clone(), lr=lr, terms=terms))
            if not math.isfinite(phi_f):  # :131-133
                break
            hist.append(phi_f)
            if dryrun:
                break
        return best, hist, trace

    def _run_lbfgs(self, x0, n, lrs, dryrun, record):
        """L-BFGS trials (common.py:18: ``torch.optim.LBFGS(params, lr=step_size)``, torch defaults).  The update rule lives in
        the third-party optimiser the reference calls; the oracle calls the same class with the reference's closure
        (optimization_based_attack.py:146-189) -- the device restatement under test is breaching_b200/attacks/lbfgs.py."""
        opt = self.cfg["optim"]
        x = torch.nn.Parameter(x0.detach().clone().to(self.dtype))
        optimizer = torch.optim.LBFGS([x], lr=opt["step_size"])
        best, fmin, hist, trace = x.detach().clone(), float("inf"), [], []
        lo, hi = -self.dm / self.ds, (1 - self.dm) / self.ds
        for it in range(n):
            optimizer.param_groups[0]["lr"] = lrs[it]

            def closure():
                phi, g, _, _ = self.closure_gradient(x.detach(), it, lrs[it], None)
                x.grad = g.to(x.dtype)
                return phi

            phi_f = float(optimizer.step(closure))
            with torch.no_grad():
                if cfg_get(opt, "boxed", False):
                    x.data = torch.max(torch.min(x, hi), lo)
                if phi_f < fmin:
                    fmin, best = phi_f, x.detach().clone()
            if record:
                trace.append(dict(objective=phi_f, candidate=x.detach().clone(), lr=lrs[it]))
            if not math.isfinite(phi_f):
                break
            hist.append(phi_f)
            if dryrun:
                break
        return best, hist, trace

    def score(self, x, scoring):
        """optimization_based_attack.py:191-204."""
        if scoring in ("euclidean", "cosine-similarity"):
            G, _ = self.param_gradient(x.detach(), create_graph=False)
            val = matching_objective(scoring, G, self.g, scale=1.0)
            val = float(val)
            return val if math.isfinite(val) else float("inf")
        raise ValueError(f"Scoring mechanism {scoring} not implemented.")


class MultiQueryOracle(TrialOracle):
    """Several (model, update) pairs attacked with one candidate: the closure sums the matching objective over
    ``zip(rec_model, shared_data)`` and adds the regularisers once (optimization_based_attack.py:157-162).  ``oracles[0]`` is
    built with the full attack config, the others with all regulariser scales set to zero."""

    def __init__(self, oracles):
        self.oracles = oracles
        first = oracles[0]
        self.cfg, self.dm, self.ds, self.dtype = first.cfg, first.dm, first.ds, first.dtype

    def objective_terms(self, x):
        total, terms = 0.0, {}
        for orc in self.oracles:
            val, t = orc.objective_terms(x)
            total = total + val
            for key, v in t.items():
                terms[key] = terms.get(key, 0.0) + v
        return total, terms

    def score(self, x, scoring):
        return sum(orc.score(x, scoring) for orc in self.oracles)

    def close(self):
        for orc in self.oracles:
            orc.close()


class JointTrialOracle(TrialOracle):
    """``OptimizationJointAttacker`` (optimization_with_label_attack.py:89-205): data and soft labels are optimised together.
    The closure feeds ``labels.softmax(-1)`` to the loss as class probabilities (:154), back-propagates onto both leaves
    (:162), post-processes both gradients (noise, per-tensor clipping, sign; :164-186) and one torch optimiser steps both
    (:108, common.py:5-18)."""

    def closure_gradients(self, x, ell, iteration, lr):
        from torch.nn.attention import SDPBackend, sdpa_kernel

        opt = self.cfg["optim"]
        x = x.detach().clone().requires_grad_(True)
        ell = ell.detach().clone().requires_grad_(True)
        self.labels = ell.softmax(dim=-1)
        with sdpa_kernel(SDPBackend.MATH):  # attention models: the fused CPU kernel has no double backward (shim 3)
            total, terms = self.objective_terms(x)
            gx, gl = torch.autograd.grad(total, [x, ell])
        raw = (gx.clone(), gl.clone())
        out = []
        for grad in (gx, gl):
            if (cfg_get(opt, "langevin_noise", 0.0) or 0.0) > 0:
                grad = grad + opt["langevin_noise"] * lr * torch.randn_like(grad)
            clip = cfg_get(opt, "grad_clip")
            if clip is not None:
                norm = grad.norm()
                if norm > clip:
                    grad = grad * (clip / (norm + 1e-6))
            signed = cfg_get(opt, "signed")
            if signed == "soft":
                s = 1 - iteration / opt["max_iterations"]
                grad = (grad * s).tanh() / s
            elif signed == "hard":
                grad = grad.sign()
            out.append(grad)
        return total.detach(), out[0], out[1], raw, terms

    def run_joint(self, x0, ell0, iterations=None, dryrun=False):
        """optimization_with_label_attack.py:89-143; the optimiser is the torch class the reference constructs."""
        opt = self.cfg["optim"]
        T = opt["max_iterations"]
        n = T if iterations is None else min(T, iterations)
        lrs = lr_table(opt["step_size"], cfg_get(opt, "step_size_decay"), cfg_get(opt, "warmup", 0) or 0, T, n)
        x = torch.nn.Parameter(x0.detach().clone().to(self.dtype))
        ell = torch.nn.Parameter(ell0.detach().clone().to(self.dtype))
        name = opt["optimizer"].lower()
        if name == "l-bfgs":
            optimizer = torch.optim.LBFGS([x, ell], lr=opt["step_size"])
        elif name in ("adam", "adam-safe"):
            kw = {} if name == "adam" else dict(betas=(0.5, 0.99), eps=1e-4)
            optimizer = torch.optim.Adam([x, ell], lr=opt["step_size"], **kw)
        elif name == "bert-adam":
            optimizer = torch.optim.AdamW([x, ell], lr=opt["step_size"], betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)
        else:
            optimizer = torch.optim.SGD([x, ell], lr=opt["step_size"], momentum=0.9 if name == "momgd" else 0.0, nesterov=name == "momgd")
        best, best_l, fmin, hist = x.detach().clone(), ell.detach().clone(), float("inf"), []
        lo, hi = -self.dm / self.ds, (1 - self.dm) / self.ds
        for it in range(n):
            optimizer.param_groups[0]["lr"] = lrs[it]

            def closure():
                phi, gx, gl, _, _ = self.closure_gradients(x.detach(), ell.detach(), it, lrs[it])
                x.grad, ell.grad = gx.to(x.dtype), gl.to(ell.dtype)
                return phi

            phi_f = float(optimizer.step(closure))
            with torch.no_grad():
                if cfg_get(opt, "boxed", False):
                    x.data = torch.max(torch.min(x, hi), lo)
                if phi_f < fmin:
                    fmin, best, best_l = phi_f, x.detach().clone(), ell.detach().clone()
            if not math.isfinite(phi_f):
                break
            hist.append(phi_f)
            if dryrun:
                break
        return best, best_l, hist, x.detach().clone(), ell.detach().clone()


def select_optimal(candidates, scores):
    """optimization_based_attack.py:206-218: first minimum wins; all non-finite -> zeros."""
    scores_t = torch.as_tensor(scores, dtype=torch.float32)
    val, idx = torch.min(scores_t, dim=0)
    if val.isfinite():
        return candidates[int(idx)], float(val), int(idx)
    return torch.zeros_like(candidates[int(idx)]), float(val), int(idx)


def pearlmutter_closure(model, loss_fn, gradients, x, labels, kind="pearlmutter-loss", scale=1.0, eps=1e-3, task_regularization=0.0,
                        implementation="forward"):
    """``PearlmutterEuclidean`` / ``PearlmutterCosine`` (objectives.py:279-365, 368-436, 468-493): finite-difference approximation
    of the candidate gradient, restated without the ``candidate.grad +=`` accumulation that breaks under torch >= 2.
    Returns ``(objective_value, task_loss, candidate_gradient)``; the model's parameters are restored (:330-332)."""
    params = list(model.parameters())
    original = [p.detach().clone() for p in params]
    x = x.detach().clone().requires_grad_(True)
    task_loss = loss_fn(model(x), labels)
    *G, dLdx = torch.autograd.grad(task_loss, (*params, x))
    if kind == "pearlmutter-loss":                                      # :463-466
        first = [a - b for a, b in zip(G, gradients)]
        value = 0.5 * scale * sum(r.pow(2).sum() for r in first)
    else:                                                               # :471-478
        sp = sum((a * b).sum() for a, b in zip(G, gradients))
        gn = sum(a.pow(2).sum() for a in G).sqrt()
        dn = sum(b.pow(2).sum() for b in gradients).sqrt()
        first = [b / (-gn * dn) - a * (-sp / (gn.pow(3) * dn)) for a, b in zip(G, gradients)]
        value = scale * (1 - sp / (gn * dn))
    eps_n = eps / sum(a.pow(2).sum() for a in G).sqrt()                 # :348

    def shifted(alpha):
        with torch.no_grad():
            for p, o, v in zip(params, original, first):
                p.copy_(o + alpha * v)
        (d,) = torch.autograd.grad(loss_fn(model(x), labels), (x,))
        return d

    if implementation == "forward":                                     # :349-358
        fd = (shifted(eps_n) - dLdx) / eps_n
    elif implementation == "backward":                                  # :380-389
        fd = (dLdx - shifted(-eps_n)) / eps_n
    elif implementation == "central":                                   # :405-420
        fd = (shifted(0.5 * eps_n) - shifted(-0.5 * eps_n)) / eps_n
    else:
        raise ValueError(implementation)
    with torch.no_grad():
        for p, o in zip(params, original):
            p.copy_(o)
    return value.detach(), task_loss.detach(), (fd * scale + task_regularization * dLdx).detach()


# --------------------------------------------------------------------------------------
# candidate augmentations  (attacks/auxiliaries/augmentations.py; closure: optimization_based_attack.py:149-153)
# --------------------------------------------------------------------------------------
def augment_candidate(x, steps=(), offsets=(), continuous_shift=None, circular=True, uniforms=None, colour_mean=None, colour_std=None):
    """The reference's augmentation modules applied with *given* random draws (the modules draw from torch's global generator):
    ``steps`` = [(kind, param)] with kind 1 = ``Jitter`` (:9-18, roll by ``offsets[s]``), 2 = ``Flip`` (:57-64, flipped iff
    ``offsets[s][0]``), then ``RandomTransform`` (:141-205, ``shift`` = ``continuous_shift``, ``uniforms`` = (randgen[:, 0],
    randgen[:, 1])), then ``ColorJitter`` (:70-89) with per-(image, channel) ``colour_mean`` / ``colour_std`` [N, 3, 1, 1]."""
    for (kind, _), (a, b) in zip(steps, offsets):
        if kind == 1:
            x = torch.roll(x, shifts=(int(a), int(b)), dims=(2, 3))            # :18
        elif kind == 2:
            x = torch.flip(x, dims=(3,)) if a else x                           # :64
    if continuous_shift is not None:
        S = x.shape[2]
        direct = torch.linspace(-1, 1.0, S, dtype=x.dtype).unsqueeze(0).repeat(S, 1).unsqueeze(-1)   # build_grid(S, S): k = 1 (:165-170)
        grid = torch.cat([direct, direct.transpose(1, 0)], dim=2).unsqueeze(0).repeat(x.shape[0], 1, 1, 1)
        randgen = torch.stack([torch.as_tensor(uniforms[0], dtype=x.dtype), torch.as_tensor(uniforms[1], dtype=x.dtype)], dim=1)
        delta = continuous_shift / (S - 1)                                      # :179
        grid[:, :, :, 0] = grid[:, :, :, 0] + ((randgen[:, 0] - 0.5) * 2 * delta)[:, None, None]
        grid[:, :, :, 1] = grid[:, :, :, 1] + ((randgen[:, 1] - 0.5) * 2 * delta)[:, None, None]
        if circular:                                                            # :197-199
            grid = (grid + 1) % 1 - 1
        x = F.grid_sample(x, grid, align_corners=True, mode="bilinear", padding_mode="zeros")   # :203 (self.align = True, :157)
    if colour_mean is not None:
        x = (x - colour_mean) / colour_std                                      # :89
    return x
