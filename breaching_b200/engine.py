"""ctypes binding of ``libbreaching_b200.so`` (C ABI in ``include/breaching_b200.h``).

PyTorch is used here only as plumbing: it owns the tensors whose ``data_ptr()`` is handed to the library.
There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

from . import compiler as C

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libbreaching_b200.so")


class EngineError(RuntimeError):
    pass


class TensorDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32)]


class ParamDesc(ctypes.Structure):
    _fields_ = [("numel", ctypes.c_int64), ("perm", ctypes.c_int32), ("d0", ctypes.c_int32), ("d1", ctypes.c_int32),
                ("d2", ctypes.c_int32)]


class OpDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("kind", "tin", "tout", "res", "R", "S", "stride", "pad", "w", "b", "has_bn", "relu", "gamma", "beta",
                 "bn_buffer")] + [("eps", ctypes.c_float), ("acc_in", ctypes.c_int32), ("acc_res", ctypes.c_int32),
                                  ("bn_train", ctypes.c_int32)]


class AttackCfg(ctypes.Structure):
    _fields_ = [
        ("objective", ctypes.c_int32),
        ("obj_scale", ctypes.c_float), ("task_regularization", ctypes.c_float), ("tag_scale", ctypes.c_float),
        ("mask_value", ctypes.c_float), ("angular_fudge", ctypes.c_float),
        ("optimizer", ctypes.c_int32),
        ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("adam_eps", ctypes.c_float),
        ("weight_decay", ctypes.c_float), ("momentum", ctypes.c_float),
        ("nesterov", ctypes.c_int32), ("signed_mode", ctypes.c_int32), ("boxed", ctypes.c_int32),
        ("max_iterations", ctypes.c_int32),
        ("langevin_noise", ctypes.c_float), ("grad_clip", ctypes.c_float),
        ("noise_seed", ctypes.c_uint64),
        ("tv_scale", ctypes.c_float), ("tv_inner_exp", ctypes.c_float), ("tv_outer_exp", ctypes.c_float),
        ("tv_eps", ctypes.c_float), ("tv_double_opponents", ctypes.c_int32),
        ("norm_scale", ctypes.c_float), ("norm_p", ctypes.c_float),
        ("di_scale", ctypes.c_float), ("di_first_bn_multiplier", ctypes.c_float),
        ("feat_scale", ctypes.c_float),
        ("orthogonality", ctypes.c_int32),
        ("objective_excludes_task", ctypes.c_int32),
    ]


OBJECTIVES = {  # reference objectives.py:496-506
    "euclidean": 0, "cosine-similarity": 1, "l1": 2, "tag-euclidean": 3, "angular": 4,
    "fast-cosine-similarity": 5, "masked-cosine-similarity": 6,
}
# Pearlmutter* (objectives.py:279-365, 468-493) approximate d/dx <grad_W L, v> -- v = d(objective)/dG, *without* the scale --
# by finite differences of grad_x L along W + eps v ("forward" / "backward" / "central"); the engine's tangent sweeps compute
# that directional derivative exactly, i.e. the eps -> 0 limit of all three, at no extra cost.  (The reference versions are
# broken under torch >= 2: `candidate.grad +=` on a None gradient)
PEARLMUTTER = {"pearlmutter-loss": "euclidean", "pearlmutter-cosine": "cosine-similarity"}
OPTIMIZERS = {  # reference common.py:6-17 -> (kind, beta1, beta2, eps, weight_decay, momentum, nesterov)
    "adam": (0, 0.9, 0.999, 1e-8, 0.0, 0.0, 0),
    "adam-safe": (0, 0.5, 0.99, 1e-4, 0.0, 0.0, 0),
    "bert-adam": (1, 0.9, 0.999, 1e-6, 0.01, 0.0, 0),
    "momgd": (2, 0.0, 0.0, 0.0, 0.0, 0.9, 1),
    "gd": (2, 0.0, 0.0, 0.0, 0.0, 0.0, 0),
}

EXPORTS = [
    "bre_engine_create", "bre_engine_destroy", "bre_engine_load_model", "bre_engine_load_targets",
    "bre_engine_load_feature_targets", "bre_engine_set_local_steps", "bre_engine_begin_trial", "bre_engine_run", "bre_engine_run_timed", "bre_engine_sync",
    "bre_engine_status", "bre_engine_read_history", "bre_engine_get_best", "bre_engine_get_candidate",
    "bre_engine_score", "bre_engine_objective_and_gradient", "bre_engine_last_terms", "bre_engine_debug_param",
    "bre_engine_debug_tensor", "bre_engine_launches_per_iteration", "bre_engine_set_option", "bre_match_reduce",
    "bre_total_variation", "bre_conv_gemm", "bre_last_error", "bre_version",
    "bre_engine_load_soft_labels", "bre_engine_label_gradient", "bre_engine_set_labels",
    "bre_token_layernorm", "bre_token_attention", "bre_token_match",
    "bre_engine_param_gradients", "bre_engine_bn_batch_stats", "bre_engine_forward", "bre_image_mse",
    "bre_engine_begin_joint_trial", "bre_engine_get_joint_labels", "bre_resize_bilinear",
    "bre_engine_set_augmentations", "bre_engine_last_augmentation", "bre_augment_view",
]


def load_library(path=None):
    """Load the shared library (no compute is triggered; works without a GPU)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError(
            f"{path} not found: build it with `python -m breaching_b200.build` (there is no CPU/eager fallback)"
        )
    lib = ctypes.CDLL(path)
    lib.bre_last_error.restype = ctypes.c_char_p
    lib.bre_version.restype = ctypes.c_char_p
    lib.bre_engine_destroy.restype = None
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    P = ctypes.POINTER
    lib.bre_engine_create.argtypes = [P(TensorDesc), i32, P(OpDesc), i32, P(ParamDesc), i32, i32, P(AttackCfg), i32, P(vp)]
    lib.bre_engine_destroy.argtypes = [vp]
    lib.bre_engine_load_model.argtypes = [vp, P(vp), i32, P(vp), P(vp), i32]
    lib.bre_engine_load_targets.argtypes = [vp, P(vp), i32, vp, vp, i32, vp, vp, i32]
    lib.bre_engine_load_feature_targets.argtypes = [vp, vp, i64]
    lib.bre_engine_set_local_steps.argtypes = [vp, i32, i32, f32, vp]
    lib.bre_engine_load_soft_labels.argtypes = [vp, vp, i64]
    lib.bre_engine_label_gradient.argtypes = [vp, vp]
    lib.bre_engine_set_labels.argtypes = [vp, vp, i32]
    lib.bre_engine_begin_trial.argtypes = [vp, vp, vp, i32]
    lib.bre_engine_run.argtypes = [vp, i32]
    lib.bre_engine_sync.argtypes = [vp]
    lib.bre_engine_run_timed.argtypes = [vp, i32, P(ctypes.c_float)]
    lib.bre_engine_status.argtypes = [vp, P(i32), P(i32), P(ctypes.c_double), P(ctypes.c_double)]
    lib.bre_engine_read_history.argtypes = [vp, vp, i32]
    lib.bre_engine_get_best.argtypes = [vp, vp]
    lib.bre_engine_get_candidate.argtypes = [vp, vp]
    lib.bre_engine_score.argtypes = [vp, vp, i32, P(ctypes.c_double)]
    lib.bre_engine_objective_and_gradient.argtypes = [vp, vp, P(ctypes.c_double), vp]
    lib.bre_engine_last_terms.argtypes = [vp, P(ctypes.c_double)]
    lib.bre_engine_debug_param.argtypes = [vp, i32, i32, vp]
    lib.bre_engine_debug_tensor.argtypes = [vp, i32, i32, vp]
    lib.bre_engine_launches_per_iteration.argtypes = [vp, P(i32)]
    lib.bre_engine_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.bre_match_reduce.argtypes = [vp, vp, vp, i64, f32, P(ctypes.c_double), vp]
    lib.bre_total_variation.argtypes = [vp, vp, i32, i32, i32, f32, f32, f32, f32, i32, i32, P(ctypes.c_double), vp]
    lib.bre_conv_gemm.argtypes = [i32, i32, vp, vp, vp, vp, vp] + [i32] * 9 + [vp]
    lib.bre_token_layernorm.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp, vp, vp]
    lib.bre_token_attention.argtypes = [i32, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.bre_token_match.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.bre_engine_begin_joint_trial.argtypes = [vp, vp, vp, i64, vp, i32]
    lib.bre_engine_get_joint_labels.argtypes = [vp, i32, vp]
    lib.bre_engine_param_gradients.argtypes = [vp, vp, vp, i32, P(vp), i32, P(ctypes.c_double)]
    lib.bre_engine_bn_batch_stats.argtypes = [vp, i32, vp, vp]
    lib.bre_engine_forward.argtypes = [vp, vp, vp]
    lib.bre_engine_set_augmentations.argtypes = [vp, i32, P(i32), P(f32), i32, f32, i32, vp, vp, i32, ctypes.c_uint64]
    lib.bre_engine_last_augmentation.argtypes = [vp, P(i32), P(i32), P(f32), P(f32)]
    lib.bre_augment_view.argtypes = [vp, vp, i32, i32, i32, i32, i32, P(i32), P(i32), P(i32), f32, i32, P(f32), P(f32), vp, vp, i32, vp, vp]
    lib.bre_resize_bilinear.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.bre_image_mse.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, P(ctypes.c_double), vp]
    for name in EXPORTS:
        if name not in ("bre_last_error", "bre_version", "bre_engine_destroy"):
            getattr(lib, name).restype = ctypes.c_int
    _LIB = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise EngineError(f"{what} failed ({rc}): {lib.bre_last_error().decode()}")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32c(t, device=None):
    t = t.detach()
    if device is not None:
        t = t.to(device)
    return t.to(torch.float32).contiguous()


def make_cfg(cfg_attack, noise_seed=0):
    """Flatten the reference-style attack config into the C struct (values as the reference reads them,
    attacks/optimization_based_attack.py:29-38, :166-184; regularizers.py constructors)."""
    from .config import cfg_get

    c = AttackCfg()
    obj = cfg_attack["objective"]
    kind = obj["type"]
    if kind in PEARLMUTTER:
        impl = cfg_get(obj, "implementation", "forward")
        if impl not in ("forward", "backward", "central"):
            # "upwind" mixes one-sided differences by the sign of dL/dx reduced over the batch axis (objectives.py:440-461): it has
            # no eps -> 0 limit that a directional derivative expresses
            raise EngineError(f"{kind}: finite-difference implementation '{impl}' is not implemented by the engine")
        if cfg_get(obj, "level_gradients", False):
            raise EngineError(f"{kind}: level_gradients is not implemented by the engine")
        kind = PEARLMUTTER[kind]
        c.objective_excludes_task = 1
    if kind not in OBJECTIVES:
        raise ValueError(f"Unknown objective type {kind} given.")
    c.objective = OBJECTIVES[kind]
    c.obj_scale = float(cfg_get(obj, "scale", 1.0))
    c.task_regularization = float(cfg_get(obj, "task_regularization", 0.0) or 0.0)
    c.tag_scale = float(cfg_get(obj, "tag_scale", 0.1))
    c.mask_value = 1e-6  # objectives.py:227 hard-codes 1e-6
    c.angular_fudge = 1e-7  # objectives.py:208
    opt = cfg_attack["optim"]
    name = str(opt["optimizer"]).lower()
    if name not in OPTIMIZERS and name != "l-bfgs":
        raise ValueError(f"Invalid optimizer {opt['optimizer']} given.")
    # L-BFGS trials are driven by attacks/lbfgs.py through objective_and_gradient(); the fused on-device step is unused then
    c.optimizer, c.beta1, c.beta2, c.adam_eps, c.weight_decay, c.momentum, c.nesterov = OPTIMIZERS["gd" if name == "l-bfgs" else name]
    signed = cfg_get(opt, "signed")
    c.signed_mode = {"hard": 1, "soft": 2}.get(signed, 0) if isinstance(signed, str) else 0
    c.boxed = int(bool(cfg_get(opt, "boxed", False)))
    c.max_iterations = int(opt["max_iterations"])
    c.langevin_noise = float(cfg_get(opt, "langevin_noise", 0.0) or 0.0)
    clip = cfg_get(opt, "grad_clip")
    c.grad_clip = -1.0 if clip is None else float(clip)
    c.noise_seed = int(noise_seed) & 0xFFFFFFFFFFFFFFFF
    c.tv_eps, c.tv_inner_exp, c.tv_outer_exp, c.norm_p, c.di_first_bn_multiplier = 1e-8, 1.0, 1.0, 2.0, 10.0
    reg = cfg_get(cfg_attack, "regularization")
    if reg is not None:
        for key in reg.keys():
            r = reg[key]
            if not r["scale"] > 0:
                continue
            if key == "total_variation":
                c.tv_scale = float(r["scale"])
                c.tv_inner_exp = float(cfg_get(r, "inner_exp", 1))
                c.tv_outer_exp = float(cfg_get(r, "outer_exp", 1))
                c.tv_eps = float(cfg_get(r, "eps", 1e-8))
                c.tv_double_opponents = int(bool(cfg_get(r, "double_opponents", False)))
            elif key == "norm":
                c.norm_scale = float(r["scale"])
                c.norm_p = float(cfg_get(r, "pnorm", 2.0))
            elif key == "deep_inversion":
                c.di_scale = float(r["scale"])
                c.di_first_bn_multiplier = float(cfg_get(r, "first_bn_multiplier", 10))
            elif key == "features":
                c.feat_scale = float(r["scale"])
            elif key == "orthogonality":
                c.orthogonality = 1  # regularizers.py:169-178 never multiplies by its scale
            else:
                raise KeyError(key)
    return c


class Engine:
    """One engine = one model replica + one trial state on one GPU."""

    def __init__(self, model, input_shape, cfg_attack, device, noise_seed=0, backend=None, program=None):
        """``backend``: "tc" (tcgen05 TF32 tensor-core GEMMs, default; shapes it does not cover run on the fp32 SIMT
        kernels) or "simt" (fp32 CUDA-core GEMMs everywhere: bit-faithful fp32 products).  Default from
        ``BRE_GEMM_BACKEND``."""
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise EngineError("the engine runs on CUDA devices only (no CPU fallback)")
        self.model = model
        # ``program``: an already lowered layer program (compiler.compile_transformer for token-sequence models, whose
        # candidate is the embedding sequence [batch, seq_len, d]); default: lower the vision model with torch.fx
        self.prog = program if program is not None else C.compile_model(model, input_shape)
        self.input_shape = tuple(int(s) for s in input_shape)
        self.ccfg = make_cfg(cfg_attack, noise_seed)
        prog = self.prog
        tens = (TensorDesc * len(prog.tensors))(*[TensorDesc(t.N, t.C, t.H, t.W) for t in prog.tensors])
        pds = []
        for p in prog.params:
            if p.perm == C.PERM_OIHW_TO_OHWI:
                pds.append(ParamDesc(p.numel, 1, p.shape[0], p.shape[1], p.shape[2] * p.shape[3]))
            elif p.perm == C.PERM_LINEAR_CHW_TO_HWC:
                pds.append(ParamDesc(p.numel, 1, p.shape[0], p.perm_c, p.perm_hw))
            else:   # plain layout; d0 carries the reserved element count when a zero tail is wanted (padded vocabulary)
                pds.append(ParamDesc(p.numel, 0, int(getattr(p, "alloc_numel", 0) or 0), 0, 0))
        self._bn_modules = []
        ops = []
        mods = C.bn_modules(model, prog) if program is None else [None] * len(prog.ops)
        for op, mod in zip(prog.ops, mods):
            bn_idx = -1
            if mod is not None:
                bn_idx = len(self._bn_modules)
                self._bn_modules.append(mod)
            ops.append(OpDesc(op.kind, op.tin, op.tout, op.res, op.R, op.S, op.stride, op.pad, op.w, op.b, int(op.has_bn),
                              int(op.relu), op.gamma, op.beta, bn_idx, float(op.eps), int(op.acc_in), int(op.acc_res),
                              int(getattr(op, "bn_train", False))))
        self._keep = (tens, (OpDesc * len(ops))(*ops), (ParamDesc * len(pds))(*pds))
        handle = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        rc = self.lib.bre_engine_create(self._keep[0], len(prog.tensors), self._keep[1], len(ops), self._keep[2], len(pds),
                                        prog.logits, ctypes.byref(self.ccfg), dev_index, ctypes.byref(handle))
        _check(self.lib, rc, "bre_engine_create")
        self.h = handle
        backend = backend or os.environ.get("BRE_GEMM_BACKEND", "tc")
        if backend not in ("tc", "simt"):
            raise ValueError(f"unknown GEMM backend {backend}")
        self.backend = backend
        self.set_option("gemm_backend", 1 if backend == "tc" else 0)
        valid = getattr(prog, "logits_valid", 0)
        if valid and valid != prog.tensors[prog.logits].C:
            self.set_option("logits_valid", valid)
        self.numel = 1
        for s in self.input_shape:
            self.numel *= s

    def close(self):
        if getattr(self, "h", None):
            self.lib.bre_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---------------------------------------------------------------------------------------------
    def _ptr_array(self, tensors):
        arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        return arr

    def load_model(self, params=None, buffers_from=None):
        """``params``: list of tensors in ``model.parameters()`` order (default: the model's own)."""
        params = [_f32c(p) for p in (params if params is not None else self.model.parameters())]
        mods = self._bn_modules
        # train-mode BN (no buffers anywhere, base_attack.py:192-197) has no running statistics: placeholders, never read
        means = [_f32c(m.running_mean) if m.running_mean is not None else torch.zeros(m.num_features) for m in mods]
        vars_ = [_f32c(m.running_var) if m.running_var is not None else torch.ones(m.num_features) for m in mods]
        torch.cuda.synchronize(self.device) if any(p.is_cuda for p in params) else None
        pa, ma, va = self._ptr_array(params), self._ptr_array(means), self._ptr_array(vars_)
        rc = self.lib.bre_engine_load_model(self.h, pa, len(params), ma, va, len(mods))
        _check(self.lib, rc, "bre_engine_load_model")

    def load_targets(self, gradients, labels, mean=None, std=None, tensor_weights=None):
        grads = [_f32c(g) for g in gradients]
        labels = labels.detach().to(torch.int64).contiguous()
        if any(g.is_cuda for g in grads) or labels.is_cuda:
            torch.cuda.synchronize(self.device)
        tw = None if tensor_weights is None else _f32c(tensor_weights).cpu()
        mean_t = None if mean is None else torch.as_tensor(mean, dtype=torch.float32).flatten().contiguous().cpu()
        std_t = None if std is None else torch.as_tensor(std, dtype=torch.float32).flatten().contiguous().cpu()
        ga = self._ptr_array(grads)
        rc = self.lib.bre_engine_load_targets(self.h, ga, len(grads), _ptr(tw), _ptr(labels), labels.numel(), _ptr(mean_t),
                                              _ptr(std_t), 0 if mean_t is None else mean_t.numel())
        _check(self.lib, rc, "bre_engine_load_targets")

    def load_feature_targets(self, measured):
        """``measured``: [N, F] in torch (flatten) feature order; permuted here if the head sees a spatial map."""
        lin = [op for op in self.prog.ops if op.kind == C.OP_LINEAR][-1]
        ti = self.prog.tensors[lin.tin]
        m = _f32c(measured)
        if ti.H * ti.W > 1:
            m = m.view(ti.N, ti.C, ti.H * ti.W).permute(0, 2, 1).contiguous()
        if m.is_cuda:
            torch.cuda.synchronize(self.device)
        _check(self.lib, self.lib.bre_engine_load_feature_targets(self.h, _ptr(m), m.numel()), "bre_engine_load_feature_targets")

    def set_local_steps(self, total_images, steps, lr, labels_per_step):
        """FedAvg: ``labels_per_step`` = list of ``steps`` LongTensors of length ``data_per_step`` (the program batch)."""
        labels = torch.cat([l.detach().to(torch.int64).flatten().cpu() for l in labels_per_step]).contiguous()
        if labels.numel() != steps * self.input_shape[0]:
            raise EngineError("labels_per_step must hold data_per_step labels for every local step")
        _check(self.lib, self.lib.bre_engine_set_local_steps(self.h, int(total_images), int(steps), float(lr), _ptr(labels)),
               "bre_engine_set_local_steps")
        self.input_shape = (int(total_images), *self.input_shape[1:])
        self.numel = 1
        for s_ in self.input_shape:
            self.numel *= s_

    def begin_trial(self, candidate, lr_table):
        cand = _f32c(candidate)
        if cand.numel() != self.numel:
            raise EngineError(f"candidate has {cand.numel()} elements, engine expects {self.numel}")
        lr = torch.as_tensor(lr_table, dtype=torch.float32).contiguous().cpu()
        if cand.is_cuda:
            torch.cuda.synchronize(self.device)
        _check(self.lib, self.lib.bre_engine_begin_trial(self.h, _ptr(cand), _ptr(lr), lr.numel()), "bre_engine_begin_trial")

    def begin_joint_trial(self, candidate, label_logits, lr_table):
        """Joint data + label optimisation on the device: ``label_logits`` [N, classes] (token models: [batch, seq, vocab])."""
        cand, ell = _f32c(candidate, self.device), _f32c(label_logits, self.device)
        if cand.numel() != self.numel:
            raise EngineError(f"candidate has {cand.numel()} elements, engine expects {self.numel}")
        lr = torch.as_tensor(lr_table, dtype=torch.float32).contiguous().cpu()
        torch.cuda.synchronize(self.device)
        self._label_shape = tuple(label_logits.shape)
        _check(self.lib, self.lib.bre_engine_begin_joint_trial(self.h, _ptr(cand), _ptr(ell), ell.numel(), _ptr(lr), lr.numel()),
               "bre_engine_begin_joint_trial")

    def joint_labels(self, best=True):
        out = torch.empty(self._label_shape, dtype=torch.float32, device=self.device)
        _check(self.lib, self.lib.bre_engine_get_joint_labels(self.h, int(bool(best)), _ptr(out)), "bre_engine_get_joint_labels")
        return out

    def set_augmentations(self, plan):
        """``plan`` (attacks/augment.py ``AugmentationPlan``) or ``None`` to switch augmentations off."""
        if plan is None:
            _check(self.lib, self.lib.bre_engine_set_augmentations(self.h, 0, None, None, 0, 0.0, 0, None, None, 0, 0), "bre_engine_set_augmentations")
            return
        n = len(plan.steps)
        kinds = (ctypes.c_int32 * max(n, 1))(*[k for k, _ in plan.steps])
        params = (ctypes.c_float * max(n, 1))(*[float(p) for _, p in plan.steps])
        scale = None if plan.colour_scale is None else _f32c(plan.colour_scale, self.device)
        shift = None if plan.colour_shift is None else _f32c(plan.colour_shift, self.device)
        torch.cuda.synchronize(self.device)
        self._aug_keep = (scale, shift)
        _check(self.lib, self.lib.bre_engine_set_augmentations(self.h, n, kinds, params, int(plan.continuous_shift is not None),
                                                               float(plan.continuous_shift or 0.0), int(plan.circular), _ptr(scale), _ptr(shift),
                                                               int(plan.differentiable), int(plan.seed) & 0xFFFFFFFFFFFFFFFF),
               "bre_engine_set_augmentations")

    def last_augmentation(self):
        o1, o2 = (ctypes.c_int32 * 4)(), (ctypes.c_int32 * 4)()
        sx, sy = (ctypes.c_float * 64)(), (ctypes.c_float * 64)()
        _check(self.lib, self.lib.bre_engine_last_augmentation(self.h, o1, o2, sx, sy), "bre_engine_last_augmentation")
        n = self.input_shape[0]
        return list(o1), list(o2), list(sx)[:n], list(sy)[:n]

    def run(self, n_iters):
        _check(self.lib, self.lib.bre_engine_run(self.h, int(n_iters)), "bre_engine_run")

    def run_timed(self, n_iters):
        """Run ``n_iters`` iterations and return the device time in milliseconds (CUDA events on the engine stream)."""
        ms = ctypes.c_float()
        _check(self.lib, self.lib.bre_engine_run_timed(self.h, int(n_iters), ctypes.byref(ms)), "bre_engine_run_timed")
        return ms.value

    def sync(self):
        _check(self.lib, self.lib.bre_engine_sync(self.h), "bre_engine_sync")

    def status(self):
        rec, stop = ctypes.c_int32(), ctypes.c_int32()
        fmin, tl = ctypes.c_double(), ctypes.c_double()
        _check(self.lib, self.lib.bre_engine_status(self.h, ctypes.byref(rec), ctypes.byref(stop), ctypes.byref(fmin), ctypes.byref(tl)),
               "bre_engine_status")
        return dict(recorded=rec.value, stopped=bool(stop.value), min_objective=fmin.value, task_loss=tl.value)

    def history(self, n=None):
        n = self.status()["recorded"] if n is None else n
        out = torch.empty(max(n, 1), dtype=torch.float32)
        _check(self.lib, self.lib.bre_engine_read_history(self.h, _ptr(out), n), "bre_engine_read_history")
        return out[:n]

    def best(self, device=None):
        out = torch.empty(self.input_shape, dtype=torch.float32, device=device or self.device)
        torch.cuda.synchronize(self.device)
        _check(self.lib, self.lib.bre_engine_get_best(self.h, _ptr(out)), "bre_engine_get_best")
        return out

    def candidate(self, device=None):
        out = torch.empty(self.input_shape, dtype=torch.float32, device=device or self.device)
        torch.cuda.synchronize(self.device)
        _check(self.lib, self.lib.bre_engine_get_candidate(self.h, _ptr(out)), "bre_engine_get_candidate")
        return out

    def score(self, candidate, scoring):
        cand = _f32c(candidate)
        if cand.is_cuda:
            torch.cuda.synchronize(self.device)
        out = ctypes.c_double()
        _check(self.lib, self.lib.bre_engine_score(self.h, _ptr(cand), OBJECTIVES[scoring], ctypes.byref(out)), "bre_engine_score")
        return out.value

    def objective_and_gradient(self, candidate):
        cand = _f32c(candidate)
        if cand.is_cuda:
            torch.cuda.synchronize(self.device)
        grad = torch.empty(self.input_shape, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        val = ctypes.c_double()
        _check(self.lib, self.lib.bre_engine_objective_and_gradient(self.h, _ptr(cand), ctypes.byref(val), _ptr(grad)),
               "bre_engine_objective_and_gradient")
        return val.value, grad

    def last_terms(self):
        arr = (ctypes.c_double * 6)()
        _check(self.lib, self.lib.bre_engine_last_terms(self.h, arr), "bre_engine_last_terms")
        return dict(zip(("match", "task_loss", "total_variation", "norm", "deep_inversion", "features"), list(arr)))

    # ---- joint data / label optimisation (optimization_with_label_attack.py) --------------------------------
    def load_soft_labels(self, probabilities):
        """Class probabilities [N, classes] as the targets of the task loss (``None`` -> back to index labels)."""
        if probabilities is None:
            _check(self.lib, self.lib.bre_engine_load_soft_labels(self.h, None, 0), "bre_engine_load_soft_labels")
            return
        q = _f32c(probabilities, self.device)
        _check(self.lib, self.lib.bre_engine_load_soft_labels(self.h, _ptr(q), q.numel()), "bre_engine_load_soft_labels")

    def label_gradient(self, shape):
        """d(objective)/d(probabilities) of the last ``objective_and_gradient`` call."""
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        _check(self.lib, self.lib.bre_engine_label_gradient(self.h, _ptr(out)), "bre_engine_label_gradient")
        return out

    def set_labels(self, labels):
        lab = labels.detach().to(device=self.device, dtype=torch.int64).contiguous()
        _check(self.lib, self.lib.bre_engine_set_labels(self.h, _ptr(lab), lab.numel()), "bre_engine_set_labels")

    # ---- the steps either side of the hot path -----------------------------------------------
    def param_gradients(self, data, labels):
        """Gradient of the mean task loss w.r.t. every parameter at ``data`` (cases/users.py:148-156): list of device tensors in
        ``model.parameters()`` order and torch layout, and the loss value."""
        x = _f32c(data, self.device)
        lab = labels.detach().to(device=self.device, dtype=torch.int64).contiguous()
        outs = [torch.empty(p.shape, dtype=torch.float32, device=self.device) for p in self.prog.params]
        torch.cuda.synchronize(self.device)
        loss = ctypes.c_double()
        _check(self.lib, self.lib.bre_engine_param_gradients(self.h, _ptr(x), _ptr(lab), lab.numel(), self._ptr_array(outs), len(outs),
                                                             ctypes.byref(loss)), "bre_engine_param_gradients")
        return outs, loss.value

    def bn_batch_stats(self):
        """(mean, biased variance) per train-mode BN layer of the last forward."""
        out = []
        for j, mod in enumerate(self._bn_modules):
            m = torch.empty(mod.num_features, dtype=torch.float32, device=self.device)
            v = torch.empty_like(m)
            _check(self.lib, self.lib.bre_engine_bn_batch_stats(self.h, j, _ptr(m), _ptr(v)), "bre_engine_bn_batch_stats")
            out.append((m, v))
        return out

    def forward(self, data):
        x = _f32c(data, self.device)
        lt = self.prog.tensors[self.prog.logits]
        out = torch.empty((lt.N, lt.C), dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        _check(self.lib, self.lib.bre_engine_forward(self.h, _ptr(x), _ptr(out)), "bre_engine_forward")
        return out

    def debug_param(self, which, index):
        p = self.prog.params[index]
        out = torch.empty(p.shape, dtype=torch.float32)
        code = {"G": 0, "v": 1, "W": 2, "g": 3}[which]
        _check(self.lib, self.lib.bre_engine_debug_param(self.h, code, index, _ptr(out)), "bre_engine_debug_param")
        return out

    def debug_tensor(self, which, tid):
        t = self.prog.tensors[tid]
        out = torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32)
        code = {"val": 0, "delta": 1, "tangent": 2, "tangent_delta": 3}[which]
        _check(self.lib, self.lib.bre_engine_debug_tensor(self.h, code, tid, _ptr(out)), "bre_engine_debug_tensor")
        return out

    def launches_per_iteration(self):
        out = ctypes.c_int32()
        _check(self.lib, self.lib.bre_engine_launches_per_iteration(self.h, ctypes.byref(out)), "bre_engine_launches_per_iteration")
        return out.value

    def set_option(self, name, value):
        _check(self.lib, self.lib.bre_engine_set_option(self.h, name.encode(), int(value)), "bre_engine_set_option")


# ---- stand-alone kernels ---------------------------------------------------------------------------
def match_reduce(G, g, chunk_weights=None, mask_value=-1.0, readback=True):
    lib = load_library()
    assert G.is_cuda and g.is_cuda and G.dtype == torch.float32 and G.is_contiguous() and g.is_contiguous()
    out = (ctypes.c_double * 5)()
    stream = torch.cuda.current_stream(G.device).cuda_stream
    with torch.cuda.device(G.device):
        rc = lib.bre_match_reduce(_ptr(G), _ptr(g), _ptr(chunk_weights), G.numel(), float(mask_value),
                                  out if readback else None, ctypes.c_void_p(stream))
    _check(lib, rc, "bre_match_reduce")
    return list(out) if readback else None


def total_variation(x, scale=0.1, inner_exp=1.0, outer_exp=1.0, eps=1e-8, double_opponents=False, grad=None):
    lib = load_library()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    accumulate = grad is not None
    grad = torch.empty_like(x) if grad is None else grad
    val = ctypes.c_double()
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        rc = lib.bre_total_variation(_ptr(x), _ptr(grad), x.shape[0], x.shape[2], x.shape[3], scale, inner_exp, outer_exp,
                                     eps, int(double_opponents), int(accumulate), ctypes.byref(val), ctypes.c_void_p(stream))
    _check(lib, rc, "bre_total_variation")
    return val.value, grad


def augment_view(x, steps=(), offsets=(), continuous_shift=None, circular=True, uniforms=None, colour_scale=None, colour_shift=None,
                 transpose=False):
    """Stand-alone augmentation view (or its transpose) with explicit draws: ``steps`` = [(kind, param)], ``offsets`` = [(o1, o2)] per
    step (roll offsets; flip: (flag, 0)), ``uniforms`` = (sx, sy) lists per image for the continuous shift."""
    lib = load_library()
    x = _f32c(x).clone()
    N, C, H, W = x.shape
    out = torch.empty_like(x)
    n = len(steps)
    kinds = (ctypes.c_int32 * max(n, 1))(*[k for k, _ in steps])
    o1 = (ctypes.c_int32 * max(n, 1))(*[int(a) for a, _ in offsets])
    o2 = (ctypes.c_int32 * max(n, 1))(*[int(b) for _, b in offsets])
    sx = sy = None
    if continuous_shift is not None:
        sx, sy = (ctypes.c_float * N)(*[float(v) for v in uniforms[0]]), (ctypes.c_float * N)(*[float(v) for v in uniforms[1]])
    scratch = torch.empty_like(x)
    cs, csh = (None if colour_scale is None else _f32c(colour_scale, x.device)), (None if colour_shift is None else _f32c(colour_shift, x.device))
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        torch.cuda.synchronize(x.device)
        rc = lib.bre_augment_view(_ptr(x), _ptr(out), N, C, H, W, n, kinds, o1, o2, float(continuous_shift or 0.0), int(circular), sx, sy, _ptr(cs),
                                  _ptr(csh), int(transpose), _ptr(scratch), ctypes.c_void_p(stream))
    _check(lib, rc, "bre_augment_view")
    return out


def resize_bilinear(x, size):
    """``F.interpolate(x, size=size, mode="bilinear", align_corners=False)`` for an NCHW fp32 batch on the device."""
    lib = load_library()
    assert x.is_cuda and x.dim() == 4
    x = _f32c(x)
    Ho, Wo = (int(size), int(size)) if not isinstance(size, (tuple, list)) else (int(size[0]), int(size[1]))
    out = torch.empty((x.shape[0], x.shape[1], Ho, Wo), dtype=torch.float32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        rc = lib.bre_resize_bilinear(_ptr(x), _ptr(out), x.shape[0], x.shape[1], x.shape[2], x.shape[3], Ho, Wo, ctypes.c_void_p(stream))
    _check(lib, rc, "bre_resize_bilinear")
    return out


def image_mse(rec, ref, mean=None, std=None, clamp=True):
    """Per-example MSE of the de-normalised, [0, 1]-clamped batches (analysis/analysis.py:228-242) -> list of floats."""
    lib = load_library()
    assert rec.is_cuda and rec.shape == ref.shape and rec.dim() == 4
    rec, ref = _f32c(rec), _f32c(ref, rec.device)
    N, C, H, W = rec.shape
    mean_a = std_a = None
    if mean is not None:
        mean_a = (ctypes.c_float * C)(*[float(v) for v in torch.as_tensor(mean).flatten().tolist()])
        std_a = (ctypes.c_float * C)(*[float(v) for v in torch.as_tensor(std).flatten().tolist()])
    out = (ctypes.c_double * N)()
    stream = torch.cuda.current_stream(rec.device).cuda_stream
    with torch.cuda.device(rec.device):
        rc = lib.bre_image_mse(_ptr(rec), _ptr(ref), N, C, H * W, mean_a, std_a, int(bool(clamp)), out, ctypes.c_void_p(stream))
    _check(lib, rc, "bre_image_mse")
    return list(out)


def conv_gemm(mode, a, w, out, N, H, W, Ci, Co, R, S, stride, pad, a2=None, w2=None, backend=0):
    """mode: 0 fprop / 1 dgrad / 2 wgrad on NHWC / OHWI device tensors (see the header)."""
    lib = load_library()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    with torch.cuda.device(a.device):
        rc = lib.bre_conv_gemm(mode, backend, _ptr(a), _ptr(w), _ptr(a2), _ptr(w2), _ptr(out), N, H, W, Ci, Co, R, S, stride,
                               pad, ctypes.c_void_p(stream))
    _check(lib, rc, "bre_conv_gemm")
    return out


def token_layernorm(sweep, x, gamma, beta, stats, in1=None, in2=None, in3=None, v_gamma=None, v_beta=None, eps=1e-5, want_param_grad=False):
    """Stand-alone LayerNorm sweep (csrc/tokens.cu) on [rows, C] device tensors; returns ``out`` (and the gamma / beta
    gradients for sweep 1 with ``want_param_grad``)."""
    lib = load_library()
    rows, C = x.shape
    out = torch.empty_like(x)
    gg = torch.empty(C, device=x.device) if want_param_grad else None
    gb = torch.empty(C, device=x.device) if want_param_grad else None
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        rc = lib.bre_token_layernorm(sweep, _ptr(x), _ptr(in1), _ptr(in2), _ptr(in3), _ptr(gamma), _ptr(beta), _ptr(v_gamma), _ptr(v_beta),
                                     float(eps), rows, C, _ptr(stats), _ptr(out), _ptr(gg), _ptr(gb), ctypes.c_void_p(stream))
    _check(lib, rc, "bre_token_layernorm")
    return (out, gg, gb) if want_param_grad else out


def token_match(rec, emb, subset=None):
    """Nearest vocabulary embedding of every row of ``rec`` [rows, d] under the reference's centred similarity (base_attack.py:126-133);
    ``subset`` (int64 ids) restricts the candidates and the result indexes into it.  Device tensors; returns int64 [rows]."""
    lib = load_library()
    rec = rec.detach().to(torch.float32).contiguous()
    emb = emb.detach().to(torch.float32).contiguous()
    if subset is not None:
        subset = subset.detach().to(torch.int64).contiguous()
    V = int(subset.numel()) if subset is not None else int(emb.shape[0])
    out = torch.empty(rec.shape[0], dtype=torch.int64, device=rec.device)
    stream = torch.cuda.current_stream(rec.device).cuda_stream
    with torch.cuda.device(rec.device):
        rc = lib.bre_token_match(_ptr(rec), _ptr(emb), _ptr(subset), int(rec.shape[0]), int(rec.shape[1]), V, _ptr(out), ctypes.c_void_p(stream))
    _check(lib, rc, "bre_token_match")
    return out


def token_attention(sweep, qkv, B, T, heads, P, Pd, in1=None, in2=None, in3=None):
    """Stand-alone multi-head self-attention sweep (csrc/tokens.cu); qkv [B*T, 3 d]."""
    lib = load_library()
    d = qkv.shape[1] // 3
    out = torch.empty(qkv.shape[0], d if sweep in (0, 2) else 3 * d, device=qkv.device)
    stream = torch.cuda.current_stream(qkv.device).cuda_stream
    with torch.cuda.device(qkv.device):
        rc = lib.bre_token_attention(sweep, _ptr(qkv), _ptr(in1), _ptr(in2), _ptr(in3), B, T, heads, d // heads, _ptr(P), _ptr(Pd), _ptr(out),
                                     ctypes.c_void_p(stream))
    _check(lib, rc, "bre_token_attention")
    return out
