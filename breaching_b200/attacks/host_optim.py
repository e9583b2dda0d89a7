"""Host-driven optimiser steps for the attack variants whose loop cannot use the fused on-device step (several leaves:
joint data + label optimisation; several models per candidate: multi-query attacks).  The closure evaluations stay on the
engine; these are the update rules of the torch optimisers ``optimizer_lookup`` constructs (common.py:5-18) applied to small
device tensors."""
import math

import torch

# common.py:6-17 -> (kind, beta1, beta2, eps, weight_decay, momentum, nesterov)
OPTIMIZERS = {
    "adam": ("adam", 0.9, 0.999, 1e-8, 0.0, 0.0, False),
    "adam-safe": ("adam", 0.5, 0.99, 1e-4, 0.0, 0.0, False),
    "bert-adam": ("adamw", 0.9, 0.999, 1e-6, 0.01, 0.0, False),
    "momgd": ("sgd", 0.0, 0.0, 0.0, 0.0, 0.9, True),
    "gd": ("sgd", 0.0, 0.0, 0.0, 0.0, 0.0, False),
}


class LeafOptimizer:
    """torch.optim.Adam / AdamW / SGD update rules (the classes ``optimizer_lookup`` builds) for a list of device leaves."""

    def __init__(self, leaves, name):
        self.kind, self.b1, self.b2, self.eps, self.wd, self.mom, self.nesterov = OPTIMIZERS[name]
        self.leaves = leaves
        self.m = [torch.zeros_like(p) for p in leaves]
        self.v = [torch.zeros_like(p) for p in leaves]
        self.t = 0

    def step(self, grads, lr):
        self.t += 1
        for p, g, m, v in zip(self.leaves, grads, self.m, self.v):
            if self.kind in ("adam", "adamw"):
                if self.kind == "adamw":
                    p.mul_(1 - lr * self.wd)
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
                denom = v.sqrt().div_(math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-lr / bc1)
            else:
                d = g
                if self.mom != 0:
                    if self.t == 1:
                        m.copy_(g)
                    else:
                        m.mul_(self.mom).add_(g)
                    d = g.add(m, alpha=self.mom) if self.nesterov else m
                p.add_(d, alpha=-lr)
