"""Four-sweep formulation of the reference's TAG / transformer hot path (BASELINE config 5)
in plain torch on the CPU.  TEST INFRASTRUCTURE ONLY -- and the specification of the attention / LayerNorm / token-level
soft-label kernels the engine does not have yet.

Model: the reference's ``TransformerModel`` (cases/models/language_models.py:150-205) as the attack sees it -- the token
embedding replaced by ``Identity`` (base_attack.py:76-128, "run-embedding"), i.e. the candidate *is* the embedding sequence
``x [N, T, d]``; learnable positional embedding added (:133-146); ``nn.TransformerEncoder`` of post-norm layers (self-attention
without mask -- ``has_mask`` is never passed, :183-191 --, ReLU feed-forward, dropout 0); linear decoder; ``CausalLoss``
(losses.py:7-26) with the joint attacker's soft labels ``labels.softmax(-1)`` shifted by one position.

As for the convolutional path (oracle/program_interp.py) the second backward is evaluated as the
weight-direction tangent of the first:  with ``G = grad_W L`` and ``v = d objective / d G``,

    d objective / d x = d/d eps  grad_x L(x, q, W + eps v),      d objective / d q = d/d eps  grad_q L(x, q, W + eps v)

which needs a forward sweep (F), a backward sweep (B: G and the deltas), a tangent-forward sweep (TF: tangents of every
activation for dW = v, dx = dq = 0) and a tangent-backward sweep (TB: tangents of the deltas).  Rules per op:

  linear  y = a W^T + b     B: da = dy W, G_W = dy^T a, G_b = sum dy       TF: y' = a' W^T + a v_W^T + v_b
                            TB: da' = dy' W + dy v_W
  layer norm  y = g xh + b  B: u = t - mean(t) - xh mean(t xh), t = dy g;  dx = u / sigma;  G_g = sum dy xh, G_b = sum dy
                            TF: xh' = (x' - mean(x') - xh mean(xh x')) / sigma;   y' = v_g xh + g xh' + v_b
                            TB: t' = dy' g + dy v_g;  u' = t' - mean(t') - xh' mean(t xh) - xh mean(t' xh + t xh')
                                dx' = u' / sigma - u mean(xh x') / sigma^2
  attention  S = Q K^T s, P = softmax(S), O = P V
                            B: dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(dP P)), dQ = dS K s, dK = dS^T Q s
                            TF: S' = (Q' K^T + Q K'^T) s, P' = P (S' - rowsum(P S')), O' = P' V + P V'
                            TB: dV' = P'^T dO + P^T dO';  dP' = dO' V^T + dO V'^T
                                dS' = P' (dP - r) + P (dP' - r'),  r = rowsum(dP P), r' = rowsum(dP' P + dP P')
                                dQ' = (dS' K + dS K') s;  dK' = (dS'^T Q + dS^T Q') s
  soft-label CE             B: dz = (p - q) / M      TB seed: dz' = p (z' - <p, z'>) / M      d objective/dq = -(z' - <p, z'>) / M

``tests/test_transformer_interp.py`` checks all of it in float64 against autograd's double backward through the actual
``nn.TransformerEncoder`` modules.
"""
import math

import torch


def _ln_forward(x, gamma, beta, eps):
    mu = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    inv = 1.0 / torch.sqrt(var + eps)
    xh = (x - mu) * inv
    return gamma * xh + beta, xh, inv


def _ln_backward(dy, xh, inv, gamma):
    t = dy * gamma
    u = t - t.mean(dim=-1, keepdim=True) - xh * (t * xh).mean(dim=-1, keepdim=True)
    return u * inv, (dy * xh).sum(dim=(0, 1)), dy.sum(dim=(0, 1)), t, u


def _ln_tangent_forward(xd, xh, inv, gamma, vg, vb):
    xhd = (xd - xd.mean(dim=-1, keepdim=True) - xh * (xh * xd).mean(dim=-1, keepdim=True)) * inv
    return vg * xh + gamma * xhd + vb, xhd


def _ln_tangent_backward(dyd, dy, t, u, xd, xh, xhd, inv, gamma, vg):
    td = dyd * gamma + dy * vg
    ud = td - td.mean(dim=-1, keepdim=True) - xhd * (t * xh).mean(dim=-1, keepdim=True) \
        - xh * (td * xh + t * xhd).mean(dim=-1, keepdim=True)
    return ud * inv - u * (xh * xd).mean(dim=-1, keepdim=True) * inv * inv


class TransformerFourSweep:
    """Sweeps over ``model`` (a ``synthetic.TransformerLM``-shaped module: ``pos_embedding``, ``layers`` = the
    ``nn.TransformerEncoderLayer`` list, ``decoder``); parameters are read in ``attack_parameters()`` order."""

    def __init__(self, model, dtype=torch.float64):
        self.dtype = dtype
        self.h = model.layers[0].self_attn.num_heads
        self.eps = model.layers[0].norm1.eps
        g = lambda p: p.detach().to(dtype)  # noqa: E731
        self.pos = g(model.pos_embedding.weight)
        self.L = []
        for layer in model.layers:
            self.L.append(dict(Win=g(layer.self_attn.in_proj_weight), bin=g(layer.self_attn.in_proj_bias),
                               Wo=g(layer.self_attn.out_proj.weight), bo=g(layer.self_attn.out_proj.bias),
                               W1=g(layer.linear1.weight), b1=g(layer.linear1.bias), W2=g(layer.linear2.weight), b2=g(layer.linear2.bias),
                               g1=g(layer.norm1.weight), be1=g(layer.norm1.bias), g2=g(layer.norm2.weight), be2=g(layer.norm2.bias)))
        self.Wd, self.bd = g(model.decoder.weight), g(model.decoder.bias)

    # parameter order of the flat lists G / V used below (== model.attack_parameters())
    KEYS = ("Win", "bin", "Wo", "bo", "W1", "b1", "W2", "b2", "g1", "be1", "g2", "be2")

    def _pack(self, pos, layers, Wd, bd):
        out = [pos]
        for d in layers:
            out += [d[k] for k in self.KEYS]
        return out + [Wd, bd]

    def _unpack(self, flat):
        pos, rest = flat[0], flat[1:]
        layers = []
        for i in range(len(self.L)):
            layers.append(dict(zip(self.KEYS, rest[i * 12:(i + 1) * 12])))
        return pos, layers, rest[-2], rest[-1]

    def _heads(self, t):
        N, T, d = t.shape
        return t.view(N, T, self.h, d // self.h).transpose(1, 2)   # [N, h, T, dh]

    def _merge(self, t):
        N, h, T, dh = t.shape
        return t.transpose(1, 2).reshape(N, T, h * dh)

    # ------------------------------------------------------------------------------------------ F
    def forward(self, x, q):
        x, q = x.to(self.dtype), q.to(self.dtype)
        N, T, d = x.shape
        s = 1.0 / math.sqrt(d // self.h)
        a = x + self.pos[:T]
        st = []
        for p in self.L:
            c = dict(a=a)
            qkv = a @ p["Win"].t() + p["bin"]
            c["Q"], c["K"], c["V"] = (self._heads(t) for t in qkv.split(d, dim=-1))
            c["P"] = torch.softmax(c["Q"] @ c["K"].transpose(-1, -2) * s, dim=-1)
            c["O"] = self._merge(c["P"] @ c["V"])
            r1 = a + c["O"] @ p["Wo"].t() + p["bo"]
            c["n1"], c["xh1"], c["inv1"] = _ln_forward(r1, p["g1"], p["be1"], self.eps)
            c["f1"] = c["n1"] @ p["W1"].t() + p["b1"]
            c["hid"] = torch.relu(c["f1"])
            r2 = c["n1"] + c["hid"] @ p["W2"].t() + p["b2"]
            a, c["xh2"], c["inv2"] = _ln_forward(r2, p["g2"], p["be2"], self.eps)
            st.append(c)
        z = a @ self.Wd.t() + self.bd
        logp = torch.log_softmax(z[:, :-1], dim=-1)
        M = N * (T - 1)
        self.st, self.a_last, self.p, self.q, self.M, self.s = st, a, logp.exp(), q[:, 1:], M, s
        self.loss = -(self.q * logp).sum(dim=-1).sum() / M
        return self.loss

    # ------------------------------------------------------------------------------------------ B
    def backward(self):
        N, T, d = self.a_last.shape
        dz = torch.zeros(N, T, self.Wd.shape[0], dtype=self.dtype)
        dz[:, :-1] = (self.p - self.q) / self.M
        self.dz = dz
        GWd, Gbd = dz.flatten(0, 1).t() @ self.a_last.flatten(0, 1), dz.sum(dim=(0, 1))
        da = dz @ self.Wd
        Gl = []
        for p, c in zip(reversed(self.L), reversed(self.st)):
            g = {}
            c["dy2"] = da
            dr2, g["g2"], g["be2"], c["t2"], c["u2"] = _ln_backward(da, c["xh2"], c["inv2"], p["g2"])
            c["dff"] = dr2                                           # delta of (hid W2^T + b2) and of the residual n1
            g["W2"], g["b2"] = dr2.flatten(0, 1).t() @ c["hid"].flatten(0, 1), dr2.sum(dim=(0, 1))
            c["df1"] = (dr2 @ p["W2"]) * (c["f1"] > 0)
            g["W1"], g["b1"] = c["df1"].flatten(0, 1).t() @ c["n1"].flatten(0, 1), c["df1"].sum(dim=(0, 1))
            dn1 = dr2 + c["df1"] @ p["W1"]
            c["dy1"] = dn1
            dr1, g["g1"], g["be1"], c["t1"], c["u1"] = _ln_backward(dn1, c["xh1"], c["inv1"], p["g1"])
            c["dattn"] = dr1
            g["Wo"], g["bo"] = dr1.flatten(0, 1).t() @ c["O"].flatten(0, 1), dr1.sum(dim=(0, 1))
            dO = self._heads(dr1 @ p["Wo"])
            c["dO"] = dO
            dV = c["P"].transpose(-1, -2) @ dO
            dP = dO @ c["V"].transpose(-1, -2)
            c["dP"], c["r"] = dP, (dP * c["P"]).sum(dim=-1, keepdim=True)
            dS = c["P"] * (dP - c["r"])
            c["dS"] = dS
            dQ, dK = dS @ c["K"] * self.s, dS.transpose(-1, -2) @ c["Q"] * self.s
            dqkv = torch.cat([self._merge(dQ), self._merge(dK), self._merge(dV)], dim=-1)
            c["dqkv"] = dqkv
            g["Win"], g["bin"] = dqkv.flatten(0, 1).t() @ c["a"].flatten(0, 1), dqkv.sum(dim=(0, 1))
            da = dr1 + dqkv @ p["Win"]
            Gl.append(g)
        Gl.reverse()
        Gpos = torch.zeros_like(self.pos)
        Gpos[:T] = da.sum(dim=0)
        self.dx = da
        self.G = self._pack(Gpos, Gl, GWd, Gbd)
        return self.G

    # ------------------------------------------------------------------------------------------ TF
    def tangent_forward(self, V):
        vpos, vl, vWd, vbd = self._unpack([v.to(self.dtype) for v in V])
        N, T, d = self.a_last.shape
        ad = vpos[:T].expand(N, T, d)                      # tangent of x (and of the soft labels) is zero
        for p, v, c in zip(self.L, vl, self.st):
            c["ad"] = ad
            qkvd = ad @ p["Win"].t() + c["a"] @ v["Win"].t() + v["bin"]
            Qd, Kd, Vd = (self._heads(t) for t in qkvd.split(d, dim=-1))
            Sd = (Qd @ c["K"].transpose(-1, -2) + c["Q"] @ Kd.transpose(-1, -2)) * self.s
            Pd = c["P"] * (Sd - (c["P"] * Sd).sum(dim=-1, keepdim=True))
            Od = self._merge(Pd @ c["V"] + c["P"] @ Vd)
            c.update(Qd=Qd, Kd=Kd, Vd=Vd, Pd=Pd, Od=Od)
            r1d = ad + Od @ p["Wo"].t() + c["O"] @ v["Wo"].t() + v["bo"]
            c["r1d"] = r1d
            n1d, c["xh1d"] = _ln_tangent_forward(r1d, c["xh1"], c["inv1"], p["g1"], v["g1"], v["be1"])
            c["n1d"] = n1d
            f1d = n1d @ p["W1"].t() + c["n1"] @ v["W1"].t() + v["b1"]
            hidd = f1d * (c["f1"] > 0)
            c["hidd"] = hidd
            r2d = n1d + hidd @ p["W2"].t() + c["hid"] @ v["W2"].t() + v["b2"]
            c["r2d"] = r2d
            ad, c["xh2d"] = _ln_tangent_forward(r2d, c["xh2"], c["inv2"], p["g2"], v["g2"], v["be2"])
        self.zd = ad @ self.Wd.t() + self.a_last @ vWd.t() + vbd
        self.V = (vpos, vl, vWd, vbd)
        return self.zd

    # ------------------------------------------------------------------------------------------ TB
    def tangent_backward(self):
        """Returns (d objective / d x  [N, T, d],  d objective / d q  [N, T, vocab] -- zero at position 0)."""
        vpos, vl, vWd, vbd = self.V
        N, T, d = self.a_last.shape
        zd = self.zd[:, :-1]
        centred = zd - (self.p * zd).sum(dim=-1, keepdim=True)
        dzd = torch.zeros_like(self.dz)
        dzd[:, :-1] = self.p * centred / self.M
        dq = torch.zeros(N, T, self.Wd.shape[0], dtype=self.dtype)
        dq[:, 1:] = -centred / self.M
        dad = dzd @ self.Wd + self.dz @ vWd
        for p, v, c in zip(reversed(self.L), reversed(vl), reversed(self.st)):
            dr2d = _ln_tangent_backward(dad, c["dy2"], c["t2"], c["u2"], c["r2d"], c["xh2"], c["xh2d"], c["inv2"], p["g2"], v["g2"])
            df1d = (dr2d @ p["W2"] + c["dff"] @ v["W2"]) * (c["f1"] > 0)
            dn1d = dr2d + df1d @ p["W1"] + c["df1"] @ v["W1"]
            dr1d = _ln_tangent_backward(dn1d, c["dy1"], c["t1"], c["u1"], c["r1d"], c["xh1"], c["xh1d"], c["inv1"], p["g1"], v["g1"])
            dOd = self._heads(dr1d @ p["Wo"] + c["dattn"] @ v["Wo"])
            dVd = c["Pd"].transpose(-1, -2) @ c["dO"] + c["P"].transpose(-1, -2) @ dOd
            dPd = dOd @ c["V"].transpose(-1, -2) + c["dO"] @ c["Vd"].transpose(-1, -2)
            rd = (dPd * c["P"] + c["dP"] * c["Pd"]).sum(dim=-1, keepdim=True)
            dSd = c["Pd"] * (c["dP"] - c["r"]) + c["P"] * (dPd - rd)
            dQd = (dSd @ c["K"] + c["dS"] @ c["Kd"]) * self.s
            dKd = (dSd.transpose(-1, -2) @ c["Q"] + c["dS"].transpose(-1, -2) @ c["Qd"]) * self.s
            dqkvd = torch.cat([self._merge(dQd), self._merge(dKd), self._merge(dVd)], dim=-1)
            dad = dr1d + dqkvd @ p["Win"] + c["dqkv"] @ v["Win"]
        return dad, dq
