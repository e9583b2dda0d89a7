"""The four-sweep formulation for the transformer / TAG path (oracle/transformer_interp.py: attention, LayerNorm, token-level
soft-label cross-entropy, label-leaf gradient) against autograd's double backward through the real ``nn.TransformerEncoder``
modules, in float64.  This is the CPU specification of the kernels SURVEY section 8 rows a15 / a16 still need."""
import pytest
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

from breaching_b200 import synthetic
from oracle.transformer_interp import TransformerFourSweep


def _relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


@pytest.mark.parametrize("nlayers,T", [(1, 5), (3, 8)])
def test_transformer_four_sweeps_match_double_backward(nlayers, T):
    torch.manual_seed(3)
    N, vocab, d, heads, dff = 2, 37, 16, 4, 24
    model = synthetic.TransformerLM(vocab, d, heads, dff, nlayers, max_positions=16).double().eval()
    with torch.no_grad():  # non-trivial LayerNorm affine parameters and biases
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    params = model.attack_parameters()
    g = [0.3 * torch.randn_like(p) for p in params]                        # the "shared gradient" to be matched
    x = torch.randn(N, T, d, dtype=torch.double, requires_grad=True)      # candidate in embedding space
    q = torch.softmax(torch.randn(N, T, vocab, dtype=torch.double), dim=-1).requires_grad_(True)   # labels.softmax(-1)

    with sdpa_kernel(SDPBackend.MATH):   # the fused CPU attention kernel has no double backward
        loss = synthetic.causal_loss(model(inputs_embeds=x), q)
        G = torch.autograd.grad(loss, params, create_graph=True)
        phi = 0.5 * sum((a - b).pow(2).sum() for a, b in zip(G, g))
        dx_ref, dq_ref = torch.autograd.grad(phi, [x, q])

    fs = TransformerFourSweep(model)
    assert abs(float(fs.forward(x.detach(), q.detach())) - float(loss)) < 1e-12
    G_fs = fs.backward()
    assert len(G_fs) == len(G)
    for a, b in zip(G_fs, G):
        assert a.shape == b.shape and _relerr(a, b.detach()) < 1e-10
    V = [a - b for a, b in zip(G_fs, g)]                                    # d phi / d G for the euclidean objective
    fs.tangent_forward(V)
    dx, dq = fs.tangent_backward()
    assert _relerr(dx, dx_ref) < 1e-9, _relerr(dx, dx_ref)
    assert _relerr(dq[:, 1:], dq_ref[:, 1:]) < 1e-9, _relerr(dq[:, 1:], dq_ref[:, 1:])
    assert dq[:, 0].abs().max().item() == 0.0 and dq_ref[:, 0].abs().max().item() == 0.0   # position 0 is never a target


def test_transformer_four_sweeps_reproduce_the_reference_tag_closure():
    """Same formulation against the unmodified reference: the TAG joint attacker's closure (tag.yaml: tag-euclidean objective,
    candidate in embedding space, token-level soft labels) on the miniature config-5 fixture -- objective value, gradient
    w.r.t. the candidate embeddings and w.r.t. the label logits (chained through the softmax)."""
    import sys, os

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import case_from_fixture, cfg_from_fixture, load_golden
    from oracle.program_interp import objective_direction

    fx = load_golden("trial_joint_tag_transformer.pt")
    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    names = [n for n, _ in model.named_parameters()]
    g = [t.double() for t in shared[0]["gradients"]]
    g.pop(names.index("encoder.weight"))
    fs = TransformerFourSweep(model.double())
    q = fx["l0"].double().softmax(dim=-1)
    fs.forward(fx["x0"].double(), q)
    G = fs.backward()
    o = cfg.objective
    val, V = objective_direction(o.type, G, g, scale=o.scale, tag_scale=o.tag_scale, scale_scheme=o.scale_scheme)
    assert abs(float(val) - fx["objective0"]) < 1e-5 * abs(fx["objective0"])
    fs.tangent_forward(V)
    dx, dq = fs.tangent_backward()
    dl = q * (dq - (q * dq).sum(dim=-1, keepdim=True))
    assert _relerr(dx.float(), fx["raw_grad_x0"]) < 1e-4, _relerr(dx.float(), fx["raw_grad_x0"])
    assert _relerr(dl.float(), fx["raw_grad_l0"]) < 1e-4, _relerr(dl.float(), fx["raw_grad_l0"])


def test_transformer_layer_program_reproduces_the_reference_tag_closure():
    """``compiler.compile_transformer`` (posadd / linear / attention / layernorm / residual / ReLU ops over row x feature
    tensors) executed by the generic four-sweep program interpreter -- the contract the CUDA engine will implement for
    config 5 -- against the reference's TAG closure on the fixture."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import case_from_fixture, cfg_from_fixture, load_golden
    from breaching_b200 import compiler
    from oracle import program_interp as PI

    fx = load_golden("trial_joint_tag_transformer.pt")
    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    names = [n for n, _ in model.named_parameters()]
    g = [t.double() for t in shared[0]["gradients"]]
    g.pop(names.index("encoder.weight"))
    B, T, d = fx["x0"].shape
    prog = compiler.compile_transformer(model, B, T, pad_vocab=False)   # the torch interpreter runs the un-padded program
    assert prog.seq_len == T and len(prog.params) == len(g)

    class _Params:  # the interpreter reads parameters in program order (token embedding removed)
        def parameters(self):
            return [p for n, p in model.named_parameters() if n != "encoder.weight"]

        def named_modules(self):
            return model.named_modules()

    it = PI.ProgramInterpreter(_Params(), prog)
    q = fx["l0"].double().softmax(dim=-1)
    o = cfg.objective
    val, dx, loss, G = it.matching_gradient(fx["x0"].double(), q, g, o.type, scale=o.scale, tag_scale=o.tag_scale,
                                            scale_scheme=o.scale_scheme)
    assert abs(float(val) - fx["objective0"]) < 1e-5 * abs(fx["objective0"])
    assert _relerr(dx.float(), fx["raw_grad_x0"]) < 1e-4
    dl = q * (it.dq - (q * it.dq).sum(dim=-1, keepdim=True))
    assert _relerr(dl.float(), fx["raw_grad_l0"]) < 1e-4
    # and the direct restatement agrees to round-off
    fs = TransformerFourSweep(model.double())
    fs.forward(fx["x0"].double(), q)
    for a, b in zip(G, fs.backward()):
        assert _relerr(a, b) < 1e-10


def test_config5_program_has_the_survey_dimensions():
    """BASELINE config 5 (TransformerModel(50257, 96, 8, 1536, 3), 32 tokens): 39 gradient tensors / 5 975 761 parameters once
    the token embedding is removed, lowered to 3 x (attention + 2 LayerNorm + 4 GEMMs) + decoder."""
    from breaching_b200 import compiler

    model = synthetic.TransformerLM(50257, 96, 8, 1536, 3)
    params = model.attack_parameters()
    assert len(params) == 39 and sum(p.numel() for p in params) == 5_975_761
    prog = compiler.compile_transformer(model, 1, 32)
    kinds = [op.kind for op in prog.ops]
    assert kinds.count(compiler.OP_ATTENTION) == 3 and kinds.count(compiler.OP_LAYERNORM) == 6 and kinds.count(compiler.OP_POSADD) == 1
    assert kinds.count(compiler.OP_LINEAR) == 13 and len(prog.params) == 39
    # the vocabulary is padded to the GEMM tile width for the engine (50 257 -> 50 304 logit columns, 50 257 real classes)
    assert prog.tensors[prog.logits].N == 32 and prog.tensors[prog.logits].C == 50304 and prog.logits_valid == 50257 and prog.seq_len == 32
    dec = [p for p in prog.params if p.shape == (50257, 96)][0]
    assert dec.alloc_numel == 50304 * 96 and [p for p in prog.params if p.shape == (50257,)][0].alloc_numel == 50304
    plain = compiler.compile_transformer(model, 1, 32, pad_vocab=False)
    assert plain.tensors[plain.logits].C == 50257 and all(p.alloc_numel == 0 for p in plain.params)
    assert all(op.S == 32 for op in prog.ops if op.kind in (compiler.OP_ATTENTION, compiler.OP_POSADD))


def _config5_inputs(fx):
    gen = torch.Generator().manual_seed(fx["l0_seed"])
    x0 = torch.randn(list(fx["x0"].shape), generator=gen)
    l0 = torch.randn([fx["x0"].shape[0], fx["x0"].shape[1], fx["case"]["ntokens"]], generator=gen)
    assert torch.equal(x0, fx["x0"]) and abs(float(l0.double().sum()) - fx["l0_checksum"]) < 1e-6 * fx["l0_abs_checksum"]
    return x0, l0


def test_full_size_config5_closure_through_the_layer_program():
    """BASELINE config 5 at full size (50 257 tokens, 96 dims, 8 heads, 3 layers, 32 positions): the lowered layer program
    evaluated by the four-sweep interpreter against the closure of the reference's TAG attacker (objective, candidate
    gradient, sampled label-logit gradient + its norm)."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import case_from_fixture, cfg_from_fixture, load_golden
    from breaching_b200 import compiler
    from oracle import program_interp as PI
    from oracle.program_interp import objective_direction  # noqa: F401

    fx = load_golden("trial_joint_tag_config5.pt")
    x0, l0 = _config5_inputs(fx)
    model, loss_fn, payload, shared, true = case_from_fixture(fx)
    cfg = cfg_from_fixture(fx)
    names = [n for n, _ in model.named_parameters()]
    g = [t.double() for t in shared[0]["gradients"]]
    g.pop(names.index("encoder.weight"))
    prog = compiler.compile_transformer(model, 1, 32, pad_vocab=False)

    class _Params:
        def parameters(self):
            return [p for n, p in model.named_parameters() if n != "encoder.weight"]

        def named_modules(self):
            return model.named_modules()

    it = PI.ProgramInterpreter(_Params(), prog)
    q = l0.double().softmax(dim=-1)
    o = cfg.objective
    val, dx, loss, G = it.matching_gradient(x0.double(), q, g, o.type, scale=o.scale, tag_scale=o.tag_scale, scale_scheme=o.scale_scheme)
    assert abs(float(val) - fx["objective0"]) < 2e-5 * abs(fx["objective0"])
    assert abs(float(loss) - fx["task_loss0"]) < 2e-5 * abs(fx["task_loss0"])
    assert _relerr(dx.float(), fx["raw_grad_x0"]) < 2e-4
    dl = q * (it.dq - (q * it.dq).sum(dim=-1, keepdim=True))
    assert _relerr(dl[:, :, ::97].float(), fx["raw_grad_l0_sample"]) < 2e-4
    assert abs(float(dl.norm()) - fx["raw_grad_l0_norm"]) < 2e-4 * fx["raw_grad_l0_norm"]
