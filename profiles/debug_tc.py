"""Debug helper: tcgen05 back end vs torch on a few shapes, with an error-pattern dump (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def report(name, out, ref):
    out, ref = out.double(), ref.double()
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"{name}: rel err {rel:.3e}  (|ref| {ref.norm().item():.3e}, |out| {out.norm().item():.3e})")
    if rel > 5e-3:
        bad = (out - ref).abs() > 1e-2 * ref.abs().max()
        print("   bad fraction", bad.float().mean().item())
        flat_out, flat_ref = out.reshape(-1, out.shape[-1]), ref.reshape(-1, ref.shape[-1])
        rows_bad = bad.reshape(-1, out.shape[-1]).any(dim=1).nonzero().flatten()[:16].tolist()
        cols_bad = bad.reshape(-1, out.shape[-1]).any(dim=0).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows_bad, "first bad cols", cols_bad)
        print("   out[0,:8]", flat_out[0, :8].tolist())
        print("   ref[0,:8]", flat_ref[0, :8].tolist())
        ratio = (flat_out[:4, :4] / flat_ref[:4, :4])
        print("   ratio[:4,:4]", ratio.tolist())
    return rel


def run(N, H, W, Ci, Co, R, st, pd, dual=False):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Ci, H, W, generator=g).to(DEV)
    w = (torch.randn(Co, Ci, R, R, generator=g) * 0.1).to(DEV)
    Ho, Wo = (H + 2 * pd - R) // st + 1, (W + 2 * pd - R) // st + 1
    dy = torch.randn(N, Co, Ho, Wo, generator=g).to(DEV)
    x2 = torch.randn(N, Ci, H, W, generator=g).to(DEV)
    w2 = (torch.randn(Co, Ci, R, R, generator=g) * 0.1).to(DEV)
    dy2 = torch.randn(N, Co, Ho, Wo, generator=g).to(DEV)
    wo, w2o = w.permute(0, 2, 3, 1).contiguous(), w2.permute(0, 2, 3, 1).contiguous()
    tag = f"[{N},{H},{W},{Ci}->{Co},k{R},s{st},p{pd}{',dual' if dual else ''}]"
    res = {}
    for mode, name in [(0, "fprop"), (1, "dgrad"), (2, "wgrad")]:
        try:
            if mode == 0:
                out = torch.zeros(N, Ho, Wo, Co, device=DEV)
                E.conv_gemm(0, nhwc(x), wo, out, N, H, W, Ci, Co, R, R, st, pd, a2=nhwc(x2) if dual else None,
                            w2=w2o if dual else None, backend=1)
                ref = F.conv2d(x.double(), w.double(), stride=st, padding=pd)
                if dual:
                    ref = ref + F.conv2d(x2.double(), w2.double(), stride=st, padding=pd)
                res[name] = report(name + tag, out, nhwc(ref))
            elif mode == 1:
                out = torch.zeros(N, H, W, Ci, device=DEV)
                E.conv_gemm(1, nhwc(dy), wo, out, N, H, W, Ci, Co, R, R, st, pd, a2=nhwc(dy2) if dual else None,
                            w2=w2o if dual else None, backend=1)
                ref = torch.nn.grad.conv2d_input((N, Ci, H, W), w.double(), dy.double(), stride=st, padding=pd)
                if dual:
                    ref = ref + torch.nn.grad.conv2d_input((N, Ci, H, W), w2.double(), dy2.double(), stride=st, padding=pd)
                res[name] = report(name + tag, out, nhwc(ref))
            else:
                if dual:
                    continue
                out = torch.zeros(Co, R, R, Ci, device=DEV)
                E.conv_gemm(2, nhwc(x), nhwc(dy), out, N, H, W, Ci, Co, R, R, st, pd, backend=1)
                ref = torch.nn.grad.conv2d_weight(x.double(), (Co, Ci, R, R), dy.double(), stride=st, padding=pd)
                res[name] = report(name + tag, out, ref.permute(0, 2, 3, 1))
        except Exception as exc:  # noqa: BLE001
            print(name + tag, "->", type(exc).__name__, str(exc)[:200])
    torch.cuda.synchronize()
    return res


if __name__ == "__main__":
    run(1, 8, 16, 32, 64, 1, 1, 0)       # one M tile, one k-block: the simplest possible case
    run(1, 8, 16, 64, 128, 1, 1, 0)      # two k-blocks, wgrad eligible (Co = 128)
    run(1, 16, 16, 64, 128, 3, 1, 1)     # 3x3, multiple tiles, split-K
    run(1, 56, 56, 64, 64, 3, 1, 1)
    run(1, 28, 28, 128, 128, 3, 1, 1, dual=True)
    run(1, 14, 14, 256, 512, 3, 2, 1)
    run(1, 7, 7, 512, 512, 3, 1, 1, dual=True)
    run(2, 14, 14, 256, 1024, 1, 1, 0)
