"""One-process GEMM suite: correctness (vs torch fp64) + device time (CUDA-graph replay) for representative
ResNet-18/50 shapes on both back ends.  usage: gemm_suite.py [backends e.g. 01] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
dev = "cuda:0"
backends = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "01")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
SHAPES = [  # mode, N, H, W, Ci, Co, R, stride, pad, dual
    (0, 1, 56, 56, 64, 64, 3, 1, 1, 1), (0, 1, 28, 28, 128, 128, 3, 1, 1, 1), (0, 1, 14, 14, 256, 256, 3, 1, 1, 1),
    (0, 1, 7, 7, 512, 512, 3, 1, 1, 1), (0, 1, 56, 56, 64, 128, 3, 2, 1, 1), (0, 1, 56, 56, 64, 64, 3, 1, 1, 0),
    (1, 1, 56, 56, 64, 64, 3, 1, 1, 1), (1, 1, 28, 28, 128, 128, 3, 1, 1, 1), (1, 1, 14, 14, 256, 256, 3, 1, 1, 1),
    (1, 1, 7, 7, 512, 512, 3, 1, 1, 1), (1, 1, 28, 28, 128, 256, 3, 2, 1, 1), (1, 1, 56, 56, 64, 128, 1, 2, 0, 1),
    (2, 1, 28, 28, 128, 128, 3, 1, 1, 0), (2, 1, 14, 14, 256, 256, 3, 1, 1, 0), (2, 1, 7, 7, 512, 512, 3, 1, 1, 0),
    (0, 8, 56, 56, 64, 64, 3, 1, 1, 1), (0, 8, 14, 14, 256, 256, 3, 1, 1, 1), (1, 8, 28, 28, 128, 128, 3, 1, 1, 1),
    (2, 8, 14, 14, 256, 256, 3, 1, 1, 0), (0, 1, 8, 16, 64, 128, 1, 1, 0, 0),
]


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (mode, N, H, W, Ci, Co, R, st, pd, dual) in SHAPES:
    g = torch.Generator().manual_seed(0)
    Ho, Wo = (H + 2 * pd - R) // st + 1, (W + 2 * pd - R) // st + 1
    x, x2 = (torch.randn(N, H, W, Ci, generator=g).to(dev) for _ in range(2))
    w, w2 = ((torch.randn(Co, R, R, Ci, generator=g) * 0.1).to(dev) for _ in range(2))
    dy, dy2 = (torch.randn(N, Ho, Wo, Co, generator=g).to(dev) for _ in range(2))
    nchw = lambda t: t.permute(0, 3, 1, 2).double()  # noqa: E731
    flops = 2.0 * N * Ho * Wo * Co * R * R * Ci * (2 if dual else 1)
    if mode == 0:
        out = torch.zeros(N, Ho, Wo, Co, device=dev)
        ref = F.conv2d(nchw(x), nchw(w), stride=st, padding=pd)
        if dual:
            ref = ref + F.conv2d(nchw(x2), nchw(w2), stride=st, padding=pd)
        ref = ref.permute(0, 2, 3, 1)
        mk = lambda be: (lambda: E.conv_gemm(0, x, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=x2 if dual else None, w2=w2 if dual else None, backend=be))  # noqa: E731
    elif mode == 1:
        out = torch.zeros(N, H, W, Ci, device=dev)
        ref = torch.nn.grad.conv2d_input((N, Ci, H, W), nchw(w), nchw(dy), stride=st, padding=pd)
        if dual:
            ref = ref + torch.nn.grad.conv2d_input((N, Ci, H, W), nchw(w2), nchw(dy2), stride=st, padding=pd)
        ref = ref.permute(0, 2, 3, 1)
        mk = lambda be: (lambda: E.conv_gemm(1, dy, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=dy2 if dual else None, w2=w2 if dual else None, backend=be))  # noqa: E731
    else:
        out = torch.zeros(Co, R, R, Ci, device=dev)
        ref = torch.nn.grad.conv2d_weight(nchw(x), (Co, Ci, R, R), nchw(dy), stride=st, padding=pd).permute(0, 2, 3, 1)
        mk = lambda be: (lambda: E.conv_gemm(2, x, dy, out, N, H, W, Ci, Co, R, R, st, pd, backend=be))  # noqa: E731
    line = f"mode {mode} [{N},{H},{W},{Ci}->{Co},k{R},s{st}{',dual' if dual else ''}] {flops / 1e9:6.3f} GF:"
    for be in backends:
        try:
            out.zero_()
            fn = mk(be)
            fn()
            torch.cuda.synchronize()
            err = ((out.double() - ref).norm() / ref.norm()).item()
            us = timed(fn)
            line += f"  be{be}: {us:7.2f} us {flops / us / 1e6:7.2f} TF/s err {err:.1e}"
        except Exception as exc:  # noqa: BLE001
            line += f"  be{be}: {type(exc).__name__} {str(exc)[:60]}"
    print(line, flush=True)
