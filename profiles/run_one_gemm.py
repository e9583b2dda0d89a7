"""Run one conv GEMM shape repeatedly (for ncu captures / CUDA-event timing).
usage: run_one_gemm.py mode backend N H W Ci Co R stride pad [dual] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

mode, backend, N, H, W, Ci, Co, R, st, pd = [int(v) for v in sys.argv[1:11]]
dual = int(sys.argv[11]) if len(sys.argv) > 11 else 0
reps = int(sys.argv[12]) if len(sys.argv) > 12 else 20
dev = "cuda:0"
Ho, Wo = (H + 2 * pd - R) // st + 1, (W + 2 * pd - R) // st + 1
x = torch.randn(N, H, W, Ci, device=dev)
x2 = torch.randn(N, H, W, Ci, device=dev)
w = torch.randn(Co, R, R, Ci, device=dev) * 0.1
w2 = torch.randn(Co, R, R, Ci, device=dev) * 0.1
dy = torch.randn(N, Ho, Wo, Co, device=dev)
dy2 = torch.randn(N, Ho, Wo, Co, device=dev)
if mode == 0:
    out = torch.empty(N, Ho, Wo, Co, device=dev)
    fn = lambda: E.conv_gemm(0, x, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=x2 if dual else None, w2=w2 if dual else None, backend=backend)
    flops = 2.0 * N * Ho * Wo * Co * R * R * Ci * (2 if dual else 1)
elif mode == 1:
    out = torch.empty(N, H, W, Ci, device=dev)
    fn = lambda: E.conv_gemm(1, dy, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=dy2 if dual else None, w2=w2 if dual else None, backend=backend)
    flops = 2.0 * N * Ho * Wo * Co * R * R * Ci * (2 if dual else 1)
else:
    out = torch.empty(Co, R, R, Ci, device=dev)
    fn = lambda: E.conv_gemm(2, x, dy, out, N, H, W, Ci, Co, R, R, st, pd, backend=backend)
    flops = 2.0 * N * Ho * Wo * Co * R * R * Ci
for _ in range(3):
    fn()
torch.cuda.synchronize()
# replay from a CUDA graph so that the number is device time, not the host's launch rate
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
ms = e0.elapsed_time(e1) / reps
print(f"mode {mode} backend {backend} shape {' '.join(sys.argv[3:11])} dual {dual}: {ms * 1e3:.2f} us/launch, {flops / ms / 1e9:.2f} TFLOP/s (graph replay of {reps} launches, L2-warm)")
