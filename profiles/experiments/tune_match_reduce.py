"""Launch-shape sweep of the matching-reduction kernel (blocks per SM x chunks in flight), measured like bench.py does:
16 launches over 4 rotating 91 MB buffer pairs replayed from one CUDA graph (cold in L2), CUDA events."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

lib = E.load_library()
lib.bre_debug_match_config.argtypes = [ctypes.c_int, ctypes.c_int]
lib.bre_debug_match_config.restype = None
dev = torch.device("cuda:0")
P = 11_380_173
pairs = [(torch.randn(P, device=dev), torch.randn(P, device=dev)) for _ in range(4)]
ref = [(pairs[0][0].double() * pairs[0][1].double()).sum().item(), (pairs[0][0].double() ** 2).sum().item()]
for bps in (2, 3, 4, 6, 8):
    for unroll in (4, 8):
        lib.bre_debug_match_config(bps, unroll)
        sums = E.match_reduce(*pairs[0])
        assert abs(sums[0] - ref[0]) < 1e-6 * abs(ref[1]) and abs(sums[1] - ref[1]) < 1e-9 * ref[1], (sums, ref)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for i in range(16):
                    E.match_reduce(*pairs[i % 4], readback=False)
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / 16)
        print(f"blocks/SM {bps} unroll {unroll}: {best * 1e3:.2f} us/launch, {8 * P / best / 1e6:.0f} GB/s")
