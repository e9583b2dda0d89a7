"""Phase timeline of one tcgen05 GEMM CTA (SM-clock timestamps of CTA (0,0,0), see TC_MARK in igemm_tc.cu).

Builds a -DBRE_TC_TRACE variant of the library next to the product one (in the build container: `python
profiles/experiments/tc_trace.py --build-only`), then on the GPU runs each shape a few times back to back and prints the
deltas of the last launch.  usage: tc_trace.py [--build-only]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from breaching_b200 import build as B  # noqa: E402

TRACE_LIB = os.path.join(B.LIBDIR, "libbreaching_b200_trace.so")


def build_trace():
    os.makedirs(B.LIBDIR, exist_ok=True)
    flags = [f for f in B.NVCC_FLAGS if f not in ("--use_fast_math=false", "-Xptxas", "-v")]
    srcs = [os.path.join(B.CSRC, s) for s in B.SOURCES]
    subprocess.run([B._nvcc(), *flags, "-DBRE_TC_TRACE", "-shared", "-o", TRACE_LIB, *srcs, "-lcudart", "-lcuda"], check=True)


if "--build-only" in sys.argv:
    build_trace()
    print("built", TRACE_LIB)
    sys.exit(0)

import torch  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

lib = E.load_library(TRACE_LIB)
lib.bre_debug_tc_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
dev = "cuda:0"
NAMES = ["entry", "setup done", "pdl_wait done", "initial TMAs issued", "first stage full", "last commit issued", "accumulator done",
         "partials parked", "cluster sync 1", "reduced+stored", "exit", "last stage full"]
SHAPES = [  # mode, N, H, W, Ci, Co, R, stride, pad, dual
    (0, 1, 56, 56, 64, 64, 3, 1, 1, 1), (0, 1, 14, 14, 256, 256, 3, 1, 1, 1), (0, 1, 7, 7, 512, 512, 3, 1, 1, 1),
    (1, 1, 28, 28, 128, 128, 3, 1, 1, 1), (2, 1, 14, 14, 256, 256, 3, 1, 1, 0), (0, 8, 56, 56, 64, 64, 3, 1, 1, 1),
]
for (mode, N, H, W, Ci, Co, R, st, pd, dual) in SHAPES:
    Ho, Wo = (H + 2 * pd - R) // st + 1, (W + 2 * pd - R) // st + 1
    x, x2 = torch.randn(N, H, W, Ci, device=dev), torch.randn(N, H, W, Ci, device=dev)
    w, w2 = torch.randn(Co, R, R, Ci, device=dev) * 0.1, torch.randn(Co, R, R, Ci, device=dev) * 0.1
    dy, dy2 = torch.randn(N, Ho, Wo, Co, device=dev), torch.randn(N, Ho, Wo, Co, device=dev)
    if mode == 0:
        out = torch.empty(N, Ho, Wo, Co, device=dev)
        fn = lambda: E.conv_gemm(0, x, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=x2 if dual else None, w2=w2 if dual else None, backend=1)
    elif mode == 1:
        out = torch.empty(N, H, W, Ci, device=dev)
        fn = lambda: E.conv_gemm(1, dy, w, out, N, H, W, Ci, Co, R, R, st, pd, a2=dy2 if dual else None, w2=w2 if dual else None, backend=1)
    else:
        out = torch.empty(Co, R, R, Ci, device=dev)
        fn = lambda: E.conv_gemm(2, x, dy, out, N, H, W, Ci, Co, R, R, st, pd, backend=1)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    assert lib.bre_debug_tc_trace(buf) == 0
    t = list(buf)
    ghz = 1.965
    order = [0, 1, 2, 3, 4, 11, 5, 6, 7, 8, 9, 10]
    line = ", ".join(f"{NAMES[i]} {((t[i] - t[0]) / ghz / 1e3):.2f}" for i in order if t[i] >= t[0] and t[i] != 0)
    print(f"mode {mode} [{N},{H},{W},{Ci}->{Co},k{R},s{st}{',dual' if dual else ''}] us since entry (SM clock @ {ghz} GHz): {line}")
