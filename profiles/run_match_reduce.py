"""Launch the matching-reduction kernel a few times on ResNet-18-sized (11.38 M float) cold buffers (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from breaching_b200 import engine as E  # noqa: E402

P = 11_380_173
pairs = [(torch.randn(P, device="cuda:0"), torch.randn(P, device="cuda:0")) for _ in range(4)]
for i in range(8):
    E.match_reduce(*pairs[i % 4], readback=False)
torch.cuda.synchronize()
print(E.match_reduce(*pairs[0]))
