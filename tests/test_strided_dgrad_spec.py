"""CPU check of the per-parity-class formulation of the strided data gradient (oracle/strided_dgrad.py): the tensor-map
corners, filter offsets and output scatter that a TMA-fed strided dgrad kernel needs, evaluated through an emulation of the
im2col traversal and compared with torch.nn.grad.conv2d_input."""
import pytest
import torch

from oracle.strided_dgrad import class_plan, dgrad_by_classes


@pytest.mark.parametrize("N,H,W,Ci,Co,R,st,pd", [
    (2, 14, 14, 4, 6, 3, 2, 1),      # ResNet 3x3 stride 2
    (1, 14, 14, 4, 6, 1, 2, 0),      # 1x1 stride-2 downsample (three of four classes receive nothing)
    (1, 22, 22, 3, 5, 7, 2, 3),      # the 7x7 stem
    (2, 9, 11, 3, 4, 5, 3, 2),       # stride 3, odd sizes
    (1, 10, 10, 2, 3, 3, 3, 0),
    (3, 7, 9, 2, 2, 3, 2, 1),        # tiles that run across rows and images (tile = 16 below)
])
def test_class_decomposition_equals_conv2d_input(N, H, W, Ci, Co, R, st, pd):
    gen = torch.Generator().manual_seed(H * 100 + R)
    Ho, Wo = (H + 2 * pd - R) // st + 1, (W + 2 * pd - R) // st + 1
    dout = torch.randn(N, Co, Ho, Wo, dtype=torch.double, generator=gen)
    w = torch.randn(Co, Ci, R, R, dtype=torch.double, generator=gen)
    ref = torch.nn.grad.conv2d_input((N, Ci, H, W), w, dout, stride=st, padding=pd)
    got = dgrad_by_classes(dout, w, (H, W), st, pd, tile=16)
    assert (got - ref).abs().max().item() < 1e-12
    for e in range(st):   # tensor-map constraints of cuTensorMapEncodeIm2col for rank 4: corners within [-128, 127], non-empty box
        plan = class_plan(H, Ho, R, st, pd, e)
        if plan is not None and plan["T"] > 0:
            assert -128 <= plan["L"] <= 127 and -128 <= plan["U"] <= 127 and Ho + plan["U"] - plan["L"] == plan["Hc"] > 0
