"""TEST INFRASTRUCTURE -- CPU/PyTorch restatement of the reference's CT encoder (the CNN branch,
scripts/training/partial_modality_training.py:179-190; identical in final_multimodal.py).  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may import this module; the product path never does.

Pinned: tests/test_oracle.py checks ``reference_cnn`` against tests/golden/ct_encoder.npz, which
oracle/gen_golden.py produced by running the reference's own ``PartialModalityNet`` class (AST-extracted).
"""
import torch
import torch.nn.functional as F
from torch import nn


def reference_cnn():
    """partial_modality_training.py:179-190 (USE_MONAI = False)."""
    return nn.Sequential(
        nn.Conv3d(1, 32, 3, stride=2, padding=1), nn.BatchNorm3d(32), nn.ReLU(),
        nn.Conv3d(32, 64, 3, stride=2, padding=1), nn.BatchNorm3d(64), nn.ReLU(),
        nn.Conv3d(64, 128, 3, stride=2, padding=1), nn.BatchNorm3d(128), nn.ReLU(),
        nn.AdaptiveAvgPool3d(1),
    )


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def matched_forward(seq, ct):
    """The same network with the operands of convolutions 2 and 3 rounded to bf16 (straight-through gradient): the
    arithmetic of the tensor-core path, evaluated by PyTorch."""
    x = ct
    for i in (0, 3, 6):
        conv, bn = seq[i], seq[i + 1]
        x = conv(x) if i == 0 else F.conv3d(_RoundBf16.apply(x), _RoundBf16.apply(conv.weight), conv.bias, stride=2, padding=1)
        x = F.relu(bn(x))
    return seq[9](x)
