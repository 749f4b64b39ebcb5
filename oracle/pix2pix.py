"""ORACLE (test infrastructure only — the product never imports this): plain-torch fp32 functional restatement of the BOOST
merge network, part of SURVEY.md §8a row D9.

Follows /root/reference:
  pix2pix/models/networks.py:444-543          UnetGenerator(input_nc 2, output_nc 1, num_downs 10, ngf 64, norm 'none'):
                                              recursive UnetSkipConnectionBlock — down = [LeakyReLU(0.2), Conv 4x4/2 pad 1],
                                              up = [ReLU, ConvTranspose 4x4/2 pad 1], skip = cat([x, block(x)]); the outermost
                                              block has no leading LeakyReLU and ends in Tanh; with norm 'none' every norm layer
                                              is Identity and (use_bias = norm == InstanceNorm = False) only the outermost
                                              ConvTranspose carries a bias
  pix2pix/models/pix2pix4depth_model.py:96-116  set_input (min-max normalise both estimates, to [-1, 1], stack) + forward
The nn.LeakyReLU(0.2, True) / nn.ReLU(True) at the head of each inner block are IN-PLACE on the tensor that is also
concatenated as the skip, so the skip operand is the activated tensor (leaky-relu'd by the child's down path, then relu'd
by the parent's up path) — restated explicitly below.
state_dict keys: the generator's (`netG.state_dict()`: "model.model.0.weight", "model.model.1.model...").
Pinned by tests/test_oracle_pin.py::test_pix2pix_unet_oracle_equals_reference.  Round 1 ships only this oracle."""
from __future__ import annotations

import torch
import torch.nn.functional as F

NUM_DOWNS = 10


def _block(x, sd, prefix, depth):
    """One UnetSkipConnectionBlock at recursion depth `depth` (0 = outermost ... NUM_DOWNS-1 = innermost).  Returns what the
    block's nn.Sequential returns (before the parent's concatenation); `x` is modified the way the in-place activations do."""
    outermost, innermost = depth == 0, depth == NUM_DOWNS - 1
    if outermost:
        h = F.conv2d(x, sd[prefix + "model.0.weight"], None, stride=2, padding=1)
        h = _child(h, sd, prefix + "model.1.", depth + 1)
        h = F.relu(h)
        h = F.conv_transpose2d(h, sd[prefix + "model.3.weight"], sd[prefix + "model.3.bias"], stride=2, padding=1)
        return torch.tanh(h)
    # inner blocks: [LeakyReLU(in place), Conv, (Identity norm)] + [submodule] + [ReLU(in place), ConvT, (Identity norm)]
    if innermost:
        h = F.conv2d(x, sd[prefix + "model.1.weight"], None, stride=2, padding=1)
        h = F.relu(h)
        return F.conv_transpose2d(h, sd[prefix + "model.3.weight"], None, stride=2, padding=1)
    h = F.conv2d(x, sd[prefix + "model.1.weight"], None, stride=2, padding=1)
    h = _child(h, sd, prefix + "model.3.", depth + 1)
    h = F.relu(h)
    return F.conv_transpose2d(h, sd[prefix + "model.5.weight"], None, stride=2, padding=1)


def _child(x, sd, prefix, depth):
    """forward() of a non-outermost block: cat([x, model(x)], 1) where model's leading LeakyReLU already rewrote x in place."""
    x = F.leaky_relu(x, 0.2)
    return torch.cat([x, _block(x, sd, prefix, depth)], 1)


def unet(sd, x):
    """netG(real_A): x [B,2,H,W] (H, W multiples of 1024 in practice; any multiple of 2^10 works) -> [B,1,H,W] in (-1, 1)."""
    return _block(x, sd, "model.", 0)


def merge_input(outer, inner):
    """Pix2Pix4DepthModel.set_input: two float32 [H,W] estimates -> real_A [1,2,H,W]."""
    inner = torch.from_numpy(inner).unsqueeze(0).unsqueeze(0)
    outer = torch.from_numpy(outer).unsqueeze(0).unsqueeze(0)
    inner = (inner - torch.min(inner)) / (torch.max(inner) - torch.min(inner))
    outer = (outer - torch.min(outer)) / (torch.max(outer) - torch.min(outer))
    inner = inner * 2 - 1                # Pix2Pix4DepthModel.normalize
    outer = outer * 2 - 1
    return torch.cat((outer, inner), 1)
