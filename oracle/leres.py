"""ORACLE (test infrastructure only — the product never imports this): plain-torch fp32 functional restatement of LeReS
inference, SURVEY.md §8a row D8 (and the base estimator of BOOST, D9).

Follows /root/reference:
  src/depthmap_generation.py:406-440          estimateleres, scale_torch (BGR flip, cv2 bilinear resize, ImageNet normalise,
                                              cv2 INTER_CUBIC resize of the prediction back to the image size)
  lib/multi_depth_model_woauxi.py:6-32        RelDepthModel -> DepthModel(encoder 'resnext101_stride32x8d', Decoder)
  lib/Resnext_torch.py:60-220                 ResNeXt-101 32x8d: 7x7/2 stem, max-pool, bottlenecks [3, 4, 23, 3] with 32-group 3x3
                                              convs (width = planes * 8 / 64 * 32), stride on the 3x3, features after each stage
  lib/network_auxi.py:16-62,95-215            Decoder: FTB (conv3x3, in-place ReLU, [conv, BN, ReLU, conv] residual branch, ReLU),
                                              FFM (FTB, add, FTB, 2x bilinear align_corners=True), AO (conv, BN, ReLU, conv, 2x up)
BatchNorm is evaluated in inference mode from the running statistics of the checkpoint (eps 1e-5).
state_dict keys are the reference module's (`RelDepthModel.state_dict()`, i.e. "depth_model.encoder_modules.encoder...").
Pinned by a strict load of a seeded state_dict into the reference module (tests/test_oracle_pin.py::test_leres_*).
Round 1 ships only this oracle for D8; the CUDA path (grouped 3x3 convs, folded BN) is round-2 work."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = (3, 4, 23, 3)
GROUPS = 32
ENC = "depth_model.encoder_modules.encoder."
DEC = "depth_model.decoder_modules."


def _bn(x, sd, key):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"], False, 0.0, 1e-5)


def _bottleneck(x, sd, p, stride, has_down):
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1, groups=GROUPS), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    if has_down:
        x = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


def encoder(x, sd):
    """ResNet._forward_impl: features after layer1..layer4 (1/4, 1/8, 1/16, 1/32)."""
    x = F.relu(_bn(F.conv2d(x, sd[ENC + "conv1.weight"], stride=2, padding=3), sd, ENC + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, nblocks in enumerate(LAYERS, start=1):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 1) else 1
            x = _bottleneck(x, sd, f"{ENC}layer{li}.{bi}", stride, bi == 0)
        feats.append(x)
    return feats


def _conv3(x, sd, key):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=1)


def _ftb(x, sd, p):
    # network_auxi.py:110-114: the branch starts with ReLU(inplace=True), which overwrites the tensor that is also the skip
    # operand of `x + self.conv_branch(x)` — the residual is therefore relu(conv1(x)), not conv1(x)
    x = F.relu(_conv3(x, sd, p + ".conv1"))
    b = _conv3(x, sd, p + ".conv_branch.1")
    b = _conv3(F.relu(_bn(b, sd, p + ".conv_branch.2")), sd, p + ".conv_branch.4")
    return F.relu(x + b)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def _ffm(low, high, sd, p):
    return _up2(_ftb(_ftb(low, sd, p + ".ftb1") + high, sd, p + ".ftb2"))


def decoder(feats, sd):
    x = _conv3(_ftb(feats[3], sd, DEC + "conv"), sd, DEC + "conv1")
    x = _up2(x)
    x = _ffm(feats[2], x, sd, DEC + "ffm2")
    x = _ffm(feats[1], x, sd, DEC + "ffm1")
    x = _ffm(feats[0], x, sd, DEC + "ffm0")
    a = DEC + "outconv.adapt_conv"
    x = _conv3(F.relu(_bn(_conv3(x, sd, a + ".0"), sd, a + ".1")), sd, a + ".3")
    return _up2(x)


def forward(sd, x):
    """RelDepthModel.depth_model: x [B,3,H,W] (ImageNet-normalised, H and W multiples of 32) -> [B,1,H,W]."""
    return decoder(encoder(x, sd), sd)


def preprocess(img, w, h):
    """estimateleres + scale_torch (src/depthmap_generation.py:406-440) on the image estimateleres receives: channel flip,
    cv2.resize (bilinear, in the array's own dtype) to (w, h), float32, ToTensor WITHOUT a /255 (torchvision only transposes float
    arrays), ImageNet mean / std."""
    import cv2
    img = np.asarray(img)
    a = cv2.resize(img[:, :, ::-1].copy(), (w, h))
    t = torch.from_numpy(a.astype(np.float32).transpose(2, 0, 1).copy())
    mean = torch.tensor((0.485, 0.456, 0.406)).view(3, 1, 1)
    std = torch.tensor((0.229, 0.224, 0.225)).view(3, 1, 1)
    return ((t - mean) / std).unsqueeze(0)


@torch.no_grad()
def estimateleres(img, sd, w=448, h=448):
    """The reference function itself (:406-421) for whatever image array it is handed."""
    import cv2
    img = np.asarray(img)
    pred = forward(sd, preprocess(img, w, h)).squeeze().cpu().numpy()
    return cv2.resize(pred, (img.shape[1], img.shape[0]), interpolation=cv2.INTER_CUBIC)


@torch.no_grad()
def get_raw_prediction(rgb_uint8, sd, w=448, h=448):
    """ModelHolder.get_raw_prediction for model type 0 (res101): (float32 [H,W], invert=True).  The holder hands estimateleres
    `cv2.cvtColor(image, COLOR_BGR2RGB) / 255.0` (:381), i.e. a float64 image in [0, 1] with swapped channels, which
    estimateleres swaps back."""
    import cv2
    img = cv2.cvtColor(np.asarray(rgb_uint8), cv2.COLOR_BGR2RGB) / 255.0
    return estimateleres(img, sd, w, h), True
