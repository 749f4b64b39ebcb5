"""ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-PyTorch fp32 functional restatement of Depth-Anything-V2 inference.

Follows (reference @ /root/reference):
  src/depthmap_generation.py:375-403  get_raw_prediction (BGR<->RGB swap + /255)
  src/depthmap_generation.py:548-559  estimatedepthanything_v2 (*255.1 re-quantisation, second swap, final bilinear resize)
  ddepth_anything_v2/depth_anything_v2/dpt.py:196-221   image2tensor (Resize lower_bound x14 INTER_CUBIC, ImageNet normalise)
  ddepth_anything_v2/depth_anything_v2/util/transform.py:48-116  Resize.get_size
  ddepth_anything_v2/depth_anything_v2/dinov2.py:179-216,297-321  pos-embed, tokens, intermediate layers (+ final norm)
  ddepth_anything_v2/depth_anything_v2/dinov2_layers/{attention.py:49-62, block.py:82-107, mlp.py:35-41, layer_scale.py:27}
  ddepth_anything_v2/depth_anything_v2/dpt.py:117-150,176-184     DPTHead.forward, DepthAnythingV2.forward
  ddepth_anything_v2/depth_anything_v2/util/blocks.py:57-80,123-148  ResidualConvUnit, FeatureFusionBlock
Operates directly on a state_dict with the reference's key names, so the same weights drive the reference module
(tests/test_oracle_pin.py pins this file to it), this oracle and the CUDA path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {
    'vits': dict(embed_dim=384, depth=12, heads=6, features=64, out_channels=[48, 96, 192, 384], layers=[2, 5, 8, 11]),
    'vitb': dict(embed_dim=768, depth=12, heads=12, features=128, out_channels=[96, 192, 384, 768], layers=[2, 5, 8, 11]),
    'vitl': dict(embed_dim=1024, depth=24, heads=16, features=256, out_channels=[256, 512, 1024, 1024], layers=[4, 11, 17, 23]),
}
MEAN = np.array([0.485, 0.456, 0.406])
STD = np.array([0.229, 0.224, 0.225])


def constrain_to_multiple_of(x, multiple_of=14, min_val=0, max_val=None):
    y = (np.round(x / multiple_of) * multiple_of).astype(int)
    if max_val is not None and y > max_val:
        y = (np.floor(x / multiple_of) * multiple_of).astype(int)
    if y < min_val:
        y = (np.ceil(x / multiple_of) * multiple_of).astype(int)
    return int(y)


def get_size_lower_bound(width, height, target, multiple_of=14):
    """transform.py:61-104 with keep_aspect_ratio=True, resize_method='lower_bound', width=height=target."""
    scale_height = target / height
    scale_width = target / width
    if scale_width > scale_height:
        scale_height = scale_width
    else:
        scale_width = scale_height
    new_height = constrain_to_multiple_of(scale_height * height, multiple_of, min_val=target)
    new_width = constrain_to_multiple_of(scale_width * width, multiple_of, min_val=target)
    return new_width, new_height


def preprocess(pil_or_rgb_uint8, input_size):
    """uint8 RGB image -> float32 tensor [1,3,h',w'] exactly as the reference builds it (channel quirk included)."""
    import cv2
    rgb = np.asarray(pil_or_rgb_uint8)
    img = cv2.cvtColor(rgb, cv2.COLOR_BGR2RGB) / 255.0                     # depthmap_generation.py:381
    img2 = cv2.cvtColor((img * 255.1).astype('uint8'), cv2.COLOR_BGR2RGB)  # :550
    h, w = img2.shape[:2]
    image = cv2.cvtColor(img2, cv2.COLOR_BGR2RGB) / 255.0                  # dpt.py:213
    nw, nh = get_size_lower_bound(image.shape[1], image.shape[0], input_size)
    image = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_CUBIC)
    image = (image - MEAN) / STD
    image = np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)
    return torch.from_numpy(image).unsqueeze(0), (h, w)


def _interp_pos(sd, npatch, w, h, dim, patch=14, offset=0.1):
    pos_embed = sd['pretrained.pos_embed'].float()
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    w0, h0 = w // patch + offset, h // patch + offset
    sqrt_N = math.sqrt(N)
    sx, sy = float(w0) / sqrt_N, float(h0) / sqrt_N
    patch_pos = F.interpolate(patch_pos.reshape(1, int(sqrt_N), int(sqrt_N), dim).permute(0, 3, 1, 2),
                              scale_factor=(sx, sy), mode="bicubic", antialias=False)
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def backbone_features(sd, x, cfg):
    """x: [B,3,H,W] float32 -> list of 4 final-normed patch-token tensors [B, Np, C] (cls dropped)."""
    B, _, Himg, Wimg = x.shape
    C, heads = cfg['embed_dim'], cfg['heads']
    t = F.conv2d(x, sd['pretrained.patch_embed.proj.weight'].float(), sd['pretrained.patch_embed.proj.bias'].float(), stride=14)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd['pretrained.cls_token'].float().expand(B, -1, -1), t), dim=1)
    # dinov2.py:213 passes (w, h) = x.shape[2:], i.e. (H, W) of the tensor, into interpolate_pos_encoding
    t = t + _interp_pos(sd, t.shape[1] - 1, Himg, Wimg, C)
    feats = []
    for i in range(cfg['depth']):
        p = f'pretrained.blocks.{i}.'
        h = F.layer_norm(t, (C,), sd[p + 'norm1.weight'].float(), sd[p + 'norm1.bias'].float(), 1e-6)
        qkv = F.linear(h, sd[p + 'attn.qkv.weight'].float(), sd[p + 'attn.qkv.bias'].float())
        N = t.shape[1]
        qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, N, C)
        o = F.linear(o, sd[p + 'attn.proj.weight'].float(), sd[p + 'attn.proj.bias'].float())
        t = t + sd[p + 'ls1.gamma'].float() * o
        h = F.layer_norm(t, (C,), sd[p + 'norm2.weight'].float(), sd[p + 'norm2.bias'].float(), 1e-6)
        h = F.gelu(F.linear(h, sd[p + 'mlp.fc1.weight'].float(), sd[p + 'mlp.fc1.bias'].float()))
        h = F.linear(h, sd[p + 'mlp.fc2.weight'].float(), sd[p + 'mlp.fc2.bias'].float())
        t = t + sd[p + 'ls2.gamma'].float() * h
        if i in cfg['layers']:
            o = F.layer_norm(t, (C,), sd['pretrained.norm.weight'].float(), sd['pretrained.norm.bias'].float(), 1e-6)
            feats.append(o[:, 1:])
    return feats


def _conv(sd, key, x, stride=1, padding=0):
    b = sd.get(key + '.bias')
    return F.conv2d(x, sd[key + '.weight'].float(), None if b is None else b.float(), stride=stride, padding=padding)


def _rcu(sd, key, x):
    out = F.relu(x)
    out = _conv(sd, key + '.conv1', out, padding=1)
    out = F.relu(out)
    out = _conv(sd, key + '.conv2', out, padding=1)
    return out + x


def _fusion(sd, key, x0, x1=None, size=None):
    output = x0
    if x1 is not None:
        output = output + _rcu(sd, key + '.resConfUnit1', x1)
    output = _rcu(sd, key + '.resConfUnit2', output)
    if size is None:
        output = F.interpolate(output, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        output = F.interpolate(output, size=size, mode="bilinear", align_corners=True)
    return _conv(sd, key + '.out_conv', output)


def head(sd, feats, patch_h, patch_w, return_intermediates=False):
    out = []
    for i, x in enumerate(feats):
        B = x.shape[0]
        x = x.permute(0, 2, 1).reshape(B, x.shape[-1], patch_h, patch_w)
        x = _conv(sd, f'depth_head.projects.{i}', x)
        if i == 0:
            x = F.conv_transpose2d(x, sd['depth_head.resize_layers.0.weight'].float(), sd['depth_head.resize_layers.0.bias'].float(), stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd['depth_head.resize_layers.1.weight'].float(), sd['depth_head.resize_layers.1.bias'].float(), stride=2)
        elif i == 3:
            x = _conv(sd, 'depth_head.resize_layers.3', x, stride=2, padding=1)
        out.append(x)
    l1, l2, l3, l4 = out
    l1rn = _conv(sd, 'depth_head.scratch.layer1_rn', l1, padding=1)
    l2rn = _conv(sd, 'depth_head.scratch.layer2_rn', l2, padding=1)
    l3rn = _conv(sd, 'depth_head.scratch.layer3_rn', l3, padding=1)
    l4rn = _conv(sd, 'depth_head.scratch.layer4_rn', l4, padding=1)
    p4 = _fusion(sd, 'depth_head.scratch.refinenet4', l4rn, size=l3rn.shape[2:])
    p3 = _fusion(sd, 'depth_head.scratch.refinenet3', p4, l3rn, size=l2rn.shape[2:])
    p2 = _fusion(sd, 'depth_head.scratch.refinenet2', p3, l2rn, size=l1rn.shape[2:])
    p1 = _fusion(sd, 'depth_head.scratch.refinenet1', p2, l1rn)
    o = _conv(sd, 'depth_head.scratch.output_conv1', p1, padding=1)
    o = F.interpolate(o, (int(patch_h * 14), int(patch_w * 14)), mode="bilinear", align_corners=True)
    o = F.relu(_conv(sd, 'depth_head.scratch.output_conv2.0', o, padding=1))
    o = F.relu(_conv(sd, 'depth_head.scratch.output_conv2.2', o))
    if return_intermediates:
        return o, dict(l1rn=l1rn, l2rn=l2rn, l3rn=l3rn, l4rn=l4rn, p4=p4, p3=p3, p2=p2, p1=p1)
    return o


def forward(sd, x, encoder, return_intermediates=False):
    """DepthAnythingV2.forward (dpt.py:176-184): x [B,3,H,W] float32 -> depth [B,H,W] float32."""
    cfg = CONFIGS[encoder]
    ph, pw = x.shape[-2] // 14, x.shape[-1] // 14
    feats = backbone_features(sd, x, cfg)
    r = head(sd, feats, ph, pw, return_intermediates)
    if return_intermediates:
        d, inter = r
        inter['feats'] = feats
        return F.relu(d).squeeze(1), inter
    return F.relu(r).squeeze(1)


@torch.no_grad()
def get_raw_prediction(pil_or_rgb_uint8, sd, encoder, net_size=518):
    """ModelHolder.get_raw_prediction for model types 12-14: -> (float32 [H,W], invert=False)."""
    image, (h, w) = preprocess(pil_or_rgb_uint8, net_size)
    depth = forward(sd, image, encoder)
    depth = F.interpolate(depth[:, None], (h, w), mode="bilinear", align_corners=True)[0, 0]
    return depth.cpu().numpy(), False
