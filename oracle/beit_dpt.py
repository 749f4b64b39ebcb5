"""ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-PyTorch fp32 functional restatement of MiDaS 3.1 DPT-BEiT inference.

PARITY UNPINNED: the reference builds this network on timm (`beit_large_patch16_512` / `_384`, requirements.txt:8
pins timm~=0.9.2), which is absent from /root/reference and from this image, so the restatement cannot be executed
against the real module here.  It follows the reference's own forward overrides line by line — every forward method of
timm's Beit is replaced by the reference (dmidas/backbones/beit.py:18-129), so only constructor shapes / parameter
names and `gen_relative_position_index` come from timm:
  src/depthmap_generation.py:375-403,455-499  get_raw_prediction + estimatemidas (Resize "minimal" x32 INTER_CUBIC,
                                               mean = std = 0.5, bicubic align_corners=False back to image size)
  dmidas/transforms.py:48-231                  Resize.get_size / NormalizeImage / PrepareForNet
  dmidas/backbones/beit.py:18-26,29-62,65-91,94-107,110-129   patch embed, rel-pos bias, attention, block, features
  dmidas/backbones/utils.py:28-39,83-124,144-249               ProjectReadout, forward_adapted_unflatten, reassemble
  dmidas/dpt_depth.py:31-166, dmidas/blocks.py:136-166,322-441 DPT decoder, head
  timm 0.9 `gen_relative_position_index`: nrd=(2Wh-1)(2Ww-1)+3; idx[1:,1:]=(dh+Wh-1)(2Ww-1)+(dw+Ww-1);
                                          idx[0,:]=nrd-3; idx[:,0]=nrd-2; idx[0,0]=nrd-1
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {
    'beitl16_512': dict(embed_dim=1024, depth=24, heads=16, features=256, out_channels=[256, 512, 1024, 1024],
                        hooks=[5, 11, 17, 23], window=32, net=512),
    'beitl16_384': dict(embed_dim=1024, depth=24, heads=16, features=256, out_channels=[256, 512, 1024, 1024],
                        hooks=[5, 11, 17, 23], window=24, net=384),
    # small test-only configuration with the same structure
    'beit_tiny': dict(embed_dim=128, depth=4, heads=2, features=64, out_channels=[64, 64, 128, 128],
                      hooks=[0, 1, 2, 3], window=4, net=64),
    # MiDaS 3.0 dpt_large_384 (model type 3): timm vit_large_patch16_384 driven by the reference's forward_flex
    # (dmidas/backbones/vit.py:12-79,107-118): absolute position embedding (bilinear resize of the 24x24 grid), no relative
    # position bias, no LayerScale, qkv with a full bias; same hooks / readout / decoder as the BEiT models
    'vitl16_384': dict(embed_dim=1024, depth=24, heads=16, features=256, out_channels=[256, 512, 1024, 1024],
                       hooks=[5, 11, 17, 23], window=24, net=384, family='vit'),
    'vit_tiny': dict(embed_dim=128, depth=4, heads=2, features=64, out_channels=[64, 64, 128, 128],
                     hooks=[0, 1, 2, 3], window=4, net=64, family='vit'),
}


def constrain_to_multiple_of(x, multiple_of=32, min_val=0, max_val=None):
    y = (np.round(x / multiple_of) * multiple_of).astype(int)
    if max_val is not None and y > max_val:
        y = (np.floor(x / multiple_of) * multiple_of).astype(int)
    if y < min_val:
        y = (np.ceil(x / multiple_of) * multiple_of).astype(int)
    return int(y)


def get_size_minimal(width, height, net_w, net_h, multiple_of=32):
    """dmidas/transforms.py get_size with keep_aspect_ratio=True, resize_method='minimal'."""
    scale_height = net_h / height
    scale_width = net_w / width
    if abs(1 - scale_width) < abs(1 - scale_height):
        scale_height = scale_width
    else:
        scale_width = scale_height
    return constrain_to_multiple_of(scale_width * width, multiple_of), constrain_to_multiple_of(scale_height * height, multiple_of)


def preprocess(rgb_uint8, net_w, net_h):
    import cv2
    img = cv2.cvtColor(np.asarray(rgb_uint8), cv2.COLOR_BGR2RGB) / 255.0   # depthmap_generation.py:381 (channels swapped)
    nw, nh = get_size_minimal(img.shape[1], img.shape[0], net_w, net_h)
    x = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_CUBIC)
    x = (x - np.array([0.5, 0.5, 0.5])) / np.array([0.5, 0.5, 0.5])
    x = np.ascontiguousarray(np.transpose(x, (2, 0, 1))).astype(np.float32)
    return torch.from_numpy(x).unsqueeze(0)


def gen_relative_position_index(window_size):
    num_relative_distance = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) + 3
    window_area = window_size[0] * window_size[1]
    coords = torch.stack(torch.meshgrid([torch.arange(window_size[0]), torch.arange(window_size[1])], indexing='ij'))
    coords_flatten = torch.flatten(coords, 1)
    relative_coords = coords_flatten[:, :, None] - coords_flatten[:, None, :]
    relative_coords = relative_coords.permute(1, 2, 0).contiguous()
    relative_coords[:, :, 0] += window_size[0] - 1
    relative_coords[:, :, 1] += window_size[1] - 1
    relative_coords[:, :, 0] *= 2 * window_size[1] - 1
    idx = torch.zeros(size=(window_area + 1,) * 2, dtype=relative_coords.dtype)
    idx[1:, 1:] = relative_coords.sum(-1)
    idx[0, 0:] = num_relative_distance - 3
    idx[0:, 0] = num_relative_distance - 2
    idx[0, 0] = num_relative_distance - 1
    return idx


def rel_pos_bias(table, old_window, new_window):
    """dmidas/backbones/beit.py:29-62 — returns [heads, N, N]."""
    old_h, old_w = 2 * old_window[0] - 1, 2 * old_window[1] - 1
    new_h, new_w = 2 * new_window[0] - 1, 2 * new_window[1] - 1
    old_nrd = old_h * old_w + 3
    new_nrd = new_h * new_w + 3
    sub = table[:old_nrd - 3].reshape(1, old_w, old_h, -1).permute(0, 3, 1, 2)
    new_sub = F.interpolate(sub, size=(int(new_h), int(new_w)), mode="bilinear")
    new_sub = new_sub.permute(0, 2, 3, 1).reshape(new_nrd - 3, -1)
    new_table = torch.cat([new_sub, table[old_nrd - 3:]])
    idx = gen_relative_position_index(new_window)
    n = new_window[0] * new_window[1] + 1
    return new_table[idx.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()


def backbone_hooks(sd, x, cfg):
    """beit_forward_features with the four forward hooks; returns the raw block outputs [B, N, C]."""
    B, _, H, W = x.shape
    C, heads = cfg['embed_dim'], cfg['heads']
    p = 'pretrained.model.'
    t = F.conv2d(x, sd[p + 'patch_embed.proj.weight'].float(), sd[p + 'patch_embed.proj.bias'].float(), stride=16)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd[p + 'cls_token'].float().expand(B, -1, -1), t), dim=1)
    window = (H // 16, W // 16)
    outs = []
    for i in range(cfg['depth']):
        b = p + f'blocks.{i}.'
        h = F.layer_norm(t, (C,), sd[b + 'norm1.weight'].float(), sd[b + 'norm1.bias'].float(), 1e-6)
        q_bias = sd[b + 'attn.q_bias'].float()
        qkv_bias = torch.cat((q_bias, torch.zeros_like(q_bias), sd[b + 'attn.v_bias'].float()))   # k has no bias (beit.py:70-74)
        qkv = F.linear(h, sd[b + 'attn.qkv.weight'].float(), qkv_bias)
        N = t.shape[1]
        qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q = q * (C // heads) ** -0.5
        attn = q @ k.transpose(-2, -1)
        attn = attn + rel_pos_bias(sd[b + 'attn.relative_position_bias_table'].float(), (cfg['window'], cfg['window']), window).unsqueeze(0)
        attn = attn.softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, N, -1)
        o = F.linear(o, sd[b + 'attn.proj.weight'].float(), sd[b + 'attn.proj.bias'].float())
        t = t + sd[b + 'gamma_1'].float() * o
        h = F.layer_norm(t, (C,), sd[b + 'norm2.weight'].float(), sd[b + 'norm2.bias'].float(), 1e-6)
        h = F.gelu(F.linear(h, sd[b + 'mlp.fc1.weight'].float(), sd[b + 'mlp.fc1.bias'].float()))
        h = F.linear(h, sd[b + 'mlp.fc2.weight'].float(), sd[b + 'mlp.fc2.bias'].float())
        t = t + sd[b + 'gamma_2'].float() * h
        if i in cfg['hooks']:
            outs.append(t)
    return outs


def vit_resize_pos_embed(posemb, gs_h, gs_w):
    """dmidas/backbones/vit.py:16-31 (_resize_pos_embed, start_index 1): bilinear, align_corners=False."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    gs_old = int(np.sqrt(len(grid)))
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gs_h, gs_w), mode="bilinear")
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_h * gs_w, -1)
    return torch.cat([tok, grid], dim=1)


def vit_backbone_hooks(sd, x, cfg):
    """forward_flex (dmidas/backbones/vit.py:34-79) over timm's VisionTransformer blocks (pre-norm, qkv with bias, GELU MLP,
    LayerNorm eps 1e-6) with the four forward hooks; returns the raw block outputs [B, N, C]."""
    B, _, H, W = x.shape
    C, heads = cfg['embed_dim'], cfg['heads']
    p = 'pretrained.model.'
    pos = vit_resize_pos_embed(sd[p + 'pos_embed'].float(), H // 16, W // 16)
    t = F.conv2d(x, sd[p + 'patch_embed.proj.weight'].float(), sd[p + 'patch_embed.proj.bias'].float(), stride=16)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd[p + 'cls_token'].float().expand(B, -1, -1), t), dim=1) + pos
    outs = []
    for i in range(cfg['depth']):
        b = p + f'blocks.{i}.'
        h = F.layer_norm(t, (C,), sd[b + 'norm1.weight'].float(), sd[b + 'norm1.bias'].float(), 1e-6)
        qkv = F.linear(h, sd[b + 'attn.qkv.weight'].float(), sd[b + 'attn.qkv.bias'].float())
        N = t.shape[1]
        q, k, v = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4).unbind(0)
        attn = ((q * (C // heads) ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, N, -1)
        t = t + F.linear(o, sd[b + 'attn.proj.weight'].float(), sd[b + 'attn.proj.bias'].float())
        h = F.layer_norm(t, (C,), sd[b + 'norm2.weight'].float(), sd[b + 'norm2.bias'].float(), 1e-6)
        h = F.gelu(F.linear(h, sd[b + 'mlp.fc1.weight'].float(), sd[b + 'mlp.fc1.bias'].float()))
        t = t + F.linear(h, sd[b + 'mlp.fc2.weight'].float(), sd[b + 'mlp.fc2.bias'].float())
        if i in cfg['hooks']:
            outs.append(t)
    return outs


def _conv(sd, key, x, stride=1, padding=0):
    b = sd.get(key + '.bias')
    return F.conv2d(x, sd[key + '.weight'].float(), None if b is None else b.float(), stride=stride, padding=padding)


def _rcu(sd, key, x):
    out = _conv(sd, key + '.conv1', F.relu(x), padding=1)
    out = _conv(sd, key + '.conv2', F.relu(out), padding=1)
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


def forward(sd, x, name, return_features=False):
    """DPTDepthModel.forward: x [B,3,H,W] -> [B,H,W].  With return_features also the six activations ZoeDepth's MidasCore
    hooks (dzoedepth/models/base_models/midas.py:298-316): output_conv child 3 (the 32-channel ReLU), layer4_rn, refinenet4..1."""
    cfg = CONFIGS[name]
    B, _, H, W = x.shape
    gh, gw = H // 16, W // 16
    hooks = vit_backbone_hooks(sd, x, cfg) if cfg.get('family') == 'vit' else backbone_hooks(sd, x, cfg)
    layers = []
    for j, t in enumerate(hooks, start=1):
        a = f'pretrained.act_postprocess{j}.'
        readout = t[:, 0].unsqueeze(1).expand_as(t[:, 1:])
        f = torch.cat((t[:, 1:], readout), -1)
        f = F.gelu(F.linear(f, sd[a + '0.project.0.weight'].float(), sd[a + '0.project.0.bias'].float()))
        f = f.transpose(1, 2).reshape(B, -1, gh, gw)
        f = _conv(sd, a + '3', f)
        if j == 1:
            f = F.conv_transpose2d(f, sd[a + '4.weight'].float(), sd[a + '4.bias'].float(), stride=4)
        elif j == 2:
            f = F.conv_transpose2d(f, sd[a + '4.weight'].float(), sd[a + '4.bias'].float(), stride=2)
        elif j == 4:
            f = _conv(sd, a + '4', f, stride=2, padding=1)
        layers.append(f)
    l1 = _conv(sd, 'scratch.layer1_rn', layers[0], padding=1)
    l2 = _conv(sd, 'scratch.layer2_rn', layers[1], padding=1)
    l3 = _conv(sd, 'scratch.layer3_rn', layers[2], padding=1)
    l4 = _conv(sd, 'scratch.layer4_rn', layers[3], padding=1)
    p4 = _fusion(sd, 'scratch.refinenet4', l4, size=l3.shape[2:])
    p3 = _fusion(sd, 'scratch.refinenet3', p4, l3, size=l2.shape[2:])
    p2 = _fusion(sd, 'scratch.refinenet2', p3, l2, size=l1.shape[2:])
    p1 = _fusion(sd, 'scratch.refinenet1', p2, l1)
    o = _conv(sd, 'scratch.output_conv.0', p1, padding=1)
    o = F.interpolate(o, scale_factor=2, mode="bilinear", align_corners=True)
    o = F.relu(_conv(sd, 'scratch.output_conv.2', o, padding=1))
    feats = [o, l4, p4, p3, p2, p1]
    o = F.relu(_conv(sd, 'scratch.output_conv.4', o))
    if return_features:
        return o.squeeze(1), feats
    return o.squeeze(1)


@torch.no_grad()
def get_raw_prediction(rgb_uint8, sd, name, net_w, net_h):
    """ModelHolder.get_raw_prediction for model types 1 / 2 / 3 -> (float32 [H,W], invert=False)."""
    img = np.asarray(rgb_uint8)
    x = preprocess(img, net_w, net_h)
    pred = forward(sd, x, name)
    pred = F.interpolate(pred.unsqueeze(1), size=img.shape[:2], mode="bicubic", align_corners=False).squeeze()
    return pred.cpu().numpy(), False
