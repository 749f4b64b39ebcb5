"""ORACLE / TEST + BENCH DATA.  Seeded synthetic state_dicts in the reference's checkpoint layout.

No checkpoint is reachable offline, and default-initialised DA-v2 weights give an identically-zero output after the
trailing ReLUs, so fixtures use a scaled random init that keeps activations O(1) through 24 blocks and
the decoder and yields a non-degenerate, strictly varying depth.  Key names and shapes are exactly those of
DepthAnythingV2(**cfg).state_dict() (checked with strict load in tests/test_oracle_pin.py).
"""
from __future__ import annotations

import torch

from .dav2 import CONFIGS


def make_dav2_state_dict(encoder='vits', seed=0, num_pos=37 * 37 + 1, dtype=torch.float32):
    cfg = CONFIGS[encoder]
    C, depth, Fch, oc = cfg['embed_dim'], cfg['depth'], cfg['features'], cfg['out_channels']
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).to(dtype)

    sd = {}
    sd['pretrained.cls_token'] = rn(1, 1, C, std=0.02)
    sd['pretrained.pos_embed'] = rn(1, num_pos, C, std=0.1)
    sd['pretrained.mask_token'] = torch.zeros(1, C, dtype=dtype)
    sd['pretrained.patch_embed.proj.weight'] = rn(C, 3, 14, 14, std=(1.0 / 588) ** 0.5)
    sd['pretrained.patch_embed.proj.bias'] = rn(C, std=0.02)
    for i in range(depth):
        p = f'pretrained.blocks.{i}.'
        sd[p + 'norm1.weight'] = rn(C, std=0.05, mean=1.0)
        sd[p + 'norm1.bias'] = rn(C, std=0.02)
        sd[p + 'attn.qkv.weight'] = rn(3 * C, C, std=C ** -0.5)
        sd[p + 'attn.qkv.bias'] = rn(3 * C, std=0.02)
        sd[p + 'attn.proj.weight'] = rn(C, C, std=C ** -0.5)
        sd[p + 'attn.proj.bias'] = rn(C, std=0.02)
        sd[p + 'ls1.gamma'] = rn(C, std=0.02, mean=0.2)
        sd[p + 'norm2.weight'] = rn(C, std=0.05, mean=1.0)
        sd[p + 'norm2.bias'] = rn(C, std=0.02)
        sd[p + 'mlp.fc1.weight'] = rn(4 * C, C, std=C ** -0.5)
        sd[p + 'mlp.fc1.bias'] = rn(4 * C, std=0.02)
        sd[p + 'mlp.fc2.weight'] = rn(C, 4 * C, std=(4 * C) ** -0.5)
        sd[p + 'mlp.fc2.bias'] = rn(C, std=0.02)
        sd[p + 'ls2.gamma'] = rn(C, std=0.02, mean=0.2)
    sd['pretrained.norm.weight'] = rn(C, std=0.05, mean=1.0)
    sd['pretrained.norm.bias'] = rn(C, std=0.02)

    h = 'depth_head.'
    for i in range(4):
        sd[h + f'projects.{i}.weight'] = rn(oc[i], C, 1, 1, std=C ** -0.5)
        sd[h + f'projects.{i}.bias'] = rn(oc[i], std=0.02)
    sd[h + 'resize_layers.0.weight'] = rn(oc[0], oc[0], 4, 4, std=oc[0] ** -0.5)
    sd[h + 'resize_layers.0.bias'] = rn(oc[0], std=0.02)
    sd[h + 'resize_layers.1.weight'] = rn(oc[1], oc[1], 2, 2, std=oc[1] ** -0.5)
    sd[h + 'resize_layers.1.bias'] = rn(oc[1], std=0.02)
    sd[h + 'resize_layers.3.weight'] = rn(oc[3], oc[3], 3, 3, std=(9 * oc[3]) ** -0.5)
    sd[h + 'resize_layers.3.bias'] = rn(oc[3], std=0.02)
    for i in range(4):
        sd[h + f'scratch.layer{i + 1}_rn.weight'] = rn(Fch, oc[i], 3, 3, std=(9 * oc[i]) ** -0.5)
    for i in range(1, 5):
        r = h + f'scratch.refinenet{i}.'
        sd[r + 'out_conv.weight'] = rn(Fch, Fch, 1, 1, std=Fch ** -0.5)
        sd[r + 'out_conv.bias'] = rn(Fch, std=0.02)
        for u in ('resConfUnit1', 'resConfUnit2'):
            for cv in ('conv1', 'conv2'):
                sd[r + f'{u}.{cv}.weight'] = rn(Fch, Fch, 3, 3, std=(2.0 / (9 * Fch)) ** 0.5 * 0.7)
                sd[r + f'{u}.{cv}.bias'] = rn(Fch, std=0.02)
    sd[h + 'scratch.output_conv1.weight'] = rn(Fch // 2, Fch, 3, 3, std=(9 * Fch) ** -0.5)
    sd[h + 'scratch.output_conv1.bias'] = rn(Fch // 2, std=0.02)
    sd[h + 'scratch.output_conv2.0.weight'] = rn(32, Fch // 2, 3, 3, std=(2.0 / (9 * (Fch // 2))) ** 0.5)
    sd[h + 'scratch.output_conv2.0.bias'] = rn(32, std=0.05, mean=0.1)
    sd[h + 'scratch.output_conv2.2.weight'] = rn(1, 32, 1, 1, std=0.08, mean=0.05)
    sd[h + 'scratch.output_conv2.2.bias'] = torch.full((1,), 0.5, dtype=dtype)
    return sd


def make_beit_dpt_state_dict(name='beit_tiny', seed=0, dtype=torch.float32):
    """Seeded synthetic state_dict in the layout of the MiDaS 3.1 checkpoints (dpt_beit_large_512.pt): keys
    `pretrained.model.*` (timm Beit), `pretrained.act_postprocess{1..4}.*`, `scratch.*`."""
    from .beit_dpt import CONFIGS as BCFG
    cfg = BCFG[name]
    C, depth, heads, Fch, oc, win = cfg['embed_dim'], cfg['depth'], cfg['heads'], cfg['features'], cfg['out_channels'], cfg['window']
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).to(dtype)

    sd = {}
    p = 'pretrained.model.'
    sd[p + 'cls_token'] = rn(1, 1, C, std=0.5)
    sd[p + 'patch_embed.proj.weight'] = rn(C, 3, 16, 16, std=(1.0 / 768) ** 0.5 * 2.0)
    sd[p + 'patch_embed.proj.bias'] = rn(C, std=0.02)
    nrd = (2 * win - 1) * (2 * win - 1) + 3
    vit = cfg.get('family') == 'vit'
    if vit:
        sd[p + 'pos_embed'] = rn(1, win * win + 1, C, std=0.1)
        sd[p + 'norm.weight'] = rn(C, std=0.05, mean=1.0)
        sd[p + 'norm.bias'] = rn(C, std=0.02)
    for i in range(depth):
        b = p + f'blocks.{i}.'
        sd[b + 'norm1.weight'] = rn(C, std=0.05, mean=1.0)
        sd[b + 'norm1.bias'] = rn(C, std=0.02)
        if vit:     # no LayerScale: keep the residual branches small through the weights instead
            sd[b + 'attn.qkv.bias'] = rn(3 * C, std=0.02)
        else:
            sd[b + 'gamma_1'] = rn(C, std=0.02, mean=0.2)
            sd[b + 'gamma_2'] = rn(C, std=0.02, mean=0.2)
            sd[b + 'attn.q_bias'] = rn(C, std=0.02)
            sd[b + 'attn.v_bias'] = rn(C, std=0.02)
            sd[b + 'attn.relative_position_bias_table'] = rn(nrd, heads, std=1.0)
        sd[b + 'attn.qkv.weight'] = rn(3 * C, C, std=C ** -0.5)
        sd[b + 'attn.proj.weight'] = rn(C, C, std=C ** -0.5 * (0.2 if vit else 1.0))
        sd[b + 'attn.proj.bias'] = rn(C, std=0.02)
        sd[b + 'norm2.weight'] = rn(C, std=0.05, mean=1.0)
        sd[b + 'norm2.bias'] = rn(C, std=0.02)
        sd[b + 'mlp.fc1.weight'] = rn(4 * C, C, std=C ** -0.5)
        sd[b + 'mlp.fc1.bias'] = rn(4 * C, std=0.02)
        sd[b + 'mlp.fc2.weight'] = rn(C, 4 * C, std=(4 * C) ** -0.5 * (0.2 if vit else 1.0))
        sd[b + 'mlp.fc2.bias'] = rn(C, std=0.02)
    for j in range(1, 5):
        a = f'pretrained.act_postprocess{j}.'
        sd[a + '0.project.0.weight'] = rn(C, 2 * C, std=(2 * C) ** -0.5)
        sd[a + '0.project.0.bias'] = rn(C, std=0.02)
        sd[a + '3.weight'] = rn(oc[j - 1], C, 1, 1, std=C ** -0.5 * 1.5)
        sd[a + '3.bias'] = rn(oc[j - 1], std=0.02)
    sd['pretrained.act_postprocess1.4.weight'] = rn(oc[0], oc[0], 4, 4, std=oc[0] ** -0.5)
    sd['pretrained.act_postprocess1.4.bias'] = rn(oc[0], std=0.02)
    sd['pretrained.act_postprocess2.4.weight'] = rn(oc[1], oc[1], 2, 2, std=oc[1] ** -0.5)
    sd['pretrained.act_postprocess2.4.bias'] = rn(oc[1], std=0.02)
    sd['pretrained.act_postprocess4.4.weight'] = rn(oc[3], oc[3], 3, 3, std=(9 * oc[3]) ** -0.5)
    sd['pretrained.act_postprocess4.4.bias'] = rn(oc[3], std=0.02)
    for i in range(4):
        sd[f'scratch.layer{i + 1}_rn.weight'] = rn(Fch, oc[i], 3, 3, std=(9 * oc[i]) ** -0.5)
    for i in range(1, 5):
        r = f'scratch.refinenet{i}.'
        sd[r + 'out_conv.weight'] = rn(Fch, Fch, 1, 1, std=Fch ** -0.5)
        sd[r + 'out_conv.bias'] = rn(Fch, std=0.02)
        for u in ('resConfUnit1', 'resConfUnit2'):
            for cv in ('conv1', 'conv2'):
                sd[r + f'{u}.{cv}.weight'] = rn(Fch, Fch, 3, 3, std=(2.0 / (9 * Fch)) ** 0.5 * 0.7)
                sd[r + f'{u}.{cv}.bias'] = rn(Fch, std=0.02)
    sd['scratch.output_conv.0.weight'] = rn(Fch // 2, Fch, 3, 3, std=(9 * Fch) ** -0.5)
    sd['scratch.output_conv.0.bias'] = rn(Fch // 2, std=0.02)
    sd['scratch.output_conv.2.weight'] = rn(32, Fch // 2, 3, 3, std=(2.0 / (9 * (Fch // 2))) ** 0.5)
    sd['scratch.output_conv.2.bias'] = rn(32, std=0.05, mean=0.1)
    sd['scratch.output_conv.4.weight'] = rn(1, 32, 1, 1, std=0.08, mean=0.05)
    sd['scratch.output_conv.4.bias'] = torch.full((1,), 0.5, dtype=dtype)
    return sd


def make_zoedepth_head_state_dict(feat_ch=256, out_conv_ch=32, seed=0, dtype=torch.float32, gain=1.0):
    """Seeded ZoeDepth-NK metric-head weights in the checkpoint layout of dzoedepth/models/zoedepth_nk/zoedepth_nk_v1.py
    (keys as `ZoeDepthNK.state_dict()` minus "core."): 64 bins, 128-d bin embedding, 16 attractors per level (the count the
    reference really instantiates, see oracle/zoedepth.py), 4-layer / 4-head / 1024-ffn router."""
    g = torch.Generator().manual_seed(seed)

    def w(*shape, s=0.05):
        return (torch.randn(*shape, generator=g) * s).to(dtype)

    sd = {}

    def conv(key, cout, cin):
        # gain > 1: fan-in scaled weights times `gain`, so that bin centres, attractor points and the log-binomial
        # parameters really vary over the image (the 0.05-std default gives an almost constant depth map)
        sd[key + ".weight"] = w(cout, cin, 1, 1) if gain == 1.0 else w(cout, cin, 1, 1, s=gain * cin ** -0.5)
        sd[key + ".bias"] = w(cout, s=0.02 if gain == 1.0 else 0.3)

    def lin(key, cout, cin):
        sd[key + ".weight"] = w(cout, cin)
        sd[key + ".bias"] = w(cout, s=0.02)

    E, emb, nb = 128, 128, 64
    conv("conv2", feat_ch, feat_ch)
    conv("patch_transformer.embedding_convPxP", E, feat_ch)
    for i in range(4):
        p = f"patch_transformer.transformer_encoder.layers.{i}"
        sd[p + ".self_attn.in_proj_weight"] = w(3 * E, E)
        sd[p + ".self_attn.in_proj_bias"] = w(3 * E, s=0.02)
        lin(p + ".self_attn.out_proj", E, E)
        lin(p + ".linear1", 1024, E)
        lin(p + ".linear2", E, 1024)
        for n in ("norm1", "norm2"):
            sd[p + f".{n}.weight"] = 1.0 + w(E, s=0.02)
            sd[p + f".{n}.bias"] = w(E, s=0.02)
    lin("mlp_classifier.0", 128, E)
    lin("mlp_classifier.2", 2, 128)
    for name in ("nyu", "kitti"):
        conv(f"seed_bin_regressors.{name}._net.0", emb // 2, feat_ch)
        conv(f"seed_bin_regressors.{name}._net.2", nb, emb // 2)
        for i in range(4):
            conv(f"attractors.{name}.{i}._net.0", emb, emb)
            conv(f"attractors.{name}.{i}._net.2", 16, emb)
        bott = (out_conv_ch + emb) // 4
        conv(f"conditional_log_binomial.{name}.mlp.0", bott, out_conv_ch + emb)
        conv(f"conditional_log_binomial.{name}.mlp.2", 4, bott)
    conv("seed_projector._net.0", emb // 2, feat_ch)
    conv("seed_projector._net.2", emb, emb // 2)
    for i in range(4):
        conv(f"projectors.{i}._net.0", emb // 2, feat_ch)
        conv(f"projectors.{i}._net.2", emb, emb // 2)
    return sd


def make_leres_state_dict(seed=0, dtype=torch.float32):
    """Seeded synthetic state_dict with the keys and shapes of the reference's RelDepthModel('resnext101').state_dict()
    (lib/multi_depth_model_woauxi.py, lib/Resnext_torch.py, lib/network_auxi.py; `num_batches_tracked` buffers left out):
    ResNeXt-101 32x8d encoder + FTB / FFM / AO decoder.  Fan-in scaled filters, BatchNorm statistics near identity, bn3 scaled
    down so 33 residual blocks keep activations O(1) (the recipe tests/test_oracle_pin.py uses for the reference module)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, co, ci, k, bias=False):
        sd[key + '.weight'] = (torch.randn(co, ci, k, k, generator=g) * (1.0 / (ci * k * k) ** 0.5)).to(dtype)
        if bias:
            sd[key + '.bias'] = (0.05 * torch.randn(co, generator=g)).to(dtype)

    def bn(key, c, gain=1.0):
        sd[key + '.weight'] = (gain + 0.05 * torch.randn(c, generator=g)).to(dtype)
        sd[key + '.bias'] = (0.05 * torch.randn(c, generator=g)).to(dtype)
        sd[key + '.running_mean'] = (0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[key + '.running_var'] = (0.5 + torch.rand(c, generator=g)).to(dtype)

    E = 'depth_model.encoder_modules.encoder.'
    conv(E + 'conv1', 64, 3, 7)
    bn(E + 'bn1', 64)
    inplanes = 64
    for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), (3, 4, 23, 3)), start=1):
        width = planes * 8 // 64 * 32
        for bi in range(nb):
            p = f'{E}layer{li}.{bi}'
            conv(p + '.conv1', width, inplanes, 1)
            bn(p + '.bn1', width)
            conv(p + '.conv2', width, width // 32, 3)
            bn(p + '.bn2', width)
            conv(p + '.conv3', planes * 4, width, 1)
            bn(p + '.bn3', planes * 4, gain=0.3)
            if bi == 0:
                conv(p + '.downsample.0', planes * 4, inplanes, 1)
                bn(p + '.downsample.1', planes * 4)
            inplanes = planes * 4
    D = 'depth_model.decoder_modules.'

    def ftb(p, ci, cm):
        conv(p + '.conv1', cm, ci, 3, bias=True)
        conv(p + '.conv_branch.1', cm, cm, 3, bias=True)
        bn(p + '.conv_branch.2', cm)
        conv(p + '.conv_branch.4', cm, cm, 3, bias=True)
    ftb(D + 'conv', 2048, 512)
    conv(D + 'conv1', 256, 512, 3, bias=True)
    for k, ci in (('ffm2', 1024), ('ffm1', 512), ('ffm0', 256)):
        ftb(D + k + '.ftb1', ci, 256)
        ftb(D + k + '.ftb2', 256, 256)
    conv(D + 'outconv.adapt_conv.0', 128, 256, 3, bias=True)
    bn(D + 'outconv.adapt_conv.1', 128)
    conv(D + 'outconv.adapt_conv.3', 1, 128, 3, bias=True)
    return sd


def make_pix2pix_state_dict(seed=0, dtype=torch.float32):
    """Seeded synthetic state_dict with the keys and shapes of the BOOST merge generator, pix2pix `define_G(2, 1, 64, 'unet_1024',
    'none')` (pix2pix/models/networks.py:444-543): 10 nested UnetSkipConnectionBlocks, only the outermost ConvTranspose has a bias.
    Gain 1.3 / sqrt(fan_in) (fan-in of a stride-2 4x4 transposed conv = 4 taps x input channels) keeps a signal through the 20
    un-normalised layers without saturating the final tanh: output std ~0.4, input sensitivity ~6 (max) / ~1 (rms)."""
    g = torch.Generator().manual_seed(seed)
    ch = [(2, 64), (64, 128), (128, 256), (256, 512)] + [(512, 512)] * 6
    sd = {}
    prefix = "model."
    for d, (cin, cout) in enumerate(ch):
        if d == 0:
            kd, ku, child = prefix + "model.0.weight", prefix + "model.3.weight", prefix + "model.1."
        elif d == 9:
            kd, ku, child = prefix + "model.1.weight", prefix + "model.3.weight", None
        else:
            kd, ku, child = prefix + "model.1.weight", prefix + "model.5.weight", prefix + "model.3."
        sd[kd] = (torch.randn(cout, cin, 4, 4, generator=g) * (1.3 / (cin * 16) ** 0.5)).to(dtype)
        up_in = cout if d == 9 else 2 * cout
        up_out = 1 if d == 0 else cin
        sd[ku] = (torch.randn(up_in, up_out, 4, 4, generator=g) * (1.3 / (up_in * 4) ** 0.5)).to(dtype)
        if d == 0:
            sd[prefix + "model.3.bias"] = (0.05 * torch.randn(1, generator=g)).to(dtype)
        prefix = child
    return sd
