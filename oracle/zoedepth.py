"""ORACLE (test infrastructure only — the product never imports this): plain-torch fp32 functional restatement of the
ZoeDepth-NK metric head, the part of SURVEY.md §8a row D7 that sits on top of the MiDaS DPT-BEiT-L-384 core.

Follows /root/reference:
  dzoedepth/models/zoedepth_nk/zoedepth_nk_v1.py:159-243   ZoeDepthNK.forward (router, seed bins, attractors, log-binomial)
  dzoedepth/models/layers/patch_transformer.py:29-92       PatchTransformerEncoder (router embedding)
  dzoedepth/models/layers/localbins_layers.py               SeedBinRegressorUnnormed, Projector
  dzoedepth/models/layers/attractor.py:26-38,127-208        inv_attractor, AttractorLayerUnnormed (memory_efficient, kind "mean")
  dzoedepth/models/layers/dist_layers.py:29-121             log_binom, LogBinomial, ConditionalLogBinomial
with the hyper-parameters of dzoedepth/models/zoedepth_nk/config_zoedepth_nk.json (softplus bin centres, inverse attractors
alpha / gamma as the layers actually apply them, 64 bins, 128-d bin embedding, attractors [16, 8, 4, 1], temperatures 0.0212 .. 50).

Input: the six activations MidasCore hands to the head (`out = [out_conv (32 ch), l4_rn (256), r4, r3, r2, r1 (256 each)]`,
dzoedepth/models/base_models/midas.py:258-276) and the head's state_dict (keys as in the reference module, minus "core.").
Pinned against the reference module itself, built around a stub core, by tests/test_oracle_pin.py::test_zoedepth_head_*.
Round 1 ships only this oracle for D7; the CUDA head is round-2 work."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CONFIG = dict(
    bin_conf=[dict(name="nyu", n_bins=64, min_depth=1e-3, max_depth=10.0), dict(name="kitti", n_bins=64, min_depth=1e-3, max_depth=80.0)],
    bin_embedding_dim=128, n_attractors=[16, 8, 4, 1], attractor_alpha=1000, attractor_gamma=2, min_temp=0.0212, max_temp=50.0,
    router_dim=128, router_heads=4, router_layers=4, router_ffn=1024)


def _conv1x1(x, sd, prefix):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _mlp2(x, sd, prefix, act):
    """Conv1x1 -> act -> Conv1x1 (the `_net` of Projector / SeedBinRegressor / attractor layers: indices 0 and 2)."""
    return _conv1x1(act(_conv1x1(x, sd, prefix + "._net.0")), sd, prefix + "._net.2")


def _positional_encoding_1d(S, E, device):
    position = torch.arange(0, S, dtype=torch.float32, device=device).unsqueeze(1)
    index = torch.arange(0, E, 2, dtype=torch.float32, device=device).unsqueeze(0)
    div_term = torch.exp(index * (-torch.log(torch.tensor(10000.0, device=device)) / E))
    pe = position * div_term
    return torch.cat([torch.sin(pe), torch.cos(pe)], dim=1)          # [S, E]


def _encoder_layer(x, sd, p, heads):
    """nn.TransformerEncoderLayer defaults: post-norm, ReLU feed-forward, eval mode.  x: [S, N, E]."""
    S, N, E = x.shape
    qkv = F.linear(x, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    hd = E // heads

    def split(t):      # [S, N, E] -> [N * heads, S, hd]
        return t.contiguous().view(S, N * heads, hd).transpose(0, 1)
    q, k, v = split(q), split(k), split(v)
    att = torch.softmax((q * (1.0 / math.sqrt(hd))) @ k.transpose(-2, -1), dim=-1) @ v
    att = att.transpose(0, 1).contiguous().view(S, N, E)
    att = F.linear(att, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
    x = F.layer_norm(x + att, (E,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    ff = F.linear(F.relu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return F.layer_norm(x + ff, (E,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])


def router_logits(x_d0, sd):
    """patch_transformer(x)[0] -> mlp_classifier: [N, 2] domain logits (zoedepth_nk_v1.py:186-188)."""
    emb = _conv1x1(x_d0, sd, "patch_transformer.embedding_convPxP").flatten(2)      # patch size 1
    emb = F.pad(emb, (1, 0)).permute(2, 0, 1)                                       # zero "class token" first; [S, N, E]
    S, N, E = emb.shape
    emb = emb + _positional_encoding_1d(S, E, emb.device).unsqueeze(1).to(emb.dtype)
    for i in range(CONFIG["router_layers"]):
        emb = _encoder_layer(emb, sd, f"patch_transformer.transformer_encoder.layers.{i}", CONFIG["router_heads"])
    e0 = emb[0]
    return F.linear(F.relu(F.linear(e0, sd["mlp_classifier.0.weight"], sd["mlp_classifier.0.bias"])), sd["mlp_classifier.2.weight"], sd["mlp_classifier.2.bias"])


def _inv_attractor(dx):
    """attractor.py:26-38.  The layers call `dist(dx)` WITHOUT their alpha / gamma (attractor.py:99-106,188-196), so the
    jit function's defaults apply — alpha = 300, gamma = 2 — whatever config_zoedepth_nk.json says (attractor_alpha 1000)."""
    return dx.div(1 + 300.0 * dx.pow(2))


def _attractor_unnormed(x, b_prev, prev_emb, sd, prefix):
    """AttractorLayerUnnormed.forward with interpolate=True, memory_efficient=True, kind='mean', attractor_type='inv'.
    The model passes its per-level attractor count as the layer's SECOND positional argument, which is `n_bins`
    (zoedepth_nk_v1.py:139-141 vs attractor.py:128), so every level really has the default 16 attractors; the count is
    read from the checkpoint tensor here."""
    n_attractors = sd[prefix + "._net.2.weight"].shape[0]
    x = x + F.interpolate(prev_emb, x.shape[-2:], mode="bilinear", align_corners=True)
    A = F.softplus(_mlp2(x, sd, prefix, F.relu))
    h, w = A.shape[-2:]
    b_centers = F.interpolate(b_prev, (h, w), mode="bilinear", align_corners=True)
    delta = torch.zeros_like(b_centers)
    for i in range(n_attractors):
        delta += _inv_attractor(A[:, i, ...].unsqueeze(1) - b_centers)
    delta = delta / n_attractors
    b_new = b_centers + delta
    return b_new, b_new


def _log_binom(n, k, eps=1e-7):
    n = n + eps
    k = k + eps
    return n * torch.log(n) - k * torch.log(k) - (n - k) * torch.log(n - k + eps)


def _conditional_log_binomial(x, cond, sd, prefix, n_classes):
    pt = F.softplus(_conv1x1(F.gelu(_conv1x1(torch.cat((x, cond), dim=1), sd, prefix + ".mlp.0")), sd, prefix + ".mlp.2"))
    p, t = pt[:, :2, ...], pt[:, 2:, ...]
    p = p + 1e-4
    p = p[:, 0, ...] / (p[:, 0, ...] + p[:, 1, ...])
    t = t + 1e-4
    t = t[:, 0, ...] / (t[:, 0, ...] + t[:, 1, ...])
    t = (CONFIG["max_temp"] - CONFIG["min_temp"]) * t.unsqueeze(1) + CONFIG["min_temp"]
    xx = p.unsqueeze(1)
    k_idx = torch.arange(0, n_classes, device=x.device).view(1, -1, 1, 1)
    K_minus_1 = torch.tensor([float(n_classes - 1)], device=x.device).view(1, -1, 1, 1)
    one_minus = torch.clamp(1 - xx, 1e-4, 1)
    xx = torch.clamp(xx, 1e-4, 1)
    y = _log_binom(K_minus_1, k_idx) + k_idx * torch.log(xx) + (n_classes - 1 - k_idx) * torch.log(one_minus)
    return torch.softmax(y / t, dim=1)


def metric_head(features, sd):
    """features = [out_conv, btlnck, x_block4, x_block3, x_block2, x_block1] (fp32 NCHW).  Returns (metric_depth [N,1,H,W],
    domain_logits [N,2], chosen head name)."""
    outconv, btlnck, x_blocks = features[0], features[1], features[2:]
    x = _conv1x1(btlnck, sd, "conv2")
    logits = router_logits(x, sd)
    vote = torch.softmax(logits.sum(dim=0, keepdim=True), dim=-1)
    name = ["nyu", "kitti"][int(torch.argmax(vote, dim=-1).squeeze().item())]
    conf = [c for c in CONFIG["bin_conf"] if c["name"] == name][0]
    b_prev = F.softplus(_mlp2(x, sd, f"seed_bin_regressors.{name}", F.relu))          # SeedBinRegressorUnnormed
    prev_emb = _mlp2(x, sd, "seed_projector", F.relu)
    b_centers = b_emb = None
    for i, xb in enumerate(x_blocks):
        b_emb = _mlp2(xb, sd, f"projectors.{i}", F.relu)
        b_prev, b_centers = _attractor_unnormed(b_emb, b_prev, prev_emb, sd, f"attractors.{name}.{i}")
        prev_emb = b_emb
    b_centers = F.interpolate(b_centers, outconv.shape[-2:], mode="bilinear", align_corners=True)
    b_emb = F.interpolate(b_emb, outconv.shape[-2:], mode="bilinear", align_corners=True)
    probs = _conditional_log_binomial(outconv, b_emb, sd, f"conditional_log_binomial.{name}", conf["n_bins"])
    return torch.sum(probs * b_centers, dim=1, keepdim=True), logits, name


# ---------------------------------------------------------------------------------------------------------------------
# test-time augmentation wrapper: dzoedepth/models/depth_model.py:57-129 (DepthModel.infer = flip aug over pad aug)
# ---------------------------------------------------------------------------------------------------------------------
def infer_with_pad_aug(model_fn, x, pad_input=True, fh=3.0, fw=3.0):
    """model_fn(x [B,3,H,W]) -> metric depth [B,1,h,w].  Reflect-pads by int(sqrt(H/2)*fh) rows / int(sqrt(W/2)*fw) columns,
    runs the model, brings the output to the padded input size (bicubic, align_corners=False) and crops the padding away."""
    import numpy as np
    assert x.dim() == 4 and x.shape[1] == 3
    pad_h = pad_w = 0
    if pad_input:
        pad_h = int(np.sqrt(x.shape[2] / 2) * fh)
        pad_w = int(np.sqrt(x.shape[3] / 2) * fw)
        padding = [pad_w, pad_w]
        if pad_h > 0:
            padding += [pad_h, pad_h]
        x = F.pad(x, padding, mode="reflect")
    out = model_fn(x)
    if out.shape[-2:] != x.shape[-2:]:
        out = F.interpolate(out, size=(x.shape[2], x.shape[3]), mode="bicubic", align_corners=False)
    if pad_input:
        if pad_h > 0:
            out = out[:, :, pad_h:-pad_h, :]
        if pad_w > 0:
            out = out[:, :, :, pad_w:-pad_w]
    return out


def infer(model_fn, x, pad_input=True, with_flip_aug=True):
    """DepthModel.infer: mean of the prediction and the un-flipped prediction of the horizontally flipped input."""
    out = infer_with_pad_aug(model_fn, x, pad_input)
    if not with_flip_aug:
        return out
    out_flip = infer_with_pad_aug(model_fn, torch.flip(x, dims=[3]), pad_input)
    return (out + torch.flip(out_flip, dims=[3])) / 2


# ---------------------------------------------------------------------------------------------------------------------
# the whole D7 path: estimatezoedepth (src/depthmap_generation.py:443-452) -> DepthModel.infer_pil -> MidasCore -> head.
# The DPT-BEiT-L-384 core is oracle/beit_dpt.py (parity unpinned: timm is absent, see its header); everything after the
# core's activations is pinned above.
# ---------------------------------------------------------------------------------------------------------------------
def prep_for_midas(x, net_w, net_h):
    """PrepForMidas (midas.py:175-186) as ZoeDepth-NK configures it: keep_aspect_ratio (force_keep_ar), multiple of 32,
    resize_method "minimal", bilinear align_corners=True, then (x - 0.5) / 0.5."""
    from oracle import beit_dpt
    w, h = beit_dpt.get_size_minimal(x.shape[3], x.shape[2], net_w, net_h, 32)
    x = F.interpolate(x, (int(h), int(w)), mode="bilinear", align_corners=True)
    return (x - 0.5) / 0.5


@torch.no_grad()
def get_raw_prediction(rgb_uint8, sd, net_w=512, net_h=384, core_name="beitl16_384"):
    """ModelHolder.get_raw_prediction for model type 9 (zoedepth_nk): (float32 metric depth [H,W], invert=True).
    sd: ZoeDepth checkpoint layout — MiDaS weights under "core.core.", head weights at the top level."""
    import numpy as np
    from oracle import beit_dpt
    core_sd = {k[len("core.core."):]: v for k, v in sd.items() if k.startswith("core.core.")}
    head_sd = {k: v for k, v in sd.items() if not k.startswith("core.")}

    def model_fn(x):
        xin = prep_for_midas(x, net_w, net_h)
        _, feats = beit_dpt.forward(core_sd, xin, core_name, return_features=True)
        return metric_head(feats, head_sd)[0]

    img = np.asarray(rgb_uint8)
    x = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).float().div(255.0).unsqueeze(0)      # transforms.ToTensor
    out = infer(model_fn, x, pad_input=True, with_flip_aug=True)
    return out.squeeze().cpu().numpy(), True
