"""CPU: the BEiT trunk of oracle/beit_dpt.py against an INDEPENDENT implementation of the same architecture.

timm (the library the reference builds its BEiT on, requirements.txt:8) is absent from this image, so the BEiT oracle
cannot be pinned to the reference module itself.  HuggingFace transformers ships its own BEiT (`BeitModel`, a port of
the original microsoft/unilm code, the same source timm's Beit derives from).  Loading the oracle's timm-layout
weights into it and comparing the hidden states after every block checks everything timm contributes to the reference
path: parameter layout (q_bias / v_bias, no k bias), the relative-position index rule (`gen_relative_position_index`),
LayerScale placement and pre-norm residual order.  The reference's own overrides (dmidas/backbones/beit.py:29-62: table
resize to the current window) reduce to the stock path at the native resolution, which is what is compared here; a
non-native window additionally checks the oracle's bilinear table resize against HF's `interpolate_pos_encoding`."""
import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _hf_model(cfg, sd):
    from transformers import BeitConfig, BeitModel
    C, depth, heads, win = cfg['embed_dim'], cfg['depth'], cfg['heads'], cfg['window']
    hc = BeitConfig(hidden_size=C, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * C,
                    image_size=16 * win, patch_size=16, use_relative_position_bias=True,
                    use_shared_relative_position_bias=False, use_absolute_position_embeddings=False,
                    layer_scale_init_value=0.1, use_mean_pooling=True, hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0, drop_path_rate=0.0, use_mask_token=False, layer_norm_eps=1e-6,
                    hidden_act="gelu")
    m = BeitModel(hc, add_pooling_layer=False).eval()
    p = 'pretrained.model.'
    new = {'embeddings.cls_token': sd[p + 'cls_token'],
           'embeddings.patch_embeddings.projection.weight': sd[p + 'patch_embed.proj.weight'],
           'embeddings.patch_embeddings.projection.bias': sd[p + 'patch_embed.proj.bias']}
    for i in range(depth):
        b, h = p + f'blocks.{i}.', f'encoder.layer.{i}.'
        qkv = sd[b + 'attn.qkv.weight']
        new[h + 'attention.attention.query.weight'] = qkv[:C]
        new[h + 'attention.attention.key.weight'] = qkv[C:2 * C]
        new[h + 'attention.attention.value.weight'] = qkv[2 * C:]
        new[h + 'attention.attention.query.bias'] = sd[b + 'attn.q_bias']
        new[h + 'attention.attention.value.bias'] = sd[b + 'attn.v_bias']
        new[h + 'attention.attention.relative_position_bias.relative_position_bias_table'] = sd[b + 'attn.relative_position_bias_table']
        new[h + 'attention.output.dense.weight'] = sd[b + 'attn.proj.weight']
        new[h + 'attention.output.dense.bias'] = sd[b + 'attn.proj.bias']
        new[h + 'intermediate.dense.weight'] = sd[b + 'mlp.fc1.weight']
        new[h + 'intermediate.dense.bias'] = sd[b + 'mlp.fc1.bias']
        new[h + 'output.dense.weight'] = sd[b + 'mlp.fc2.weight']
        new[h + 'output.dense.bias'] = sd[b + 'mlp.fc2.bias']
        new[h + 'lambda_1'] = sd[b + 'gamma_1']
        new[h + 'lambda_2'] = sd[b + 'gamma_2']
        new[h + 'layernorm_before.weight'] = sd[b + 'norm1.weight']
        new[h + 'layernorm_before.bias'] = sd[b + 'norm1.bias']
        new[h + 'layernorm_after.weight'] = sd[b + 'norm2.weight']
        new[h + 'layernorm_after.bias'] = sd[b + 'norm2.bias']
    res = m.load_state_dict({k: v.float() for k, v in new.items()}, strict=False)
    # everything the HF module owns must have been provided (buffers such as the index are rebuilt by HF itself)
    assert not [k for k in res.missing_keys if 'relative_position_index' not in k], res.missing_keys
    assert not res.unexpected_keys
    return m


@pytest.mark.parametrize("hw", [(64, 64), (96, 64), (64, 128)])
def test_beit_trunk_equals_hf_beit(hw):
    from oracle import beit_dpt, synth_weights
    cfg = beit_dpt.CONFIGS['beit_tiny']
    sd = synth_weights.make_beit_dpt_state_dict('beit_tiny', seed=5)
    m = _hf_model(cfg, sd)
    x = torch.randn(2, 3, hw[0], hw[1], generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(pixel_values=x, output_hidden_states=True, interpolate_pos_encoding=hw != (64, 64)).hidden_states
        got = beit_dpt.backbone_hooks(sd, x, cfg)       # hooks = every block for beit_tiny
    assert len(got) == cfg['depth']
    for i, g in enumerate(got):
        w = want[i + 1]
        err = float((g - w).abs().max()) / float(w.abs().max())
        assert err < 2e-5, (hw, i, err)


def test_relative_position_index_equals_hf():
    """timm's gen_relative_position_index (restated in the oracle and in the product) == HF's generate_relative_position_index."""
    from oracle import beit_dpt
    from transformers.models.beit.modeling_beit import BeitRelativePositionBias
    from transformers import BeitConfig
    for win in [(4, 4), (3, 5), (24, 24), (32, 48)]:
        hf = BeitRelativePositionBias(BeitConfig(num_attention_heads=2), window_size=win)
        want = hf.generate_relative_position_index(win)
        got = beit_dpt.gen_relative_position_index(win)
        assert torch.equal(want, got), win


# ---------------------------------------------------------------------------------------------------------------------
# dpt_large_384 (model type 3): the ViT-L/16 trunk of the oracle against HuggingFace's ViTModel at the native resolution
# (the reference resizes the position grid bilinearly, HF bicubically, so only the native grid is comparable)
# ---------------------------------------------------------------------------------------------------------------------
def test_vit_trunk_equals_hf_vit():
    from transformers import ViTConfig, ViTModel
    from oracle import beit_dpt, synth_weights
    cfg = beit_dpt.CONFIGS['vit_tiny']
    sd = synth_weights.make_beit_dpt_state_dict('vit_tiny', seed=6)
    C, depth, heads, win = cfg['embed_dim'], cfg['depth'], cfg['heads'], cfg['window']
    hc = ViTConfig(hidden_size=C, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * C, image_size=16 * win,
                   patch_size=16, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, layer_norm_eps=1e-6, hidden_act="gelu", qkv_bias=True)
    m = ViTModel(hc, add_pooling_layer=False).eval()
    p = 'pretrained.model.'
    new = {'embeddings.cls_token': sd[p + 'cls_token'], 'embeddings.position_embeddings': sd[p + 'pos_embed'],
           'embeddings.patch_embeddings.projection.weight': sd[p + 'patch_embed.proj.weight'],
           'embeddings.patch_embeddings.projection.bias': sd[p + 'patch_embed.proj.bias'],
           'layernorm.weight': sd[p + 'norm.weight'], 'layernorm.bias': sd[p + 'norm.bias']}
    for i in range(depth):
        b, h = p + f'blocks.{i}.', f'encoder.layer.{i}.'
        qw, qb = sd[b + 'attn.qkv.weight'], sd[b + 'attn.qkv.bias']
        for k, nm in enumerate(('query', 'key', 'value')):
            new[h + f'attention.attention.{nm}.weight'] = qw[k * C:(k + 1) * C]
            new[h + f'attention.attention.{nm}.bias'] = qb[k * C:(k + 1) * C]
        new[h + 'attention.output.dense.weight'] = sd[b + 'attn.proj.weight']
        new[h + 'attention.output.dense.bias'] = sd[b + 'attn.proj.bias']
        new[h + 'intermediate.dense.weight'] = sd[b + 'mlp.fc1.weight']
        new[h + 'intermediate.dense.bias'] = sd[b + 'mlp.fc1.bias']
        new[h + 'output.dense.weight'] = sd[b + 'mlp.fc2.weight']
        new[h + 'output.dense.bias'] = sd[b + 'mlp.fc2.bias']
        new[h + 'layernorm_before.weight'] = sd[b + 'norm1.weight']
        new[h + 'layernorm_before.bias'] = sd[b + 'norm1.bias']
        new[h + 'layernorm_after.weight'] = sd[b + 'norm2.weight']
        new[h + 'layernorm_after.bias'] = sd[b + 'norm2.bias']
    res = m.load_state_dict({k: v.float() for k, v in new.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = m(pixel_values=x, output_hidden_states=True).hidden_states
        got = beit_dpt.vit_backbone_hooks(sd, x, cfg)
    for i, g in enumerate(got):
        err = float((g - want[i + 1]).abs().max()) / float(want[i + 1].abs().max())
        assert err < 2e-5, (i, err)
