"""Tool (not a test): where does the fp16-operand error of the DA-v2 forward come from?  Compares the engine's
intermediate buffers with the fp32 oracle's, stage by stage.  usage: python tests/diag_precision.py [vits|vitb|vitl] [net]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from depthmap_b200.depthmap_generation import DepthAnythingV2Engine
    from oracle import dav2 as odav2
    from oracle import synth_weights
    from synth import synth_rgb
    enc = sys.argv[1] if len(sys.argv) > 1 else 'vits'
    net = int(sys.argv[2]) if len(sys.argv) > 2 else 140
    dev = torch.device("cuda")
    sd = synth_weights.make_dav2_state_dict(enc, seed=1)
    eng = DepthAnythingV2Engine(sd, enc, dev)
    img = synth_rgb(net, net, 3)
    got = eng.forward_batch(torch.from_numpy(img[None]).to(dev), net).cpu().numpy()[0]
    b = eng._bufs
    x, (h, w) = odav2.preprocess(img, net)
    with torch.no_grad():
        want, inter = odav2.forward(sd, x, enc, return_intermediates=True)
    cfg = odav2.CONFIGS[enc]
    Fch = cfg['features']

    def rep(name, ours, ref):
        ours = ours.float().cpu().numpy()
        ref = ref.float().cpu().numpy()
        rng = float(ref.max() - ref.min())
        rms = float(np.sqrt((ref ** 2).mean()))
        e = np.abs(ours - ref)
        print(f"{name:10s} max/range {e.max() / rng:.3e}  mean/range {e.mean() / rng:.3e}  max/rms {e.max() / rms:.3e}  mean/rms {e.mean() / rms:.3e}  range {rng:.3f} rms {rms:.3f}")

    for i in range(4):
        rep(f"feat{i}", b['feat'][i].reshape(1, -1, cfg['embed_dim']), inter['feats'][i])
    for i, k in enumerate(['l1rn', 'l2rn', 'l3rn', 'l4rn']):
        rep(k, b['l'][i][..., :Fch].permute(0, 3, 1, 2), inter[k])
    for i, k in enumerate(['p4', 'p3', 'p2', 'p1']):
        rep(k, b['path'][i][..., :Fch].permute(0, 3, 1, 2), inter[k])
    rep("d(net)", b['d'], want)
    rng = float(want.max() - want.min())
    print("final", np.abs(got - want[0].numpy()).max() / rng if got.shape == tuple(want.shape[1:]) else "resized")


if __name__ == "__main__":
    main()
