"""Tool for ncu: one eager forward of the BOOST merge U-Net (and one LeReS forward at 896) between cudaProfilerStart / Stop.
usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/profile_unet.py"""
import os
import sys

os.environ["DEPTHMAP_B200_UNET_GRAPH"] = "0"
os.environ["DEPTHMAP_B200_LERES_GRAPH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from depthmap_b200.boost import UnetMergeEngine
    from depthmap_b200.depthmap_generation import LeresEngine
    from oracle import synth_weights
    dev = torch.device("cuda")
    which = sys.argv[1] if len(sys.argv) > 1 else "unet"
    if which == "unet":
        eng = UnetMergeEngine(synth_weights.make_pix2pix_state_dict(seed=1), dev)
        x2 = torch.rand(1024, 1024, 2, device=dev) * 2 - 1
        fn = lambda: eng.forward(x2)
    elif which == "patch":          # one BOOST patch: LeReS at 448 / 896, two merge-net forwards and every glue kernel, all eager
        from depthmap_b200.boost import BoostPipeline
        pipe = BoostPipeline(LeresEngine(synth_weights.make_leres_state_dict(seed=2), dev), UnetMergeEngine(synth_weights.make_pix2pix_state_dict(seed=1), dev), dev, 0)
        img = torch.rand(3, 1024, 1024, device=dev)
        base = torch.rand(1024, 1024, device=dev)
        upd = base.clone()

        def fn():
            mapped, sums = pipe.fitted_patch(img, base, (100, 100, 600, 600), 448)
            pipe.blend(upd, mapped, sums, (100, 100, 600, 600))
    else:
        eng = LeresEngine(synth_weights.make_leres_state_dict(seed=2), dev)
        img = torch.rand(3, 1024, 1024, device=dev)
        fn = lambda: eng.forward_batch(None, 896, 896, planar=(img, (0, 0, 900, 900)))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", which)


if __name__ == "__main__":
    main()
