"""Golden values of the reference's cheap evaluation metrics (TEST INFRASTRUCTURE; needs /root/reference).

Runs the UNMODIFIED ``evaluation/metrics/{mse,psnr,motion_masked_mse,vgg_cosine_similarity}.py`` on seeded inputs (VGG19
with the seeded stand-in weights of ``caddy_oracle.make_vgg_weights``) -> tests/golden/metrics.npz.
usage: python oracle/make_metric_golden.py [--input-pipeline]   (the flag writes tests/golden/input_pipeline.npz instead)"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import caddy_oracle as O          # noqa: E402
from oracle import ref_harness as R           # noqa: E402

SEED, SHAPE = 123, (2, 4, 3, 32, 48)


def inputs():
    g = torch.Generator().manual_seed(SEED)
    ref = torch.rand(SHAPE, generator=g)
    gen = (ref + 0.1 * torch.randn(SHAPE, generator=g)).clamp(0, 1)
    return ref, gen


def main():
    R.install_shims()
    sys.path.insert(0, R.REF_ROOT)
    from evaluation.metrics.mse import MSE
    from evaluation.metrics.psnr import PSNR
    from evaluation.metrics.motion_masked_mse import MotionMaskedMSE
    from evaluation.metrics.vgg_cosine_similarity import VGGCosineSimilarity
    ref, gen = inputs()
    out = {"mse": MSE()(ref, gen).numpy(), "psnr": PSNR()(ref, gen).numpy(), "psnr_range255": PSNR()(ref * 255, gen * 255, range=255.0).numpy(),
           "motion_masked_mse": MotionMaskedMSE()(ref, gen).numpy()}
    vcs = VGGCosineSimilarity()
    sd = O.make_vgg_weights()
    vcs.vgg.load_state_dict({k: v for k, v in _slice_names(vcs.vgg, sd).items()}, strict=True)
    with torch.no_grad():
        out["vgg_cosine"] = vcs(ref, gen).numpy()
    # sequence helpers of training/losses.py used by the evaluator
    from training.losses import MotionLossWeightMaskCalculator, SequenceLossEvaluator, StatesLoss
    out["weight_mask_same"] = MotionLossWeightMaskCalculator(0.3).compute_weight_mask(ref, gen).numpy()
    out["weight_mask_short"] = MotionLossWeightMaskCalculator(0.0).compute_weight_mask(ref, gen[:, 1:]).numpy()
    ev = SequenceLossEvaluator(StatesLoss())
    avg, terms = ev(ref, gen)
    out["seq_same_avg"], out["seq_same_terms"] = avg.numpy(), terms.numpy()
    avg, terms = ev(ref, gen[:, 1:])
    out["seq_short_avg"], out["seq_short_terms"] = avg.numpy(), terms.numpy()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print({k: v.reshape(-1)[:3] for k, v in out.items()})


def _slice_names(vgg, features_sd):
    """The reference's Vgg19 regroups torchvision's features into slice1..slice5 keeping the layer indices."""
    want = vgg.state_dict()
    out = {}
    for k in want:
        idx = k.split(".")[1]
        out[k] = features_sd[f"features.{idx}.{k.split('.')[2]}"]
    return out


if __name__ == "__main__" and "--input-pipeline" not in sys.argv:
    main()


def input_pipeline_golden():
    """dataset/transforms.py:90-108 (get_final_transforms) on seeded uint8 frames: crop + ToTensor + Normalize(0.5, 0.5)."""
    from PIL import Image
    R.install_shims()
    sys.path.insert(0, R.REF_ROOT)
    from dataset.transforms import TransformsGenerator
    rng = np.random.RandomState(7)
    frames = rng.randint(0, 256, size=(3, 60, 90, 3), dtype=np.uint8)
    crop = [10, 5, 74, 53]                                  # left, upper, right, lower -> 64 x 48
    cfg = {"data": {"crop": crop}, "model": {"representation_network": {"target_input_size": [64, 48]}}}
    tf = TransformsGenerator.get_final_transforms(cfg)["train"]
    out = np.stack([tf(Image.fromarray(f)).numpy() for f in frames])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "input_pipeline.npz")
    np.savez_compressed(path, frames=frames, crop=np.array(crop), out=out)
    print("input pipeline golden", out.shape, float(out.min()), float(out.max()))


def input_resize_golden():
    """dataset/transforms.py:15-32 + 90-108 when the cropped frame does NOT have the target size: PIL crop, PIL
    ``resize(BILINEAR)`` (antialiased, 8-bit fixed point), ToTensor, Normalize -> tests/golden/input_resize.npz."""
    from PIL import Image
    R.install_shims()
    sys.path.insert(0, R.REF_ROOT)
    from dataset.transforms import TransformsGenerator
    rng = np.random.RandomState(11)
    out = {}
    cases = {"down": ((2, 60, 90, 3), [10, 5, 74, 53], [32, 20]),        # 64 x 48 crop -> 32 x 20 (w, h): both axes shrink
             "up": ((2, 30, 40, 3), None, [64, 48]),                      # no crop, both axes grow
             "mixed": ((2, 50, 70, 3), [3, 1, 67, 49], [48, 48]),         # 64 x 48 crop -> 48 x 48: width shrinks, height kept
             "bair": ((1, 64, 64, 3), None, [256, 256])}                  # BAIR's native 64 x 64 frames to the 256 x 256 model input
    for name, (shape, crop, size) in cases.items():
        frames = rng.randint(0, 256, size=shape, dtype=np.uint8)
        cfg = {"data": {"crop": crop}, "model": {"representation_network": {"target_input_size": size}}}
        tf = TransformsGenerator.get_final_transforms(cfg)["train"]
        res = np.stack([tf(Image.fromarray(f)).numpy() for f in frames])
        rs = TransformsGenerator.check_and_resize(crop, size)
        out[f"{name}.frames"] = frames
        out[f"{name}.crop"] = np.array(crop if crop is not None else [-1, -1, -1, -1])
        out[f"{name}.size"] = np.array(size)
        out[f"{name}.u8"] = np.stack([np.asarray(rs(Image.fromarray(f))) for f in frames])
        out[f"{name}.out"] = res
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "input_resize.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__" and "--input-pipeline" in sys.argv:
    input_pipeline_golden()
    input_resize_golden()
