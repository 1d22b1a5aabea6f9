#!/usr/bin/env python
"""Generates tests/golden/zero123.npz + zero123_keys.json by executing the reference's OWN network classes
(read from /root/reference at generation time only, unmodified):

  * ``UNetModel``  extern/ldm_zero123/modules/diffusionmodules/openaimodel.py:429-842
  * ``Encoder``    extern/ldm_zero123/modules/diffusionmodules/model.py:380-495
  * ``DiagonalGaussianDistribution``  extern/ldm_zero123/modules/distributions/distributions.py:24-69

The package ``extern.ldm_zero123`` is imported from /root/reference with its heavy, absent dependencies stubbed:
``extern.ldm_zero123.util`` (imports matplotlib / torchvision / PIL only for logging helpers; the networks use
``exists`` / ``default`` / ``instantiate_from_config`` from it) and ``omegaconf.listconfig.ListConfig`` (a type check
in UNetModel.__init__).  Nothing of the reference's arithmetic is replaced.

Two fixtures:
  1. REDUCED width (a constructor argument, not a source change): every parameter randomised from a seeded generator
     in state_dict order, one forward each -> inputs / outputs.  The test fills dreammesh4d_b200.zero123 modules the same
     way (same names, same order, same shapes = strict state_dict compatibility) and compares outputs.
  2. FULL YAML configuration (load/zero123/sd-objaverse-finetune-c_concat-256.yaml:28-60) instantiated on the meta
     device: the (name, shape) list of its state_dict -> zero123_keys.json; the test requires the full-size
     dreammesh4d_b200.zero123 model to expose exactly that list (a real Zero123 checkpoint then loads strict=True).
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference")


def import_reference():
    util = types.ModuleType("extern.ldm_zero123.util")
    util.exists = lambda v: v is not None
    util.default = lambda v, d: v if v is not None else (d() if callable(d) else d)

    def instantiate_from_config(config):
        raise NotImplementedError("not needed by the network classes")
    util.instantiate_from_config = instantiate_from_config
    oc = types.ModuleType("omegaconf")
    lc = types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = type("ListConfig", (list,), {})
    oc.listconfig = lc
    sys.modules.setdefault("omegaconf", oc)
    sys.modules.setdefault("omegaconf.listconfig", lc)
    sys.path.insert(0, str(REF))
    import extern.ldm_zero123  # noqa: F401  (namespace package from the reference tree)
    sys.modules["extern.ldm_zero123.util"] = util
    from extern.ldm_zero123.modules.diffusionmodules.model import Encoder
    from extern.ldm_zero123.modules.diffusionmodules.openaimodel import UNetModel
    from extern.ldm_zero123.modules.distributions.distributions import DiagonalGaussianDistribution
    return UNetModel, Encoder, DiagonalGaussianDistribution


sys.path.insert(0, str(OUT.parents[1]))
from tests.helpers import seeded_fill as fill  # noqa: E402  (the test fills the product modules the same way)


UNET_SMALL = dict(image_size=8, in_channels=8, out_channels=4, model_channels=32, attention_resolutions=[2, 1],
                  num_res_blocks=1, channel_mult=[1, 2, 2], num_heads=4, use_spatial_transformer=True,
                  transformer_depth=1, context_dim=24, use_checkpoint=False, legacy=False, use_fp16=False)
ENC_SMALL = dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
                 num_res_blocks=1, attn_resolutions=[], dropout=0.0)
UNET_FULL = dict(image_size=32, in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                 num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                 transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False, use_fp16=True)
ENC_FULL = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def main():
    UNetModel, Encoder, DGD = import_reference()
    torch.manual_seed(0)
    blob = {}
    # ---- reduced-width numerics ----
    unet = UNetModel(**UNET_SMALL).eval()
    fill(unet, 11)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 8, 8, 8, generator=g)
    t = torch.tensor([3, 500, 977])
    ctx1 = torch.randn(3, 1, 24, generator=g)         # Zero123: one context token
    ctx3 = torch.randn(3, 3, 24, generator=g)         # general cross-attention path
    with torch.no_grad():
        blob["unet_x"], blob["unet_t"], blob["unet_ctx1"], blob["unet_ctx3"] = x, t, ctx1, ctx3
        blob["unet_y1"] = unet(x, t, context=ctx1)
        blob["unet_y3"] = unet(x, t, context=ctx3)
    enc = Encoder(**ENC_SMALL).eval()
    fill(enc, 12)
    img = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    with torch.no_grad():
        h = enc(img)
        blob["enc_x"], blob["enc_y"] = img, h
        post = DGD(h)
        blob["post_mean"], blob["post_std"] = post.mean, post.std
        torch.manual_seed(77)
        blob["post_sample_seed77"] = post.sample()
    keys = {"unet_small": [[k, list(v.shape)] for k, v in unet.state_dict().items()],
            "enc_small": [[k, list(v.shape)] for k, v in enc.state_dict().items()]}
    # ---- full-size key lists (meta device: no memory) ----
    with torch.device("meta"):
        keys["unet_full"] = [[k, list(v.shape)] for k, v in UNetModel(**UNET_FULL).state_dict().items()]
        keys["enc_full"] = [[k, list(v.shape)] for k, v in Encoder(**ENC_FULL).state_dict().items()]
    np.savez_compressed(OUT / "zero123.npz", **{k: v.detach().numpy() for k, v in blob.items()})
    (OUT / "zero123_keys.json").write_text(json.dumps(keys))
    print("wrote zero123.npz and zero123_keys.json:", {k: len(v) for k, v in keys.items()},
          "params full unet", sum(int(np.prod(s)) for _, s in keys["unet_full"]))


if __name__ == "__main__":
    main()
