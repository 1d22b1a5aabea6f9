"""Zero123 denoiser / first-stage encoder against the reference's OWN classes.

tests/golden/zero123.npz + zero123_keys.json were produced by executing ``UNetModel`` / ``Encoder`` /
``DiagonalGaussianDistribution`` from /root/reference (tests/golden/make_zero123_golden.py).  CPU, fp32."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import zero123 as Z
from tests import helpers as Hh

GOLD = Path(__file__).resolve().parent / "golden"
TOL = 2e-5          # fp32 CPU, different but equivalent operation order (fused attention, matmul-as-1x1-conv)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "zero123.npz"), json.loads((GOLD / "zero123_keys.json").read_text())


def keys_of(m):
    return [[k, list(v.shape)] for k, v in m.state_dict().items()]


SMALL_UNET = Z.UNetConfig(in_channels=8, out_channels=4, model_channels=32, attention_resolutions=(2, 1), num_res_blocks=1,
                          channel_mult=(1, 2, 2), num_heads=4, context_dim=24)
SMALL_ENC = Z.EncoderConfig(in_channels=3, ch=32, ch_mult=(1, 2, 2), num_res_blocks=1, z_channels=4)


def test_state_dict_names_and_shapes_match_the_reference_at_full_size(gold):
    """The full YAML configuration (860M-parameter UNet, 34M-parameter encoder): same parameter names, order and
    shapes as the reference modules => a Zero123 checkpoint loads with strict=True."""
    _, keys = gold
    with torch.device("meta"):
        unet, enc = Z.Zero123UNet(), Z.Zero123Encoder()
    assert keys_of(unet) == keys["unet_full"]
    assert keys_of(enc) == keys["enc_full"]
    assert sum(int(np.prod(s)) for _, s in keys["unet_full"]) == 859_532_484


def test_latent_diffusion_prefixes():
    with torch.device("meta"):
        m = Z.Zero123Model()
    names = list(m.state_dict().keys())
    assert all(n.startswith(("model.diffusion_model.", "first_stage_model.encoder.", "first_stage_model.quant_conv.",
                             "cc_projection.")) for n in names)
    assert m.cc_projection.weight.shape == (768, 772)


def test_unet_reproduces_reference_unetmodel(gold):
    g, keys = gold
    unet = Z.Zero123UNet(SMALL_UNET).eval()
    assert keys_of(unet) == keys["unet_small"]
    Hh.seeded_fill(unet, 11)
    x, t = torch.from_numpy(g["unet_x"]), torch.from_numpy(g["unet_t"])
    with torch.no_grad():
        y1 = unet(x, t, torch.from_numpy(g["unet_ctx1"]))       # single-token shortcut
        y3 = unet(x, t, torch.from_numpy(g["unet_ctx3"]))       # general cross-attention
    assert Hh.rel_linf(y1.numpy(), g["unet_y1"]) <= TOL
    assert Hh.rel_linf(y3.numpy(), g["unet_y3"]) <= TOL


def test_encoder_and_posterior_reproduce_reference(gold):
    g, keys = gold
    enc = Z.Zero123Encoder(SMALL_ENC).eval()
    assert keys_of(enc) == keys["enc_small"]
    Hh.seeded_fill(enc, 12)
    with torch.no_grad():
        h = enc(torch.from_numpy(g["enc_x"]))
    assert Hh.rel_linf(h.numpy(), g["enc_y"]) <= TOL
    post = Z.DiagonalGaussian(h)
    assert Hh.rel_linf(post.mean.numpy(), g["post_mean"]) <= TOL
    assert Hh.rel_linf(post.std.numpy(), g["post_std"]) <= TOL
    torch.manual_seed(77)                                        # same global-generator draw as the reference's sample()
    assert Hh.rel_linf(post.sample().numpy(), g["post_sample_seed77"]) <= TOL


def test_flop_counters_match_a_hook_count():
    """unet_flops / encoder_flops (used for the tensor-pipe roofline in bench.py) against torch's own operator-level FLOP
    counter (convolutions, matmuls, attention products of the modules as executed)."""
    from torch.utils.flop_counter import FlopCounterMode
    unet = Z.Zero123UNet(SMALL_UNET).eval()
    n, h, w = 2, 8, 8
    with FlopCounterMode(display=False) as fc, torch.no_grad():
        unet(torch.zeros(n, 8, h, w), torch.zeros(n, dtype=torch.long), torch.zeros(n, 1, 24))
    est, counted = Z.unet_flops(SMALL_UNET, n, h, w), float(fc.get_total_flops())
    assert 0.94 * est <= counted <= 1.06 * est, (est, counted)
    enc = Z.Zero123Encoder(SMALL_ENC).eval()
    with FlopCounterMode(display=False) as fc, torch.no_grad():
        enc(torch.zeros(n, 3, 32, 32))
    est, counted = Z.encoder_flops(SMALL_ENC, n, 32, 32), float(fc.get_total_flops())
    assert 0.94 * est <= counted <= 1.06 * est, (est, counted)
