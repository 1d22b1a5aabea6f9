"""GPU: the channels-last GroupNorm (+ channel bias, + SiLU) kernels against torch's own ops, and the Zero123
networks running on them against the golden vectors produced by the reference's UNetModel / Encoder."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dreammesh4d_b200 import zero123 as Z
from dreammesh4d_b200.nhwc import GroupNormAct, groupnorm_nhwc
from tests import helpers as Hh
from tests.test_zero123 import SMALL_ENC, SMALL_UNET

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = Path(__file__).resolve().parent / "golden"


def torch_ref(x, gamma, beta, G, eps, silu, cb):
    if cb is not None:
        x = x + cb[:, :, None, None]
    y = F.group_norm(x, G, gamma, beta, eps)
    return F.silu(y) if silu else y


@pytest.mark.parametrize("shape,G", [((2, 320, 32, 32), 32), ((3, 128, 64, 48), 32), ((2, 1920, 8, 8), 32), ((1, 2560, 4, 4), 32),
                                     ((2, 64, 7, 5), 32), ((1, 128, 256, 256), 32)])
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("bias", [False, True])
def test_groupnorm_nhwc_fp32_matches_torch(shape, G, silu, bias):
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 1.7 + 0.4).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gamma, beta = (1 + 0.3 * torch.randn(shape[1], generator=g)).to(DEV), (0.2 * torch.randn(shape[1], generator=g)).to(DEV)
    cb = torch.randn(shape[0], shape[1], generator=g).to(DEV) if bias else None
    gy = torch.randn(shape, generator=g).to(DEV)
    y = groupnorm_nhwc(x, gamma, beta, G, 1e-5, silu, cb)
    assert y.is_contiguous(memory_format=torch.channels_last)
    (dx,) = torch.autograd.grad(y, x, gy)
    xr = x.detach().double().requires_grad_(True)
    yr = torch_ref(xr, gamma.double(), beta.double(), G, 1e-5, silu, None if cb is None else cb.double())
    (dxr,) = torch.autograd.grad(yr, xr, gy.double())
    assert Hh.rel_linf(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) <= 2e-5
    assert Hh.rel_linf(dx.cpu().numpy(), dxr.cpu().numpy()) <= 1e-4


def test_groupnorm_nhwc_fp16_is_as_accurate_as_torch_fp16():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 640, 16, 16, generator=g).to(DEV).half().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    m = GroupNormAct(32, 640, silu=True).to(DEV).half()
    Hh.seeded_fill(m, 3)
    for p in m.parameters():
        p.requires_grad_(False)
    cb = torch.randn(4, 640, generator=g).to(DEV)
    gy = torch.randn(x.shape, generator=g).to(DEV).half()
    y = m(x, chan_bias=cb)
    (dx,) = torch.autograd.grad(y, x, gy)
    xr = x.detach().double().requires_grad_(True)
    yr = torch_ref(xr, m.weight.double(), m.bias.double(), 32, 1e-5, True, cb.double())
    (dxr,) = torch.autograd.grad(yr, xr, gy.double())
    yt = F.silu(F.group_norm(x.detach() + cb.half()[:, :, None, None], 32, m.weight, m.bias, 1e-5))     # torch's own fp16 path
    ours = Hh.rel_linf(y.detach().float().cpu().numpy(), yr.detach().cpu().numpy())
    theirs = Hh.rel_linf(yt.float().cpu().numpy(), yr.detach().cpu().numpy())
    assert ours <= max(2e-3, 1.5 * theirs)
    assert Hh.rel_linf(dx.float().cpu().numpy(), dxr.cpu().numpy()) <= 4e-3


def test_zero123_networks_on_the_fused_kernels_reproduce_the_reference():
    """Same golden vectors as tests/test_zero123.py (reference UNetModel / Encoder outputs), now on CUDA where GroupNorm,
    SiLU and the time-embedding add run in the fused channels-last kernel."""
    g = np.load(GOLD / "zero123.npz")
    torch.backends.cudnn.allow_tf32 = False          # compare in true fp32 (TF32 convolutions alone cost ~5e-4)
    torch.backends.cuda.matmul.allow_tf32 = False
    unet = Z.Zero123UNet(SMALL_UNET).eval()
    Hh.seeded_fill(unet, 11)
    unet = unet.to(DEV).to(memory_format=torch.channels_last)
    enc = Z.Zero123Encoder(SMALL_ENC).eval()
    Hh.seeded_fill(enc, 12)
    enc = enc.to(DEV).to(memory_format=torch.channels_last)
    for p in list(unet.parameters()) + list(enc.parameters()):
        p.requires_grad_(False)
    d = lambda k: torch.from_numpy(g[k]).to(DEV)
    from dreammesh4d_b200 import _lib
    _lib.profile_enable(True)
    _lib.profile_collect()
    with torch.no_grad():
        y1 = unet(d("unet_x"), d("unet_t"), d("unet_ctx1"))
        y3 = unet(d("unet_x"), d("unet_t"), d("unet_ctx3"))
        h = enc(d("enc_x"))
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    assert prof.get("groupnorm_nhwc_forward_kernels", (0, 0))[1] > 20          # the fused path really ran
    assert Hh.rel_linf(y1.cpu().numpy(), g["unet_y1"]) <= 1e-4
    assert Hh.rel_linf(y3.cpu().numpy(), g["unet_y3"]) <= 1e-4
    assert Hh.rel_linf(h.cpu().numpy(), g["enc_y"]) <= 1e-4
    # gradient to the image through the encoder (the SDS path): fused backward vs torch ops on the CPU
    img = d("enc_x").clone().requires_grad_(True)
    gl = torch.randn(h.shape, generator=torch.Generator().manual_seed(0)).to(DEV)
    (gi,) = torch.autograd.grad(enc(img), img, gl)
    enc_cpu = Z.Zero123Encoder(SMALL_ENC).eval()
    Hh.seeded_fill(enc_cpu, 12)
    ic = torch.from_numpy(g["enc_x"]).clone().requires_grad_(True)
    (gc,) = torch.autograd.grad(enc_cpu(ic), ic, gl.cpu())
    assert Hh.rel_linf(gi.cpu().numpy(), gc.numpy()) <= 2e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize("with_res", [False, True])
def test_bias_residual_add_matches_torch(dtype, tol, with_res):
    from dreammesh4d_b200.nhwc import bias_residual_add
    g = torch.Generator().manual_seed(11)
    shape = (3, 132, 17, 9)
    h = torch.randn(shape, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    res = torch.randn(shape, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True) if with_res else None
    bias = torch.randn(shape[1], generator=g).to(DEV, dtype)
    out = bias_residual_add(h, bias, res)
    ref = h.float() + bias.float()[None, :, None, None] + (res.float() if with_res else 0.0)
    assert out.dtype == dtype and Hh.rel_linf(out.detach().float().cpu(), ref.detach().cpu()) <= tol
    gy = torch.randn(shape, generator=g).to(DEV, dtype)
    grads = torch.autograd.grad(out, [h] + ([res] if with_res else []), gy)
    for gr in grads:
        assert torch.equal(gr, gy)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3)])
def test_geglu_matches_torch(dtype, tol):
    from dreammesh4d_b200.nhwc import geglu
    g = torch.Generator().manual_seed(12)
    proj = (2.0 * torch.randn(2, 37, 2 * 320, generator=g)).to(DEV, dtype).requires_grad_(True)
    out = geglu(proj)
    a, gate = proj.float().chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    assert out.shape == (2, 37, 320) and Hh.rel_linf(out.detach().float().cpu(), ref.detach().cpu()) <= tol
    gy = torch.randn(out.shape, generator=g).to(DEV, dtype)
    (gp,) = torch.autograd.grad(out, proj, gy)
    (gr,) = torch.autograd.grad(ref, proj, gy.float())
    assert Hh.rel_linf(gp.float().cpu(), gr.float().cpu()) <= 5 * tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize("mode", ["none", "full", "broadcast"])
@pytest.mark.parametrize("C", [320, 1280, 36])
def test_add_layernorm_matches_torch(dtype, tol, mode, C):
    from dreammesh4d_b200.nhwc import add_layernorm
    g = torch.Generator().manual_seed(C)
    B, N = 3, 50
    ln = torch.nn.LayerNorm(C).to(DEV, dtype)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.3 * torch.randn(C, generator=g)); ln.bias.copy_(0.2 * torch.randn(C, generator=g))
    x = (1.5 * torch.randn(B, N, C, generator=g)).to(DEV, dtype)
    delta = None if mode == "none" else torch.randn(B, N if mode == "full" else 1, C, generator=g).to(DEV, dtype)
    if mode == "broadcast":
        delta = delta.expand(B, N, C)
    with torch.no_grad():
        x_new, y = add_layernorm(x, delta, ln)
    with torch.no_grad():
        x_ref = x if delta is None else x + delta
        y_ref = F.layer_norm(x_ref.float(), (C,), ln.weight.float(), ln.bias.float(), ln.eps)
    assert torch.equal(x_new, x_ref)
    assert y.dtype == dtype and Hh.rel_linf(y.float().cpu(), y_ref.cpu()) <= tol
    # with gradients recorded the torch path is taken and differentiable
    xg = x.clone().requires_grad_(True)
    _, y2 = add_layernorm(xg, delta, ln)
    y2.float().sum().backward()
    assert xg.grad is not None
