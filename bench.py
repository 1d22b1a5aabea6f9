#!/usr/bin/env python
"""bench.py — the DreamMesh4D dynamic-stage hot path on B200.

Metric (BASELINE.json): rasterize fwd+bwd Gaussians/s @512x512, 8 views (+ train-step ms as context)
    value = P * n_views * n_gpus / t(fwd+bwd)          [Gaussians/s, whole job]
Default workload (--config c3) at every N: BASELINE config C3 geometry — 100k-face sphere, 300k surface-bound
Gaussians, 8 views per rank per step, each view at its own timestamp (8 attribute sets), 512x512, white bg, one
3-channel pass with colour + depth + alpha + means2D gradients.  Weak scaling (every rank renders its
own 8 cameras, the reference's DDP semantics); for N>1 the gradients of the shared attributes are summed with one NCCL
all-reduce per step, captured in the step's CUDA graph.  For N>1 the same line also carries the STRONG-scaling form
of C3 (the 8 views sharded over the ranks) under "strong_scaling".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c3|c2|c4|c5]
N>1 is launched by torchrun (one process per GPU).  Prints ONE JSON line on rank 0.
Other configurations (not driver lines; results under profiles/):
  --config c2   train step at BASELINE config C2 (50k faces / 150k Gaussians, 4 views, SDS)
  --config c4   rasterizer microbench: 1M free Gaussians, 1024x1024, 16 cameras
  --config c5   skinning microbench: 200k vertices / 512 control nodes / 600k Gaussians, GB/s vs roofline
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 for N>1 launches; the CPU arm is meant to use every host core.  Set before any
    # OpenMP runtime (torch's, the oracle's) initialises.
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np
import torch

WORKLOAD = "C3: 100k-face sphere, 300k surface-bound Gaussians, 8 views x 8 timestamps per rank, 512x512, 1 pass (3ch) fwd+bwd"
METRIC = "rasterize fwd+bwd Gaussians/s @512x512, 8 views/GPU"
N_FACES, G_PER_FACE, H, W, VIEWS = 100_000, 3, 512, 512, int(os.environ.get("DM4D_VIEWS", "8"))   # DM4D_VIEWS: tuning only
M_NODES, K_NBR = 1000, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--small", action="store_true", help="tiny workload for a functional check (not a bench number)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--kernels-only", action="store_true", help="tuning aid: print step time + per-kernel times and stop "
                    "(no e2e / cpu / train-step legs; not a bench line)")
    ap.add_argument("--no-train", action="store_true", help="skip the train-step / SDS context legs")
    return ap.parse_args()


def bench_config(P: int, small: bool = False) -> dict:
    """`config` of the JSON line — identical for the GPU arm and the reference (CPU) arm."""
    return {"workload": WORKLOAD if not small else "SMALL functional check (not a bench number)",
            "P": P, "H": H, "W": W, "views_per_gpu": VIEWS, "channels": 3,
            "gradients": "colour + depth + alpha images -> means3D, means2D, scales, rotations, opacities, colours",
            "l2": "GPU arm: flushed between steps (256 MiB write, outside the per-step events); CPU arm: n/a"}


# ------------------------------------------------------------------------------------------------
# workload (synthetic)
# ------------------------------------------------------------------------------------------------
def build_scene(small: bool, n_faces: int = N_FACES, views: int = VIEWS, m_nodes: int = M_NODES):
    from dreammesh4d_b200 import synthetic
    n_faces = 2_000 if small else n_faces
    scene = synthetic.make_sugar_scene(n_faces, g=G_PER_FACE)
    graph = synthetic.make_deform_graph(scene.verts, 64 if small else m_nodes, K_NBR, seed=0)
    node = synthetic.random_node_attrs(views, graph.node_xyz.shape[0], seed=1)
    return scene, graph, node


def build_cameras(rank: int, views: int = VIEWS):
    from dreammesh4d_b200 import synthetic
    from dreammesh4d_b200.camera import get_cam_info_gaussian
    c2w, fovy = synthetic.random_orbit_cameras(views, seed=2 + rank)
    return get_cam_info_gaussian(c2w, fovy, fovy)


def cams_c2w_fovy(rank: int, views: int = VIEWS):
    from dreammesh4d_b200 import synthetic
    return synthetic.random_orbit_cameras(views, seed=2 + rank), synthetic.random_orbit_cameras(views, seed=102 + rank)


def time_events(run, steps: int, before=None):
    """Per-step CUDA-event times (ms) on the current stream; `before` runs outside the events."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if before is not None:
            before()
        ev[i][0].record()
        run(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def stats(ms):
    s = sorted(ms)
    return {"ms_median": s[len(s) // 2], "ms_mean": sum(s) / len(s), "ms_max": s[-1], "ms_min": s[0], "steps": len(s)}


# ------------------------------------------------------------------------------------------------
# SDS / train-step context legs ("train-step ms" half of BASELINE's metric)
# ------------------------------------------------------------------------------------------------
_ZERO123 = {}


def zero123_model(dev):
    """One random-weight Zero123 (UNet 860M + first-stage encoder, fp16 like the reference's half_precision_weights)
    per process; weights are unavailable offline, shapes / FLOPs / traffic are the real model's."""
    if dev not in _ZERO123:
        from dreammesh4d_b200 import zero123
        torch.backends.cudnn.benchmark = True
        _ZERO123[dev] = zero123.build_random(device=dev, dtype=torch.float16, seed=0)
    return _ZERO123[dev]


def sds_tensor_leg(dev, views: int, steps: int = 10):
    """Tensor-pipe context of row A9: the UNet evaluation (batch 2 x views, fp16, no grad) and the first-stage encoder
    forward + backward-to-the-image (batch views, 256x256), each as a CUDA-graph replay timed with CUDA events;
    TFLOP/s against MEASURED_PEAKS.json:bf16_tflops_sustained."""
    from dreammesh4d_b200 import zero123
    m = zero123_model(dev)
    cfg = m.cfg
    n = 2 * views
    x = torch.randn(n, 4, 32, 32, device=dev, dtype=torch.float16)
    t = torch.randint(20, 500, (n,), device=dev)
    cond = {"c_concat": [torch.randn(n, 4, 32, 32, device=dev, dtype=torch.float16)],
            "c_crossattn": [torch.randn(n, 1, 768, device=dev, dtype=torch.float16)]}
    img = torch.rand(views, 3, 256, 256, device=dev, requires_grad=True)
    g_lat = torch.randn(views, 4, 32, 32, device=dev)

    def unet():
        with torch.no_grad():
            return m.apply_model(x, t, cond)

    def enc():
        lat = m.get_first_stage_encoding(m.encode_first_stage((img * 2 - 1).half())).float()
        (gi,) = torch.autograd.grad(lat, img, g_lat)
        return gi

    out = {}
    for name, fn, flops in (("unet_fwd", unet, zero123.unet_flops(cfg.unet, n, 32, 32)),
                            ("encoder_fwd_bwd", enc, 2.0 * zero123.encoder_flops(cfg.encoder, views, 256, 256))):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        ms = stats(time_events(lambda i: g.replay(), steps))["ms_median"]
        out[name] = {"ms": ms, "tflops": flops / (ms * 1e-3) / 1e12, "flops": flops}
    peak = measured_peaks()[2]
    for v in out.values():
        v["frac_of_sustained_bf16_peak"] = round(v["tflops"] / peak, 4)
    out["peak_tflops"] = peak
    out["what"] = (f"Zero123 UNet forward (batch {n}, 8x32x32 latents, fp16, 860M parameters, random weights of the YAML's shapes) and "
                   f"first-stage encoder forward + backward to the image (batch {views}, 256x256, fp16): PyTorch tensor-core "
                   "matmuls / cuDNN convolutions, one CUDA-graph replay each, CUDA events; FLOPs = 2 x MACs of the executed modules "
                   "(zero123.unet_flops / encoder_flops; backward counted as 1 x forward: weights are frozen)")
    return out


def train_step_ms(dev, dist, steps: int, n_faces: int, views: int, n_frames: int, label: str, small: bool = False):
    """One dynamic-stage optimizer step on the hot path INCLUDING the Zero123 SDS term (sugar_4dgen.py:397-429):
    substep "zero123": `views` random cameras -> SDS loss (x 0.1, sugar_dynamic_dg.yaml:136) ; substep "ref": the same
    frames from the reference camera -> MSE rgb (x 5000) + mask (x 500) + ARAP (x 10) + mesh normal consistency (x 100).
    HexPlane/MLP deformation -> fused skinning -> 6-channel batched rasterizer -> fused post-ops -> losses -> backward
    -> gradient exchange -> fused Adam."""
    from dreammesh4d_b200 import synthetic
    from dreammesh4d_b200.arap import ARAPEnergy, face_pairs, mesh_normal_consistency
    from dreammesh4d_b200.deformation import HexPlaneDeformation
    from dreammesh4d_b200.geometry import DynamicSuGaRGeometry
    from dreammesh4d_b200.renderer import DiffGaussianBatchRenderer
    from dreammesh4d_b200.sds import TemporalStableZero123SDS
    from dreammesh4d_b200.trainstep import DynamicStageStep, GraphedDynamicStageStep
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    scene, graph, _ = build_scene(small, n_faces, views)
    P = scene.n_gaussians
    cams = cams_c2w_fovy(rank, views)
    model = zero123_model(dev)
    gen = torch.Generator().manual_seed(7)
    guidance = TemporalStableZero123SDS(model, torch.randn(n_frames, 1, 768, generator=gen).to(dev),
                                        torch.randn(n_frames, 4, 32, 32, generator=gen).to(dev), guidance_scale=3.0,
                                        min_step_percent=0.02, max_step_percent=0.5)     # sugar_dynamic_dg.yaml:115-117
    guidance.update_step(0, 0)

    def build(exchange):
        torch.manual_seed(0)
        net = HexPlaneDeformation().to(dev)
        with torch.no_grad():
            for head, sdev in ((net.deformation_net.pos_deform, 0.01), (net.deformation_net.rotations_deform, 0.05),
                               (net.deformation_net.scales_deform, 0.01), (net.deformation_net.opacity_deform, 0.3)):
                head.feature_out[1].weight.normal_(0, sdev)
        geo = DynamicSuGaRGeometry(scene, graph, net).to(dev)
        ren = DiffGaussianBatchRenderer(geo)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), eps=1e-15, capturable=True, fused=True)
        arap = ARAPEnergy(geo._points.detach(), geo._surface_mesh_faces)
        pairs = face_pairs(geo._surface_mesh_faces)

        def loss_fn(out, batch):       # the loss terms that are live in sugar_dynamic_dg.yaml:135-158
            if batch["guidance"] == "zero123":
                return 0.1 * guidance(out["comp_rgb"], batch["elevation"], batch["azimuth"], batch["camera_distances"],
                                      batch["frame_indices"])["loss_sds"]
            timed = geo._timed
            return 5000.0 * torch.nn.functional.mse_loss(out["comp_rgb"], batch["rgb"]) + \
                500.0 * torch.nn.functional.mse_loss(out["comp_mask"], batch["mask"]) + \
                10.0 * arap(timed["verts"], timed["vert_rot"]).sum() + 100.0 * mesh_normal_consistency(timed["verts"], pairs)
        return ren, DynamicStageStep(geo, ren, opt, loss_fn, exchange=exchange)

    hh, ww = (128, 128) if small else (H, W)
    frames = torch.randperm(n_frames, generator=gen)[:views].sort()[0]
    ts = torch.linspace(0, 1, n_frames + 2)[1:-1][frames]
    batches = []
    for kind, (c2w, fovy) in zip(("zero123", "ref"), cams):
        focal = 0.5 * hh / torch.tan(0.5 * fovy)
        j, i = torch.meshgrid(torch.arange(hh, dtype=torch.float32) + 0.5, torch.arange(ww, dtype=torch.float32) + 0.5, indexing="ij")
        dirs = torch.stack([(i[None] - ww / 2) / focal[:, None, None], -(j[None] - hh / 2) / focal[:, None, None],
                            -torch.ones(len(fovy), hh, ww)], dim=-1)
        rays_d = (dirs[..., None, :] * c2w[:, None, None, :3, :3]).sum(-1)
        rays_o = c2w[:, None, None, :3, 3].expand_as(rays_d)
        b = {"guidance": kind, "c2w": c2w.to(dev), "fovy": fovy.to(dev), "height": hh, "width": ww, "timestamp": ts.to(dev),
             "frame_indices": frames.to(dev), "rays_o": rays_o.contiguous().to(dev), "rays_d": rays_d.contiguous().to(dev),
             "elevation": (torch.rand(views, generator=gen) * 90 - 10).to(dev), "azimuth": (torch.rand(views, generator=gen) * 360 - 180).to(dev),
             "camera_distances": torch.full((views,), 3.8, device=dev)}
        if kind == "ref":
            b.update({"rgb": torch.rand(views, hh, ww, 3, device=dev), "mask": (torch.rand(views, hh, ww, 1, device=dev) > 0.5).float()})
        batches.append(b)

    res = {"label": label, "P": P, "views_per_substep": views, "n_frames": n_frames, "substeps": 2, "includes_sds": True,
           "what": (f"optimizer step of the dynamic stage on the hot path at {label}: substep zero123 = {views} random views -> Zero123 "
                    "SDS (encoder fwd+bwd at 256^2, UNet fwd at batch 2x views, fp16, random weights of the YAML's shapes); substep ref = "
                    "the same frames from the reference camera -> MSE rgb+mask + ARAP + normal consistency; HexPlane lookup (fused "
                    "kernel) + MLP heads (PyTorch) -> fused skinning -> 6-channel rasterizer -> fused post-ops -> backward -> gradient "
                    "exchange -> fused Adam; CUDA events per step")}
    modes = ["dense"] if world == 1 else ["dense", "node_gather"]
    best = None
    for exchange in modes:
        ren, stepper = build(exchange)
        stepper(batches, 0)                                   # sizes the binning workspace with one read-back
        ren.capacity = int(ren.last_state.status()[0] * 1.3) + 4096
        for i in range(3):
            stepper(batches, i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        eager = stats(time_events(lambda i: stepper(batches, i), steps))
        entry = {"eager": eager}
        if not os.environ.get("DM4D_BENCH_NO_TRAIN_GRAPH"):
            try:
                graphed = GraphedDynamicStageStep(stepper, batches)
                for _ in range(2):
                    graphed(batches)
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                gs = stats(time_events(lambda i: graphed(batches), steps))
                if bool(torch.isfinite(graphed.loss)):
                    entry["graph"] = gs
                else:
                    entry["graph_error"] = "non-finite loss"
                stepper.check_overflow()
            except Exception as e:      # keep the eager number, say why
                entry["graph_error"] = f"{type(e).__name__}: {e}"[:300]
        res[f"exchange_{exchange}"] = entry
        t = entry.get("graph", eager)["ms_median"]
        if best is None or t < best[0]:
            best = (t, exchange, "one CUDA-graph replay per optimizer step" if "graph" in entry else "eager")
    res.update({"ms_median": best[0], "exchange": best[1] if world > 1 else "none (single GPU)", "launch": best[2]})
    return res


def ref_equiv_ms(dev_in, vp, gC, gD, gA, P, steps: int):
    """Same 8 views, same inputs and image gradients through bench_ref_equiv (per-view launch sequence with the
    num_rendered read-back, global radix sort, block-wide walks, per-pixel atomics): a same-GPU reference point for the
    classic rasterizer structure.  Our transcription — not the upstream code."""
    from bench_ref_equiv.ref_equiv import RefEquivView
    view = RefEquivView(P, H, W, vp.device)
    means, rots = dev_in["means"].detach(), dev_in["rots"].detach()
    scales, opac, cols = dev_in["scales"].detach(), dev_in["opac"].detach(), dev_in["cols"].detach()
    vps = vp.clone()
    vps[:, 38] = 0.0                    # every per-view call sees a single attribute set
    rows = [vps[v].contiguous() for v in range(VIEWS)]

    def step(_=0):
        n = 0
        for v in range(VIEWS):
            n += view.forward_backward(means[v], scales, rots[v], opac, cols, rows[v], gC[v], gD[v], gA[v])
        return n

    for _ in range(2):
        n_r = step()
    torch.cuda.synchronize()
    ms = stats(time_events(step, steps))["ms_mean"]
    return {"ms_per_step": ms, "value": P * VIEWS / (ms * 1e-3), "unit": "Gaussians/s", "num_rendered": n_r, "steps": steps,
            "what": "OUR transcription of the upstream rasterizer structure (bench_ref_equiv/: per-view launches + num_rendered "
                    "read-back, one global CUB radix sort, 256-instance block-wide walks without sub-tile culling, ten global "
                    "atomics per (pixel, instance) in the backward; projection math shared with the product), same GPU, same "
                    "inputs, CUDA events; not the upstream code"}


def dropin_ms(dev_in, normals, cams, gC, gD, gA, dev, steps: int):
    """The ZERO-EDIT path: exactly what the reference's renderer does per step through the drop-in module
    (diff_sugar_rasterizer_temporal.py:144,169-178,202-211): one GaussianRasterizer per view, called twice (RGB with the
    means2D gradient holder, then normals as colours), 8 views, then the backward.  Eager, per-call, CUDA events."""
    from dreammesh4d_b200 import rasterizer as R
    V, PV, campos, tanx, tany = cams
    Vd, PVd, cd = V.to(dev), PV.to(dev), campos.to(dev)
    bg = torch.ones(3, device=dev)
    per_view = [[dev_in["means"][v].detach().clone().requires_grad_(True), dev_in["rots"][v].detach().clone().requires_grad_(True),
                 normals[v].detach().clone().requires_grad_(True)] for v in range(VIEWS)]
    shared = [dev_in[k].detach().clone().requires_grad_(True) for k in ("scales", "opac", "cols")]
    gN = gC.flip(1).contiguous()

    def step(normal_grads):
        outs, grads = [], []
        for v in range(VIEWS):
            m, q, nrm = per_view[v]
            s = R.GaussianRasterizationSettings(H, W, float(tanx[v]), float(tany[v]), bg, 1.0, Vd[v], PVd[v], 0, cd[v], False, False)
            rast = R.GaussianRasterizer(s)
            m2d = torch.zeros_like(m, requires_grad=True)
            c, radii, d, a = rast(means3D=m, means2D=m2d, opacities=shared[1], colors_precomp=shared[2], scales=shared[0], rotations=q)
            n, _, _, _ = rast(means3D=m, means2D=torch.zeros_like(m), opacities=shared[1], colors_precomp=nrm, scales=shared[0], rotations=q)
            outs += [c, d, a]
            grads += [gC[v], gD[v], gA[v]]
            if normal_grads:
                outs.append(n)
                grads.append(gN[v])
        torch.autograd.backward(outs, grads)
        for t in shared + [x for pv in per_view for x in pv]:
            t.grad = None

    res = {}
    for name, ng in (("rgb_bwd_normal_fwd_only", False), ("rgb_and_normal_bwd", True)):
        for _ in range(3):
            step(ng)
        torch.cuda.synchronize()
        res[name] = stats(time_events(lambda i: step(ng), steps))
    res["what"] = ("zero-edit drop-in path: per view one GaussianRasterizer, two calls (RGB + means2D holder, then normals; the second "
                   "call re-uses the first call's projection / binning / sort), 8 views, forward + backward; eager, per-call launches, "
                   "grow-only capacities (no num_rendered read-back after the first call); 'rgb_bwd_normal_fwd_only' = the gradient "
                   "flow of the shipped YAML (16 forwards, 8 backwards), 'rgb_and_normal_bwd' = both passes differentiated")
    return res


def gaussian_sets_gpu(scene, graph, node, dev):
    """Per-timestamp Gaussian sets produced by the product path (fused skinning kernels), on the GPU."""
    from dreammesh4d_b200 import skinning, synthetic
    d = lambda t: t.to(dev)
    faces = d(scene.faces.int())
    rest_q, _ = skinning.sugar_rest_frames(d(scene.verts), faces, d(scene.complex_rot), scene.g)
    with torch.no_grad():
        means, rots, normals, _, _ = skinning.skin_gaussians(*[d(t) for t in node], d(scene.verts), faces,
                                                             d(graph.nbr_idx.int()), d(graph.nbr_w), d(scene.bary), rest_q)
    scales = torch.cat([torch.full((scene.n_gaussians, 1), scene.thickness), scene.log_scales.exp()], dim=-1)
    return {"means3D": means, "rotations": rots, "normals": normals, "scales": d(scales),
            "opacities": d(torch.sigmoid(scene.densities)), "colors": d(scene.sh_dc[:, 0] * synthetic.C0 + 0.5)}


def gaussian_sets_oracle(scene, graph, node):
    """Same sets from the CPU oracle (used only by the CPU legs; never by the product path)."""
    from oracle import skin_oracle
    with torch.no_grad():
        return skin_oracle.deform_gaussians(scene, graph, *node)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peaks():
    """(HBM GB/s, source, sustained bf16 TFLOP/s)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)", float(j.get("bf16_tflops_sustained", 1400.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1400.0


# per-kernel ALGORITHMIC bytes per launch; R = instances in the launch, n = views*P, px = views*H*W
def algorithmic_bytes(kernel: str, n: int, R: int, px: int, tiles: int) -> float:
    return {
        "preprocess_kernel": 56.0 * n + 56.0 * n + 4.0 * R,          # read attributes; write record+radius+rect; count atomics
        "scan_tiles_kernel": 12.0 * tiles,
        "scatter_kernel": 8.0 * n + 8.0 * R + 4.0 * R,              # rect+depth; key write; cursor atomics
        "sort_pack_kernel": 8.0 * R + 48.0 * R + 48.0 * R,           # keys; record gather; stream write
        "render_forward_kernel": 48.0 * R + 24.0 * px,               # stream; colour(12)+depth+alpha+n_contrib
        "render_backward_kernel": 48.0 * R + 28.0 * px + 48.0 * n,   # stream; dL(20)+n_contrib+alpha; accumulator rows
        "preprocess_backward_kernel": 56.0 * n + 48.0 * n + 52.0 * n,   # attributes; accumulator rows; 40 B grads + 12 B means2D
    }.get(kernel, 0.0)


# SM issue peak: 4 schedulers x 1 warp instruction per clock per SM
def issue_peak_per_s(sm_mhz: float, sms: int = 148) -> float:
    return sms * 4 * sm_mhz * 1e6


def dbg(msg):
    if os.environ.get("DM4D_BENCH_DEBUG"):
        log(f"[rank {os.environ.get('RANK', '0')}] {msg}")


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def bind_to_gpu_numa_node(local_rank: int):
    """Multi-rank runs: pin this process to the CPUs of its GPU's NUMA node BEFORE any pinned host buffer is allocated, so
    the e2e leg's 235 MB/step of host traffic stays on the memory controller next to the GPU's PCIe root (first-touch
    placement).  Without it every rank's buffers land wherever torchrun happened to start the process and the host side,
    not PCIe, bounds the e2e number (2.9 -> 14.8 ms/step from N=1 to N=8 in round 1).  Returns a description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        sysfs = Path("/sys/bus/pci/devices") / f"{dom[-4:]}:{rest}".lower()
        node = int((sysfs / "numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"NUMA node {node} ({len(cpus)} CPUs)"
    except Exception:
        return None


def dist_setup():
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        where = bind_to_gpu_numa_node(local_rank)
        if where is not None:
            sys.stderr.write(f"[bench] rank {rank}: bound to {where}\n")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world, dev, dist


def make_runner(fn, warmup: int, no_graph: bool):
    """Warm up `fn` (>= 3 calls, incl. any collective inside it) and, unless --no-graph, capture it into a CUDA graph: the
    whole step — forward, backward, exchange — becomes one launch, so host jitter cannot open gaps between kernels."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(warmup, 3)):
            out = fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if no_graph:
        return fn, out, False
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    graph.replay()
    torch.cuda.synchronize()
    return graph.replay, out, True


def run_ours(args):
    import warnings
    warnings.filterwarnings("ignore", message=".*AccumulateGrad node's stream.*")
    log("importing")
    from dreammesh4d_b200 import _lib
    from dreammesh4d_b200 import rasterizer as R

    rank, local_rank, world, dev, dist = dist_setup()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N bench.py --gpus N)")

    log("building workload")
    scene, graph, node = build_scene(args.small)
    cams = build_cameras(rank)
    gs_dev = gaussian_sets_gpu(scene, graph, node, dev)
    gs = {k: v.cpu() for k, v in gs_dev.items()}
    means, rots, normals = gs["means3D"], gs["rotations"], gs["normals"]      # [8,P,3], [8,P,4]
    scales, opac, cols = gs["scales"], gs["opacities"], gs["colors"]         # shared [P,k]
    P = means.shape[1]
    V, PV, campos, tanx, tany = cams
    bg = torch.ones(VIEWS, 3)

    # ---- device-resident copies for `value` ----
    d = lambda t: t.to(dev).contiguous()
    vp = R.make_view_params(d(V), d(PV), d(campos), tanx, tany, d(bg), set_index=torch.arange(VIEWS))
    g = torch.Generator().manual_seed(1234 + rank)
    gC_h = torch.randn(VIEWS, 3, H, W, generator=g).pin_memory()
    gD_h = torch.randn(VIEWS, 1, H, W, generator=g).mul_(0.1).pin_memory()
    gA_h = torch.randn(VIEWS, 1, H, W, generator=g).pin_memory()
    host_in = {k: v.contiguous().pin_memory() for k, v in
               dict(means=means, rots=rots, scales=scales, opac=opac, cols=cols).items()}
    dev_in = {k: d(v).requires_grad_(True) for k, v in host_in.items()}
    gC, gD, gA = d(gC_h), d(gD_h), d(gA_h)

    def capacity_for(inp, vpar):
        st = []
        with torch.no_grad():
            R.rasterize_batch(inp["means"], inp["opac"], inp["scales"], inp["rots"], inp["cols"], vpar, H, W, state_out=st)
        return int(st[0].status()[0] * 1.25) + 4096

    capacity = capacity_for(dev_in, vp)       # one synchronous run tells R; afterwards the path is free of host syncs

    def make_step(inp, vpar, gCv, gDv, gAv, cap, distinct=True):
        """forward + backward of a rank's views through the public API (CUDA-graph capturable), followed by the path's one
        exchange step: NCCL sum of the shared-attribute gradients (same stream, inside the graph)."""
        screen = torch.zeros(vpar.shape[0], P, 3, device=dev, requires_grad=True)     # viewspace_points holder (temporal.py:108-113)

        def step():
            out_state = []
            color, radii, depth, alpha = R.rasterize_batch(inp["means"], inp["opac"], inp["scales"], inp["rots"], inp["cols"],
                                                           vpar, H, W, means2D=screen, capacity=cap, distinct_sets=distinct,
                                                           state_out=out_state)
            torch.autograd.backward([color, depth, alpha], [gCv, gDv, gAv])
            grads = {k: inp[k].grad for k in inp}
            grads["means2D"] = screen.grad
            for k in inp:
                inp[k].grad = None
            screen.grad = None
            shared = None
            if dist is not None:
                shared = torch.cat([grads["scales"].reshape(-1), grads["opac"].reshape(-1), grads["cols"].reshape(-1)])
                dist.all_reduce(shared)
            return color, depth, alpha, grads, shared, out_state[0]
        return step

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import gc
    gc.collect()
    gc.freeze()          # keep the cyclic GC out of the timed regions

    def timed(run, n):
        barrier()
        ms = time_events(lambda i: run(), n, before=flush_buf.zero_)
        barrier()
        return ms

    log("warm-up + capture")
    # ---- timed: K steps, CUDA events per step on the launch stream, L2 flushed between steps ----
    local_step = make_step(dev_in, vp, gC, gD, gA, capacity)
    exchange_in_graph = True
    try:
        replay_step, step_out, graphed = make_runner(local_step, args.warmup, args.no_graph)
    except Exception as e:      # a collective that cannot be captured: eager step (exchange still inside, warmed up)
        log(f"graph capture failed ({type(e).__name__}: {e}); running the step eagerly")
        torch.cuda.synchronize()
        replay_step, step_out, graphed = make_runner(local_step, args.warmup, True)
        exchange_in_graph = False

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms = timed(replay_step, args.steps)
    n_r, overflow = step_out[-1].status()
    if overflow:
        raise SystemExit("bin capacity overflow during the timed region — result invalid")
    log(f"timed region done: {sum(step_ms) / args.steps:.3f} ms/step")

    # ---- strong scaling (C3 as BASELINE names it: the SAME 8 views sharded over the ranks) ----
    strong = None
    if world > 1 and VIEWS % world == 0:
        from dreammesh4d_b200.dist import shard_views
        mine = shard_views(VIEWS, rank, world)
        cams0 = build_cameras(0)                                  # every rank shards rank 0's cameras
        V0, PV0, c0, tx0, ty0 = cams0
        vps = R.make_view_params(d(V0[mine]), d(PV0[mine]), d(c0[mine]), tx0[mine], ty0[mine], d(bg[:len(mine)]),
                                 set_index=torch.arange(len(mine)))
        sub = {"means": dev_in["means"].detach()[mine.to(dev)].clone().requires_grad_(True),
               "rots": dev_in["rots"].detach()[mine.to(dev)].clone().requires_grad_(True),
               **{k: dev_in[k].detach().clone().requires_grad_(True) for k in ("scales", "opac", "cols")}}
        cap_s = capacity_for(sub, vps)
        sstep = make_step(sub, vps, gC[:len(mine)].contiguous(), gD[:len(mine)].contiguous(), gA[:len(mine)].contiguous(), cap_s)
        try:
            sreplay, sout, sgraphed = make_runner(sstep, args.warmup, args.no_graph or not exchange_in_graph)
            s_ms = timed(sreplay, args.steps)
            t = torch.tensor([sum(s_ms)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_per = float(t[0]) / args.steps
            strong = {"scaling": "strong", "views_total": VIEWS, "views_per_gpu": len(mine), "ms_per_step": s_per,
                      "value": P * VIEWS / (s_per * 1e-3), "unit": "Gaussians/s", "steps": args.steps,
                      "what": "C3 as BASELINE names it: the SAME 8 views sharded round-robin over the ranks (dist.shard_views), one NCCL "
                              "all-reduce of the shared-attribute gradients per step inside the step's CUDA graph; value = P x 8 / t, "
                              "max over ranks; divide by the N=1 `value` for the strong-scaling speed-up"}
            log(f"strong scaling: {s_per:.3f} ms/step")
        except Exception as e:
            log(f"strong-scaling leg skipped: {type(e).__name__}: {e}")

    # ---- per-kernel device times (CUDA events around every launch inside libdm4d), same steps, eager ----
    _lib.profile_enable(True)
    _lib.profile_collect()
    prof_steps = max(3, min(args.steps, 10))
    for _ in range(prof_steps):
        flush_buf.zero_()
        local_step()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    launches_per_step = int(sum(n for _, n in prof.values())) // prof_steps
    # the tile sort is one profiling scope but five kernel launches (one per tier: 1024x16, 1024x8, 512x8, 256x8, 128x8 keys)
    launches_per_step += 4 * (int(prof.get("sort_pack_kernel", (0.0, 0))[1]) // prof_steps)
    if args.kernels_only:
        sampler.result()
        if rank == 0:
            print(json.dumps({"tuning": True, "ms_per_step": sum(step_ms) / args.steps,
                              "kernels_ms": {k: round(ms / max(n, 1), 4) for k, (ms, n) in prof.items()}}), flush=True)
        return None

    # ---- e2e: same work through the public API with HOST (pinned) buffers, every copy inside the timed region ----
    # HostStreamedRasterStep software-pipelines consecutive steps over three streams (H2D of step i+1 | fwd+bwd of
    # step i | D2H of step i-1).  The region is timed as a whole (first H2D .. last D2H) and divided by the step count.
    from dreammesh4d_b200.streaming import HostStreamedRasterStep
    streamed = HostStreamedRasterStep(host_in, {"gC": gC_h, "gD": gD_h, "gA": gA_h}, vp, H, W, capacity,
                                      reduce_fn=(lambda flat: dist.all_reduce(flat)) if dist is not None else None)
    h2d, d2h = streamed.h2d_bytes, streamed.d2h_bytes
    log("e2e")
    e_steps = max(3, args.steps)
    if args.no_graph or dist is not None:      # multi-rank e2e region stays eager
        streamed.run_many(3)                   # warm-up
        torch.cuda.synchronize()
        run_region = lambda: streamed.run_many(e_steps)
    else:                                      # the whole K-step pipelined region is ONE graph: immune to host jitter
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            streamed.run_many(3)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        region = torch.cuda.CUDAGraph()
        with torch.cuda.graph(region):
            streamed.run_many(e_steps)
        region.replay()
        torch.cuda.synchronize()
        run_region = region.replay
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_region()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    log(f"e2e done: {e2e_ms / e_steps:.3f} ms/step")
    if streamed.overflowed():
        raise SystemExit("bin capacity overflow in the e2e region — result invalid")
    clocks = sampler.result()

    # ---- the zero-edit drop-in path, as the reference's renderer calls it ----
    dropin = None
    if dist is None:
        try:
            dropin = dropin_ms(dev_in, gs_dev["normals"], cams, gC, gD, gA, dev, steps=max(5, min(args.steps, 10)))
            log(f"drop-in path: {dropin['rgb_and_normal_bwd']['ms_median']:.3f} ms/step")
        except Exception as e:
            log(f"drop-in leg skipped: {type(e).__name__}: {e}")

    # ---- context: the classic pipeline on the same GPU (our transcription of the upstream structure) ----
    ref_eq = None
    if dist is None:
        try:
            ref_eq = ref_equiv_ms(dev_in, vp, gC, gD, gA, P, steps=max(3, min(args.steps, 10)))
            log(f"ref-equivalent pipeline: {ref_eq['ms_per_step']:.3f} ms/step")
            # the reference renders RGB and normals as TWO rasterizer calls per view (temporal.py:169-178, 202-211); the
            # product fuses them into one 6-channel pass: time that pass on the same views, like `value`
            nrm = gs_dev["normals"].contiguous().requires_grad_(True)
            g6 = torch.cat([gC, gC.flip(1)], dim=1).contiguous()
            vp6 = R.make_view_params(d(V), d(PV), d(campos), tanx, tany, torch.ones(VIEWS, 6, device=dev), set_index=torch.arange(VIEWS))

            def six():
                c6, _, dep, alp = R.rasterize_batch(dev_in["means"], dev_in["opac"], dev_in["scales"], dev_in["rots"], dev_in["cols"],
                                                    vp6, H, W, colors2=nrm, capacity=capacity, distinct_sets=True)
                torch.autograd.backward([c6, dep, alp], [g6, gD, gA])
                for t_ in list(dev_in.values()) + [nrm]:
                    t_.grad = None
            replay6, _, _ = make_runner(six, args.warmup, args.no_graph)   # same launch mode as `value`
            six_ms = sum(timed(replay6, 10)) / 10
            ref_eq["rgb_plus_normal"] = {"ours_fused_6ch_ms": six_ms, "ref_equiv_two_passes_ms": 2 * ref_eq["ms_per_step"],
                                         "speedup": round(2 * ref_eq["ms_per_step"] / six_ms, 2),
                                         "what": "what the reference's renderer asks of the rasterizer per step: RGB and normal "
                                                 "images of 8 views, forward + backward (two calls per view there, one fused "
                                                 "6-channel pass here)"}
            if dropin is not None:
                dropin["batched_fused_6ch_ms"] = six_ms
                dropin["ratio_to_batched"] = round(dropin["rgb_and_normal_bwd"]["ms_median"] / six_ms, 2)
        except Exception as e:
            log(f"ref-equivalent measurement skipped: {type(e).__name__}: {e}")

    # ---- context: full hot-path optimizer step INCLUDING the Zero123 SDS term ("train-step ms"), C3 then C2 ----
    train, train_c2, sds = None, None, None
    if not args.no_train:
        del streamed
        torch.cuda.empty_cache()
        try:
            sds = sds_tensor_leg(dev, VIEWS, steps=10)
            log(f"Zero123 UNet fwd: {sds['unet_fwd']['ms']:.2f} ms ({sds['unet_fwd']['tflops']:.0f} TFLOP/s), "
                f"encoder fwd+bwd: {sds['encoder_fwd_bwd']['ms']:.2f} ms ({sds['encoder_fwd_bwd']['tflops']:.0f} TFLOP/s)")
        except Exception as e:
            log(f"SDS tensor leg skipped: {type(e).__name__}: {e}")
        try:
            train = train_step_ms(dev, dist, steps=max(5, min(args.steps, 12)), n_faces=N_FACES, views=VIEWS, n_frames=32,
                                  label="C3 (100k faces / 300k Gaussians, 8 views per substep, 32 frames, 512^2)", small=args.small)
            log(f"train step C3: {train['ms_median']:.3f} ms")
        except Exception as e:      # context only: never fail the bench line on it
            log(f"train-step measurement skipped: {type(e).__name__}: {e}")
        if dist is None and not args.small:
            try:
                train_c2 = train_step_ms(dev, None, steps=max(5, min(args.steps, 12)), n_faces=50_000, views=4, n_frames=16,
                                         label="C2 (50k faces / 150k Gaussians, 4 views per substep, 16 frames, 512^2)")
                log(f"train step C2: {train_c2['ms_median']:.3f} ms")
            except Exception as e:
                log(f"C2 train-step measurement skipped: {type(e).__name__}: {e}")

    # ---- max over ranks ----
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    e2e_ms_per_step = e2e_ms / e_steps
    value = P * VIEWS * world / (ms_per_step * 1e-3)
    e2e_value = P * VIEWS * world / (e2e_ms_per_step * 1e-3)

    out = None
    if rank == 0:
        peak, peak_src, _ = measured_peaks()
        kern = {k: {"ms_per_launch": ms / max(n, 1), "launches": n, "share": ms / max(sum(m for m, _ in prof.values()), 1e-9)}
                for k, (ms, n) in prof.items()}
        tiles = VIEWS * ((H + 15) // 16) * ((W + 15) // 16)
        for k in kern:
            ab = algorithmic_bytes(k, VIEWS * P, n_r, VIEWS * H * W, tiles)
            if ab:
                kern[k]["hbm_frac"] = round(ab / (kern[k]["ms_per_launch"] * 1e-3) / 1e9 / peak, 4)
        dom = max(prof, key=lambda k: prof[k][0]) if prof else None
        roof = None
        if dom:
            ab = algorithmic_bytes(dom, VIEWS * P, n_r, VIEWS * H * W, tiles)
            dur = prof[dom][0] / prof[dom][1] * 1e-3
            ach = ab / dur / 1e9
            traffic, instr = None, None
            tf = ROOT / "profiles" / "traffic.json"
            if tf.exists():
                try:
                    j = json.loads(tf.read_text())
                    traffic, instr = j.get(dom), j.get("_instructions_executed", {}).get(dom)
                except Exception:
                    pass
            roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "algorithmic_bytes": ab,
                    "launch_ms": round(dur * 1e3, 4), "peak_source": peak_src}
            if instr and clocks.get("sm_mhz"):
                # the roofline that actually binds the render kernels: warp instructions issued per launch (ncu
                # smsp__inst_executed.sum of the capture named in profiles/traffic.json) against 148 SMs x 4 issue slots
                roof["issue_frac"] = round(instr / dur / issue_peak_per_s(clocks["sm_mhz"]), 4)
                roof["warp_instructions_per_launch"] = instr
                roof["issue_note"] = ("issue_frac = ncu warp instructions per launch / live launch duration / (148 SMs x 4 schedulers x SM "
                                      "clock): the kernel is instruction-issue bound, not HBM bound (DESIGN.md §6)")
        cpu = None
        if world == 1:       # reported baseline: rank 0 at N=1 only
            log("cpu baseline")
            cpu = cpu_baseline_sample(host_in, cams, P, views=VIEWS, reps=4)      # ~10 s of CPU work on 16 threads
        st = stats(step_ms)
        out = {
            "metric": METRIC, "value": value, "unit": "Gaussians/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "ms_per_step_median": st["ms_median"], "ms_per_step_max": st["ms_max"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(P, args.small),
            "run": {"num_rendered": n_r,
                    "timing": "CUDA events per step on the launch stream, summed over K steps, max over ranks (median / max of rank 0 beside it)",
                    "launch": ("one CUDA-graph replay per step (forward + backward" + (" + NCCL exchange" if world > 1 and exchange_in_graph else "") +
                               " captured through the public API)") if graphed else "eager",
                    "exchange": "none" if world == 1 else "NCCL all-reduce of the shared-attribute gradients (8.4 MB) per step, warmed up with the step"},
            "e2e": {"value": e2e_value, "unit": "Gaussians/s", "ms_per_step": e2e_ms_per_step,
                    "how": "HostStreamedRasterStep: pinned host sets + image gradients -> H2D | fwd+bwd (+exchange) | D2H of images and all gradients to pinned host, software-pipelined across consecutive steps on 3 streams; whole region timed with CUDA events (235 MB/step > L2, no flush)",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps},
            "gpu_launches": (launches_per_step + (1 if world > 1 else 0)) * args.steps,
            "kernels": kern, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "strong_scaling": strong, "dropin": dropin, "sds": sds, "train_step": train, "train_step_c2": train_c2,
            "ref_equiv": None if ref_eq is None else dict(ref_eq, speedup=round(ref_eq["ms_per_step"] / ms_per_step, 2)),
        }
    if out is not None:
        print(json.dumps(out), flush=True)
    finish(dist)
    return None


def finish(dist):
    """End of a (possibly multi-rank) run.  With NCCL collectives captured in live CUDA graphs,
    ``destroy_process_group`` can block forever in communicator teardown (seen at N=2: ten minutes until the watchdog);
    the result line is already on stdout, so synchronise, meet at a barrier and leave without the teardown."""
    if dist is None:
        return
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


# ------------------------------------------------------------------------------------------------
# other configurations (profiles/ evidence; not driver lines)
# ------------------------------------------------------------------------------------------------
def run_c2(args):
    rank, local_rank, world, dev, dist = dist_setup()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sds = sds_tensor_leg(dev, 4, steps=10)
    train = train_step_ms(dev, dist, steps=max(5, args.steps), n_faces=50_000, views=4, n_frames=16,
                          label="C2 (50k faces / 150k Gaussians, 4 views per substep, 16 frames, 512^2)")
    if rank == 0:
        print(json.dumps({"metric": "train-step ms (dynamic stage incl. Zero123 SDS)", "value": train["ms_median"], "unit": "ms",
                          "n_gpus": world, "higher_is_better": False, "dtype": "f32 (raster/skinning) + f16 (Zero123)", "data": "synthetic",
                          "config": {"workload": train["label"]}, "train_step": train, "sds": sds, "clocks": sampler.result()}), flush=True)
    finish(dist)


def run_c4(args):
    """BASELINE config 4: 1M free Gaussians, 1024x1024, 16 cameras, one 3-channel pass fwd+bwd."""
    from dreammesh4d_b200 import _lib, synthetic
    from dreammesh4d_b200 import rasterizer as R
    rank, local_rank, world, dev, dist = dist_setup()
    P4, HW, NV = (20_000, 256, 4) if args.small else (1_000_000, 1024, 16)
    means, scales, rots, opac, cols = [t.to(dev).requires_grad_(True) for t in synthetic.random_gaussians(P4, seed=0)]
    V, PV, campos, tanx, tany = build_cameras(rank, NV)
    vp = R.make_view_params(V.to(dev), PV.to(dev), campos.to(dev), tanx, tany, torch.ones(NV, 3))
    g = torch.Generator().manual_seed(3)
    gC, gD, gA = (torch.randn(NV, c, HW, HW, generator=g).to(dev) for c in (3, 1, 1))
    st = []
    with torch.no_grad():
        R.rasterize_batch(means, opac, scales, rots, cols, vp, HW, HW, state_out=st)
    n_r = st[0].status()[0]
    cap = int(n_r * 1.1) + 4096
    del st
    screen = torch.zeros(NV, P4, 3, device=dev, requires_grad=True)

    def step():
        c, _, dpt, a = R.rasterize_batch(means, opac, scales, rots, cols, vp, HW, HW, means2D=screen, capacity=cap)
        torch.autograd.backward([c, dpt, a], [gC, gD, gA])
        for t in (means, scales, rots, opac, cols, screen):
            t.grad = None

    replay, _, graphed = make_runner(step, args.warmup, args.no_graph)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = time_events(lambda i: replay(), args.steps, before=flush.zero_)
    _lib.profile_enable(True)
    _lib.profile_collect()
    for _ in range(3):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    clocks = sampler.result()
    peak, peak_src, _ = measured_peaks()
    tiles = NV * ((HW + 15) // 16) ** 2
    kern = {}
    for k, (t_ms, n) in prof.items():
        ab = algorithmic_bytes(k, NV * P4, n_r, NV * HW * HW, tiles)
        kern[k] = {"ms_per_launch": t_ms / n, "hbm_frac": round(ab / (t_ms / n * 1e-3) / 1e9 / peak, 4) if ab else None}
    s = stats(ms)
    print(json.dumps({"metric": "rasterize fwd+bwd Gaussians/s @1024x1024, 16 cameras (BASELINE config 4)",
                      "value": P4 * NV / (s["ms_mean"] * 1e-3), "unit": "Gaussians/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": s["ms_mean"], "ms_per_step_median": s["ms_median"],
                      "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "C4: 1M free Gaussians (U(ball r=0.5), log-uniform scales 1e-3..1e-2), 1024x1024, 16 orbit cameras, "
                                             "1 pass (3ch) fwd+bwd with colour+depth+alpha+means2D gradients", "P": P4, "H": HW, "W": HW,
                                 "views": NV, "num_rendered": n_r, "l2": "flushed between steps", "launch": "CUDA graph" if graphed else "eager"},
                      "kernels": kern, "clocks": clocks, "peak_source": peak_src}), flush=True)


def run_c5(args):
    """BASELINE config 5: LBS/DQ skinning + per-face Gaussian update, 200k vertices / 512 control nodes / 600k Gaussians."""
    from dreammesh4d_b200 import _lib, skinning, synthetic
    rank, local_rank, world, dev, dist = dist_setup()
    Vn, Fn, M, T = (5_000, 5_000, 64, 2) if args.small else (200_000, 200_000, 512, 1)
    verts, faces = synthetic.uv_sphere(2 * (Vn - 2))                  # closed sphere with Vn vertices, 2(Vn-2) faces
    faces = faces[torch.randperm(faces.shape[0], generator=torch.Generator().manual_seed(0))[:Fn].sort()[0]]   # 200k-face subset
    scene = synthetic.make_sugar_scene(2 * (Vn - 2), g=3)
    scene.faces = faces
    P5 = Fn * 3
    scene.log_scales, scene.complex_rot = scene.log_scales[:P5], scene.complex_rot[:P5]
    graph = synthetic.make_deform_graph(scene.verts, M, K_NBR, seed=0)
    d = lambda t: t.to(dev)
    fi = d(faces.int())
    rq, _ = skinning.sugar_rest_frames(d(scene.verts), fi, d(scene.complex_rot), 3)
    static = (d(scene.verts), fi, d(graph.nbr_idx.int()), d(graph.nbr_w), d(scene.bary), rq)
    peak, peak_src, _ = measured_peaks()
    V_, K = scene.verts.shape[0], K_NBR
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def measure(method, T):
        node = [d(t).requires_grad_(True) for t in synthetic.random_node_attrs(T, M, seed=1)]
        gm, gr, gn = (torch.randn(T, P5, k, device=dev) for k in (3, 4, 3))

        def step():
            means, rots, normals, vx, vr = skinning.skin_gaussians(*node, *static, method=method)
            torch.autograd.backward([means, rots, normals], [gm, gr, gn])
            for t in node:
                t.grad = None

        replay, _, _ = make_runner(step, args.warmup, args.no_graph)
        ms = stats(time_events(lambda i: replay(), args.steps, before=flush.zero_))
        _lib.profile_enable(True)
        _lib.profile_collect()
        for _ in range(5):
            flush.zero_()
            step()
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        # algorithmic bytes per timestamp: fwd reads 12V + 8KV + 68M + 12F + 16P(rest quat) ; writes 28V + 40P
        fwd = T * (12 * V_ + 8 * K * V_ + 68 * M + 12 * Fn + 16 * P5 + 28 * V_ + 40 * P5)
        bwd = T * (40 * P5 + 12 * Fn + 16 * P5 + 28 * V_ + 28 * V_ + 12 * V_ + 8 * K * V_ + 68 * M + 68 * M)
        return {"timestamps": T, "fwd_bwd_us": ms["ms_median"] * 1e3, "algorithmic_bytes": fwd + bwd,
                "GBps": (fwd + bwd) / (ms["ms_median"] * 1e-3) / 1e9, "hbm_frac": round((fwd + bwd) / (ms["ms_median"] * 1e-3) / 1e9 / peak, 4),
                "kernels_us": {k: round(t_ms / n * 1e3, 2) for k, (t_ms, n) in prof.items()}}

    out = {method: measure(method, T) for method in ("hybrid", "lbs", "dqs")}
    T8 = 2 * T if args.small else 8
    batched = {method: measure(method, T8) for method in ("hybrid",)}
    print(json.dumps({"metric": "skinning + per-face Gaussian update fwd+bwd (BASELINE config 5)", "value": out["hybrid"]["GBps"],
                      "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"C5: {V_} vertices, {M} control nodes, K={K}, {Fn} faces, {P5} Gaussians, {T} timestamp(s), "
                                             "fused skinning forward + backward (all node-attribute gradients); one CUDA-graph replay per step, "
                                             "L2 flushed between steps"},
                      "methods": out, "batched_timestamps": batched,
                      "batched_note": f"the step of the hot path deforms all its timestamps in one launch sequence (8 at C3, 4 at C2): the same "
                                      f"microbench at {T8} timestamps", "peak_GBps": peak, "peak_source": peak_src}), flush=True)


# ------------------------------------------------------------------------------------------------
# CPU oracle legs
# ------------------------------------------------------------------------------------------------
def cpu_baseline_sample(host_in, cams, P, views: int, reps: int = 1):
    """Times the CPU oracle (oracle/raster_oracle.c, OpenMP on every host core) on `views` of the
    same workload: forward + backward. Reported baseline, not the target."""
    from oracle.raster_oracle import RasterOracle, cpu_threads
    V, PV, campos, tanx, tany = cams
    g = np.random.default_rng(0)
    gC = g.standard_normal((3, H, W)).astype(np.float32)
    gD = (0.1 * g.standard_normal((1, H, W))).astype(np.float32)
    gA = g.standard_normal((1, H, W)).astype(np.float32)
    o = RasterOracle(P, H, W, 3, "f32")
    t0 = time.perf_counter()
    done = 0
    for _ in range(reps):
        for v in range(views):
            o.forward(host_in["means"][v].numpy(), host_in["scales"].numpy(), host_in["rots"][v].numpy(),
                      host_in["opac"].numpy(), host_in["cols"].numpy(), V[v].numpy(), PV[v].numpy(), float(tanx[v]),
                      float(tany[v]), np.ones(3, np.float32))
            o.backward(gC, gD, gA)
            done += 1
    dt = time.perf_counter() - t0
    return {"value": P * done / dt, "unit": "Gaussians/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"{done} view(s) of the same workload ({P} Gaussians, {H}x{W}), fwd+bwd, {dt:.2f} s",
            "seconds": dt}


def run_reference(args):
    """`--impl reference`: the reference's rasterizer is CUDA-only and not obtainable here,
    so this arm times the CPU oracle port on all host cores; each step = the 8 views of the same workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    scene, graph, node = build_scene(args.small)
    cams = build_cameras(0)
    gs = gaussian_sets_oracle(scene, graph, node)
    host_in = dict(means=gs["means3D"], rots=gs["rotations"], scales=gs["scales"], opac=gs["opacities"], cols=gs["colors"])
    P = host_in["means"].shape[1]
    for _ in range(min(args.warmup, 1)):
        cpu_baseline_sample(host_in, cams, P, views=1)
    steps = max(1, min(args.steps, 6))
    t0 = time.perf_counter()
    for _ in range(steps):           # one step = the 8 views of the workload, like a step of the GPU arm
        cpu_baseline_sample(host_in, cams, P, views=VIEWS)
    dt = time.perf_counter() - t0
    value = P * VIEWS * steps / dt
    from oracle.raster_oracle import cpu_threads
    cpu = {"value": value, "unit": "Gaussians/s", "cores": cpu_threads(), "kind": "port",
           "sample": f"each step = the {VIEWS} views of the workload ({P} Gaussians, {H}x{W}) fwd+bwd on the CPU oracle; {steps} steps "
                     f"(bounded: at most 6)"}
    return {"impl": "reference", "metric": METRIC, "value": value,
            "unit": "Gaussians/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(P, args.small),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(int(os.environ.get("DM4D_BENCH_WATCHDOG_S", "420")), exit=True)   # never hang a GPU box
    args = parse()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 for N>1 launches: the CPU arm is meant to use every host core
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    # stdout must carry exactly ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1, so route fd 1
    # to stderr while running and keep the real stdout for the result line.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            out = run_reference(args)
        else:
            out = {"c3": run_ours, "c2": run_c2, "c4": run_c4, "c5": run_c5}[args.config](args)
    except BaseException:
        import traceback
        log("FAILED rank %s:\n%s" % (os.environ.get("RANK", "0"), traceback.format_exc()))
        raise
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
