import sys; sys.path.insert(0,'.')
import numpy as np, torch
from dreammesh4d_b200 import synthetic
from dreammesh4d_b200.camera import get_cam_info_gaussian
from dreammesh4d_b200.geometry import DynamicSuGaRGeometry
from dreammesh4d_b200.renderer import DiffGaussianBatchRenderer
from dreammesh4d_b200 import rasterizer as R
from oracle.raster_oracle import RasterOracle
from tests.test_renderer_gpu import make_rays
DEV='cuda'
B,H,W=1,256,256
scene = synthetic.make_sugar_scene(10_000, g=3)
graph = synthetic.make_deform_graph(scene.verts, 16, 4)
geo = DynamicSuGaRGeometry(scene, graph, None, static_learnable=True).to(DEV)
ren = DiffGaussianBatchRenderer(geo)
c2w, fovy = synthetic.random_orbit_cameras(B, seed=5)
rays_o, rays_d = make_rays(c2w, fovy, H, W)
batch = {"c2w": c2w.to(DEV), "fovy": fovy.to(DEV), "height": H, "width": W, "rays_o": rays_o.to(DEV), "rays_d": rays_d.to(DEV)}
geo.update_step(0,0)
out = ren.batch_forward(batch)
cpu=lambda t:t.detach().cpu()
Vm, PV, campos, tanx, tany = get_cam_info_gaussian(c2w, fovy, fovy)
P=scene.n_gaussians
for C in (6,3):
    o = RasterOracle(P, H, W, C, "f32")
    feat = torch.cat([cpu(geo.get_points_rgb()), cpu(geo.get_gs_normals)], dim=1)[:, :C]
    c, r, d, a = o.forward(cpu(geo.get_xyz).numpy(), cpu(geo.get_scaling).numpy(), cpu(geo.get_rotation).numpy(), cpu(geo.get_opacity).numpy(), feat.numpy(), Vm[0].numpy(), PV[0].numpy(), float(tanx[0]), float(tany[0]), np.ones(C, np.float32))
    ok=~o.ambiguous
    got = cpu(out["comp_rgb"]).permute(0,3,1,2)[0].numpy()
    err = np.abs(got - np.clip(c[:3],0,1))*ok[None]
    idx = np.unravel_index(err.argmax(), err.shape)
    print("C",C,"max err", err.max(), "at", idx, "n>1e-4", (err>1e-4).sum(), "ambig", (~ok).sum(), "alpha", a[0][idx[1:]], "ncontrib", o.n_contrib[idx[1:]])
    ga = cpu(out["comp_mask"]).permute(0,3,1,2)[0].numpy()
    print("  alpha err", (np.abs(ga-a)*ok[None]).max())
# direct rasterize_batch with 3 channels same inputs
vp = R.make_view_params(Vm.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, torch.ones(1,6,device=DEV))
col, rad, dep, alp = R.rasterize_batch(geo.get_xyz.detach(), geo.get_opacity.detach(), geo.get_scaling.detach(), geo.get_rotation.detach(), geo.get_points_rgb().detach(), vp, H, W, colors2=geo.get_gs_normals.detach())
print("direct vs batch_forward rgb", (col[0,:3].clamp(0,1).cpu()-cpu(out["comp_rgb"]).permute(0,3,1,2)[0]).abs().max().item())
print("view params:", Vm[0], get_cam_info_gaussian(c2w.to(DEV), fovy.to(DEV), fovy.to(DEV))[0][0].cpu()-Vm[0])
