import sys; sys.path.insert(0,'.')
import torch, bench
from dreammesh4d_b200 import rasterizer as R
dev=torch.device('cuda',0)
scene, graph, node = bench.build_scene(False); cams = bench.build_cameras(0)
gs = bench.gaussian_sets_gpu(scene, graph, node, dev)
V, PV, campos, tanx, tany = cams
vp = R.make_view_params(V.to(dev), PV.to(dev), campos.to(dev), tanx, tany, torch.ones(8,3,device=dev), set_index=torch.arange(8))
st=[]
R.rasterize_batch(gs["means3D"], gs["opacities"], gs["scales"], gs["rotations"], gs["colors"], vp, 512, 512, state_out=st)
cnts=[]
for v in range(8):
    ranges,_,nc = st[0].export_view(v)
    cnts.append((ranges[:,1]-ranges[:,0]).cpu())
c=torch.cat(cnts).float()
print("tiles", c.numel(), "nonempty", int((c>0).sum()), "mean(nonempty)", c[c>0].mean().item(), "max", c.max().item())
for q in (0.5,0.9,0.99,0.999): print("q",q, torch.quantile(c[c>0], q).item())
for thr in (1024,2048,4096,8192): print(">",thr, int((c>thr).sum()))
