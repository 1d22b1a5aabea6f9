#!/usr/bin/env python
"""Generates tests/golden/arap.npz by executing the reference's own ARAPCoach
(custom/threestudio-dreammesh4d/utils/arap_utils.py, unmodified source, read from /root/reference at generation
time only) on a small seeded mesh with supplied per-vertex rotations.  Stand-ins are installed only for the modules
the file imports but the exercised path does not use (open3d, threestudio.utils.typing)."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
ROOT = OUT.parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/custom/threestudio-dreammesh4d/utils/arap_utils.py")


class _Ann:
    def __getitem__(self, k): return self


def main():
    typing_stub = types.ModuleType("threestudio.utils.typing")
    for n in ("Float", "Int", "Tensor", "Dict", "List", "Optional", "Union", "Any", "Tuple"):
        setattr(typing_stub, n, _Ann())
    sys.modules.update({"open3d": types.ModuleType("open3d"), "threestudio": types.ModuleType("threestudio"),
                        "threestudio.utils": types.ModuleType("threestudio.utils"), "threestudio.utils.typing": typing_stub})
    ns = {"__name__": "ref_arap"}
    exec(compile(REF.read_text(), str(REF), "exec"), ns)
    ARAPCoach = ns["ARAPCoach"]

    from dreammesh4d_b200 import synthetic
    from oracle.skin_oracle import q_act
    verts, faces = synthetic.uv_sphere(264)
    g = torch.Generator().manual_seed(0)
    verts = (verts + 0.02 * torch.randn(verts.shape, generator=g)).float()
    coach = ARAPCoach(verts, faces.numpy(), torch.device("cpu"))
    T = 2
    xp = verts[None] + 0.05 * torch.randn(T, *verts.shape, generator=g)
    q = torch.nn.functional.normalize(torch.cat([0.2 * torch.randn(T, verts.shape[0], 3, generator=g),
                                                 torch.ones(T, verts.shape[0], 1)], dim=-1), dim=-1)
    I = torch.eye(3)
    Rm = torch.stack([q_act(q[..., None, :].expand(-1, -1, 3, -1), I.expand(T, verts.shape[0], 3, 3))], 0)[0].transpose(-1, -2)
    energy = torch.stack([coach.compute_arap_energy(xyz_prime=xp[t], vert_rotations=Rm[t]) for t in range(T)])
    np.savez_compressed(OUT / "arap.npz", verts=verts.numpy(), faces=faces.numpy(), verts_def=xp.numpy(), rot_xyzw=q.numpy(),
                        rot_matrix=Rm.numpy(), energy=energy.numpy(), edge_cot_weights=coach.edge_cot_weights.numpy())
    print("wrote arap.npz energy", energy.tolist())


if __name__ == "__main__":
    main()
