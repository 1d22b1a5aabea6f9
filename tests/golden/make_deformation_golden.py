#!/usr/bin/env python
"""Generates tests/golden/deformation.npz by executing the reference's own deformation network
(custom/threestudio-dreammesh4d/geometry/deformation.py, the WHOLE file, unmodified, read from /root/reference at
generation time only): ModelHiddenParams -> DeformationNetwork(args) on a reduced plane configuration (a config value,
not a source change), every parameter randomised, then DeformationNetwork.forward_dynamic_delta(points, times).
The archive holds the complete state_dict (names + tensors), the inputs and the four outputs: the test loads the
state_dict into dreammesh4d_b200.deformation.HexPlaneDeformation with strict=True (checkpoint compatibility) and
compares the outputs."""
from argparse import ArgumentParser
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/custom/threestudio-dreammesh4d/geometry/deformation.py")


def main():
    ns = {"__name__": "ref_deformation"}
    exec(compile(REF.read_text(), str(REF), "exec"), ns)
    args = ns["ModelHiddenParams"](ArgumentParser())
    args.kplanes_config = dict(args.kplanes_config, output_coordinate_dim=8, resolution=[6, 7, 8, 5])
    args.multires = [1, 2]
    args.no_do = False                         # exercise all four heads
    torch.manual_seed(0)
    net = ns["DeformationNetwork"](args)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.3 + (0.5 if p.dim() == 4 else 0.0))
    M, T = 9, 3
    xyz = torch.rand(M, 3, generator=g) * 1.6 - 0.8
    ts = torch.tensor([0.15, 0.5, 0.9])
    outs = []
    for t in ts:            # dynamic_sugar.py:430-436: one call per timestamp, time = 2 t - 1 broadcast over the nodes
        time = (t * 2 - 1).expand(M, 1)
        outs.append(net.forward_dynamic_delta(xyz, time))
    dx, dr, ds, do = (torch.stack([o[k] for o in outs]) for k in range(4))
    blob = {"xyz": xyz, "ts": ts, "dx": dx, "dr": dr, "ds": ds, "do": do}
    sd = net.state_dict()
    blob.update({"sd::" + k: v for k, v in sd.items()})
    np.savez_compressed(OUT / "deformation.npz", **{k: v.detach().numpy() for k, v in blob.items()})
    print("wrote deformation.npz:", len(sd), "state entries; feat_dim", net.deformation_net.grid.feat_dim)


if __name__ == "__main__":
    main()
