"""Camera matrices: dreammesh4d_b200.camera (batched, free of host synchronisation) against
golden vectors produced by executing the reference's own get_cam_info_gaussian / convert_pose /
get_projection_matrix_gaussian (tests/golden/make_camera_golden.py)."""
from pathlib import Path

import numpy as np
import torch

from dreammesh4d_b200.camera import get_cam_info_gaussian

GOLD = Path(__file__).resolve().parent / "golden" / "camera.npz"


def test_camera_block_matches_reference_code():
    z = np.load(GOLD)
    c2w, fovy = torch.from_numpy(z["c2w"]), torch.from_numpy(z["fovy"])
    view, full, center, tanx, tany = get_cam_info_gaussian(c2w, fovy, fovy, znear=0.1, zfar=100.0)
    assert np.abs(view.numpy() - z["world_view"]).max() <= 1e-6
    assert np.abs(full.numpy() - z["full_proj"]).max() <= 2e-6 * np.abs(z["full_proj"]).max()
    assert np.abs(center.numpy() - z["center"]).max() <= 2e-6
    assert torch.allclose(tanx, torch.tan(fovy * 0.5)) and torch.allclose(tany, torch.tan(fovy * 0.5))
    # the camera centre is the translation of c2w (the y/z flip does not move it)
    assert np.abs(center.numpy() - z["c2w"][:, :3, 3]).max() <= 2e-6
