#!/usr/bin/env python
"""Generates tests/golden/api_calls.json: how the reference's two renderers call the rasterizer module — the keyword
arguments of every ``GaussianRasterizationSettings(...)`` construction and ``rasterizer(...)`` call, the names imported
from ``diff_gaussian_rasterization`` and the arity of the unpacked return — extracted by AST from
custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_{temporal,normal}.py (read at generation time only).
tests/test_dropin_api.py checks the drop-in module accepts exactly these call shapes."""
import ast
import json
from pathlib import Path

OUT = Path(__file__).resolve().parent
REFDIR = Path("/root/reference/custom/threestudio-dreammesh4d/renderer")


def main():
    blob = {}
    for fname in ("diff_sugar_rasterizer_temporal.py", "diff_sugar_rasterizer_normal.py"):
        tree = ast.parse((REFDIR / fname).read_text())
        rec = {"imports": [], "settings_kwargs": [], "forward_kwargs": [], "return_arity": []}
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module == "diff_gaussian_rasterization":
                rec["imports"] += [a.name for a in node.names]
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
                if node.func.id == "GaussianRasterizationSettings":
                    rec["settings_kwargs"].append([k.arg for k in node.keywords])
                if node.func.id == "rasterizer":
                    rec["forward_kwargs"].append([k.arg for k in node.keywords])
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and isinstance(node.value.func, ast.Name) \
                    and node.value.func.id == "rasterizer" and isinstance(node.targets[0], ast.Tuple):
                rec["return_arity"].append(len(node.targets[0].elts))
        blob[fname] = rec
    (OUT / "api_calls.json").write_text(json.dumps(blob, indent=1) + "\n")
    print(json.dumps(blob, indent=1))


if __name__ == "__main__":
    main()
