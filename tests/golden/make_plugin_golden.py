#!/usr/bin/env python
"""Generates tests/golden/plugin_api.json from the reference's sources (AST only, nothing is executed):
  * the ``Config`` dataclass fields (name, default literal) of the four registered classes and their registered names,
  * every attribute the two systems read on ``self.geometry`` / ``self.renderer``
    (custom/threestudio-dreammesh4d/system/sugar_4dgen.py, sugar_static.py, base.py)."""
import ast
import json
from pathlib import Path

P = Path("/root/reference/custom/threestudio-dreammesh4d")
OUT = Path(__file__).resolve().parent / "plugin_api.json"
CLASSES = {"geometry/sugar.py": "SuGaRModel", "geometry/dynamic_sugar.py": "DynamicSuGaRModel",
           "renderer/diff_sugar_rasterizer_temporal.py": "DiffGaussian", "renderer/diff_sugar_rasterizer_normal.py": "DiffSuGaR"}


def config_of(path: Path, cls_name: str):
    tree = ast.parse(path.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    reg = next(ast.literal_eval(d.args[0]) for d in cls.decorator_list if isinstance(d, ast.Call) and getattr(d.func, "attr", "") == "register")
    cfg = next(n for n in cls.body if isinstance(n, ast.ClassDef) and n.name == "Config")
    fields = []
    for st in cfg.body:
        if isinstance(st, ast.AnnAssign) and isinstance(st.target, ast.Name):
            try:
                default = ast.literal_eval(st.value)
            except Exception:
                default = ast.unparse(st.value)
            fields.append([st.target.id, list(default) if isinstance(default, tuple) else default])
    return reg, fields


def attrs_on(path: Path, owner: str):
    found = set()
    for node in ast.walk(ast.parse(path.read_text())):
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Attribute) and node.value.attr == owner and \
                isinstance(node.value.value, ast.Name) and node.value.value.id == "self":
            found.add(node.attr)
    return sorted(found)


def main():
    out = {"classes": {}, "system_calls": {}}
    for rel, name in CLASSES.items():
        reg, fields = config_of(P / rel, name)
        out["classes"][reg] = {"class": name, "file": rel, "config": fields}
    for rel in ("system/sugar_4dgen.py", "system/sugar_static.py", "system/base.py"):
        out["system_calls"][rel] = {"geometry": attrs_on(P / rel, "geometry"), "renderer": attrs_on(P / rel, "renderer")}
    OUT.write_text(json.dumps(out, indent=1))
    print({k: len(v["config"]) for k, v in out["classes"].items()}, out["system_calls"])


if __name__ == "__main__":
    main()
