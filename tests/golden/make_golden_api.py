"""tests/golden/reference_api.json: the public surface of the reference's path modules, read from its SOURCE with ``ast``
(no import, so the absent third-party libraries do not matter): for every module-level function and every method of
``NDMPS`` the positional parameter names in order and the source text of their defaults.

Build container only (needs /root/reference):
    python tests/golden/make_golden_api.py
"""
import ast
import json
from pathlib import Path

SRC = Path("/root/reference/src/imgcompressionmps")
MODULES = ["core/ndmps.py", "utils/core.py", "utils/metrics.py", "utils/filetools.py", "evaluation/benchmark.py"]
OUT = Path(__file__).resolve().parent / "reference_api.json"


def signature(fn: ast.FunctionDef):
    args = fn.args
    names = [a.arg for a in args.posonlyargs + args.args]
    defaults = [None] * (len(names) - len(args.defaults)) + [ast.unparse(d) for d in args.defaults]
    return {"params": names, "defaults": defaults, "kwonly": [a.arg for a in args.kwonlyargs],
            "decorators": [ast.unparse(d) for d in fn.decorator_list]}


def main():
    api = {}
    for rel in MODULES:
        tree = ast.parse((SRC / rel).read_text())
        mod = {"functions": {}, "classes": {}}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
                mod["functions"][node.name] = signature(node)
            elif isinstance(node, ast.ClassDef):
                mod["classes"][node.name] = {m.name: signature(m) for m in node.body
                                             if isinstance(m, ast.FunctionDef) and (not m.name.startswith("_") or m.name == "__init__")}
        api[rel[:-3].replace("/", ".")] = mod
    OUT.write_text(json.dumps(api, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, {k: (len(v["functions"]), {c: len(m) for c, m in v["classes"].items()}) for k, v in api.items()})


if __name__ == "__main__":
    main()
