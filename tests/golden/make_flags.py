#!/usr/bin/env python
"""Extract the command-line interface of the reference's entry points (option strings, dest, default) into
tests/golden/entry_flags.json.  Run in the build container, where /root/reference exists; the fixture travels,
the reference does not.  Parsing is static (ast): the reference scripts import tensorboardX / imageio at top level
and cannot be imported here."""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "entry_flags.json")


def const(node, env):
    try:
        return eval(compile(ast.Expression(node), "<flag>", "eval"), {"__builtins__": {}}, env)
    except Exception:
        return "<expr>"


def flags_of(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            opts = [a.value for a in node.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
            kw = {k.arg: const(k.value, {}) for k in node.keywords if k.arg in ("default", "dest", "action", "nargs")}
            out.append({"options": opts, **{k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else "<expr>")
                                            for k, v in kw.items()}})
    return out


if __name__ == "__main__":
    res = {name: flags_of(os.path.join(REF, name + ".py")) for name in ("Train_Stage1_K", "Train_Stage1_Kslow", "Train_Stage2_K", "Test_KITTI")}
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    print({k: len(v) for k, v in res.items()}, "->", OUT)
