"""Extract the influent scenario DATA tables of the reference (buffer_tank3.py:18-1197) into
gym_sbr2_b200/data/influent_tables.npz.  Authoring container only (reads /root/reference).

The reference stores eight scenarios (switch 0..7) as literal 48-point mean profiles per component plus the
flow, a std of 0.1*mean for Ss, Xi, Xs, Xbh, Snh, Snd, Xnd and q (0 for the rest), and draws
np.random.randn(48) once (switch 0) or twice (switch 1..7, first draw discarded).  Only the numbers are taken
(parsed from the AST, nothing is executed); the generator itself is re-implemented in gym_sbr2_b200/influent.py.

Also extracted: the one live branch of buffer_tank2.py (`SBR-v0` / `SBR-v1`; `switch` is forced to 1, :18): 96-point
mean AND standard-deviation profiles per component and for the flow (keys bt2_mean, bt2_std [14, 96]).
"""
import ast
import os
import sys

import numpy as np

REF = os.path.join(os.environ.get("SBR_REFERENCE_ROOT", "/root/reference"), "gym_SBR", "envs", "buffer_tank3.py")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_sbr2_b200", "data",
                   "influent_tables.npz")
ORDER = ["q", "si", "ss", "xi", "xs", "xbh", "xba", "xp", "so", "sno", "snh", "snd", "xnd", "salk"]


def branch_tables(body):
    means, stds, draws = {}, {}, 0
    for node in ast.walk(ast.Module(body=body, type_ignores=[])):
        if not isinstance(node, ast.Assign) or len(node.targets) != 1 or not isinstance(node.targets[0], ast.Name):
            continue
        name = node.targets[0].id
        v = node.value
        if name.endswith("_m") and isinstance(v, ast.Call) and getattr(v.func, "attr", "") == "array":
            means[name[:-2]] = np.array(ast.literal_eval(v.args[0]), dtype=float)
        elif (name.endswith("_m") and isinstance(v, ast.BinOp) and isinstance(v.op, ast.Mult)
              and isinstance(v.left, ast.Constant) and isinstance(v.right, ast.Call)):
            # e.g. `ss_m = 1.5*np.array([...])` (carbon-/N-rich scenarios): same product as the reference forms
            means[name[:-2]] = v.left.value * np.array(ast.literal_eval(v.right.args[0]))
        elif name.endswith("_s"):
            if isinstance(v, ast.Constant):
                stds[name[:-2]] = float(v.value)
            elif isinstance(v, ast.BinOp) and isinstance(v.op, ast.Mult) and isinstance(v.left, ast.Constant):
                assert v.right.id == name[:-2] + "_m"
                stds[name[:-2]] = float(v.left.value)
        elif name == "rnd":
            draws += 1
    return means, stds, draws


def buffer_tank2_tables():
    ref2 = os.path.join(os.path.dirname(REF), "buffer_tank2.py")
    tree = ast.parse(open(ref2).read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "buffer_tank"][0]
    node = [n for n in fn.body if isinstance(n, ast.If)][0]
    assert ast.unparse(node.test) == "switch == 0" and not isinstance(node.orelse[0], ast.If)
    arr = {}
    for n in node.orelse:                                  # the `else` branch = switch 1
        if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Name) and isinstance(n.value, ast.Call) \
                and getattr(n.value.func, "attr", "") == "array":
            arr[n.targets[0].id] = np.array(ast.literal_eval(n.value.args[0]), dtype=float)
    mean = np.stack([arr[name + "_m"] for name in ORDER])
    std = np.stack([arr[name + "_s"] for name in ORDER])
    assert mean.shape == (14, 96) and std.shape == (14, 96)
    return mean, std


def main():
    tree = ast.parse(open(REF).read())
    fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "buffer_tank"][0]
    node = [n for n in fn.body if isinstance(n, ast.If)][0]
    branches = {}
    while True:
        k = node.test.comparators[0].value
        branches[k] = node.body
        if len(node.orelse) == 1 and isinstance(node.orelse[0], ast.If):
            node = node.orelse[0]
        else:
            break
    assert sorted(branches) == list(range(8)), sorted(branches)
    mean = np.zeros((8, 14, 48))
    frac = np.zeros((8, 14))
    draws = np.zeros(8, dtype=np.int64)
    for k in range(8):
        m, s, d = branch_tables(branches[k])
        for j, name in enumerate(ORDER):
            assert m[name].shape == (48,), (k, name, m[name].shape)
            mean[k, j] = m[name]
            frac[k, j] = s[name]
        draws[k] = d
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    bt2_mean, bt2_std = buffer_tank2_tables()
    np.savez(OUT, mean=mean, std_frac=frac, draws=draws, order=np.array(ORDER), bt2_mean=bt2_mean, bt2_std=bt2_std)
    print("wrote", OUT, "draws", draws.tolist(), "std fractions", frac[0].tolist())


if __name__ == "__main__":
    sys.exit(main())
