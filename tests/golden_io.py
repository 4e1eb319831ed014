"""Helpers to read the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def graph_case_names(z=None):
    z = z or load("graph_cases.npz")
    return sorted({k.split("/")[0] for k in z.files})


def graph_case(z, name):
    mask = z[f"{name}/mask"]
    unc = z[f"{name}/unc"]
    c = dict(depth=z[f"{name}/depth"], mask=None if mask.size == 0 else mask,
             unc=None if unc.size == 0 else unc, res=tuple(float(v) for v in z[f"{name}/res"]),
             conn=str(z[f"{name}/conn"]))
    ref = {k.split("/ref/")[1]: z[k] for k in z.files if k.startswith(f"{name}/ref/")}
    return c, ref


def model_case_keys(z=None):
    z = z or load("model_cases.npz")
    return sorted({k.split("/")[0] for k in z.files if "/" in k})
