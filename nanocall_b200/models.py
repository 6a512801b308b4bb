"""Builtin pore-model tables (data extracted from the reference by
tools/extract_builtin_models.py; names/strands as Builtin_Model.cpp:1-19,
src/builtin_models/builtin_model_names.inl, builtin_model_strands.inl)."""
import json
import os
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_builtin_models():
    """-> list of dict(name, strand, table float32[4096,4] = level_mean, level_stdv, sd_mean, sd_stdv)."""
    with open(os.path.join(_DATA, "builtin_models.json")) as f:
        meta = json.load(f)
    raw = np.fromfile(os.path.join(_DATA, "builtin_models.bin"), dtype="<f4")
    tables = raw.reshape(len(meta["models"]), meta["n_states"], 4)
    return [dict(name=m["name"], strand=m["strand"], table=np.ascontiguousarray(tables[i]))
            for i, m in enumerate(meta["models"])]


def builtin_model(name):
    """Look a model up by its full name or by the prefix the CLI uses (e.g. 'r73.t')."""
    for m in load_builtin_models():
        if m["name"] == name or m["name"].startswith(name):
            return m
    raise KeyError(name)
