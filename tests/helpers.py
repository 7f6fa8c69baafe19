"""Shared helpers for the parity tests (golden fixtures, hashing, clip regeneration)."""
import glob
import hashlib
import json
import os

import numpy as np

from find_motion_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        fx = json.load(f)
    kw = dict(fx["kwargs"])
    if kw.get("mask_areas") is not None:
        kw["mask_areas"] = [tuple(tuple(p) for p in a) for a in kw["mask_areas"]]
    fx["kwargs"] = kw
    return fx


def golden_clip(fx):
    c = fx["clip"]
    return synth.make_clip(c["W"], c["H"], c["n"], c["seed"], fps=c["fps"])


DECISION_KEYS = ("movement", "counter", "decay", "cache_len", "wrote", "n_flush")


def check_record(name, t, rec, gold, planes=None):
    """Compare one frame's record (oracle- or GPU-produced) with the golden trace entry."""
    where = f"{name} frame {t}"
    if planes is not None:
        for key in ("gray", "blur", "thresh", "bg"):
            if key in planes and planes[key] is not None:
                assert sha(planes[key]) == gold[key], f"{where}: plane {key} differs from the reference"
    assert [float(a) for a in rec["areas"]] == gold["areas"], f"{where}: contour areas"
    assert [list(b) for b in rec["boxes"]] == [list(b) for b in gold["boxes"]], f"{where}: bounding boxes"
    for key in DECISION_KEYS:
        assert rec[key] == gold[key], f"{where}: {key} {rec[key]} != {gold[key]}"
