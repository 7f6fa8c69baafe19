"""CPU: the restated oracle reproduces every golden trace recorded from the real reference."""
import pytest

from oracle import restated as R
from tests import helpers


@pytest.mark.parametrize("name", helpers.golden_names())
def test_oracle_matches_reference_trace(name):
    fx = helpers.load_golden(name)
    clip = helpers.golden_clip(fx)
    c = fx["clip"]
    so = R.StreamOracle(c["W"], c["H"], **fx["kwargs"])
    for key in ("gaussian", "min_area", "max_area", "cache_frames", "min_movement_frames", "scale"):
        assert so.p[key] == fx["params"][key], key
    assert len(fx["trace"]) == c["n"]
    for t, gold in enumerate(fx["trace"]):
        rec = so.process(clip[t], keep_planes=True)
        helpers.check_record(name, t, rec, gold, rec["planes"])
    assert so.dec.wrote_frames == fx["result"][0]
