"""Golden vectors generated from the reference's own code (tests/golden/make_golden.py) replayed
against the plain-C oracle and the host instance of the kernel arithmetic."""
import pytest

import oraclelib as ol


@pytest.fixture(scope="module")
def golden():
    return ol.load_golden()


@pytest.mark.parametrize("mode", list(ol.GOLDEN_MODES))
@pytest.mark.parametrize("impl", ["oracle", "hostsim"])
def test_cpu_implementations_reproduce_golden(golden, mode, impl):
    g = golden
    fn = ol.oracle().hmrt_oracle_trace if impl == "oracle" else ol.hostsim().hostsim_trace
    for i, cam in enumerate(g["cams"]):
        got = ol.cpu_trace(fn, g["pyramid"], g["color_map"], g["coarse"], int(g["levels"]), int(g["W"]), int(g["H"]), cam,
                           ol.golden_opts(g, mode))
        ol.assert_same_trace(got, ol.golden_expected(g, mode, i), f"golden {mode}/{i} vs {impl}")


def test_golden_is_not_trivial(golden):
    g = golden
    hit_frac = [(g[f"hits_ramp_{i}"][..., 3] & 1).mean() for i in range(len(g["cams"]))]
    assert max(hit_frac) > 0.9 and min(hit_frac) < 0.7
    assert any((g[f"hits_shadow_{i}"][..., 3] & 8).any() for i in range(len(g["cams"])))
