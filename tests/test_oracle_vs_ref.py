"""Pins the plain-C restatement (oracle/hmrt_oracle.c) to the reference's OWN code compiled for
the host (oracle/_ref, built from /root/reference by oracle/build_ref.sh): colours, castRay's
final ray position and the mirror/hit flags must be bit-identical."""
import numpy as np
import pytest

import oraclelib as ol

pytestmark = pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built (reference tree absent)")


@pytest.mark.parametrize("name", list(ol.SCENES))
@pytest.mark.parametrize("mode", ["ramp", "shadow", "colormap+shadow"])
def test_restatement_equals_reference(name, mode):
    sc = ol.scene(name)
    W, H = 160, 120
    for cam in ol.cameras_for(sc, 6):
        opts = ol.make_opts(sc["max_height"], use_color_map="colormap" in mode, shadows="shadow" in mode)
        a = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, sc["pyramid"], sc["color_map"], sc["coarse"], sc["levels"], W, H, cam, opts)
        b = ol.cpu_trace(ol.ref().hmrt_ref_trace, sc["pyramid"], sc["color_map"], sc["coarse"], sc["levels"], W, H, cam, opts)
        ol.assert_same_trace(a, b, f"{name}/{mode}")


def test_reference_launcher_is_float_math():
    assert ol.ref().hmrt_ref_float_math() == 1
    assert ol.ref_dpow().hmrt_ref_float_math() == 0


def test_double_pow_variant_agreement():
    """The Linux double-pow/floor meaning of the reference differs from the
    canonical fp32 meaning only in the rounding of tX/tZ: report-level check, >= 99.9 % hit cells."""
    sc = ol.scene("r1024_l8")
    W, H = 320, 240
    cam = ol.cameras_for(sc, 2)[1]
    opts = ol.make_opts(sc["max_height"])
    _, ha = ol.cpu_trace(ol.ref().hmrt_ref_trace, sc["pyramid"], None, sc["coarse"], sc["levels"], W, H, cam, opts)
    _, hb = ol.cpu_trace(ol.ref_dpow().hmrt_ref_trace, sc["pyramid"], None, sc["coarse"], sc["levels"], W, H, cam, opts)
    ca, _ = ol.hit_cells(ha, sc["r0"])
    cb, _ = ol.hit_cells(hb, sc["r0"])
    assert (ca == cb).mean() >= 0.999


def test_max_height_termination():
    """max_height takes part in ray termination (CudaKernel.cu:153), not only in colouring."""
    sc = ol.scene("r512_l4")
    cam = ol.make_camera((256.3, 0.5 * sc["max_height"], 255.1), (0.3, 0.25, 1.0))
    for mh in (sc["max_height"], 0.6 * sc["max_height"]):
        opts = ol.make_opts(mh)
        a = ol.cpu_trace(ol.oracle().hmrt_oracle_trace, sc["pyramid"], None, sc["coarse"], sc["levels"], 96, 64, cam, opts)
        b = ol.cpu_trace(ol.ref().hmrt_ref_trace, sc["pyramid"], None, sc["coarse"], sc["levels"], 96, 64, cam, opts)
        ol.assert_same_trace(a, b, f"max_height={mh}")
