"""Generates tests/golden/ray_golden.npz from the REFERENCE'S OWN code (oracle/_ref/libhmrt_ref.so,
i.e. /root/reference/GPUHeightmapRaytracer/src/CudaKernel.cu:1-286 compiled for the host by
oracle/build_ref.sh).  Run in the development container (the reference tree is not on the GPU
box):   python tests/golden/make_golden.py
The fixture stores the INPUT bytes (finest level, colour map, cameras) next to the reference's
outputs, so nothing depends on libm or numpy's RNG when it is replayed.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oraclelib as ol  # noqa: E402

R0, LEVELS, W, H = 256, 6, 64, 48


def main():
    ref = ol.ref()
    assert ref is not None and ref.hmrt_ref_float_math() == 1, "needs oracle/_ref (run oracle/build_ref.sh)"
    fin = ol.sines_terrain(R0, seed=7)
    fin[100:104, 60:120] += 25.0  # a ridge so that shadows and silhouettes appear at this small size
    cmap = np.random.default_rng(11).integers(0, 256, (R0, R0, 3), dtype=np.uint8)
    pyr = ol.pyramid_from_finest(fin, LEVELS)
    mh = float(fin.max())
    poses = [((30.37, mh + 15.0, 30.79), (1.0, -0.4, 1.0)), ((128.2, 1.3 * mh, 20.6), (0.0, -0.5, 1.0)),
             ((230.1, mh + 5.0, 128.2), (-1.0, -0.15, 0.1)), ((128.0, 2.0 * mh, 128.0), (0.1, -1.0, 0.1)),
             ((100.4, 45.0, 200.3), (0.4, 0.1, -1.0)), ((200.5, mh + 2.0, 220.5), (-0.8, -0.05, -0.6))]
    out = dict(finest=fin, color_map=cmap, levels=np.int32(LEVELS), W=np.int32(W), H=np.int32(H), max_height=np.float32(mh))
    cams = np.zeros((len(poses), 9), np.float32)
    for i, (pos, fwd) in enumerate(poses):
        cam = ol.make_camera(pos, fwd)
        cams[i] = list(cam.frame_dim) + list(cam.forward) + list(cam.position)
        for mode, (uc, sh) in {"ramp": (False, False), "shadow": (False, True), "cmap": (True, False)}.items():
            opts = ol.make_opts(mh, use_color_map=uc, shadows=sh)
            rgb, hits = ol.cpu_trace(ref.hmrt_ref_trace, pyr, cmap, R0 >> (LEVELS - 1), LEVELS, W, H, cam, opts)
            out[f"rgb_{mode}_{i}"] = rgb
            if mode != "cmap":  # the walk does not depend on the colouring mode
                out[f"hits_{mode}_{i}"] = hits.view(np.uint32).reshape(H, W, 4)
    out["cameras"] = cams
    out["light_dir"] = np.array(list(ol.make_opts(mh).light_dir), np.float32)
    np.savez_compressed(HERE / "ray_golden.npz", **out)
    print("wrote", HERE / "ray_golden.npz")


if __name__ == "__main__":
    main()
