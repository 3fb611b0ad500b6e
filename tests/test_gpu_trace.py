"""Parity of the CUDA ray traversal (through the C ABI) against the CPU oracle: bit-exact colours,
castRay end positions, flags and step counts; golden vectors from the reference's own code;
sharding / batching / store-path properties; full-size (4K, 16384^2) properties + sampled rows."""
import ctypes as C

import numpy as np
import pytest
import torch

import oraclelib as ol

pytestmark = pytest.mark.gpu


def _oracle(sc, W, H, cam, opts, rows=None):
    return ol.cpu_trace(ol.oracle().hmrt_oracle_trace, sc["pyramid"], sc.get("color_map"), sc["coarse"], sc["levels"], W, H,
                        cam, opts, rows=rows)


@pytest.mark.parametrize("name", list(ol.SCENES))
@pytest.mark.parametrize("mode", ["ramp", "shadow", "colormap+shadow"])
def test_cuda_equals_oracle(cuda_ctx, name, mode):
    import gpulib

    sc = ol.scene(name, seed=5)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 320, 200
    cams = ol.cameras_for(sc, 6)
    opts = ol.make_opts(sc["max_height"], use_color_map="colormap" in mode, shadows="shadow" in mode)
    rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)  # six frames, one launch
    for i, cam in enumerate(cams):
        ol.assert_same_trace((rgb[i], hits[i]), _oracle(sc, W, H, cam, opts), f"{name}/{mode}/cam{i}")


@pytest.mark.parametrize("mode", list(ol.GOLDEN_MODES))
def test_cuda_reproduces_reference_golden(cuda_ctx, mode):
    """tests/golden/ray_golden.npz was produced by the reference's own castRay (oracle/_ref)."""
    import gpulib

    g = ol.load_golden()
    sc = dict(pyramid=g["pyramid"], color_map=g["color_map"], coarse=g["coarse"], levels=int(g["levels"]),
              max_height=float(g["max_height"]))
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    rgb, hits = gpulib.gpu_trace(cuda_ctx, int(g["W"]), int(g["H"]), g["cams"], ol.golden_opts(g, mode))
    for i in range(len(g["cams"])):
        ol.assert_same_trace((rgb[i], hits[i]), ol.golden_expected(g, mode, i), f"golden {mode}/{i}")


def test_reference_acceptance_bars(cuda_ctx):
    """BASELINE.json bars, stated explicitly (they are implied by bit-exactness): hit cell >= 99.9 %,
    hit distance within 1e-4 relative on matching cells, colour within 1/255."""
    import gpulib

    sc = ol.scene("r1024_l8", seed=9)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 640, 480
    cam = ol.cameras_for(sc, 2)[1]
    opts = ol.make_opts(sc["max_height"])
    rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts)
    ergb, ehits = _oracle(sc, W, H, cam, opts)
    cg, hg = ol.hit_cells(hits[0], sc["r0"])
    ce, he = ol.hit_cells(ehits, sc["r0"])
    assert (cg == ce).mean() >= 0.999
    both = hg & he & (cg == ce)
    org = np.array(list(cam.position), np.float64)

    def dist(h):
        x = np.where((h["flags"] & 2) != 0, sc["r0"] - h["x"].astype(np.float64), h["x"])
        z = np.where((h["flags"] & 4) != 0, sc["r0"] - h["z"].astype(np.float64), h["z"])
        return np.sqrt((x - org[0]) ** 2 + (h["y"] - org[1]) ** 2 + (z - org[2]) ** 2)

    dg, de = dist(hits[0])[both], dist(ehits)[both]
    assert (np.abs(dg - de) <= 1e-4 * de).all()  # tolerance from BASELINE.json's north star
    assert (np.abs(rgb[0].astype(int) - ergb.astype(int)) <= 1).all()
    assert (rgb[0] != ergb).any(axis=-1).mean() <= 0.05  # >= 95 % pixel-exact


@pytest.mark.parametrize("W,H", [(333, 97), (64, 48), (1920, 1080)])
def test_store_paths_and_ragged_frames(cuda_ctx, W, H):
    """W % 16 != 0 takes the byte-store path, W % 16 == 0 the 128-bit path; H % 8 != 0 leaves a ragged tile."""
    import gpulib

    sc = ol.scene("r512_l4", seed=2)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    cam = ol.cameras_for(sc, 2)[1]
    opts = ol.make_opts(sc["max_height"])
    rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts)
    ol.assert_same_trace((rgb[0], hits[0]), _oracle(sc, W, H, cam, opts), f"{W}x{H}")
    rgb2, _ = gpulib.gpu_trace(cuda_ctx, W, H, [cam], opts, hits=False)  # non-instrumented kernel
    assert (rgb2 == rgb).all()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_tile_sharding_equals_whole_frame(cuda_ctx, world):
    import gpulib
    from hmrt import dist as hd

    sc = ol.scene("r512_l4", seed=2)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 256, 203
    cams = ol.cameras_for(sc, 3)
    whole, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, ol.make_opts(sc["max_height"], shadows=True), hits=False)
    parts = []
    for r in range(world):
        o = ol.make_opts(sc["max_height"], shadows=True, tile_first=r, tile_stride=world)
        p, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, o, hits=False)
        parts.append(torch.from_numpy(p))
    for f in range(len(cams)):
        got = hd.assemble_frame([p[f] for p in parts], H, W).numpy()
        assert (got == whole[f]).all()


def test_trace_host_matches_device_and_reference_call_order(cuda_ctx):
    import gpulib
    import hmrt

    sc = ol.scene("r512_l4", seed=2)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 640, 360
    cams = ol.cameras_for(sc, 4)
    opts = ol.make_opts(sc["max_height"])
    dev, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts, hits=False)
    host = torch.empty((len(cams), H, W, 3), dtype=torch.uint8).pin_memory()
    cuda_ctx.trace_host(W, H, cams, opts, host)
    assert (host.numpy() == dev).all()
    # freeDeviceVariables then rayTrace must fail loudly, not render garbage
    cuda_ctx.clear_heightmap()
    with pytest.raises(hmrt.HmrtError) as e:
        cuda_ctx.trace(W, H, cams, opts)
    assert e.value.code == -2


@pytest.fixture(params=[2, 1, 0], ids=["streamed", "per-group", "auto"])
def host_variant(request, cuda_ctx):
    """The schedules of hmrt_trace_host: one launch + per-segment stream waits, one launch per frame group, and the default mix."""
    cuda_ctx.set_host_variant(request.param)
    yield request.param
    cuda_ctx.set_host_variant(0)


@pytest.mark.parametrize("W,H,n_frames,stride", [(3840, 2160, 1, 1), (3840, 2157, 3, 1), (3840, 2160, 5, 2), (2560, 1083, 2, 1), (200, 43, 3, 3)])
def test_trace_host_large_frames_split_last_launch(cuda_ctx, host_variant, W, H, n_frames, stride):
    """hmrt_trace_host, streamed: one launch whose row segments release their device->host copies as they complete;
    per-group: launches of >= 4 M rays, the LAST single-frame launch cut into four tile ranges with their own copies.
    Either way the host buffer must equal the device rendering of hmrt_trace byte for byte -- whole frames, ragged last
    row tile (H % 8 != 0), row-tile sharding, frames with fewer strips than segments."""
    import gpulib

    sc = ol.scene("r1024_l8", seed=3)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    cams = ol.cameras_for(sc, n_frames)
    for first in range(stride):
        opts = ol.make_opts(sc["max_height"], shadows=True, tile_first=first, tile_stride=stride)
        dev, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts, hits=False)
        host = torch.full(dev.shape, 77, dtype=torch.uint8).pin_memory()
        cuda_ctx.trace_host(W, H, cams, opts, host)
        assert (host.numpy() == dev).all()


def test_trace_host_begin_wait_two_calls_in_flight(cuda_ctx, host_variant):
    """hmrt_trace_host_begin / _wait: two calls in flight into two host buffers (different cameras, different sizes) deliver what
    the synchronous call delivers; the state errors; a synchronous call collects calls still in flight."""
    import gpulib
    import hmrt

    sc = ol.scene("r1024_l8", seed=5)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    opts = ol.make_opts(sc["max_height"], shadows=True)
    W, H = 1920, 1083
    batches = [ol.cameras_for(sc, 1), ol.cameras_for(sc, 3), ol.cameras_for(sc, 5)[2:], ol.cameras_for(sc, 4)[1:]]  # the second call grows the framebuffers
    want = [gpulib.gpu_trace(cuda_ctx, W, H, cams, opts, hits=False)[0] for cams in batches]
    with pytest.raises(hmrt.HmrtError) as e:
        cuda_ctx.trace_host_wait()
    assert e.value.code == -2  # nothing in flight
    hosts = [torch.full(w.shape, 99, dtype=torch.uint8).pin_memory() for w in want]
    for rounds in range(2):
        for h in hosts:
            h.fill_(99)
        cuda_ctx.trace_host_begin(W, H, batches[0], opts, hosts[0])
        for k in range(1, len(batches)):
            cuda_ctx.trace_host_begin(W, H, batches[k], opts, hosts[k])
            if k == 1:
                with pytest.raises(hmrt.HmrtError) as e:
                    cuda_ctx.trace_host_begin(W, H, batches[k], opts, hosts[k])
                assert e.value.code == -2  # two already in flight
            cuda_ctx.trace_host_wait()
            assert (hosts[k - 1].numpy() == want[k - 1]).all(), (rounds, k)
        cuda_ctx.trace_host_wait()
        assert (hosts[-1].numpy() == want[-1]).all()
    # a synchronous call while one is in flight: both are complete when it returns
    hosts[0].fill_(1)
    hosts[1].fill_(1)
    cuda_ctx.trace_host_begin(W, H, batches[0], opts, hosts[0])
    cuda_ctx.trace_host(W, H, batches[1], opts, hosts[1])
    assert (hosts[0].numpy() == want[0]).all() and (hosts[1].numpy() == want[1]).all()
    with pytest.raises(hmrt.HmrtError):
        cuda_ctx.trace_host_wait()


def test_many_frames_in_one_call_equal_single_frame_calls(cuda_ctx, host_variant):
    """More frames than fit in the kernel parameters (48) travel through the device-side frame table; device and host
    output of a 70-frame call must equal 70 single-frame calls."""
    import gpulib

    sc = ol.scene("r512_l4", seed=6)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H, n = 96, 64, 70
    base = ol.cameras_for(sc, 6)
    cams = []
    for i in range(n):  # 70 distinct cameras: the scene cameras, shifted
        c = base[i % len(base)]
        cams.append(ol.make_camera((c.position[0] + 1.5 * (i // 7), c.position[1] + 0.25 * i, c.position[2] - 0.75 * (i // 7)), tuple(c.forward)))
    opts = ol.make_opts(sc["max_height"], shadows=True)
    many, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts, hits=False)
    for i in (0, 1, 47, 48, 49, 69):
        one, _ = gpulib.gpu_trace(cuda_ctx, W, H, [cams[i]], opts, hits=False)
        assert (many[i] == one[0]).all(), i
    host = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
    cuda_ctx.trace_host(W, H, cams, opts, host)
    assert (host.numpy() == many).all()


def test_error_codes(cuda_ctx):
    import gpulib
    import hmrt

    sc = ol.scene("r256_l1")
    sc["color_map"] = None
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    cam = ol.cameras_for(sc, 1)[0]
    with pytest.raises(hmrt.HmrtError):  # colour-map mode without a colour map
        cuda_ctx.trace(64, 48, [cam], ol.make_opts(sc["max_height"], use_color_map=True))
    with pytest.raises(hmrt.HmrtError):  # W < 2: (W - 1) divides in viewToGridSpace
        cuda_ctx.trace(1, 48, [cam], ol.make_opts(sc["max_height"]))
    # a tile range past the frame renders nothing and is not an error
    out, _ = cuda_ctx.trace(64, 48, [cam], ol.make_opts(sc["max_height"], tile_first=6, tile_stride=8))
    assert out.shape[1] == 0


def test_production_kernel_equals_operation_by_operation_walk(cuda_ctx):
    """hmrt_set_trace_variant(1) runs the walk that mirrors CudaKernel.cu:121-177 line by line (IEEE divides,
    floorf, level tables); the production kernel (FADD.RM floor, 3-op division proven by tests/cuda/divcheck.cu,
    carried level state) must agree with it bit for bit, including degenerate axis-aligned rays."""
    import gpulib

    sc = ol.scene("r1024_l8", seed=4)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 641, 360  # odd width: the centre column has dir.x == 0 for an axis-aligned camera
    cams = ol.cameras_for(sc, 6) + [ol.make_camera((512.0, 1.4 * sc["max_height"], 512.0), (0.0, -0.4, 1.0)),
                                    ol.make_camera((512.5, 1.4 * sc["max_height"], 300.0), (1.0, -0.3, 0.0)),
                                    ol.make_camera((512.5, 3.0 * sc["max_height"], 512.5), (0.0, -1.0, 1e-6))]
    opts = ol.make_opts(sc["max_height"], shadows=True, use_color_map=True)
    try:
        cuda_ctx.set_trace_variant(1)
        slow = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)
    finally:
        cuda_ctx.set_trace_variant(0)
    fast = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)
    for i in range(len(cams)):
        ol.assert_same_trace((fast[0][i], fast[1][i]), (slow[0][i], slow[1][i]), f"variant cam{i}")
    # and both equal the oracle on the degenerate cameras
    for i in (6, 7, 8):
        ol.assert_same_trace((fast[0][i], fast[1][i]), _oracle(sc, W, H, cams[i], opts), f"degenerate cam{i}")


def test_full_size_4k_over_16384(cuda_ctx):
    """BASELINE config 3 at full size: 3840x2160 over a 16384^2 map (1.43 GB pyramid, built on the GPU by
    the product's own mip kernel).  Size-independent properties + sampled rows against the oracle."""
    import gpulib

    r0, levels = 16384, 8
    coarse = r0 >> (levels - 1)
    res, idx, total = ol.pyramid_layout(coarse, levels)
    g = torch.Generator(device="cuda").manual_seed(1)
    pyr = torch.zeros(total, dtype=torch.float32, device="cuda")
    fin = pyr[idx[0]:].view(r0, r0)
    xs = torch.arange(r0, device="cuda", dtype=torch.float32)
    fin.copy_(200 + 120 * torch.sin(xs[None, :] * 0.0021) * torch.cos(xs[:, None] * 0.0017) + 40 * torch.sin(xs[None, :] * 0.013 + xs[:, None] * 0.011))
    fin.add_(torch.rand((r0, r0), device="cuda", generator=g) * 4).clamp_(min=0)
    cuda_ctx.build_mips(pyr, coarse, levels)
    torch.cuda.synchronize()
    mh = float(fin.max())
    cuda_ctx.set_heightmap(pyr, None, coarse, levels, mh)
    W, H = 3840, 2160
    cams = [ol.make_camera((r0 / 2 + 0.4, 2.0 * mh + 1500, r0 / 2 - 0.3), (0.6, -0.3, 0.8)),
            ol.make_camera((r0 * 0.2, 3000.0, r0 * 0.25), (0.7, -0.15, 0.7))]
    opts = ol.make_opts(mh, shadows=True)
    whole, hits = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)
    # (1) determinism + instrumented == plain kernel
    again, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts, hits=False)
    assert (again == whole).all()
    # (2) 8-way row-tile sharding reassembles to the same frames
    from hmrt import dist as hd
    parts = []
    for r in range(8):
        p, _ = gpulib.gpu_trace(cuda_ctx, W, H, cams, ol.make_opts(mh, shadows=True, tile_first=r, tile_stride=8), hits=False)
        parts.append(torch.from_numpy(p))
    for f in range(2):
        assert (hd.assemble_frame([p[f] for p in parts], H, W).numpy() == whole[f]).all()
    # (3) sampled rows against the CPU oracle at full size (bit-exact)
    host_pyr = pyr.cpu().numpy()
    sc = dict(pyramid=host_pyr, coarse=coarse, levels=levels)
    hit_frac = (hits["flags"] & 1).mean()
    assert 0.3 < hit_frac <= 1.0
    for f, cam in enumerate(cams):
        for r0_, r1_ in [(0, 4), (1077, 1083), (2156, 2160)]:
            ergb, ehits = _oracle(sc, W, H, cam, opts, rows=(r0_, r1_))
            ol.assert_same_trace((whole[f][r0_:r1_], hits[f][r0_:r1_]), (ergb[r0_:r1_], ehits[r0_:r1_]), f"4K cam{f} rows {r0_}")


def test_rcp_inrange_exhaustive():
    """The production walk's reciprocal (hmrt::rcp_rn_inrange: MUFU.RCP + one Newton step, no range test) has the bits of
    __frcp_rn for EVERY float with 2^-40 <= |d| <= 2, the range fast_div_ok admits (tests/cuda/rcpcheck.cu)."""
    import shutil
    import subprocess
    from pathlib import Path

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    here = Path(__file__).resolve().parent
    exe = here / "_build" / "rcpcheck"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if Path(nvcc).exists():
        (here / "_build").mkdir(exist_ok=True)
        subprocess.run([nvcc, "-O3", "-gencode", "arch=compute_100a,code=sm_100a", str(here / "cuda" / "rcpcheck.cu"), "-o", str(exe)],
                       check=True, cwd=here / "cuda")
    assert exe.exists(), "tests/_build/rcpcheck is missing and nvcc is not available to build it"
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches vs __frcp_rn: 0" in out.stdout


def test_height_ramp_outside_0_255_follows_the_host_build(cuda_ctx):
    """`max_height` BELOW the data (the reference's `-` key lowers it by 100 per press, main.cpp:839-844): the ramp of
    getHeightColorValue (CudaKernel.cu:38-56) leaves [0, 255] and the float -> unsigned char conversion decides.  The
    canonical meaning here is the reference's HOST build (truncate to int, keep the low byte: wrap-around), which is what
    oracle/_ref executes; the reference's own CUDA build would saturate instead -- a documented, deliberate choice
    (DESIGN.md section 3).  max_height also ends rising rays earlier (:153): both effects must match."""
    import gpulib

    sc = ol.scene("r512_l4", seed=8)
    keep = gpulib.upload_scene(cuda_ctx, sc)  # noqa: F841
    W, H = 256, 160
    cams = ol.cameras_for(sc, 4)
    lib = ol.ref() or ol.oracle()
    fn = lib.hmrt_ref_trace if ol.ref() is not None else lib.hmrt_oracle_trace
    for frac in (0.5, 0.3):
        opts = ol.make_opts(sc["max_height"] * frac, shadows=True)
        rgb, hits = gpulib.gpu_trace(cuda_ctx, W, H, cams, opts)
        wrapped = 0
        for i, cam in enumerate(cams):
            ergb, ehits = ol.cpu_trace(fn, sc["pyramid"], None, sc["coarse"], sc["levels"], W, H, cam, opts)
            ol.assert_same_trace((rgb[i], hits[i]), (ergb, ehits), f"max_height x {frac}, cam {i}")
            y = hits[i]["y"][(hits[i]["flags"] & 1) != 0]
            wrapped += int((y * 2 / (sc["max_height"] * frac) > 2.0).sum())
        assert wrapped > 100, "the case must actually drive the ramp out of range"
