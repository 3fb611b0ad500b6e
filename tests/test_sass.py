"""Static checks on the compiled sm_100a code (cuobjdump; no GPU needed).

* the library really carries sm_100a SASS for every kernel the design names;
* the production walk uses the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2) and the round-down floor;
* ptxas did NOT contract a packed multiply into a following packed add: every walk step has exactly one
  FFMA2.RM (floor) and three other FFMA2 (boundary, remainder, quotient) -- an extra FFMA2 is how the
  mul.rn.f32x2 + add.rn.f32x2 contraction showed up in an earlier build;
* there is no double-precision arithmetic and no IEEE-divide slow path inside the air loop of the walk.
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / "gpu-heightmap-raytracer_b200" / "csrc" / "libhmrt.so"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not LIB.exists(), reason="needs cuobjdump and a built libhmrt.so")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].split("/*")[0].strip())
    assert "sm_100a" in out or "sm_100" in out
    return funcs


def test_every_kernel_is_present(sass):
    names = " ".join(sass)
    for k in ["trace_persistent_kernel", "top_level_max_kernel", "scatter_las_kernel", "scatter_xyz_kernel", "rx_bin_kernel",
              "rx_apply_kernel", "rx_gather_mips_kernel", "rx_barrier_kernel", "build_mips_fused_kernel", "build_mip_level_kernel", "resolve_colors_kernel", "locality_probe_kernel",
              "compose_window_kernel", "compose_colors_kernel"]:
        assert k in names, f"{k} missing from libhmrt.so"


def test_packed_walk_is_not_contracted(sass):
    prod = {n: ins for n, ins in sass.items() if "trace_persistent_kernel" in n and ("ELi1E" in n or "ELi2E" in n)}
    assert len(prod) == 12  # {hits, no hits, no hits + segment notify (hmrt_trace_host)} x {pow2, generic} x {tile-granular tail, none}
    for name, ins in prod.items():
        floor_rm = sum(i.startswith("FFMA2.RM") for i in ins)
        ffma2 = sum(i.startswith("FFMA2 ") for i in ins)
        assert floor_rm >= 3, name                      # two air half-steps + the general loop (x2 with shadows)
        assert ffma2 == 3 * floor_rm, f"{name}: {ffma2} FFMA2 for {floor_rm} walk steps -- a packed mul/add pair was contracted"
        assert any(i.startswith("FMUL2") for i in ins) and any(i.startswith("FADD2") for i in ins)
        assert not any(re.match(r"D(FMA|ADD|MUL)\b", i) for i in ins), f"{name}: fp64 arithmetic on the traced path"
    # the air loop of the production kernel: four steps per trip, no register copies, no constant-bank loads
    for name, ins in prod.items():
        if "ILb0E" not in name or name.endswith("Lb1EEEvNS_11TraceParamsE"):
            continue  # the instrumented (hits) variants also count iterations inside the loop; the notify variants are the host path
        rm = [k for k, i in enumerate(ins) if i.startswith("FFMA2.RM")]
        end = next(k for k in range(rm[3], len(ins)) if "BRA" in ins[k])
        body = [re.sub(r"^@!?U?P\d+\s+", "", i) for i in ins[rm[0]:end + 1]]
        assert sum(i.startswith("FFMA2") for i in body) == 16, name
        assert not any(re.match(r"(MOV|IMAD\.MOV|LDC|LDCU)\b", i) for i in body), f"{name}: copies / constant loads inside the air loop"
        assert len(body) <= 66, f"{name}: {len(body)} instructions per four air steps"


def test_mip_kernel_uses_wide_loads_and_shuffles(sass):
    ins = next(v for n, v in sass.items() if "build_mips_fused_kernel" in n)
    assert sum("LDG.E.128" in i for i in ins) == 8
    assert any(i.startswith("SHFL") for i in ins)


def test_trace_stores_are_128_bit(sass):
    ins = next(v for n, v in sass.items() if "trace_persistent_kernelILb0ELi2E" in n)
    assert any(i.startswith("STG.E.128") for i in ins)


def test_bin_kernel_streams_through_tma(sass):
    """The tile-binning pass stages its records with TMA bulk copies signalled on an mbarrier (csrc/rasterx.cu)."""
    # the usual instantiation: no colour keys, aligned 20-byte records, power-of-two cell size, 8 records per thread
    ins = next(v for n, v in sass.items() if "rx_bin_kernelILb0ELb1ELb1ELi8E" in n)
    assert any(i.startswith("UBLKCP") for i in ins), "no bulk-copy (TMA) instruction in rx_bin_kernel"
    assert any(i.startswith("SYNCS") for i in ins), "no mbarrier instruction in rx_bin_kernel"
    assert any(i.startswith("ATOMS") for i in ins)        # the per-tile histogram lives in shared memory
    assert not any(i.startswith("ATOMG") for i in ins)    # no global atomics with a return value on the binning path
    # every shared-memory access is an LDS / STS (no generic LD / ST of the sorted pairs); pairs leave as 64-bit stores
    assert not any(re.search(r"(^|\s)(LD|ST)\.E", i) for i in ins), "generic load/store in rx_bin_kernel"
    assert any("STS.64" in i for i in ins) and any("STG.E.64" in i for i in ins)
    # 24 instantiations: keys x record alignment x cell-size form x step size
    assert sum("rx_bin_kernel" in n for n in sass) == 24


def test_tolerance_walk_has_no_air_loop(sass):
    """Variant 2 replaces the air loop by one closed-form step: its FFMA2.RM count drops to the general loop's."""
    exact = next(v for n, v in sass.items() if "trace_persistent_kernelILb0ELi2ELb0E" in n)
    jump = next(v for n, v in sass.items() if "trace_persistent_kernelILb0ELi3ELb0E" in n)
    assert sum(i.startswith("FFMA2.RM") for i in jump) < sum(i.startswith("FFMA2.RM") for i in exact)
