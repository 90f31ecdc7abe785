"""Build-time guard on the hot kernels' resource usage (no GPU needed): the register counts fix how many CTAs an SM
holds, and a growing stack frame means spills in the pixel loops -- both were tuned by measurement, so
a change that silently moves them should fail here rather than show up as a slower bench.  Also checks that every
kernel carries sm_100a code only."""
import re
import shutil
import subprocess

import pytest

from lasr_b200 import _build

# kernel (mangled-name fragment) -> (max registers, max stack bytes)
BUDGET = {
    "raster_forward_kernelILb1ELi1E": (64, 32),          # LASR configuration, one colour set: 4 CTAs of 8 warps per SM
    "raster_forward_kernelILb1ELi3E": (80, 96),          # three colour sets: 3 CTAs per SM
    "raster_backward_kernelILb1ELb0ELi1ELi0E": (128, 0),  # 4 CTAs of 4 warps per SM, no spills
    "raster_backward_kernelILb1ELb0ELi3ELi1E": (128, 64),  # the hot path's instantiation (two live sets)
    "raster_backward_kernelILb1ELb0ELi3ELi0E": (168, 0),  # three live sets: 3 CTAs per SM, no spills
    "lbs_forward_kernel": (64, 0),
    "stage_resample_kernel": (96, 0),
    "chamfer_nn_kernelILi1E": (48, 0),
}


@pytest.fixture(scope="module")
def usage():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = _build.build_library()
    txt = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for name, reg, stack in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", txt):
        out[name] = (int(reg), int(stack))
    assert out, "no kernels found in the library"
    return txt, out


def test_hot_kernels_stay_within_their_register_and_stack_budgets(usage):
    _, out = usage
    for frag, (max_reg, max_stack) in BUDGET.items():
        hits = {n: v for n, v in out.items() if frag in n}
        assert hits, f"kernel {frag} not found in liblasr_b200.so"
        for n, (reg, stack) in hits.items():
            assert reg <= max_reg, f"{n}: {reg} registers > {max_reg}"
            assert stack <= max_stack, f"{n}: {stack} bytes of stack > {max_stack}"


def test_library_is_sm_100a_only(usage):
    txt, _ = usage
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    assert archs == {"sm_100a"}, archs


def test_forward_kernel_stages_its_records_with_the_tma_unit():
    """The LASR-configuration forward kernels stage a tile's face records and boxes in shared memory with 1-D bulk
    copies completing on mbarriers (cp.async.bulk ... mbarrier::complete_tx -> UBLKCP, mbarrier arrive / try_wait ->
    SYNCS): the async-proxy path of Hopper / Blackwell, visible in the SASS of the product library."""
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib = _build.build_library()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per_kernel = re.split(r"\n\s*Function : ", sass)
    fwd = [k for k in per_kernel if k.startswith("_ZN4lasr21raster_forward_kernelILb1E")]
    assert len(fwd) == 3                                            # K = 1, 2, 3
    for k in fwd:
        assert "UBLKCP" in k and "SYNCS.ARRIVE.TRANS64" in k and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in k, k[:80]
    generic = [k for k in per_kernel if k.startswith("_ZN4lasr21raster_forward_kernelILb0E")]
    assert generic and "UBLKCP" not in generic[0]                   # the reference-faithful generic modes keep plain loads
