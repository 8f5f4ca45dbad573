"""Static checks on the compiled device code (cuobjdump, no GPU): the pipelined step kernel really is a bulk-copy /
mbarrier pipeline, and its row loop has not silently grown (a build flag once cost 17 % this way)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "giggles_b200", "libgiggles_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="cuobjdump or library missing")


def _functions():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=300).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name, body = f.split("\n", 1)
        out[name.strip()] = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    return out


@pytest.fixture(scope="module")
def functions():
    return _functions()


def _row_loop(lines):
    """Length (instructions) of the innermost loop that stores four 128-bit rows (the 4x unrolled row loop)."""
    best = None
    for i, l in enumerate(lines):
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?BRA\s+0x([0-9a-f]+)", l)
        if m and int(m.group(2), 16) < int(m.group(1), 16):
            n = (int(m.group(1), 16) - int(m.group(2), 16)) // 16
            if sum("STG.E.128" in x for x in lines[i - n:i + 1]) == 4:
                best = n if best is None else min(best, n)
    return best


def test_compiled_for_sm_100a_only():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs


def test_pipelined_kernel_uses_bulk_copies_and_mbarriers(functions):
    names = [n for n in functions if "step_tma_kernel" in n]
    assert len(names) >= 20
    for n in names:
        text = "\n".join(functions[n])
        assert "UBLKCP" in text, n            # cp.async.bulk global -> shared
        assert "SYNCS" in text, n             # mbarrier arrive / try_wait
        assert "LDS.128" in text or "LDS.64" in text, n


def test_headline_row_loops_stay_small(functions):
    # step_tma_kernel<FWD, G=32, NS=1, MT=false, CW=8, TWO=false, V=4>: the R = 88 bench kernels
    fwd = next(n for n in functions if re.search(r"step_tma_kernelILb1ELi32ELi1ELb0ELi8ELb0ELi4E", n))
    bwd = next(n for n in functions if re.search(r"step_tma_kernelILb0ELi32ELi1ELb0ELi8ELb0ELi4E", n))
    assert _row_loop(functions[fwd]) <= 240      # 4 rows of 88 values: ~55 instructions per row (class-segmented loop)
    assert _row_loop(functions[bwd]) <= 250      # ~60 per row


def test_no_local_memory_spills_in_headline_kernels(functions):
    for pat in (r"step_tma_kernelILb1ELi32ELi1ELb0ELi8ELb0ELi4E", r"step_tma_kernelILb0ELi32ELi1ELb0ELi8ELb0ELi4E"):
        n = next(x for x in functions if re.search(pat, x))
        text = "\n".join(functions[n])
        assert "STL" not in text and "LDL" not in text, n


def test_wide_panel_kernel_is_a_bulk_copy_pipeline_without_spills(functions):
    # wide_step_kernel<FWD, TWO> (hmm_wide_kernel.cuh): the R > 128 path; one float4 of state per lane, so it must not spill
    names = [n for n in functions if "wide_step_kernel" in n]
    assert len(names) == 4
    for n in names:
        text = "\n".join(functions[n])
        assert "UBLKCP" in text and "SYNCS" in text and "LDS.128" in text and "STG.E.128" in text, n
        assert "STL" not in text and "LDL" not in text, n


def test_small_panel_pair_kernel_does_not_spill_at_ten_haplotypes(functions):
    # pair_wave_kernel<10, FWD>: configs[0]'s panel
    names = [n for n in functions if re.search(r"pair_wave_kernelILi10E", n)]
    assert len(names) == 2
    for n in names:
        text = "\n".join(functions[n])
        assert "STL" not in text and "LDL" not in text, n
        # the step arguments live in shared memory, so the column pointers are generic: LD.E.128 / ST.E.128
        assert re.search(r"\bLDG?\.E\.128", text) and re.search(r"\bSTG?\.E\.128", text) and "SHFL.BFLY" in text, n


def test_column_owning_kernel_for_mid_size_panels(functions):
    # colown_step_kernel<FWD, TWO> (hmm_colown_kernel.cuh): the backward step of configs[1]'s 88-haplotype panel by default
    names = [n for n in functions if "colown_step_kernel" in n]
    assert len(names) == 4
    for n in names:
        text = "\n".join(functions[n])
        assert "UBLKCP" in text and "SYNCS" in text and "LDS.128" in text and "STG.E.128" in text, n
        assert "STL" not in text and "LDL" not in text, n
