"""CPU checks of two pieces of host-visible data the GPU path depends on: the coefficient table of the device exp
(csrc/device_common.cuh: VP_EXP_CONSTANTS) and the per-unit dependency scan of the build script."""
import math
import os
import re
from fractions import Fraction

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "varpro_b200", "csrc")


def _exp_constants():
    txt = open(os.path.join(CSRC, "device_common.cuh")).read()
    body = txt.split("#define VP_EXP_CONSTANTS", 1)[1].split("static __constant__", 1)[0]
    body = re.sub(r"/\*.*?\*/", "", body).replace("\\", " ")
    vals = [float(t) for t in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", body)]
    assert len(vals) == 14, vals
    return vals


def test_device_exp_table_reproduces_exp_to_rounding_level():
    """exp(a) = 2^k (1 + r + r^2 P9(r)), r = a - k ln2 (two-term ln 2): evaluated here in exact rational arithmetic with
    the table of the header, the scheme must agree with exp to a few 1e-17 over the reduced range and beyond -- i.e.
    the coefficients (generated with mpmath) were transcribed correctly; fp64 rounding adds <= 1 ulp on the device."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    c = _exp_constants()
    assert c[0] == 1.4426950408889634 and c[1] == 6755399441055744.0
    ln2 = Fraction(-c[2]) + Fraction(-c[3])
    assert abs(mp.mpf(ln2.numerator) / ln2.denominator - mp.log(2)) < mp.mpf("1e-32")
    worst = mp.mpf(0)
    for a in [x * 0.0137 for x in range(-3000, 3001)] + [-700.0 + 1e-9, 699.999, 1e-300, -1e-300]:
        k = round(a * c[0])  # rint(a / ln 2), as the shifter addition does
        r = Fraction(a) - k * ln2
        p = Fraction(c[4])
        for ci in c[5:]:
            p = p * r + Fraction(ci)
        p = (p * r + 1) * r + 1
        approx = mp.mpf(p.numerator) / p.denominator * mp.mpf(2) ** k
        worst = max(worst, abs(approx / mp.e ** mp.mpf(a) - 1))
    assert worst < mp.mpf("4e-17"), worst


def test_build_dependency_scan_follows_the_instantiation_parts():
    """build.py recompiles an object when one of ITS headers changes: inst.cu includes one kernel header per part."""
    from varpro_b200 import build as vb
    units = {os.path.basename(o): (src, flags) for o, src, flags in vb._units()}

    def deps(name):
        src, flags = units[name]
        part = next((f.split("=")[1] for f in flags if f.startswith("-DVP_INST_PART=")), None)
        return {os.path.basename(p) for p in vb._includes(src, part)}

    batch, queue, simt = deps("inst_f64_3_3_batch.o"), deps("inst_f64_3_2_queue1.o"), deps("inst_f64_3_2_simt.o")
    assert "batch_fit_kernel.cuh" in batch and "fit_queue_kernel.cuh" not in batch and "dmma_tile.cuh" not in batch
    assert {"fit_queue_kernel.cuh", "fit_kernel_dmma.cuh", "dmma_tile.cuh", "lm_step.cuh"} <= queue and "batch_fit_kernel.cuh" not in queue
    assert "stream_kernel.cuh" in simt and "fit_kernel_dmma.cuh" not in simt
    for d in (batch, queue, simt):
        assert {"inst.cu", "kernel_tables.h", "device_common.cuh", "varpro_b200.h"} <= d
    # every kernel group of kernel_tables.h has a translation unit, and the stamps differ between parts
    assert len([u for u in units if u.startswith("inst_")]) == len(vb._groups())
    assert vb._stamp(*units["inst_f64_3_3_batch.o"]) != vb._stamp(*units["inst_f64_3_2_queue1.o"])
