"""Expected results of the bench's sharded reductions, computed on the CPU by the oracle (test infrastructure) over the
WHOLE synthetic raster, chunk by chunk — so that every bench run can assert its GPU result against a constant instead
of trusting the generator's luck. Writes tests/golden/bench_expected.json.

    python tests/golden/make_bench_expected.py        # ~ minutes of CPU

config 4: f32 32768^2, seed 0xEC40, uniform real in [-1e4, 1e4) (synth kind 2): total-order min / max bits and the
count of cells (statistics' count); the same for the 16384^2 raster the 2-GPU test uses."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from erased_cells_b200 import synth  # host generator only (numpy)
from erased_cells_b200.api import CellType
from oracle import oracle as orc

orc.build()


def f32_raster_min_max(side, seed, chunk=1 << 24):
    n = side * side
    kmin = kmax = None
    for off in range(0, n, chunk):
        a = synth.host(CellType.Float32, min(chunk, n - off), seed, index_offset=off, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
        mn, mx = orc.tight_min_max(a)
        if kmin is None or orc.value_cmp(mn, kmin) < 0:
            kmin = mn
        if kmax is None or orc.value_cmp(mx, kmax) > 0:
            kmax = mx
    return {"side": side, "seed": seed, "min_bits": hex(kmin.bits), "max_bits": hex(kmax.bits), "count": n}


out = {"c4_f32_32768": f32_raster_min_max(32768, 0xEC40), "c4_f32_16384": f32_raster_min_max(16384, 0xEC40)}
with open(os.path.join(ROOT, "tests", "golden", "bench_expected.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
