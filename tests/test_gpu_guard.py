"""Out-of-bounds write check without compute-sanitizer (closed on this pool): the parity suite's kernels rerun in a
subprocess with EC_DEBUG_GUARD=1 — every device block then has 256-byte red zones hugging its first and last byte,
verified on free — over ragged sizes around every tile boundary."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import sys
    import numpy as np
    sys.path.insert(0, %r)
    import erased_cells_b200 as ec
    from erased_cells_b200 import CellBuffer, CellType, CellValue, Mask, MaskedCellBuffer, NoData, synth
    L = ec.lib()
    sizes = [1, 31, 33, 127, 4095, 4096, 4097, 8191, 16384 + 5, 32768 * 2 + 77, 65536 * 3 + 1]
    launches0 = L.ec_kernel_launches()
    for n in sizes:
        for ct in CellType:
            a = CellBuffer.from_vec(synth.host(ct, n, 1 + int(ct)))
            b = CellBuffer.from_vec(synth.host(ct, n + 3, 2 + int(ct)))
            for op in range(4):
                a._bin(op, b); a._bin(op, CellValue(CellType.Float64, 0.5))
            (-a); a.clone(); a.min_max(); a.normalized_difference(b); a.binary_scalar(3, b, 2, 0.5); a == b
            m = Mask.new(synth.host(CellType.UInt8, n, n) < 128)
            (~m); (m & m); (m | m); m.counts(); m.to_vec()
            ma = MaskedCellBuffer.from_buffer_with_nodata(a, NoData.default(ct))
            (ma - MaskedCellBuffer(a, m)); ma.min_max(); MaskedCellBuffer(a, m).min_max(); a.statistics(); MaskedCellBuffer(a, m).statistics()
            for d in CellType:
                if ct.can_fit_into(d):
                    a.convert(d); MaskedCellBuffer(a, m).to_vec_with_nodata(NoData.new(d, 1))
            CellBuffer.fill(n, CellValue(ct, 3)); CellBuffer.with_defaults(n, ct)
            e = CellBuffer.from_vec(synth.host(ct, 40, 9)); e.extend(np.arange(5).astype(ct.dtype))
            with ec.lazy():
                ((a - b) / (a + b)).to_vec(); (a / b * 0.5).to_vec()
            with ec.lazy(jit=True):  # run-time specialised kernel (one build per cell type, then cached)
                ((a - b) * 0.5 + a).to_vec()
    del a, b, m, ma, e
    L.ec_synchronize()
    print("launches", L.ec_kernel_launches() - launches0, "violations", L.ec_guard_violations())
    sys.exit(1 if L.ec_guard_violations() else 0)
""") % ROOT


@pytest.mark.gpu
def test_no_out_of_bounds_writes_with_red_zones():
    env = dict(os.environ, EC_DEBUG_GUARD="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "violations 0" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_red_zones_catch_a_deliberate_overrun():
    """the guard itself works: write one cell past a wrapped sub-range that ends exactly at a block's last byte"""
    script = textwrap.dedent("""
        import sys
        import numpy as np
        sys.path.insert(0, %r)
        import erased_cells_b200 as ec
        from erased_cells_b200 import CellBuffer, CellType, CellValue
        L = ec.lib()
        b = CellBuffer.with_defaults(1000, CellType.UInt8)
        over = CellBuffer.wrap_device(CellType.UInt8, b.device_ptr() + 992, 9)   # 1 byte past the block
        over.put(8, CellValue(CellType.UInt8, 7))
        del over, b
        L.ec_synchronize()
        sys.exit(0 if L.ec_guard_violations() == 1 else 1)
    """) % ROOT
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300, env=dict(os.environ, EC_DEBUG_GUARD="1"))
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
