import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
import erased_cells_b200 as ec
from erased_cells_b200 import CellType, synth
L = ec.lib(); ec._lib.check(L.ec_init(0)); L.ec_set_min_max_cache(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ec._lib.check(L.ec_set_stream(C.c_void_p(st.cuda_stream)))
n5 = 32768 * 32768
def make(t):
    return (synth.device(CellType.UInt16, n5, 0xEC50 + t, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0),
            synth.device(CellType.UInt16, n5, 0xEC58 + t, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0))
bands = [make(t) for t in range(8)]
def ev(): return torch.cuda.Event(enable_timing=True)
def run(tiles, label, do_mm=True, keep=None):
    for rep in range(3):
        e = [ev() for _ in range(2 * len(tiles) + 1)]
        e[0].record()
        for i, t in enumerate(tiles):
            nir, red = bands[t]
            nd = nir.normalized_difference(red)
            e[2 * i + 1].record()
            if do_mm: nd.min_max()
            e[2 * i + 2].record()
        torch.cuda.synchronize()
    nd_ms = [e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(len(tiles))]
    mm_ms = [e[2 * i + 1].elapsed_time(e[2 * i + 2]) for i in range(len(tiles))]
    print(label, "total %.3f" % e[0].elapsed_time(e[-1]), "ndvi", ["%.3f" % x for x in nd_ms], "min_max", ["%.3f" % x for x in mm_ms])
from erased_cells_b200 import sharding
bd = {t: bands[t] for t in range(8)}
def c5(which):
    k0 = k1 = None
    for t in which:
        nir, red = bd[t]
        ndvi = nir.normalized_difference(red)
        mn, mx = ndvi.min_max()
        k = sharding.keys_of(mn, mx)
        k0 = k[0] if k0 is None else min(k0, k[0])
        k1 = k[1] if k1 is None else min(k1, k[1])
    return k0, k1
for rep in range(3):
    a, b = ev(), ev()
    t0 = time.perf_counter(); a.record()
    for _ in range(3): c5(range(8))
    b.record(); torch.cuda.current_stream().synchronize()
    print("bench-style c5: %.3f ms per step (events), %.3f wall" % (a.elapsed_time(b) / 3, (time.perf_counter() - t0) * 1e3 / 3))
print("cached bytes", L.ec_cached_bytes() / 1e9)
import os
print("ATTR", os.environ.get("EC_PDL_REDUCE_ATTR"), "TRIGGER", os.environ.get("EC_PDL_REDUCE_TRIGGER"))

nir, red = bands[0]
def chain():
    return (nir - red) / (nir + red)
for ov in (0, 1, 0, 1):
    L.ec_set_launch_overlap(ov)
    chain(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(3): r = chain()
    b.record(); torch.cuda.current_stream().synchronize()
    print("launch overlap", ov, "unfused NDVI chain 32768^2: %.3f ms" % (a.elapsed_time(b) / 3))
del r
small = [synth.device(CellType.UInt16, 4096 * 4096, 0x70 + i, kind=synth.INT_RANGE, lo=1, hi=40000) for i in range(16)]
def chain_small():
    out = None
    for i in range(0, 16, 2):
        out = (small[i] - small[i + 1]) / (small[i] + small[i + 1])
    return out
for ov in (0, 1, 0, 1):
    L.ec_set_launch_overlap(ov)
    chain_small(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(10): r = chain_small()
    b.record(); torch.cuda.current_stream().synchronize()
    print("launch overlap", ov, "8 unfused NDVI chains at 4096^2: %.3f ms" % (a.elapsed_time(b) / 10))
