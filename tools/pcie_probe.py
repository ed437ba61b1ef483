#!/usr/bin/env python3
"""PCIe ceiling of the box next to the e2e step's phases: raw pinned H2D / D2H bandwidth (torch copy engine), then
hark_table_from_host / filter / hark_table_to_host timed separately on the bench's e2e shape."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harkdb_b200 import hark_ffi  # noqa: E402


def main():
    out = {}
    nbytes = 4 << 30
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[name + "_gbs"] = nbytes / best / 1e9
    del h, d
    env = hark_ffi.Futhark()
    rows = 1 << 28
    t = env.synth(rows, [3] * 8, [dict(kind=0)] * 8, seed=42)
    import ctypes as C
    hin = env.lib.hark_host_alloc(rows * 32)
    hout = env.lib.hark_host_alloc(rows * 8)
    host_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_float)), shape=(rows, 8))
    host_out = np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_float)), shape=(rows, 2))
    t.to_numpy(out=host_in)
    t.free()
    preds = [(1, 0, 0, 0.5), (4, 2, 0, 0.5)]
    for chunk in (64, 16, 256):
        env.set_option("upload.chunk_mb", chunk)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dt = env.to_device(host_in, np.float32)
            env.sync()
            t1 = time.perf_counter()
            r = env.query_filter(dt, [0, 2], preds)
            env.sync()
            t2 = time.perf_counter()
            k = r.shape[0]
            r.to_numpy(out=host_out[:k])
            t3 = time.perf_counter()
            r.free(); dt.free()
        out[f"chunk{chunk}"] = {"from_host_ms": (t1 - t0) * 1e3, "from_host_gbs": rows * 32 / (t1 - t0) / 1e9,
                                "filter_ms": (t2 - t1) * 1e3, "to_host_ms": (t3 - t2) * 1e3,
                                "to_host_gbs": k * 8 / (t3 - t2) / 1e9}
    env.set_option("upload.chunk_mb", 64)

    def e2e_step(trace=None):
        t0 = time.perf_counter()
        dt = env.to_device(host_in, np.float32)
        t1 = time.perf_counter()
        r = env.query_filter(dt, [0, 2], preds)
        dt.free()
        k = r.shape[0]
        t2 = time.perf_counter()
        r.to_numpy(out=host_out[:k])
        t3 = time.perf_counter()
        r.free()
        t4 = time.perf_counter()
        if trace is not None:
            trace.append([round((b - a) * 1e3, 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))])
        return k

    def loop(tag):
        tr = []
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            e2e_step(tr)
        torch.cuda.synchronize()
        out[tag] = {"step_ms": (time.perf_counter() - t0) * 1e3 / 4, "phases_ms[from_host,filter,to_host,free]": tr}

    loop("steps_pool_only")
    big = env.synth(10 ** 9, [3] * 8, [dict(kind=0)] * 8, seed=42)      # the bench keeps its 32 GB table resident
    loop("steps_with_resident_table")
    big.free()
    loop("steps_after_free")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
