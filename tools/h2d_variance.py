#!/usr/bin/env python3
"""Fresh-process H2D timing: 6 uploads of a 2^28 x 8 f32 pinned row-major table (hark_table_from_host), with the
PCIe link state sampled while they run.  Run several times in a row to see process-to-process variance."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harkdb_b200 import hark_ffi  # noqa: E402

env = hark_ffi.Futhark()
rows = 1 << 28
hin = env.lib.hark_host_alloc(rows * 32)
host_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_float)), shape=(rows, 8))
t0 = time.perf_counter()
if len(sys.argv) > 1 and sys.argv[1] == "touch":
    host_in[:] = 1.0            # CPU first touch of every page before any DMA
touch_ms = (time.perf_counter() - t0) * 1e3
ms = []
link = []
for i in range(6):
    t0 = time.perf_counter()
    dt = env.to_device(host_in, np.float32)
    env.sync()
    ms.append(round((time.perf_counter() - t0) * 1e3, 1))
    if i == 2:
        link = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pstate",
                               "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    dt.free()
print(json.dumps({"mode": sys.argv[1:] or "plain", "touch_ms": round(touch_ms, 1), "from_host_ms": ms, "link": link}))
