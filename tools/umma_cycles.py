"""Prints tcgen05.mma issue/completion cycles for the shapes the PFN uses (diagnostic, run on the GPU box)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcp_b200 import _lib  # noqa: E402

lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for mode in (0, 1):
    for n in (32, 64):
        for ksteps in (1, 2, 4):
            for reps in (0, 1, 2, 8, 32):
                _lib.check(lib.pcp_selftest_umma_cycles(mode, n, ksteps, reps, C.c_void_p(out.data_ptr()), None), "cycles")
                torch.cuda.synchronize()
                o = out.cpu().tolist()
                nm = 3 * ksteps * reps
                print(f"mode={'TS' if mode else 'SS'} N={n} ksteps={ksteps} reps={reps} mmas={nm}: total {o[0]} cyc, issue {o[2]} cyc, "
                      f"empty commit+wait {o[1]} cyc" + (f", per-MMA {(o[0] - o[1]) / nm:.1f}" if nm else ""))
