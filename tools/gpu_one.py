"""One solve of B config-2 problems (development tool; e.g. under ncu).  python tools/gpu_one.py [B] [mode] [thresh] [timing]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if len(sys.argv) > 2:
    os.environ["OBCA_MODE"] = sys.argv[2]
if len(sys.argv) > 3:
    os.environ["OBCA_TAIL_THRESH"] = sys.argv[3]
if len(sys.argv) > 4:
    os.environ["OBCA_PHASE_TIMING"] = "1"
import obca_b200
from obca_b200 import parking, scenarios

sc = scenarios.reverse_parking_batch(B, 80, 0)
for rep in range(2 if len(sys.argv) > 4 else 1):
    r = parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                    sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
lib = obca_b200.lib()
rnd = C.c_int(0); hand = C.c_int(0); kms = (C.c_double * 5)(); pr = (C.c_ulonglong * 8)()
lib.obca_last_schedule(C.c_int(0), C.byref(rnd), C.byref(hand), kms)
lib.obca_last_profile(C.c_int(0), pr)
print(f"B={B} solve {r['time'] * 1e3:.1f} ms conv {int(r['exitflag'].sum())} iters mean {r['iters'].mean():.1f} max {r['iters'].max()} rounds {rnd.value} "
      f"handed {hand.value} ms [eval, sweep, step, tail, dualws] {[round(x, 3) for x in kms]} counters {[int(x) for x in pr]}")
