"""development: raw qprof counters after a small quadcopter batch"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import quadcopter, scenarios
sc = scenarios.quadcopter_batch(8, 100, seed=2)
o = obca_b200.default_opts(); o.max_iter = 60
r = quadcopter.quadcopter_solve_batch(sc["x0"], sc["xF"], sc["N"], sc["Ts"], sc["R"], sc["obs"], sc["xWS"], 1.0, 1, o)
q = (C.c_ulonglong * 8)(); obca_b200.lib().obca_debug_qprof(q)
print("iters", r["iters"], "exit", r["exitflag"], "qprof", list(q))
