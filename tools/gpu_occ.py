"""development: per-kernel event times of the rounds (rounds only, full batch) for several CTA-per-SM caps of k_pk_eval / k_pk_step"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obca_b200
from obca_b200 import parking, scenarios
sc = scenarios.reverse_parking_batch(4096, 80, 0)
def run():
    return parking.parking_solve_batch(sc["x0"], sc["xF"], 80, sc["Ts"], sc["L"], sc["ego"], sc["XYbounds"], 3, sc["vOb"], sc["A"], sc["b"],
                                       sc["rx"], sc["ry"], sc["ryaw"], 0, sc["xWS"], sc["uWS"])
os.environ["OBCA_MODE"] = "0"; os.environ["OBCA_PHASE_TIMING"] = "1"
for key, vals in (("OBCA_EVAL_OCC", (3, 2, 1)), ("OBCA_STEP_OCC", (5, 3, 2, 1))):
    for v in vals:
        os.environ[key] = str(v)
        run(); r = run()
        rnd = C.c_int(0); hand = C.c_int(0); kms = (C.c_double * 5)()
        obca_b200.lib().obca_last_schedule(C.c_int(0), C.byref(rnd), C.byref(hand), kms)
        print(f"{key}={v}: total {r['time']*1e3:.1f} ms rounds {rnd.value} handed {hand.value}; ms [eval, sweep, step, tail] = {[round(x, 2) for x in kms][:4]}", flush=True)
    os.environ.pop(key)
