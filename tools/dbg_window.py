import os, sys, time, ctypes as C
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.getcwd())
import examinimd_b200 as emd
L = emd.lib()
app = emd.App(["-il", "input/in.lj", "--comm-type", "SERIAL", "--neigh-type", "CSR", "--force-iteration", "NEIGH_HALF", "--region", "80", "80", "80"])
def timed(fn):
    ms = C.c_float()
    emd.check(L.emd_ctx_tic(app.ctx)); fn(); emd.check(L.emd_ctx_toc(app.ctx, C.byref(ms)))
    return ms.value
app.run(5); app.thermo(); app.advance(14)
for k in range(4):
    print("run(20) window", k, "%.3f ms" % timed(lambda: app.run(20)), flush=True)
for k in range(3):
    print("advance(20) window", k, "%.3f ms" % timed(lambda: app.advance(20)), flush=True)
for k in range(3):
    print("run(10) window", k, "%.3f ms" % timed(lambda: app.run(10)), flush=True)
for k in range(3):
    t0 = time.time(); r = app.thermo(); print("thermo() host wall %.3f ms" % (1e3 * (time.time() - t0)))
