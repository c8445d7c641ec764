import os, sys, ctypes as C, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import examinimd_b200 as emd
L = emd.lib()
os.environ["EMD_HOST_LATTICE"] = "1"
app = emd.App(["-il", "input/in.lj", "--comm-type", "SERIAL", "--region", "10", "10", "10"])
h = app.download(); app.close()
n = h["v"].shape[0]
class Lat(C.Structure):
    _fields_ = [("i0", C.c_longlong * 3), ("n", C.c_int * 3), ("fcc", C.c_int), ("a", C.c_double), ("offset", C.c_double * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3)]
a = (4.0 / 0.8442) ** (1.0 / 3.0)
lat = Lat()
dom = a * 10
for d in range(3):
    lat.i0[d] = int(0.0 / dom * 10 - 0.5); i1 = int(dom / dom * 10 + 0.5); lat.n[d] = i1 - lat.i0[d] + 1
    lat.offset[d] = 0.0; lat.lo[d] = 0.0; lat.hi[d] = dom
lat.fcc = 1; lat.a = a
ctx = emd.Context()
P = C.c_void_p
cnt = C.c_int()
emd.check(L.emd_lattice_count(ctx.handle, C.byref(lat), C.byref(cnt)))
print("count", cnt.value, n)
x = torch.zeros((n, 3), dtype=torch.float64, device="cuda"); v = torch.zeros_like(x); q = torch.zeros(n, dtype=torch.float64, device="cuda")
typ = torch.zeros(n, dtype=torch.int32, device="cuda"); ids = torch.zeros(n, dtype=torch.int32, device="cuda")
mass = torch.tensor([2.0], dtype=torch.float64, device="cuda")
emd.check(L.emd_lattice_fill(ctx.handle, C.byref(lat), 87287, 0, P(mass.data_ptr()), P(x.data_ptr()), P(v.data_ptr()), P(q.data_ptr()), P(typ.data_ptr()), P(ids.data_ptr())))
torch.cuda.synchronize()
xr, vr = x.cpu().numpy(), v.cpu().numpy()
print("x equal", np.array_equal(xr, h["x"]))
# numpy pipeline on the device's raw velocities
m = 2.0
tm = px = py = pz = 0.0
for i in range(n):
    tm += m; px += m * vr[i, 0]; py += m * vr[i, 1]; pz += m * vr[i, 2]
vs = vr - np.array([px / tm, py / tm, pz / tm])
T = 0.0
for i in range(n):
    T += (vs[i, 0] * vs[i, 0] + vs[i, 1] * vs[i, 1] + vs[i, 2] * vs[i, 2]) * m
dof = 3 * n - 3
T *= 1.0 / (1.0 * dof * 1.0)
sc = math.sqrt(1.4 / T)
vf = vs * sc
print("numpy pipeline on device raw v == host path:", np.array_equal(vf, h["v"]), "mismatches", (vf != h["v"]).sum())
# host-side hash for atom 0 to check the raw value
def uniform(seed):
    IA, IM, IQ, IR = 16807, 2147483647, 127773, 2836
    k = seed // IQ
    seed = IA * (seed - k * IQ) - IR * k
    if seed < 0: seed += IM
    return seed, (1.0 / IM) * seed
import struct
def raw(i):
    hsh = 0
    data = struct.pack("<i", 87287) + struct.pack("<ddd", *xr[i])
    for b in data:
        sb = b - 256 if b > 127 else b
        hsh = (hsh + sb) & 0xffffffff
        hsh = (hsh + (hsh << 10)) & 0xffffffff
        hsh ^= hsh >> 6
    hsh = (hsh + (hsh << 3)) & 0xffffffff; hsh ^= hsh >> 11; hsh = (hsh + (hsh << 15)) & 0xffffffff
    seed = hsh & 0x7ffffff or 1
    for _ in range(5): seed, _u = uniform(seed)
    out = []
    for _ in range(3):
        seed, u = uniform(seed); out.append((u - 0.5) / math.sqrt(2.0))
    return out
bad = 0
for i in range(0, n, 97):
    r = raw(i)
    if list(vr[i]) != r: bad += 1; print("raw mismatch", i, list(vr[i]), r)
print("raw checked, bad =", bad)
