"""Offline model of the shared-memory bank conflicts of the tile LJ kernel (kernels/tiles.cu).
Builds the tile staging order and the ELL rows for an oracle liquid state, then counts LDS.64
wavefronts per warp instruction for candidate thread mappings.  LDS.64 rule: a warp is served as
two half-warps; within one, lanes reading the same 8-byte word are broadcast, distinct words in
the same bank pair (word index mod 16) serialise."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "tests"))
from oracle_py import OracleMD

def main(region=(20, 20, 20), tile=(2, 2, 4), steps=25):
    md = OracleMD.from_deck(REPO / "input" / "in.lj", "CSR", "NEIGH_FULL", region=region)
    md.step(steps)
    md.stage("exchange", "bin_sort", "halo", "bin_all", "neigh")
    x = md.arr("x"); n = md.geti("N_local")
    g = md.geom(); nbx, nby, nbz = g["nbinx"], g["nbiny"], g["nbinz"]
    bc = md.arr("bincount").reshape(nbx, nby, nbz); bo = md.arr("binoffsets").reshape(nbx, nby, nbz); pv = md.arr("permute")
    cut2 = md.getd("neigh_cutoff") ** 2
    tx, ty, tz = tile
    res = {"A_thread_per_atom": [0, 0], "B_8lane_group": [0, 0], "C_16lane_group": [0, 0], "D_thread_per_atom_bank_rotated": [0, 0],
           "E_rotate_kernel": [0, 0], "F_schedule_kernel": [0, 0]}
    cols = {"E_rotate_kernel": [0, 0], "F_schedule_kernel": [0, 0]}  # [columns, longest row] per warp
    ntile = 0
    for bx0 in range(1, nbx - 1, tx):
        for by0 in range(1, nby - 1, ty):
            for bz0 in range(1, nbz - 1, tz):
                ntile += 1
                if ntile % 7: continue  # sample
                # staged cells
                start = {}; slots_j = []
                for cx in range(bx0 - 1, bx0 + tx + 1):
                    for cy in range(by0 - 1, by0 + ty + 1):
                        for cz in range(bz0 - 1, bz0 + tz + 1):
                            ok = 0 <= cx < nbx and 0 <= cy < nby and 0 <= cz < nbz
                            cnt = bc[cx, cy, cz] if ok else 0
                            start[(cx, cy, cz)] = (len(slots_j), cnt)
                            if cnt: slots_j.extend(pv[bo[cx, cy, cz]: bo[cx, cy, cz] + cnt])
                slots_j = np.array(slots_j); xs = x[slots_j]
                rows = []
                for cx in range(bx0, min(bx0 + tx, nbx - 1)):
                    for cy in range(by0, min(by0 + ty, nby - 1)):
                        for cz in range(bz0, min(bz0 + tz, nbz - 1)):
                            s0, cnt = start[(cx, cy, cz)]
                            for k in range(cnt):
                                own = s0 + k
                                if slots_j[own] >= n: rows.append(np.zeros(0, int)); continue
                                cand = []
                                for dx in (-1, 0, 1):
                                    for dy in (-1, 0, 1):
                                        for dz in (-1, 0, 1):
                                            b, c = start[(cx + dx, cy + dy, cz + dz)]
                                            cand.extend(range(b, b + c))
                                cand = np.array(cand)
                                d = xs[cand] - xs[own]
                                keep = ((d * d).sum(1) <= cut2) & (cand != own)
                                rows.append(cand[keep])
                def wavefronts(addr_halves):
                    w = 0
                    for half in addr_halves:
                        half = np.unique(half[half >= 0])
                        if half.size == 0: continue
                        w += np.bincount(half % 16, minlength=16).max()
                    return w
                # A: warp = 32 consecutive atoms, step q
                for w0 in range(0, len(rows), 32):
                    grp = rows[w0:w0 + 32]
                    maxn = max((len(r) for r in grp), default=0)
                    for q in range(maxn):
                        a = np.array([r[q] if q < len(r) else -1 for r in grp] + [-1] * (32 - len(grp)))
                        res["A_thread_per_atom"][0] += wavefronts([a[:16], a[16:]]); res["A_thread_per_atom"][1] += (a >= 0).sum()
                # D: thread per atom, each row reordered so that lane l prefers bank (q + l) mod 16 at step q
                def rotate(r, lane):
                    buckets = [list(r[r % 16 == b]) for b in range(16)]
                    out = []
                    for q in range(len(r)):
                        b = (q + lane) % 16
                        if not buckets[b]:
                            b = max(range(16), key=lambda k: len(buckets[k]))
                        out.append(buckets[b].pop(0))
                    return np.array(out, int)
                for w0 in range(0, len(rows), 32):
                    grp = [rotate(r, l) for l, r in enumerate(rows[w0:w0 + 32])]
                    maxn = max((len(r) for r in grp), default=0)
                    for q in range(maxn):
                        a = np.array([r[q] if q < len(r) else -1 for r in grp] + [-1] * (32 - len(grp)))
                        res["D_thread_per_atom_bank_rotated"][0] += wavefronts([a[:16], a[16:]]); res["D_thread_per_atom_bank_rotated"][1] += (a >= 0).sum()
                # E: kernels/tiles.cu tiles_rotate_kernel -- lane l reads bank (q + l) mod 16 at column q while that bucket lasts,
                #    else a bucket holding more than its share of what is left (first in rotation order), else any
                def rotate_kernel(r, hl):
                    buckets = [list(r[r % 16 == b]) for b in range(16)]
                    out = []
                    for q in range(len(r)):
                        pref = (q + hl) % 16
                        b = pref
                        if not buckets[pref]:
                            th = (len(r) - q + 15) // 16
                            ne = sorted((k for k in range(16) if buckets[k]), key=lambda k: (k - pref) % 16)
                            big = [k for k in ne if len(buckets[k]) > th]
                            b = big[0] if big else ne[0]
                        out.append(buckets[b].pop(0))
                    return np.array(out, int)
                # F: tiles_schedule_kernel (EMD_TILES_SCHED=full) -- lane after lane: join a slot already taken in this column if
                #    it heads one of my buckets, else the first free bank (rotation order) among my over-full buckets, else among
                #    all my non-empty ones; no free bank: idle (padding)
                def schedule_half(grp):
                    bk = [[sorted(r[r % 16 == b]) for b in range(16)] for r in grp]
                    rem = [len(r) for r in grp]
                    columns = []
                    q = 0
                    while any(rem):
                        col = [-1] * 16
                        taken = {}
                        for l in range(len(grp)):
                            if not rem[l]: continue
                            pick = next(((b, sl) for b, sl in taken.items() if bk[l][b] and bk[l][b][0] == sl), None)
                            if pick is None:
                                r0 = (q + l) % 16
                                free = sorted((b for b in range(16) if bk[l][b] and b not in taken), key=lambda b: (b - r0) % 16)
                                if free:
                                    th = (rem[l] + 15) // 16
                                    big = [b for b in free if len(bk[l][b]) > th]
                                    b = big[0] if big else free[0]
                                    pick = (b, bk[l][b][0])
                            if pick is None: continue
                            b, sl = pick
                            bk[l][b].remove(sl); rem[l] -= 1; taken[b] = sl; col[l] = sl
                        columns.append(col); q += 1
                    return columns
                for w0 in range(0, len(rows), 32):
                    grp = [rotate_kernel(r, l % 16) for l, r in enumerate(rows[w0:w0 + 32])]
                    maxn = max((len(r) for r in grp), default=0)
                    cols["E_rotate_kernel"][0] += maxn; cols["E_rotate_kernel"][1] += maxn
                    for q in range(maxn):
                        a = np.array([r[q] if q < len(r) else -1 for r in grp] + [-1] * (32 - len(grp)))
                        res["E_rotate_kernel"][0] += wavefronts([a[:16], a[16:]]); res["E_rotate_kernel"][1] += (a >= 0).sum()
                    raw = rows[w0:w0 + 32]
                    h0, h1 = schedule_half(raw[:16]), (schedule_half(raw[16:]) if len(raw) > 16 else [])
                    cols["F_schedule_kernel"][0] += max(len(h0), len(h1)); cols["F_schedule_kernel"][1] += max((len(r) for r in raw), default=0)
                    for h in (h0, h1):
                        for col in h:
                            a = np.array(col)
                            res["F_schedule_kernel"][0] += wavefronts([a]); res["F_schedule_kernel"][1] += (a >= 0).sum()
                # B: 8 lanes per atom, 4 atoms per warp
                for name, gl in (("B_8lane_group", 8), ("C_16lane_group", 16)):
                    apw = 32 // gl
                    for w0 in range(0, len(rows), apw):
                        grp = rows[w0:w0 + apw]
                        maxn = max((len(r) for r in grp), default=0)
                        for q in range(0, maxn, gl):
                            a = np.full(32, -1)
                            for gi, r in enumerate(grp):
                                seg = r[q:q + gl]
                                a[gi * gl: gi * gl + len(seg)] = seg
                            res[name][0] += wavefronts([a[:16], a[16:]]); res[name][1] += (a >= 0).sum()
    for k, (w, p) in res.items():
        extra = f"; columns = {cols[k][0] / max(cols[k][1], 1):.3f} x longest row of the warp" if k in cols else ""
        print(f"{k}: {w / p * 32:.2f} wavefronts per 32 pairs per LDS.64  (x3 arrays = {3 * w / p * 32:.1f}){extra}")

if __name__ == "__main__":
    main()
