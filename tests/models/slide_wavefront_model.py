"""numpy model of the device schedule of snow_slide (chm_b200/csrc/pbsm3d_slide.cuh): live-set frontier expansion + event-driven
wavefront rounds with the "every EARLIER face within two edges has had its turn" rule (a blocked face is parked and woken when a
face within two edges of it has had its turn).  Used by tests/test_slide_oracle.py to show, on CPU,
that the schedule reproduces the reference's sequential sweep bit for bit (same fire() arithmetic, different execution order)."""
import numpy as np


def earlier(key, g, f):
    return key[g] > key[f] or (key[g] == key[f] and g < f)


def fire(st, f, sd, sdv, swe, dsd, dmass):
    """One face's turn, the arithmetic of oracle/slide_oracle.py:SlideState.sweep (no ghosts)."""
    nb, maxD, cz, area, cosf = st.neigh, st.maxDepth, st.cz, st.area, st.cosf
    snow, snow_v, w_e = sd[f], sdv[f], swe[f]
    del_depth = snow - maxD[f]
    del_swe = w_e * (1 - maxD[f] / snow)
    z_s = cz[f] + snow_v
    w = [0.0, 0.0, 0.0]
    w_dem = 0.0
    for j in range(3):
        n = nb[f, j]
        w[j] = max(0.0, z_s - cz[f]) if n < 0 else max(0.0, z_s - (cz[n] + sdv[n]))
        w_dem += w[j]
    if w_dem == 0:
        return
    w = [x / w_dem for x in w]
    for j in range(3):
        n = nb[f, j]
        if n < 0:
            continue
        sd[n] += del_depth * (area[f] / area[n]) * w[j]
        swe[n] += del_swe * (area[f] / area[n]) * w[j]
        sdv[n] = sd[n] / cosf[f]
        dsd[n] += del_depth * area[f] * w[j]
        dmass[n] += del_swe * area[f] * w[j]
    sd[f] = maxD[f]
    sdv[f] = sd[f] / cosf[f]
    swe[f] = w_e * maxD[f] / snow
    dsd[f] -= del_depth * area[f]
    dmass[f] -= del_swe * area[f]


def run(st, snowdepthavg, snowdepthavg_vert, swe_mm):
    """Returns (dsd, dmass, rounds, live_count, fired)."""
    T, nb = st.T, st.neigh
    sd = np.array(snowdepthavg, dtype=np.float64, copy=True)
    sdv = np.array(snowdepthavg_vert, dtype=np.float64, copy=True)
    swe = np.asarray(swe_mm, dtype=np.float64) / 1000.0
    dsd, dmass = np.zeros(T), np.zeros(T)
    key = st.cz + sdv
    stamp = np.ones(T, dtype=np.int64)
    frontier = [int(f) for f in np.flatnonzero(sd > st.maxDepth)]
    stamp[frontier] = 0
    live = list(frontier)
    while frontier:
        nxt = []
        for g in frontier:
            for j in range(3):
                f = nb[g, j]
                if f >= 0 and stamp[f] == 1 and earlier(key, g, f):
                    stamp[f] = 0
                    nxt.append(int(f))
                    live.append(int(f))
        frontier = nxt
    work, rounds, fired, examined = live, 0, 0, 0
    queued = np.zeros(T, dtype=np.int64)
    while work:
        sv = rounds + 2
        before = stamp.copy()          # what "had its turn before this round" means on the device (stamps of this round don't count)
        unsettled = lambda n: before[n] == 0
        turn_now = []
        for f in work:
            if before[f] != 0:
                continue
            examined += 1
            wait = any(n >= 0 and unsettled(n) and earlier(key, n, f) for n in nb[f])
            active = sd[f] > st.maxDepth[f]
            if not wait and active:
                for n in nb[f]:
                    if n < 0:
                        continue
                    for m in nb[n]:
                        if m >= 0 and m != f and unsettled(m) and earlier(key, m, f):
                            wait = True
            if not wait:                # parked otherwise: the blocking face wakes it when it has had its turn
                turn_now.append(f)
        # faces taking their turn in one round are > 2 edges apart if they fire: any order gives the same bits; reverse it
        nxt = []
        for f in reversed(turn_now):
            if sd[f] > st.maxDepth[f]:
                fire(st, f, sd, sdv, swe, dsd, dmass)
                fired += 1
            stamp[f] = sv
        for f in turn_now:              # wake the later-ordered live faces within two edges whose turn is still to come
            for n in nb[f]:
                if n < 0:
                    continue
                for m in [n] + [int(x) for x in nb[n]]:
                    if m >= 0 and m != f and stamp[m] == 0 and earlier(key, f, m) and queued[m] != sv:
                        queued[m] = sv
                        nxt.append(m)
        work = nxt
        rounds += 1
    assert not np.any(stamp == 0), "a live face never had its turn"
    return dsd, dmass, rounds, len(live), fired
