"""Driver of flood_sim.c: activations / sweeps of tile-scheduling policies (design exploration)."""
import ctypes, sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import fieldgen

def _build():
    """gcc -O3 -shared flood_sim.c into the temp dir (rebuilt when the source is newer)."""
    import subprocess, tempfile
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flood_sim.c")
    out = os.path.join(tempfile.gettempdir(), f"ws_flood_sim_{os.getuid()}.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O3", "-shared", "-fPIC", "-o", out, src], check=True)
    return out


lib = ctypes.CDLL(_build())
def maxima(img):
    a = img.astype(np.int16)
    c = a[1:-1, 1:-1]
    ok = np.ones(c.shape, bool)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr or dc:
                ok &= a[1 + dr:a.shape[0] - 1 + dr, 1 + dc:a.shape[1] - 1 + dc] < c
    rc = np.argwhere(ok) + 1
    return np.ascontiguousarray(rc, dtype=np.int32)

def run(img, seeds, TW, TH, policy, delta, ncta=444):
    R, C = img.shape
    T = np.empty((R, C), np.uint32)
    stats = (ctypes.c_long * 8)()
    hist = np.zeros(1 << 16, np.uint32)
    t0 = time.time()
    lib.flood_sim(img.ctypes.data_as(ctypes.c_void_p), R, C, seeds.ctypes.data_as(ctypes.c_void_p), len(seeds), TW, TH,
                  policy, delta, ncta, T.ctypes.data_as(ctypes.c_void_p), stats, hist.ctypes.data_as(ctypes.c_void_p), hist.size)
    return T, list(stats)[:7], hist[:stats[0]], time.time() - t0

if __name__ == "__main__":
    S = int(sys.argv[1]); kind = sys.argv[2]; sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 16.0
    img = {"uniform": lambda: fieldgen.uniform(S, S, 0), "smooth": lambda: fieldgen.smooth(S, S, sigma, 0),
           "cgps": lambda: fieldgen.cgps_like(S, S, 0)}[kind]()
    seeds = maxima(img)
    print(kind, S, "seeds", len(seeds))
    ref = None
    for (TW, TH, pol, d) in [(64, 32, 0, 0), (64, 32, 1, 0), (64, 32, 1, 2), (64, 32, 1, 8), (64, 32, 2, 64), (64, 32, 2, 4096),
                             (128, 64, 0, 0), (128, 64, 1, 2), (128, 128, 0, 0)]:
        T, st, hist, dt = run(img, seeds, TW, TH, pol, d)
        if ref is None: ref = T
        assert np.array_equal(T, ref)
        nt = ((S + TW - 1) // TW) * ((S + TH - 1) // TH)
        print(f"tile {TW}x{TH} policy {pol} delta {d}: sweeps {st[0]} acts {st[1]} ({st[1]/nt:.2f}/tile) gs {st[2]} ({st[2]/max(st[1],1):.2f}/act) "
              f"rounds {st[3]} small-sweeps {st[5]} changed-px/px {st[4]/S/S:.2f}  area-work {st[1]*TW*TH/S/S:.2f}  [{dt:.1f}s]")
