"""Where the end-to-end time goes: H2D, list build, nl_pairs_to_host (by host thread count, i rebuilt or copied), plain D2H rates,
and the two host decoders on their own."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, C, L = make_positions(n, 10)
Xh = torch.from_numpy(X).pin_memory()
pbc = (True, True, True)
Lb = nl._lib.lib()


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return 1e3 * min(t)


print("H2D 240 MB: %.2f ms" % timed(lambda: Xh.to("cuda", non_blocking=True)))
Xd = Xh.cuda()
print("list (no R): %.2f ms" % timed(lambda: nl.neighbour_list(Xd, CUTOFF, C, pbc)))
pl = nl.neighbour_list(Xd, CUTOFF, C, pbc)
P = nl.npairs(pl)
buf = nl.HostPairBuffers(P + 1024, n)
for nt in (4, 6, 8, 10, 12, 16):
    print("to_host nthreads=%2d, fraction of i copied 0 / 0.2 / 0.3 / 0.4 / 0.5 / 1: " % nt +
          " ".join("%.2f" % timed(lambda: nl.to_host(pl, out=buf, nthreads=nt, i_copy_fraction=f)) for f in (0.0, 0.2, 0.3, 0.4, 0.5, 1.0)) + " ms")
print("plain D2H j (%.2f GB): %.2f ms" % (4e-9 * P, timed(lambda: buf.j[:P].copy_(pl.j, non_blocking=True))))
print("plain D2H S (%.2f GB): %.2f ms" % (12e-9 * P, timed(lambda: buf.S[:P].copy_(pl.S, non_blocking=True))))

first = buf.first[:n + 1].numpy()
codes = buf.host_scratch.numpy()
for nt in (1, 4, 8, 16):
    q = ((P + nt - 1) // nt + 3) & ~3
    def run(fn):
        th = [threading.Thread(target=fn, args=(min(P, q * t), min(P, q * (t + 1)))) for t in range(nt)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        return 1e3 * (time.perf_counter() - t0)
    te = min(run(lambda a, b: Lb.nl_host_expand_rows(0, first.ctypes.data, None, n, a, b, buf.i.data_ptr())) for _ in range(3))
    tu = min(run(lambda a, b: Lb.nl_host_unpack_shifts(0, codes.ctypes.data, a, b, buf.S.data_ptr())) for _ in range(3))
    print("host decoders, %2d threads: expand i %.2f ms (%.1f GB/s)  unpack S %.2f ms (%.1f GB/s)" % (nt, te, 4e-6 * P / te, tu, 12e-6 * P / tu))
# memset-like upper bound of host stores
a = np.empty(P * 3, dtype=np.int32)
t0 = time.perf_counter(); a.fill(0); print("numpy fill 3.1 GB, 1 thread: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
