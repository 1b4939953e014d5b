"""Strong scaling of ONE host process over the GPUs of a box: a fixed sweep of perturbed candidates through
aff_predict_batch_multi (static split, one host thread + private streams per device, records written into
the caller's array) on 1 / 2 / 4 / 8 devices.  Host buffers in and out: plan compile, H2D, kernels, D2H.
usage: python tools/strong_scaling.py [n ...]     (default 100000 1000000)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affaction_b200 as ab  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [100000, 1000000]
ndev = ab.lib().aff_device_count()
scene = ab.HostScene()
warm = ab.HostBatch(scene)
warm.add_perturbed("get fresh_tomatoes", 64 * max(ndev, 1))
ab.predict_batch_multi(scene, warm, list(range(ndev)))          # CUDA context + module load on every device
print("| candidates | devices | seconds | rollouts/s | speed-up | efficiency |")
print("|---|---|---|---|---|---|")
for n in sizes:
    b = ab.HostBatch(scene)
    t0 = time.perf_counter()
    b.add_perturbed("get fresh_tomatoes", n)
    gen = time.perf_counter() - t0
    base = None
    ref = None
    for g in (1, 2, 4, 8):
        if g > ndev:
            break
        t0 = time.perf_counter()
        res = ab.predict_batch_multi(scene, b, list(range(g)))
        dt = time.perf_counter() - t0
        key = [(r.idx, r.success, r.code, r.first_fail_step) for r in res[::max(1, n // 1000)]]
        if ref is None:
            ref, base = key, dt
        assert key == ref, "verdicts must not depend on the number of devices"
        print("| %d | %d | %.3f | %.0f | %.2f | %.2f |" % (n, g, dt, n / dt, base / dt, base / dt / g), flush=True)
    print("(candidate generation on the host, outside the timed call: %.1f s for %d)" % (gen, n), flush=True)
