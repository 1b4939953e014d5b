"""Small fixed workload for ncu: N perturbed get rollouts, L launches.
usage: python tools/profile_run.py [n] [launches] [command] [kernel: warp|sweep|auto]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affaction_b200 as ab
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cmd = sys.argv[3] if len(sys.argv) > 3 else "get fresh_tomatoes"
if len(sys.argv) > 4 and sys.argv[4] != "auto":
    os.environ["AFF_KERNEL"] = sys.argv[4]
s = ab.HostScene()
b = ab.HostBatch(s)
b.add_perturbed(cmd, n)
dev = ab.DeviceBatch(s, b, ab.default_opts())
st = torch.cuda.current_stream().cuda_stream
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
times = []
for i in range(launches):
    e0.record()
    dev.launch(st)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    print("launch %d: %.2f ms, %.0f rollouts/s" % (i, times[-1], n / times[-1] * 1e3))
res = dev.results()
print("success", sum(r.success for r in res), "of", n)
print("best %.2f ms = %.0f rollouts/s" % (min(times), n / min(times) * 1e3))
