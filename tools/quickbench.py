import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import affaction_b200 as ab
import torch
s=ab.HostScene(); 
for n in (4096, 32768):
    b=ab.HostBatch(s); t=time.time(); b.add_perturbed('get fresh_tomatoes', n); t_gen=time.time()-t
    t=time.time(); dev=ab.DeviceBatch(s,b,ab.default_opts()); t_create=time.time()-t
    st=torch.cuda.current_stream().cuda_stream
    dev.launch(st); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); dev.launch(st); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    res=dev.results()
    import collections
    print('n',n,'gen %.2fs create %.3fs kernel %.1f ms -> %.0f rollouts/s'%(t_gen,t_create,ms,n/ms*1e3),'grid',dev.grid_blocks,'smem/warp',dev.smem_bytes_per_warp,'h2d MB %.1f'%(dev.h2d_bytes/1e6), collections.Counter(r.code for r in res))
print('fp64 peak TFLOP/s', ab.lib().aff_measure_fp64_peak(0, 1.0))
