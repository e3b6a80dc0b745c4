import time, os, subprocess, sys
import torch
print(torch.__version__, torch.version.cuda, torch.backends.cuda.cufft_plan_cache.max_size)
x = torch.zeros(8, 256, 256, device='cuda')
for M in (256, 300, 1024, 1050, 1215, 1792, 336, 672):
    x = torch.zeros(8, M, M, device='cuda')
    torch.cuda.synchronize(); t=time.time(); y = torch.fft.rfft2(x); torch.cuda.synchronize(); t1=time.time()-t
    t=time.time(); y = torch.fft.rfft2(x); torch.cuda.synchronize(); t2=time.time()-t
    print('torch rfft2 M=%d first %.1f ms second %.3f ms' % (M, t1*1e3, t2*1e3), flush=True)
os.system("grep -i cufft /proc/%d/maps | awk '{print $NF}' | sort -u" % os.getpid())
os.system("du -sh ~/.nv/ComputeCache 2>/dev/null; ls ~/.nv 2>/dev/null; nproc; uptime")
