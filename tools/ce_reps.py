import sys, os, torch
sys.path.insert(0, "/root/repo")
from recommendsystemproject_b200 import ops
dev="cuda"; B,H,D=65536,4096,128
g=torch.Generator(device=dev).manual_seed(1)
u=torch.nn.functional.normalize(torch.randn(B,D,device=dev,generator=g),dim=1).requires_grad_(True)
it=torch.nn.functional.normalize(torch.randn(B,D,device=dev,generator=g),dim=1).requires_grad_(True)
pool=torch.nn.functional.normalize(torch.randn(H,D,device=dev,generator=g),dim=1).requires_grad_(True)
ids=torch.randint(1,B*50,(B,),device=dev,generator=g)
flush_buf=torch.empty(256*1024*1024,dtype=torch.uint8,device=dev)
def f():
    loss=ops.fused_inbatch_ce(u,it,ids,None,pool,0.05,precision="bf16")[0]
    loss.backward()
for _ in range(3): f()
torch.cuda.synchronize()
ts=[]
for r in range(12):
    flush_buf.fill_(1)
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); b.synchronize(); ts.append(round(a.elapsed_time(b),3))
print("flushed reps:", ts)
ts=[]
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for r in range(20): f()
b.record(); b.synchronize()
print("20 back-to-back: %.3f ms each" % (a.elapsed_time(b)/20))
import subprocess
print(subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active","--format=csv,noheader"],capture_output=True,text=True).stdout)
