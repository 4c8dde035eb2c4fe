import os, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, ntm_b200
from ntm_b200 import lib, signals
from conftest import load_ckpt
dev="cuda:0"
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
with torch.inference_mode():
    m=ntm_b200.RNN(1,64,1,False).to(dev); m.load_state_dict(load_ckpt("cfg2")); m.mode="fp32"
    m.initialize_hidden(); m.warm_start(); hw=m.hidden.clone()
    for B,T in ((1,100000),(8,100000),(148,48000),(256,48000),(1024,24000),(2048,12000),(8192,6000)):
        x=signals.stream_batch_device(B,T,dev,dur=10.0).reshape(B,1,T)
        m.hidden=hw.expand(1,B,64).contiguous(); m(x[:,:,:500])
        best=1e9
        for _ in range(2):
            m.hidden=hw.expand(1,B,64).contiguous()
            e0.record(); y=m(x); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
        print(f"fp32 B={B} T={T}: {best*1e6/T:8.1f} ns/step {B*T/best/1e6:7.3f} Gs/s",flush=True)
