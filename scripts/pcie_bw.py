import torch
dev=torch.device("cuda:0")
n=64*256*256*4
d=torch.empty(n,device=dev); h=torch.empty(n).pin_memory(); h2=torch.empty(n).pin_memory()
def t(f,reps=10):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
ms=t(lambda: h.copy_(d,non_blocking=True)); print("D2H 1 stream: %.1f GB/s"%(n*4/ms/1e6))
ms=t(lambda: d.copy_(h,non_blocking=True)); print("H2D 1 stream: %.1f GB/s"%(n*4/ms/1e6))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
half=n//2
def two():
    cur=torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h[:half].copy_(d[:half],non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(d[half:],non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms=t(two); print("D2H 2 streams: %.1f GB/s"%(n*4/ms/1e6))
def bidir():
    cur=torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
d2=torch.empty(n,device=dev)
ms=t(bidir); print("D2H+H2D concurrently: %.1f GB/s each"%(n*4/ms/1e6))
