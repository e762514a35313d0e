import importlib, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
ctx = rtnw.Context(0); hs = rtnw.HostScene("final_northstar")
nx=ny=1000; cam = hs.camera(nx, ny)
host = torch.empty(ny, nx, 3, dtype=torch.float32).pin_memory(); hn = host.numpy()
for ns in (16, 100):
    p = hs.params(nx=nx, ny=ny, ns=ns, seed=1)
    for rep in range(3):
        torch.cuda.synchronize()
        t0=time.perf_counter(); ds = ctx.upload(hs.desc_ptr); t1=time.perf_counter()
        out, st = ds.render(cam, p, out=hn); t2=time.perf_counter()
        ds.close(); torch.cuda.synchronize(); t3=time.perf_counter()
        print(f"ns {ns}: upload {1e3*(t1-t0):.2f} render(host) {1e3*(t2-t1):.2f} [kernel {st.kernel_ms:.2f} total {st.total_ms:.2f}] close+sync {1e3*(t3-t2):.2f} | step {1e3*(t3-t0):.2f} ms")
