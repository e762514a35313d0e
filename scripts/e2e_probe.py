import importlib, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
rtnw = importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")
ctx = rtnw.Context(0); hs = rtnw.HostScene("final_northstar")
nx=ny=1000; cam = hs.camera(nx, ny); p = hs.params(nx=nx, ny=ny, ns=16, seed=1)
host = torch.empty(ny, nx, 3, dtype=torch.float32).pin_memory(); hn = host.numpy()
for rep in range(3):
    t0=time.perf_counter(); ds = ctx.upload(hs.desc_ptr); t1=time.perf_counter()
    out, st = ds.render(cam, p, out=hn); t2=time.perf_counter()
    ds.close(); t3=time.perf_counter()
    print(f"upload {1e3*(t1-t0):.2f} ms  render(host) {1e3*(t2-t1):.2f} ms [kernel {st.kernel_ms:.2f} total {st.total_ms:.2f}]  close {1e3*(t3-t2):.2f} ms")
pag = np.empty((ny,nx,3), dtype=np.float32)
ds = ctx.upload(hs.desc_ptr)
t1=time.perf_counter(); out, st = ds.render(cam, p, out=pag); t2=time.perf_counter()
print(f"pageable out: render {1e3*(t2-t1):.2f} ms kernel {st.kernel_ms:.2f}")
