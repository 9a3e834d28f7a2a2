"""PCIe floor for the host-pointer calls: pinned H2D / D2H of the benchmark's byte counts."""
import torch
n_out, n_in = 4096 << 20, 1414 << 20
d = torch.empty(n_out, dtype=torch.uint8, device="cuda")
h = torch.empty(n_out, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("D2H 4.3 GB", lambda: h.copy_(d, non_blocking=True)), ("H2D 4.3 GB", lambda: d.copy_(h, non_blocking=True)),
                 ("H2D 1.41 GB", lambda: d[:n_in].copy_(h[:n_in], non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s: %.1f ms" % (name, ms))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n_in, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n_in, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); e0.record()
with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
s1.synchronize(); s2.synchronize(); e1.record(); torch.cuda.synchronize()
print("D2H 4.3 GB + H2D 1.41 GB concurrently: %.1f ms" % e0.elapsed_time(e1))
