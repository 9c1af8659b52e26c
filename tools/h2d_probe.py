import torch, time
n = 361*1024*1024
a = torch.empty(n, dtype=torch.uint8, pin_memory=True); b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
da = torch.empty(n, dtype=torch.uint8, device="cuda"); db = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one():
    with torch.cuda.stream(s1):
        da.copy_(a, non_blocking=True); db.copy_(b, non_blocking=True)
def two():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): db.copy_(b, non_blocking=True)
def d2h():
    with torch.cuda.stream(s1): a.copy_(da, non_blocking=True)
def both():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
for name, f, nbytes in (("h2d 1 stream", one, 2*n), ("h2d 2 streams", two, 2*n), ("d2h", d2h, n), ("h2d+d2h", both, 2*n)):
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, "%.1f GB/s" % (nbytes / dt / 1e9))
import os; print("cpus", os.cpu_count())
os.system(r"nvidia-smi topo -m | head -20; numactl -H 2>/dev/null | head -5; lscpu | grep -i 'numa\|model name\|^CPU(s)'")
