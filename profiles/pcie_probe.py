"""Raw pinned-memory copy bandwidth of the box (upper bound of the host-buffer e2e path)."""
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, chunks=1, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        c = n // chunks
        for i in range(chunks):
            if up:
                with torch.cuda.stream(s1): d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h2[i*c//2:(i+1)*c//2].copy_(d2[i*c//2:(i+1)*c//2], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return dt
for up, down, ch in [(1,0,1),(0,1,1),(1,1,1),(1,1,32),(1,0,32),(1,0,256)]:
    dt = run(up, down, ch); run(up, down, ch)
    print(f"up={up} down={down} chunks={ch}: {dt*1e3:.2f} ms  H2D {up*n/dt/1e9:.1f} GB/s  D2H {down*n/2/dt/1e9:.1f} GB/s")
