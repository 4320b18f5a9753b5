"""What a chunked, dependent H2D -> D2H pipeline can reach on this box (development aid): D2H of chunk i may start only after
H2D of chunk i has finished (as in HostStreamedCanonicalizer, with the kernels taken out)."""
import sys, time, torch
B, shape = 512, (3, 224, 224)
x_host = torch.rand(B, *shape).pin_memory()
z_host = torch.empty_like(x_host).pin_memory()
dev = torch.device("cuda")
xd = torch.empty(B, *shape, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(chunk, dependent=True, lag=0):
    evs = []
    n = B // chunk
    for i in range(n):
        lo, hi = i * chunk, (i + 1) * chunk
        with torch.cuda.stream(s1):
            xd[lo:hi].copy_(x_host[lo:hi], non_blocking=True)
            evs.append(s1.record_event())
        with torch.cuda.stream(s2):
            if dependent:
                s2.wait_event(evs[i])
            z_host[lo:hi].copy_(xd[lo:hi], non_blocking=True)
    torch.cuda.synchronize()
for chunk in (512, 128, 64, 32, 16):
    for dep in (False, True):
        for _ in range(2): run(chunk, dep)
        t0 = time.perf_counter()
        for _ in range(5): run(chunk, dep)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"chunk {chunk:4d} dependent={dep}: {ms:.2f} ms per 512 images each way = {B / ms * 1e3:.0f} img/s")
