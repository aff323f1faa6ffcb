"""Does source/destination alignment of the pinned chunk copies matter?  Bare H2D+D2H chunk pipeline, chunk starts shifted."""
import sys, time, torch
n = 1 << 30
h_in = torch.empty(n + 4096, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n + 4096, dtype=torch.uint8, pin_memory=True)
chunk = 32 << 20
slots = 3
d = [torch.empty(chunk + 4096, dtype=torch.uint8, device="cuda") for _ in range(slots)]
st = [torch.cuda.Stream() for _ in range(slots)]
for shift_host, shift_dev in ((0, 0), (0, 37), (0, 1), (0, 128), (128, 0), (256, 37)):
    def run():
        k = 0
        for off in range(0, n, chunk):
            s = k % slots
            a = off + (shift_host if off else 0)
            ln = min(chunk, n - a)
            with torch.cuda.stream(st[s]):
                d[s][shift_dev:shift_dev + ln].copy_(h_in[a:a + ln], non_blocking=True)
                h_out[a:a + ln].copy_(d[s][shift_dev:shift_dev + ln], non_blocking=True)
            k += 1
        torch.cuda.synchronize()
    run()
    t0 = time.perf_counter()
    for _ in range(3):
        run()
    print("host shift %4d dev shift %2d: %.1f GB/s" % (shift_host, shift_dev, 3 * n / (time.perf_counter() - t0) / 1e9), flush=True)
