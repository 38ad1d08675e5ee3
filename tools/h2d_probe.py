"""Pinned host -> device bandwidth for the e2e batch payload (154 MB): one stream vs the copy split over 2 / 4 streams."""
import torch
n = 153_904_896
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunk = (n + parts - 1) // parts
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(4):
            for i, s in enumerate(streams):
                s.wait_event(e0)
                with torch.cuda.stream(s):
                    d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
    print(f"{parts} stream(s): {4 * n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
