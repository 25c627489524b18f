import time, torch, torch.nn.functional as F, os
x = torch.randn(4, 384, 96, 96); w = torch.randn(64, 384, 7, 7)
for nt in (16, 32, 64, 128):
    torch.set_num_threads(nt)
    t = time.perf_counter(); F.conv2d(x, w, None, 2, 3); t1 = time.perf_counter() - t
    t = time.perf_counter(); F.conv2d(x, w, None, 2, 3); t2 = time.perf_counter() - t
    print(nt, "threads: first", round(t1, 3), "second", round(t2, 3), flush=True)
