#!/usr/bin/env python
"""Host<->device copy rates of this box (pinned memory, 1 GiB per direction), alone and
concurrently: the ceiling of bench.py's e2e figure, which moves 2 B in + 2 B out per pixel."""
import json

import torch


def rate(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    r = {"h2d_GBps": rate(h2d, n), "d2h_GBps": rate(d2h, n), "bidirectional_each_GBps": rate(both, n)}
    r["e2e_ceiling_Mpixel_per_s"] = r["bidirectional_each_GBps"] * 1e9 / 2 / 1e6
    print(json.dumps(r))


if __name__ == "__main__":
    main()
