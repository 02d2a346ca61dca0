"""PCIe probe for the e2e leg: pinned H2D alone, D2H alone, both at once (two streams), at the
4K float32 frame size (99.5 MB) and in 4 / 16 chunks.  Prints GB/s per direction."""
import torch
n = 3840 * 2160 * 3
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = n * 4


def run(h2d, d2h, chunks, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    step = n // chunks
    for _ in range(reps):
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for chunks in (1, 4, 16):
    run(True, True, chunks, 3)
    a, b, c = run(True, False, chunks), run(False, True, chunks), run(True, True, chunks)
    print("chunks %2d: H2D alone %.2f ms (%.1f GB/s), D2H alone %.2f ms (%.1f GB/s), both %.2f ms (%.1f GB/s each way)"
          % (chunks, a, nbytes / a / 1e6, b, nbytes / b / 1e6, c, nbytes / c / 1e6))
