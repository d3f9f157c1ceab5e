"""PCIe H2D diagnostics: how fast do pinned host buffers of the bench's sizes reach HBM, as one big copy, as the
384 per-frame pieces on one stream, and as pieces spread over 2 / 3 streams."""
import torch

MB = 1 << 20
sizes = []
for _ in range(128):
    sizes += [int(3.07 * MB), int(0.64 * MB), int(2.56 * MB)]
total = sum(sizes)
big = torch.empty(total, dtype=torch.uint8).pin_memory()
dbig = torch.empty(total, dtype=torch.uint8, device="cuda")
pieces = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
dpieces = [torch.empty(s, dtype=torch.uint8, device="cuda") for s in sizes]


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def one():
    dbig.copy_(big, non_blocking=True)


def split(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]

    def run():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for s in streams:
            s.wait_event(ev)
        for i, (src, dst) in enumerate(zip(pieces, dpieces)):
            with torch.cuda.stream(streams[i % nstreams]):
                dst.copy_(src, non_blocking=True)
        for s in streams:
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
    return run


def single():
    for src, dst in zip(pieces, dpieces):
        dst.copy_(src, non_blocking=True)


for name, fn in (("one contiguous copy", one), ("384 pieces, 1 stream", single), ("384 pieces, 2 streams", split(2)), ("384 pieces, 3 streams", split(3))):
    ms = timed(fn)
    print(f"{name}: {ms:.2f} ms  {total / ms / 1e6:.1f} GB/s")

# the same pieces as views into ONE pinned arena
offs = [0]
for s in sizes:
    offs.append(offs[-1] + s)
views = [big[offs[i]:offs[i + 1]] for i in range(len(sizes))]


def arena():
    for src, dst in zip(views, dpieces):
        dst.copy_(src, non_blocking=True)


print(f"384 pieces out of one pinned arena, 1 stream: {timed(arena):.2f} ms  {total / timed(arena) / 1e6:.1f} GB/s")
# fewer, bigger pieces
views2 = [big[offs[3 * i]:offs[3 * i + 3]] for i in range(128)]
d2 = [torch.empty(v.numel(), dtype=torch.uint8, device="cuda") for v in views2]


def arena2():
    for src, dst in zip(views2, d2):
        dst.copy_(src, non_blocking=True)


print(f"128 pieces (6.3 MB) out of one arena: {timed(arena2):.2f} ms  {total / timed(arena2) / 1e6:.1f} GB/s")
