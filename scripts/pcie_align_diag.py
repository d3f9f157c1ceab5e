"""Does the per-copy overhead of pinned H2D copies depend on size / address alignment?"""
import torch
MB = 1 << 20


def run(label, sizes, align):
    total = sum(sizes)
    host = torch.empty(total + len(sizes) * align + align, dtype=torch.uint8).pin_memory()
    dev = torch.empty(total + len(sizes) * align + align, dtype=torch.uint8, device="cuda")
    offs = []; o = 0
    for s in sizes:
        o = (o + align - 1) // align * align
        offs.append(o); o += s
    src = [host[a:a + s] for a, s in zip(offs, sizes)]; dst = [dev[a:a + s] for a, s in zip(offs, sizes)]
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for x, y in zip(src, dst):
            y.copy_(x, non_blocking=True)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{label}: {best:.2f} ms {total / best / 1e6:.1f} GB/s ({len(sizes)} copies)")


odd = []
for _ in range(128):
    odd += [3145704, 655360 - 16, 2621440 - 48]
run("odd sizes, 16 B aligned", odd, 16)
run("odd sizes, 4 KB aligned starts", odd, 4096)
al = [(s + 4095) // 4096 * 4096 for s in odd]
run("4 KB-multiple sizes, 4 KB aligned", al, 4096)
al = [(s + 65535) // 65536 * 65536 for s in odd]
run("64 KB-multiple sizes, 64 KB aligned", al, 65536)
al2 = [(s + (2 << 20) - 1) // (2 << 20) * (2 << 20) for s in odd]
run("2 MB-multiple sizes, 2 MB aligned", al2, 2 << 20)
run("one copy", [sum(odd)], 4096)
