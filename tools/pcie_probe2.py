"""Does an in-flight D2H delay an H2D submitted later on another stream (and vice versa)?  Monolithic vs 8 MiB pieces."""
import time
import torch

n = 256 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
small_h = torch.empty(4 << 20, dtype=torch.uint8).pin_memory()
small_d = torch.empty(4 << 20, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ev = lambda: torch.cuda.Event(enable_timing=True)


def copy(dst, src, stream, piece):
    with torch.cuda.stream(stream):
        if piece is None:
            dst.copy_(src, non_blocking=True)
        else:
            for o in range(0, n, piece):
                dst[o:o + piece].copy_(src[o:o + piece], non_blocking=True)


def case(first, second, piece, delay_ms=1.0):
    torch.cuda.synchronize()
    base = ev(); base.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    marks = {}
    def go(kind, stream):
        b, e = ev(), ev()
        b.record(stream)
        if kind == "h2d": copy(d_a, h_a, stream, piece)
        elif kind == "d2h": copy(h_b, d_b, stream, piece)
        elif kind == "small_h2d":
            with torch.cuda.stream(stream): small_d.copy_(small_h, non_blocking=True)
        elif kind == "small_d2h":
            with torch.cuda.stream(stream): small_h.copy_(small_d, non_blocking=True)
        e.record(stream)
        marks[kind] = (b, e)
    go(first, s1)
    time.sleep(delay_ms * 1e-3)
    go(second, s2)
    torch.cuda.synchronize()
    out = " ".join(f"{k}: {base.elapsed_time(b):6.2f}->{base.elapsed_time(e):6.2f}" for k, (b, e) in marks.items())
    print(f"piece={piece} {first} then {second}: {out}")


for piece in (None, 8 << 20):
    case("d2h", "h2d", piece)
    case("h2d", "d2h", piece)
    case("h2d", "small_h2d", piece)
    case("d2h", "small_d2h", piece)
    case("d2h", "small_h2d", piece)
    case("h2d", "small_d2h", piece)
