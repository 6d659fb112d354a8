"""Latency of ONE device-resident image per call (BASELINE configs 1 and 2 are single images) for the kernel paths."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import libiqo_b200 as iqo

CASES = [("cfg1 L3 1080p->720p", iqo.LANCZOS, 3, 1, 1920, 1080, 1280, 720),
         ("cfg4 L3 1080p->540p", iqo.LANCZOS, 3, 1, 1920, 1080, 960, 540),
         ("cfg3y L2 2160p->1080p", iqo.LANCZOS, 2, 1, 3840, 2160, 1920, 1080),
         ("L3 1000x700->333x500", iqo.LANCZOS, 3, 1, 1000, 700, 333, 500),
         ("cfg2a area 2160p->1080p", iqo.AREA, 0, 1, 3840, 2160, 1920, 1080),
         ("cfg2b linear 720p->2160p", iqo.LINEAR, 0, 1, 1280, 720, 3840, 2160)]
PATHS = [("auto", iqo.PATH_AUTO), ("no_stream", iqo.PATH_NO_STREAM)]
NS = [int(v) for v in sys.argv[1:]] or [1, 4]
for name, kind, deg, px, sw, sh, dw, dh in CASES:
    for n in NS:
        src = torch.randint(0, 256, (n, sh, sw), dtype=torch.uint8, device="cuda")
        dst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
        line = "%-26s n=%d " % (name, n)
        for pname, path in PATHS:
            with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
                r.set_path(path)
                s = torch.cuda.current_stream().cuda_stream
                for _ in range(20):
                    r.resize_batch(n, sw, sw * sh, src, dw, dw * dh, dst, s)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(200):
                    r.resize_batch(n, sw, sw * sh, src, dw, dw * dh, dst, s)
                e1.record(); torch.cuda.synchronize()
                line += " %s[%s] %.1f us" % (pname, r.last_kernel(), e0.elapsed_time(e1) * 1000 / 200)
        print(line)
