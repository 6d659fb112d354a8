"""Device-resident batches on the tensor-path kernel (PATH_MMA) against what AUTO would otherwise run (PATH_NO_MMA)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import libiqo_b200 as iqo

CASES = [("L3 3:2 1080p->720p", iqo.LANCZOS, 3, 1, 1920, 1080, 1280, 720, 512),
         ("L2 3:2 1080p->720p", iqo.LANCZOS, 2, 1, 1920, 1080, 1280, 720, 512),
         ("L4 3:2 1080p->720p", iqo.LANCZOS, 4, 1, 1920, 1080, 1280, 720, 512),
         ("L3 3:2 2160p->1440p", iqo.LANCZOS, 3, 1, 3840, 2160, 2560, 1440, 128),
         ("L3 X2:1 Y3:2 1080p->960x720", iqo.LANCZOS, 3, 1, 1920, 1080, 960, 720, 512),
         ("L3 1:2 540p->1080p", iqo.LANCZOS, 3, 1, 960, 540, 1920, 1080, 512),
         ("L3 3:4 810p->1080p", iqo.LANCZOS, 3, 1, 1440, 810, 1920, 1080, 512),
         ("L3 2:1 1080p->540p", iqo.LANCZOS, 3, 1, 1920, 1080, 960, 540, 512),
         ("L2 2:1 2160p->1080p", iqo.LANCZOS, 2, 1, 3840, 2160, 1920, 1080, 128),
         ("L3 4:1 2160p->540p", iqo.LANCZOS, 3, 1, 3840, 2160, 960, 540, 128),
         ("L3 1000x700->333x500", iqo.LANCZOS, 3, 1, 1000, 700, 333, 500, 1024),
         ("area 3:2 1080p->720p", iqo.AREA, 0, 1, 1920, 1080, 1280, 720, 512),
         ("area 3:1 2160p->720p", iqo.AREA, 0, 1, 3840, 2160, 1280, 720, 128),
         ("linear 2:3 720p->1080p", iqo.LINEAR, 0, 1, 1280, 720, 1920, 1080, 512),
         ("linear 3:2 1080p->720p", iqo.LINEAR, 0, 1, 1920, 1080, 1280, 720, 512)]
PATHS = [("mma", iqo.PATH_MMA), ("no_mma", iqo.PATH_NO_MMA)]
for name, kind, deg, px, sw, sh, dw, dh, n in CASES:
    src = torch.randint(0, 256, (n, sh, sw), dtype=torch.uint8, device="cuda")
    dst = torch.zeros((n, dh, dw), dtype=torch.uint8, device="cuda")
    line = "%-30s n=%4d " % (name, n)
    outs = []
    for pname, path in PATHS:
        with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
            r.set_path(path)
            s = torch.cuda.current_stream().cuda_stream
            try:
                for _ in range(3):
                    r.resize_batch(n, sw, sw * sh, src, dw, dw * dh, dst, s)
            except Exception as e:  # the forced path may not apply to the shape
                line += " %s[n/a]" % pname
                continue
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                r.resize_batch(n, sw, sw * sh, src, dw, dw * dh, dst, s)
            e1.record(); torch.cuda.synchronize()
            outs.append(dst.clone())
            line += " %s[%s] %.3f ms" % (pname, r.last_kernel(), e0.elapsed_time(e1) / 10)
    if len(outs) == 2:
        line += "  equal=%s" % bool(torch.equal(outs[0], outs[1]))
    print(line, flush=True)
