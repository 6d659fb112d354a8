import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, libiqo_b200 as iqo
n = 256
with iqo.Yuv420Resizer(0, 2, 3840, 2160, 1920, 1080) as r:
    src = torch.randint(0, 256, (n, r.src_frame_bytes), dtype=torch.uint8, device="cuda")
    dst = torch.zeros((n, r.dst_frame_bytes), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        r.resize(n, src, dst, s)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r.resize(n, src, dst, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = n * (r.src_frame_bytes + r.dst_frame_bytes)
    print("cfg3 256 YUV420 frames 4K->1080p Lanczos2: %.4f ms per batch, %.1f GB/s, %.1f%% of 6544, %.1f dst Gpix/s"
          % (ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / 6544, n * 3110400 / ms / 1e6))
