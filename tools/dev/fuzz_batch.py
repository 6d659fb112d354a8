"""One-off randomized parity fuzz of device-resident batches with random pitches / frame strides / base offsets."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import libiqo_b200 as iqo
from oracle_lib import oracle_resize, lcg_image, LANCZOS, AREA, LINEAR

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
PATH = iqo.PATH_STREAM if os.environ.get("FUZZ_STREAM") else iqo.PATH_AUTO  # FUZZ_STREAM=1: streaming kernels on small launches too
t0 = time.time()
stats, bad = {}, 0
while time.time() - t0 < budget:
    fam = rng.randint(0, 4)
    if fam == 0:
        dw, dh = int(rng.randint(4, 200)) * 2, int(rng.randint(8, 120)); kind, deg, px, sw, sh = LANCZOS, int(rng.choice([2, 3])), 1, 2 * dw, 2 * dh
    elif fam == 1:
        k = int(rng.randint(1, 30)) * 8; kind, deg, px, sw, sh, dw, dh = LANCZOS, 3, 1, 3 * k, int(rng.randint(16, 200)), 2 * k, int(rng.randint(16, 200))
    elif fam == 2:
        sw, sh = int(rng.randint(2, 100)) * 4, int(rng.randint(4, 100)); kind, deg, px, dw, dh = LINEAR, 0, 1, 3 * sw, int(rng.randint(sh, 3 * sh))
    else:
        dw, dh = int(rng.randint(1, 100)) * 8, int(rng.randint(4, 100)); kind, deg, px, sw, sh = AREA, 0, 1, 2 * dw, 2 * dh
    n = int(rng.randint(1, 4))
    spitch = sw + int(rng.choice([0, 4, 8, 16, 12, 1]))
    dpitch = dw + int(rng.choice([0, 4, 8, 3]))
    sfs = spitch * sh + int(rng.choice([0, 8, 16, 4]))
    dfs = dpitch * dh + int(rng.choice([0, 8, 5]))
    soff, doff = int(rng.choice([0, 16, 8, 4])), int(rng.choice([0, 8, 4, 1]))
    host = np.zeros(soff + n * sfs + 64, dtype=np.uint8)
    want = []
    for f in range(n):
        img = lcg_image(sh, spitch, seed=int(rng.randint(1, 1 << 20)))
        host[soff + f * sfs: soff + f * sfs + sh * spitch] = img.reshape(-1)
        rc, w = oracle_resize(kind, img, dw, dh, deg, px, sw=sw)
        want.append(w)
    if rc != 0:
        continue
    dsrc = torch.from_numpy(host).cuda()
    ddst = torch.full((doff + n * dfs + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    with iqo.make_resizer(kind, deg, sw, sh, dw, dh, px) as r:
        r.set_path(PATH)
        r.resize_batch(n, spitch, sfs, dsrc[soff:], dpitch, dfs, ddst[doff:], torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = r.last_kernel()
    stats[k] = stats.get(k, 0) + 1
    out = ddst.cpu().numpy()
    for f in range(n):
        got = out[doff + f * dfs: doff + f * dfs + dh * dpitch].reshape(dh, dpitch)[:, :dw]
        if not np.array_equal(got, want[f]):
            bad += 1
            print("MISMATCH", kind, deg, sw, sh, dw, dh, n, spitch, dpitch, sfs, dfs, soff, doff, k)
            break
print("cases per kernel:", stats, "mismatching cases:", bad)
