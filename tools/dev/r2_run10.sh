set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mma.py -q -x -m gpu --timeout 300 > gpurun_out/r2_mma_tests2.log 2>&1; echo "mma tests rc=$?"; tail -15 gpurun_out/r2_mma_tests2.log
python - <<'P'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch, libiqo_b200 as iqo
def run(kind, sw, sh, dw, dh, n, path):
    src = torch.randint(0,256,(n,sh,sw),dtype=torch.uint8,device='cuda'); dst=torch.zeros((n,dh,dw),dtype=torch.uint8,device='cuda')
    s=torch.cuda.current_stream().cuda_stream
    with iqo.make_resizer(kind,0,sw,sh,dw,dh) as r:
        r.set_path(path)
        for _ in range(3): r.resize_batch(n,sw,sw*sh,src,dw,dw*dh,dst,s)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(5): r.resize_batch(n,sw,sw*sh,src,dw,dw*dh,dst,s)
        e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/5
        k=r.last_kernel()
    by=n*(sw*sh+dw*dh)
    print("%-8s %dx%d->%dx%d x%d %-12s %.3f ms  %.0f GB/s frac %.3f" % ("area" if kind==1 else "linear", sw,sh,dw,dh,n,k,ms,by/ms/1e6, by/ms/1e6/6544))
for path in (iqo.PATH_NO_MMA, iqo.PATH_MMA):
    run(1,1920,1080,1280,720,1024,path)
    run(1,3840,2160,1280,720,256,path)
    run(2,1280,720,1920,1080,1024,path)
    run(2,1280,720,3200,1800,512,path)
    run(1,3840,2160,1920,1080,256,path)
    run(2,1280,720,3840,2160,256,path)
P
