python -m pytest tests -m gpu -q -x -k "fused or golden or smoke" 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in 22 23 13 14; do echo "variant $v"; PXF_CHAIN_VARIANT=$v $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'], d['clocks'])"; done
echo interpreter; PXF_NO_SPECIALIZE=1 $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'])"
python - <<'PY'
import torch, time
n=1<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,fn in (('h2d',lambda: d.copy_(h,non_blocking=True)),('d2h',lambda: h.copy_(d,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, 3*n/(time.perf_counter()-t)/1e9,'GB/s')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream(); h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); print('bidir', 6*n/(time.perf_counter()-t)/1e9,'GB/s total')
PY
