python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
python - > gpurun_out/eval_speed.log 2>&1 <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import zeroshotvideoclassification_b200 as z
torch.manual_seed(0)
model = z.get_network(z.default_opt("r2plus1d_18")).cuda().eval()
x = torch.randn(22, 1, 3, 16, 112, 112, device="cuda")
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
def folded():
    with torch.no_grad(): model(x)
def unfolded():
    model(x)
print("eval forward bs=22: folded %.2f ms, conv+scale/shift passes %.2f ms" % (bench(folded), bench(unfolded)))
PY
cat gpurun_out/pytest_gpu.log gpurun_out/eval_speed.log
