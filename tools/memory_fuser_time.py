import torch, time, sys
sys.path.insert(0, '/root/repo')
import mavlm_b200 as M
torch.manual_seed(0)
enc = M.MemoryFuser(3584, num_layers=2, num_heads=4).eval().cuda().bfloat16()
x = torch.randn(80, 196, 3584, device='cuda').bfloat16()
with torch.no_grad():
    for _ in range(2): y = enc(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y = enc(x)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
fl = 80 * 196 * (2 * 3584 * 3584 * (1 + 1) + 2 * (2 * 3584 * 3584 * 12)) + 2 * 80 * 4 * 4 * 196 * 196 * 896
print(f"MemoryFuser 7B dims, 80x196 tokens, 2 layers: {ms:.2f} ms, {fl / ms / 1e9:.0f} TFLOP/s, finite={bool(torch.isfinite(y.float()).all())}")
