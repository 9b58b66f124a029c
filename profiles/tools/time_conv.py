import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import waldo_b200 as wb
bench.load_peak()
dev = torch.device("cuda:0")
print(bench.wif_to_emb_leg(dev))
# forward + backward (d raw_output by the kernel, d weight by torch)
B, Tc, Tp, Cin, Cout, H, W = 8, 4, 1, 40, 16, 512, 1024
gen = torch.Generator(device=dev).manual_seed(4)
raw = torch.randn(B, Tc, Tp, Cin, H, W, device=dev, generator=gen).requires_grad_(True)
wgt = (torch.randn(Cout, Cin, 3, 3, device=dev, generator=gen) * 0.05).requires_grad_(True)
g = torch.randn(B * Tc * Tp, Cout, H, W, device=dev, generator=gen)
def step(fn):
    raw.grad = None; wgt.grad = None
    fn().backward(g)
def timed(fn, n=5):
    for _ in range(2): step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ours = timed(lambda: wb.wif_to_emb(raw, wgt))
ref = timed(lambda: torch.nn.functional.conv2d(raw.permute(0, 2, 1, 3, 4, 5).reshape(B * Tp * Tc, Cin, H, W), wgt, padding=1))
wgt_only = wgt.detach().requires_grad_(True)
print(f"fwd+bwd: waldo_b200 {ours:.3f} ms (d raw_output by k_conv3x3_fwd<16,5>, d weight by k_conv3x3_wgrad), stock torch {ref:.3f} ms")
