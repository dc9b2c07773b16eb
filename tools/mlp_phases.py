"""Tuning aid (GPU): phase timestamps of the fused MLP chain on the Patch04 shape."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops, _lib
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
widths = [64, 128, 256, 512, 256, 64]
M = 640000
ACT, L2 = ops.ACT_RELU, True
if len(sys.argv) > 1 and sys.argv[1] == "attention":  # theta / phi of the patch attention (model/attention.py:29-46)
    widths, M, ACT, L2 = [128, 128, 128, 128, 32], 1048576, ops.ACT_LEAKY, False
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, widths[0], device=dev, generator=g)
ws = [torch.randn(widths[i + 1], widths[i], device=dev, generator=g) / widths[i] ** 0.5 for i in range(len(widths) - 1)]
bs = [torch.randn(widths[i + 1], device=dev, generator=g) * 0.1 for i in range(len(widths) - 1)]
imgs = [ops.tc_mlp_weight_image(w) for w in ws]
for _ in range(3):
    y = ops.tc_mlp(x, imgs, bs, widths, act=ACT, l2_normalize=L2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ops.tc_mlp(x, imgs, bs, widths, act=ACT, l2_normalize=L2); e1.record(); e1.synchronize()
print(f"tc_mlp {e0.elapsed_time(e1):.3f} ms")
buf = (ctypes.c_longlong * 64)()
_lib.lib().rf_tc_mlp_debug_read(ctypes.cast(buf, ctypes.c_void_p))
t0 = buf[0]
print("input planes ready", buf[1] - t0)
for l in range(len(widths) - 1):
    print(f"layer {l}: start {buf[2 + 3 * l] - t0}  mma done {buf[3 + 3 * l] - t0} (+{buf[3 + 3 * l] - buf[2 + 3 * l]})  epilogue done {buf[4 + 3 * l] - t0} (+{buf[4 + 3 * l] - buf[3 + 3 * l]})")
