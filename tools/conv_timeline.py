"""Per-CTA timeline of the tensor-core conv (globaltimer stamps): where does a tile's time go?"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops, _lib
L = _lib.lib()
torch.manual_seed(0)
for (N, C_, H, W, K, R) in [(8, 496, 128, 240, 496, 3), (8, 256, 256, 256, 128, 3), (8, 64, 128, 128, 64, 3)]:
    x = torch.randn(N, H, W, C_, device="cuda")
    w = torch.randn(K, C_, R, R, device="cuda") / (C_ * R * R) ** 0.5
    wp = ops.pack_conv_weight_f16(w)
    pad = (R // 2,) * 4
    for _ in range(2):
        ops.conv2d(x, wp, K, R, R, 1, pad, precision="3xfp16")
    buf = torch.zeros(1 << 20, dtype=torch.int64, device="cuda")
    L.creste_conv2d_tc_debug(C.c_void_p(buf.data_ptr()))
    ops.conv2d(x, wp, K, R, R, 1, pad, precision="3xfp16")
    torch.cuda.synchronize()
    L.creste_conv2d_tc_debug(C.c_void_p(0))
    t = buf.view(-1, 8).cpu().double()
    t = t[t[:, 0] > 0]
    lead = t[t[:, 2] > 0]              # leader CTAs have the first-full stamp
    d = lambda a, b, tt=lead: float((tt[:, b] - tt[:, a]).mean()) / 1e3
    print(f"C{C_}->K{K} @{H}x{W}: {len(t)} CTAs; us: prologue {d(0,1):.1f} | first operands {d(1,2):.1f} | "
          f"main loop {d(2,3):.1f} | epilogue {d(3,4):.1f} | drain+sync {d(4,5):.1f} | total {d(0,5):.1f}; "
          f"kernel span {(float(t[:,5].max()-t[:,0].min()))/1e3:.0f} us")
