#!/usr/bin/env python
"""GPU experiment (not a bench number): which cuDNN kernel does the 7x7 32->32 full-resolution convolution of the
U-Nets (conv1b, 36 % of a whole inference step) get, per dtype / memory format / batch?"""
import json
import torch
import torch.nn.functional as F

dev = "cuda:0"
torch.backends.cudnn.benchmark = True
H, W = 1088, 1920
flops = lambda m, cin, cout, k: 2.0 * m * H * W * cin * cout * k * k
for cin, cout, k in ((32, 32, 7), (16, 32, 7), (32, 64, 5), (64, 64, 5)):
    for m in (2, 4):
        for dtype in (torch.bfloat16, torch.float16, torch.float32):
            for fmt in ("nhwc", "nchw"):
                hh, ww = (H, W) if k == 7 else (H // 2, W // 2)
                x = torch.randn(m, cin, hh, ww, device=dev, dtype=dtype)
                w = torch.randn(cout, cin, k, k, device=dev, dtype=dtype)
                if fmt == "nhwc":
                    x, w = x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
                torch.backends.cudnn.allow_tf32 = True
                for _ in range(3):
                    F.conv2d(x, w, None, 1, k // 2)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    F.conv2d(x, w, None, 1, k // 2)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                fl = 2.0 * m * hh * ww * cin * cout * k * k
                print(json.dumps({"cin": cin, "cout": cout, "k": k, "batch": m, "dtype": str(dtype).split(".")[1], "format": fmt,
                                  "ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1)}))
                del x, w
