import json, os, sys, torch
sys.path.insert(0, os.getcwd())
import ssm_b200
from ssm_b200 import q8, synthetic
dev = torch.device("cuda:0")
B, H, W = 16, 1080, 1920
x = synthetic.frames(2 * B, H, W, n_frames=1, seed=42, smooth=True, device=dev)
x = (x - x.amin()) / (x.amax() - x.amin())
images = (x.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
for _ in range(3): q8.quads_from_u8(images, order="rgb")
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20): q8.quads_from_u8(images, order="rgb")
e[1].record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("SSM_B200_LIB", "product"), "quads_from_u8_ms": e[0].elapsed_time(e[1]) / 20}))
