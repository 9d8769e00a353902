"""One profiled training step (yolov5l, 16 pairs of 640x512) for `ncu --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icafusion_b200 import Model, autograd, synth
from icafusion_b200.synth import load_synth
from icafusion_b200.trainer import TrainStep

size = sys.argv[1] if len(sys.argv) > 1 else "l"
B, H, W = int(sys.argv[2]) if len(sys.argv) > 2 else 16, 512, 640
dev = torch.device("cuda:0")
model = Model(f"yolov5{size}_Transfusion_kaist")
load_synth(model, 0)
model = model.to(dev).train()
ts = TrainStep(model, None, total_batch_size=B, imgsz=640)
rgb, ir = [(t * 255).to(torch.uint8).to(dev) for t in synth.synth_images(B, H, W, 0)]
tg = torch.zeros(4 * B, 6)
tg[:, 0] = torch.arange(4 * B) % B
tg[:, 2:4] = 0.1 + 0.8 * torch.rand(4 * B, 2)
tg[:, 4:6] = 0.03 + 0.2 * torch.rand(4 * B, 2)
tg = tg.to(dev)
for _ in range(2):
    ts(rgb, ir, tg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts(rgb, ir, tg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
