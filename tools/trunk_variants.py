#!/usr/bin/env python
"""time the cuDNN image trunk (HookModule: ResNet + 1x1 convs) under a few torch settings, B=8, 640x480."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from dagr_b200.utils.args import default_args
from dagr_b200.model.dagr import DAGR

net = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
torch.manual_seed(0)
m = DAGR(default_args("s", batch_size=8, use_image=True, img_net=net), height=480, width=640).eval().cuda()
img = torch.rand(8, 3, 480, 640, device="cuda")
trunk = m.backbone.net


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ref = [f.float().clone() for f in trunk(img)[0]]
    print("default", t(lambda: trunk(img)))
    torch.backends.cudnn.benchmark = True
    print("cudnn.benchmark", t(lambda: trunk(img)))
    trunk_cl = trunk.to(memory_format=torch.channels_last)
    img_cl = img.contiguous(memory_format=torch.channels_last)
    print("channels_last", t(lambda: trunk_cl(img_cl)))
    out = [f.float() for f in trunk_cl(img_cl)[0]]
    print("max rel diff channels_last vs default", max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref)))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    print("fp32 strict (no TF32), channels_last", t(lambda: trunk_cl(img_cl)))
    out32 = [f.float() for f in trunk_cl(img_cl)[0]]
    print("max rel diff TF32 vs strict fp32", max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, out32)))
    torch.backends.cudnn.allow_tf32 = True
    with torch.autocast("cuda", dtype=torch.bfloat16):
        print("bf16 autocast channels_last", t(lambda: trunk_cl(img_cl)))
