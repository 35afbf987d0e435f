"""Run the C2 pipeline once on a few frames (for ncu captures):  python tools/run_stage.py [frames]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import timg_b200
from timg_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
IW, IH = 3840, 2160
_, ow, oh = timg_b200.calc_fit(IW, IH, 2700, 1800, 9, 18)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx = timg_b200.Context(0, stream.cuda_stream)
frames = torch.stack([synth.frame_torch(1234 + i, IW, IH, "photo", dev) for i in range(F)])
b = timg_b200.Batch(n_frames=F, src_w=IW, src_h=IH, src_fmt=0, out_w=ow, out_h=oh, has_bg=1, bg=0xff000000, pattern=0,
                    pattern_w=0, pattern_h=0, flags=0, x_indent_cells=0, animation=0)
cap = F * 6 * 1024 * 1024
out = torch.empty(cap, dtype=torch.uint8, device=dev)
offs = torch.zeros(F + 1, dtype=torch.int64, device=dev)
L = timg_b200.lib()
for _ in range(2):
    rc = L.b200timg_sixel_batch_dev(ctx.h, C.byref(b), frames.data_ptr(), out.data_ptr(), cap, offs.data_ptr())
    assert rc == 0, L.b200timg_last_error(ctx.h)
torch.cuda.synchronize()
print("bytes/frame", int(offs[-1]) // F)
