"""Dumps a clock64 timeline of CTA 0 of the attention kernel (first 64 key blocks): python tools/att_trace.py
Roles: 0/1 = MMA warps of Q tile 0/1, 2 + 4*tile + quad = softmax warps (one per TMEM lane quadrant)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
NR = 10
buf = torch.zeros(NR * 8 * 64, dtype=torch.int64, device="cuda")
os.environ["MD_ATT_TRACE_PTR"] = str(buf.data_ptr())
from musediffusion_b200 import ops
B, L, NH = 64, 2096, 12
qkv = (torch.randn(B * L, 3 * NH * 64, device="cuda") * 0.7).to(torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, B, L, NH)
torch.cuda.synchronize()
t = buf.cpu().view(NR, 8, 64)
t0 = int(t[t > 0].min())
import numpy as np
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/att_trace_raw.npy", (t - t0).numpy())
mma = ["qk_issue(j+1)", "pv_wait_begin", "pv_pfull_seen"]
sm = ["block_top", "enter", "s_in_regs", "max_done", "exp_begin", "-", "exp_done", "pfull_arr"]   # MD_TRACE event numbers
print("softmax warps: event times relative to the top of the warp's own key block; last column = period (traced build, kPoly = 2)")
for step in range(19, 25):
    print("---- key block", step)
    for role in range(2, NR):
        base = int(t[role, 0, step])
        nxt = int(t[role, 0, step + 1])
        row = " ".join("%s=%5d" % (sm[e], int(t[role, e, step]) - base) for e in (1, 2, 3, 4, 6, 7))
        print("  tile%d quad%d @%7d  %s  period=%d" % ((role - 2) // 4, (role - 2) % 4, base - t0, row, nxt - base))
    for role in range(2):
        print("  MMA%d  " % role + " ".join("%s=%7d" % (mma[e], int(t[role, e, step]) - t0) for e in range(3)))
print("==== work-item boundary (key blocks 15..19; a work item has 17 key blocks): absolute times")
for step in list(range(15, 20)) + list(range(32, 37)) + list(range(49, 54)):
    for role in (2, 6):
        print("  blk %2d tile%d quad0: " % (step, (role - 2) // 4) + " ".join("%s=%7d" % (sm[e], int(t[role, e, step]) - t0) for e in (0, 1, 2, 3, 4, 6, 7)))
    for role in range(2):
        print("  blk %2d MMA%d  " % (step, role) + " ".join("%s=%7d" % (mma[e], int(t[role, e, step]) - t0) for e in range(3)))
