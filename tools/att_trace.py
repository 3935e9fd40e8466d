"""Dumps a clock64 timeline of CTA 0 of the attention kernel (first 64 key blocks): python tools/att_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
buf = torch.zeros(4 * 8 * 64, dtype=torch.int64, device="cuda")
os.environ["MD_ATT_TRACE_PTR"] = str(buf.data_ptr())
from musediffusion_b200 import ops
B, L, NH = 64, 2096, 12
qkv = (torch.randn(B * L, 3 * NH * 64, device="cuda") * 0.7).to(torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, B, L, NH)
torch.cuda.synchronize()
t = buf.cpu().view(4, 8, 64)
t0 = int(t[t > 0].min())
names = {0: ["qk_issue(j+1)", "pv_wait_begin", "pv_pfull_seen"], 1: ["qk_issue(j+1)", "pv_wait_begin", "pv_pfull_seen"],
         2: ["sfull_wait_begin", "sfull_seen", "ldtm_done", "max_done", "pre_turn", "turn_got", "exp_done", "pfull_arrived"],
         3: ["sfull_wait_begin", "sfull_seen", "ldtm_done", "max_done", "pre_turn", "turn_got", "exp_done", "pfull_arrived"]}
for step in range(17, 27):
    print("---- key block", step)
    ev = []
    for role in range(4):
        for e, n in enumerate(names[role]):
            v = int(t[role, e, step])
            if v:
                ev.append((v - t0, ["MMA0", "MMA1", "SM0 ", "SM1 "][role], n))
    for v, r, n in sorted(ev):
        print("  %8d  %s %s" % (v, r, n))
