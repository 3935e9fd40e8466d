"""`python -m musediffusion_b200 modification` on ONE and on TWO GPUs (torchrun, NCCL all-gather of the decoded ids) must write
the same tokens: batches are dealt round-robin to the ranks (run/sample.py:169-172) and the in-kernel noise is keyed by the
global sequence index, so the result does not depend on the GPU count.  python tools/cli_multi_gpu_check.py  (needs 2 GPUs)"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import musediff_oracle as O
from musediffusion_b200 import checkpoint as C

L = 200
d = tempfile.mkdtemp()
md = os.path.join(d, "m")
os.makedirs(md)
args = dict(C.MODEL_FIELDS, seq_len=L)
json.dump(args, open(os.path.join(md, "training_args.json"), "w"))
p = O.make_random_params(seed=31, seq_len=L)
torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items()}, os.path.join(md, "model_000000.pt"))
common = ["modification", "--model_path", os.path.join(md, "model_000000.pt"), "--step", "40", "--batch_size", "3", "--num_batches", "5",
          "--strength", "1.0"]
env = dict(os.environ, PYTHONPATH=ROOT)
r1 = subprocess.run([sys.executable, "-m", "musediffusion_b200"] + common + ["--out_dir", os.path.join(d, "one")], env=env,
                    capture_output=True, text=True, cwd=ROOT)
print(r1.stdout[-300:], r1.stderr[-600:] if r1.returncode else "")
r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                     "--master-port", "29533", "-m", "musediffusion_b200"] + common + ["--out_dir", os.path.join(d, "two")], env=env,
                    capture_output=True, text=True, cwd=ROOT)
print(r2.stdout[-300:], r2.stderr[-1200:] if r2.returncode else "")
sub = os.path.join("m", "model_000000.pt.modification.samples", "tokens.npy")
a, b = np.load(os.path.join(d, "one", sub)), np.load(os.path.join(d, "two", sub))
print("shapes", a.shape, b.shape, "identical:", bool(np.array_equal(a, b)))
assert r1.returncode == 0 and r2.returncode == 0 and a.shape == (15, L) and np.array_equal(a, b)
print("OK: 1-GPU and 2-GPU CLI runs decode to the same %d x %d token ids" % a.shape)
