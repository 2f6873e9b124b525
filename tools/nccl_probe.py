import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
print(subprocess.run("ip -o addr 2>/dev/null | awk '{print $2,$4}' ; env | grep -i nccl", shell=True, capture_output=True, text=True).stdout)
import numpy as np
if os.environ.get("PROBE_TORCH"):
    import torch
    print("torch imported", torch.__version__)
import __graft_entry__ as g
pkg = g.load_package()
syn = pkg.Synth("S3-hex", 8)
asm = pkg.Assembly(device=0)
try:
    asm.dist_init(0, 1, pkg.nccl_unique_id(), np.array([0, syn.nb], np.uint64))
    print("dist_init world=1 OK with env", os.environ.get("NCCL_SOCKET_IFNAME"))
except Exception as e:
    print("dist_init failed:", e)
