"""Repeated one-shot encodes, sessions and batches: device and host memory must level off."""
import os, sys, resource
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import draco_oxide_b200 as dxo
from draco_oxide_b200 import synth
m = synth.config2_mesh(400)
small = [synth.batch_mesh(k, 3000 + 500 * k) for k in range(32)]
def snap(tag):
    free, total = torch.cuda.mem_get_info()
    print(f"{tag:28s} device used {(total - free) / 2**20:8.1f} MiB   host RSS {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024:8.1f} MiB")
snap("start")
for r in range(4):
    for _ in range(15):
        out = bytearray(); dxo.encode(m, out)
    snap(f"after {15 * (r + 1)} encodes")
for r in range(3):
    for _ in range(5):
        s = dxo.Session(m); s.run_steps(3); s.close()
    snap(f"after {5 * (r + 1)} sessions")
for r in range(3):
    for _ in range(5):
        dxo.encode_batch(small)
    snap(f"after {5 * (r + 1)} batches")
