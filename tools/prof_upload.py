import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytrace3_b200 import Context, abi
from raytrace3_b200.scene import make_scene
ctx = Context(0)
sc = make_scene("reference")
for i in range(3):
    t0 = time.perf_counter(); d = sc.flatten(); t1 = time.perf_counter()
    ctx.upload_scene(d); t2 = time.perf_counter()
    print(f"flatten {1e3*(t1-t0):.2f} ms, upload(desc) {1e3*(t2-t1):.2f} ms")
sc2 = make_scene("cornell", textures=None)
for i in range(2):
    t1 = time.perf_counter(); ctx.upload_scene(sc2); t2 = time.perf_counter()
    print(f"cornell upload {1e3*(t2-t1):.2f} ms")
import torch
x = torch.empty(31_000_000, dtype=torch.uint8)
y = torch.empty(31_000_000, dtype=torch.uint8, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(x); torch.cuda.synchronize(); print("torch pageable H2D 31MB ms", 1e3*(time.perf_counter()-t0))
xp = x.pin_memory()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(xp); torch.cuda.synchronize(); print("torch pinned H2D 31MB ms", 1e3*(time.perf_counter()-t0))
