"""Where does bench.py's c2 sub-result lose 6 % against the same steps timed alone?  (diagnostic)"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
pkg = entry.load_package(); gen = pkg.generators
b = gen.round_to_float(gen.uniform_cube(16384, 3, seed=44))
mode = sys.argv[1] if len(sys.argv) > 1 else "default"

def c2(tag):
    out = {"mode": mode, "state": tag}
    for pdl in (1,):
        with pkg.NBodyCuda(3, 16384, 32) as ctx:
            ctx.upload(b); ctx.step(1e-4, 400); ctx.upload(b)
            ms, wall = [], []
            for _ in range(3):
                t0 = time.perf_counter(); ctx.step(1e-4, 100); wall.append((time.perf_counter() - t0) * 10)
                ms.append(ctx.last_elapsed_ms / 100)
            out[f"pdl={pdl}"] = round(min(ms), 4)
            out[f"wall_ms_per_step pdl={pdl}"] = round(min(wall), 4)
    print(json.dumps(out), flush=True)

c2("fresh")
import torch
if mode == "side":
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        x = torch.empty(1 << 20, dtype=torch.uint8, device="cuda"); x.zero_()
    torch.cuda.synchronize()
elif mode == "memset":
    x = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.cudart().cudaMemset(x.data_ptr(), 0, 1 << 20) if hasattr(torch.cuda.cudart(), "cudaMemset") else None
    torch.cuda.synchronize()
else:
    x = torch.empty(1 << 20, dtype=torch.uint8, device="cuda"); x.zero_(); torch.cuda.synchronize()
c2("after a torch kernel")
