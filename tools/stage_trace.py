"""durations of the pipeline stages of one training step (events on the stage streams) vs the sum of their kernels"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
R = 8192
eng = NGPEngine(scale=0.5, n_rays=R, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, R, seed=1000)).to(dev)
for s in range(1, 4): eng.train_step_packed(pool[s % 8], global_step=s)
eng.capture()
for s in range(4, 800): eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
N = 30
E = lambda: torch.cuda.Event(enable_timing=True)
ev = [[E() for _ in range(6)] for _ in range(N)]
main = torch.cuda.current_stream(dev)
orig_g, orig_gb, orig_gm = eng._graph, eng._graph_back, eng._graph_march
class Wrap:
    def __init__(self, g, k0, k1, tbl): self.g, self.k0, self.k1, self.tbl = g, k0, k1, tbl
    def replay(self):
        i = Wrap.i
        self.tbl[i][self.k0].record(torch.cuda.current_stream(dev)); self.g.replay(); self.tbl[i][self.k1].record(torch.cuda.current_stream(dev))
eng._graph_march = Wrap(orig_gm, 0, 1, ev); eng._graph = Wrap(orig_g, 2, 3, ev); eng._graph_back = Wrap(orig_gb, 4, 5, ev)
adam_done = [E() for _ in range(N)]
for i, s in enumerate(range(801, 801 + N)):
    Wrap.i = i
    eng.train_step_packed(pool[s % 8], global_step=s * 16 + 1)       # no occupancy updates in the trace
    adam_done[i].record(eng._comm_stream)
torch.cuda.synchronize()
import numpy as np
rows = []
for i in range(2, N - 1):
    m = ev[i]
    rows.append([m[0].elapsed_time(m[1]), m[2].elapsed_time(m[3]), m[4].elapsed_time(m[5]), m[5].elapsed_time(adam_done[i]), m[3].elapsed_time(m[4]),
                 adam_done[i].elapsed_time(ev[i + 1][2]), ev[i][2].elapsed_time(ev[i + 1][2])])
r = np.array(rows) * 1e3
names = ["march graph", "field front graph", "field back graph", "adam (+hop)", "hop front->back", "hop adam->next front", "STEP (front to front)"]
print("samples", int(eng.counter[0]))
for k, n in enumerate(names): print(f"{n:24s} {r[:, k].mean():8.1f} us")
