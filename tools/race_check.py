"""Developer check: bitwise reproducibility of the fused evaluation with two evaluations in flight (fresh operators per
step, as bench.py's e2e leg does). python tools/race_check.py [n] [reps]"""
import os, sys
import numpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gaussian-process-param-estimation_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench
from gaussian_proc import generate_correlation
from gaussian_proc._mixed_correlation import MixedCorrelation
from gaussian_proc._likelihood import _fused

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fresh = (len(sys.argv) <= 3) or sys.argv[3] != 'reuse'
pts, z, X = bench.make_inputs(n)
streams = [torch.cuda.Stream() for _ in range(2)]
cells = [(0.01, 0.104), (0.1, 0.1)]
outs = {c: [] for c in cells}
pending = []
keep = {}
for r in range(reps):
    for k, (eta, rho) in enumerate(cells):
        if len(pending) >= 2:
            c, h = pending.pop(0)
            out = h[0].cpu().numpy().copy() if h[5] is not None else h[0]
            h[5].synchronize()
            outs[c].append(h[0].cpu().numpy().copy())
        with torch.cuda.stream(streams[k]):
            if fresh or k not in keep:
                K = generate_correlation(pts, rho, 2.5, device=True)
                keep[k] = MixedCorrelation(K)
            pending.append(((eta, rho), _fused.evaluate_async(z, X, keep[k], eta, traceinv=True, drho=True)))
for c, h in pending:
    h[5].synchronize()
    outs[c].append(h[0].cpu().numpy().copy())
for c in cells:
    ref = outs[c][0]
    print(c, 'info', [int(o[4]) for o in outs[c]], 'max rel dev vs first',
          max(float(numpy.max(numpy.abs(o - ref) / (numpy.abs(ref) + 1e-300))) for o in outs[c]), 'logdet', ref[0])
