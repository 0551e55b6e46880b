import numpy as np, glob
fs = sorted(glob.glob("gpurun_out/trig_*.npz"))
d = {f: np.load(f) for f in fs}
base = d[fs[0]]
for f in fs:
    for k in base.files:
        print(f, k, "differs from", fs[0], int((d[f][k] != base[k]).sum()))
