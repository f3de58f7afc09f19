"""Which parameter tensors differ between the multi-kernel step and the megakernel after ONE eager step? (child processes)"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import iwae_b200 as ib
P, X, Y = ib.data.synthetic_binarised(120, seed=5)
m = ib.IWAE(200, 50, out_bias=ib.data.get_bias(P.mean(0)), seed=11)
opt = ib.Adam(1e-3, epsilon=1e-4)
x = torch.from_numpy(X[:20]).cuda()
g = m.grads(x, 10, 1.0, objective="iwae_elbo", seed=100)
torch.cuda.synchronize()
np.save(sys.argv[1], g.cpu().numpy())
names = [(v.name, v.offset, int(v.value.numel())) for v in m.trainable_weights]
import json; open(sys.argv[1] + ".json", "w").write(json.dumps(names))
""" % ROOT
outs = []
for tag, env in (("ref", {}), ("mega", {"IWAE_MEGA": "1"})):
    e = dict(os.environ); e.pop("IWAE_MEGA", None); e.update(env); e["IWAE_NO_GRAPH"] = "1"
    path = f"/tmp/g_{tag}.npy"
    r = subprocess.run([sys.executable, "-c", CHILD, path], env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    outs.append(np.load(path))
import json
names = json.load(open("/tmp/g_ref.npy.json"))
a, b = outs
print("max abs diff", np.max(np.abs(a - b)), "n differing", int((a != b).sum()), "of", a.size)
for name, off, n in names:
    d = np.abs(a[off:off + n] - b[off:off + n])
    if d.max() > 0:
        idx = np.nonzero(d)[0]
        print(f"{name:40s} differing {len(idx):6d}/{n}  max {d.max():.3e}  ref scale {np.abs(a[off:off+n]).max():.3e}  first idx {idx[:6]}")
