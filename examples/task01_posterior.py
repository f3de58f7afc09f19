"""task01's posterior tools on the B200 path (tasks/plot_task01.py:33-104 without the plotting):
self-normalised importance weights of L = 10,000 samples from q(z|x) for one image of a 2-D-latent IWAE,
the SNIS posterior mean, and sampling-importance-resampling of 200 indices - the same indices
`np.random.choice(L, 200, replace=False, p=al)` returns for the same RandomState.

    python examples/task01_posterior.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import iwae_b200 as ib                      # noqa: E402
from iwae_b200 import posterior             # noqa: E402


def main():
    P, X, _ = ib.data.synthetic_binarised(64)
    model = ib.IWAE(200, 2, out_bias=ib.data.get_bias(P.mean(0)))      # task01.py:84: n_latent = 2
    x, L = X[:1], 10000
    res = model(x, L)                                                   # plot_task01.py:36
    z = res["z"].numpy()[:, 0, :]                                       # [L, 2]
    log_w = (res["lpxz"].numpy() + res["lpz"].numpy() - res["lqzx"].numpy())[:, 0]      # plot_task01.py:78-83
    al = posterior.softmax_f32(log_w)                                   # scipy.special.softmax, on the device
    rs = np.random.RandomState(123)                                     # plot_task01.py:189
    idx = posterior.sir_from_randomstate(al, 200, rs)                   # plot_task01.py:88
    want = np.random.RandomState(123).choice(np.arange(L), size=200, replace=False, p=al.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), want)
    _, qzx = model.encoder(x, 1)                                        # plot_task01.py:40-42
    print("q(z|x) loc", qzx.loc.numpy()[0], "scale", qzx.scale.numpy()[0])
    print("SNIS posterior mean", res["snis_z"].numpy()[0], " effective sample size %.1f" % (1.0 / float((al ** 2).sum())))
    print("SIR: 200 of %d samples, mean of the resampled z" % L, z[idx.cpu().numpy()].mean(0), "(indices match numpy's choice)")


if __name__ == "__main__":
    main()
