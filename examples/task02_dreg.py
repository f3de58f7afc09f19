"""tasks/task02.py's driver (the doubly-reparameterised gradient estimator, reference CLI at task02.py:15-22) on the
B200-native IWAEDReG class: same loop (task02.py:200-259) - per-epoch binarisation, `model.train_step(x, k, beta,
optimizer)` without an objective argument, `val_step` on the test set every 200 steps, save_weights, k=5000 test LLH.

    python examples/task02_dreg.py --n_samples 50 --epochs 1 --n_train 2000 --n_test 200
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import iwae_b200 as ib                      # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--n_samples", type=int, default=5, help="number of importance samples")
parser.add_argument("--batch_size", type=int, default=20, help="batch size")
parser.add_argument("--epochs", type=int, default=-1, help="-1: the learning rate scheme from the paper (3280 epochs)")
parser.add_argument("--gpu", type=str, default="0", help="Choose GPU")
parser.add_argument("--n_train", type=int, default=60000)
parser.add_argument("--n_test", type=int, default=10000)
parser.add_argument("--eval_samples", type=int, default=5000)
parser.add_argument("--max_steps", type=int, default=-1)


def main(argv=None):
    args = parser.parse_args(argv)
    string = "task02_{0}".format(args.n_samples)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    np.random.seed(123)
    if args.epochs == -1:
        learning_rate_dict, epochs = ib.data.lr_schedule()
    else:
        epochs, learning_rate_dict = args.epochs, {0: 0.0001}
    (Xtrain, _), (Xtest, _) = ib.data.synthetic_mnist(args.n_train, args.n_test)
    Ntrain, Ntest = Xtrain.shape[0], Xtest.shape[0]
    steps_pr_epoch = Ntrain // args.batch_size
    total_steps = steps_pr_epoch * epochs
    model = ib.IWAEDReG(200, 100, out_bias=ib.data.get_bias(Xtrain.mean(axis=0)))      # task02.py:167-169
    optimizer = ib.Adam(learning_rate_dict[0], epsilon=1e-4)
    Xtest = ib.data.bernoullisample(Xtest)
    start, done = time.time(), False
    for epoch in range(epochs):
        Xb = ib.data.bernoullisample(Xtrain)
        perm = np.random.permutation(Ntrain)
        if args.epochs == -1 and epoch in learning_rate_dict:
            optimizer.learning_rate.assign(learning_rate_dict[epoch])
        for _step in range(steps_pr_epoch):
            step = _step + steps_pr_epoch * epoch
            beta = 1.0
            res = model.train_step(Xb[perm[_step * args.batch_size:(_step + 1) * args.batch_size]], args.n_samples, beta, optimizer)
            if step % 200 == 0:
                model.write_to_tensorboard(res, step)
                test_res = model.val_step(Xtest, args.n_samples, beta)
                took, start = time.time() - start, time.time()
                print("epoch {0}/{1}, step {2}/{3}, train ELBO: {4:.2f}, val ELBO: {5:.2f}, time: {6:.2f}"
                      .format(epoch, epochs, step, total_steps, res["iwae_elbo"].numpy(), test_res["iwae_elbo"], took))
            if 0 <= args.max_steps <= step:
                done = True
                break
        if done:
            break
    model.save_weights("/tmp/iwae/{0}/final_weights".format(string))
    per_example, total = model.llh(Xtest, args.eval_samples)          # replaces the per-image loop, task02.py:245-259
    llh = float(total) / Ntest
    print("Test-set {0} sample log likelihood estimate: {1:.4f}".format(args.eval_samples, llh))
    return llh


if __name__ == "__main__":
    main()
