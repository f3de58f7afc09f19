"""tasks/task04.py's driver (conditional IWAE with a label-conditioned learned prior; `--std_prior` gives task05's
N(0,1) prior) on the B200-native CIWAE class with the reference CLI (task04.py:34-43): per-epoch binarisation,
`model.train_step(x, y, k, beta, optimizer, objective=...)`, `val_step(Xtest, ytest, k, beta)` every 200 steps,
save_weights, k=5000 test LLH (task04.py:300-373).

    python examples/task04_ciwae.py --n_samples 50 --epochs 1 --n_train 2000 --n_test 200
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import iwae_b200 as ib                      # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--stochastic_layers", type=int, default=1, choices=[1], help="only the 1-layer CIWAE exists (task04.py:277 crashes for 2)")
parser.add_argument("--n_samples", type=int, default=5, help="number of importance samples")
parser.add_argument("--batch_size", type=int, default=20, help="batch size")
parser.add_argument("--epochs", type=int, default=-1, help="-1: the learning rate scheme from the paper (3280 epochs)")
parser.add_argument("--objective", type=str, default="iwae_elbo", choices=["vae_elbo", "iwae_elbo", "iwae_eq14", "vae_elbo_kl"])
parser.add_argument("--gpu", type=str, default="0", help="Choose GPU")
parser.add_argument("--std_prior", action="store_true", help="tasks/task05.py: N(0,1) prior instead of the learned p(z|y)")
parser.add_argument("--n_train", type=int, default=60000)
parser.add_argument("--n_test", type=int, default=10000)
parser.add_argument("--eval_samples", type=int, default=5000)
parser.add_argument("--max_steps", type=int, default=-1)


def main(argv=None):
    args = parser.parse_args(argv)
    string = "{0}_{1}_{2}_{3}".format("task05" if args.std_prior else "task04", args.objective, args.stochastic_layers, args.n_samples)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    np.random.seed(123)
    if args.epochs == -1:
        learning_rate_dict, epochs = ib.data.lr_schedule()
    else:
        epochs, learning_rate_dict = args.epochs, {0: 0.0001}
    (Xtrain, ytrain), (Xtest, ytest) = ib.data.synthetic_mnist(args.n_train, args.n_test)
    ytrain, ytest = np.asarray(ytrain, dtype=np.float32), np.asarray(ytest, dtype=np.float32)
    Ntrain, Ntest = Xtrain.shape[0], Xtest.shape[0]
    steps_pr_epoch = Ntrain // args.batch_size
    total_steps = steps_pr_epoch * epochs
    model = ib.CIWAE(200, 100, learned_prior=not args.std_prior, out_bias=ib.data.get_bias(Xtrain.mean(axis=0)))
    optimizer = ib.Adam(learning_rate_dict[0], epsilon=1e-4)
    Xtest = ib.data.bernoullisample(Xtest)
    start, done = time.time(), False
    for epoch in range(epochs):
        Xb = ib.data.bernoullisample(Xtrain)
        perm = np.random.permutation(Ntrain)                         # tf.data shuffle of (x, y) pairs, task04.py:306-309
        if args.epochs == -1 and epoch in learning_rate_dict:
            optimizer.learning_rate.assign(learning_rate_dict[epoch])
        for _step in range(steps_pr_epoch):
            step = _step + steps_pr_epoch * epoch
            idx = perm[_step * args.batch_size:(_step + 1) * args.batch_size]
            beta = 1.0
            res = model.train_step(Xb[idx], ytrain[idx], args.n_samples, beta, optimizer, objective=args.objective)
            if step % 200 == 0:
                model.write_to_tensorboard(res, step)
                test_res = model.val_step(Xtest, ytest, args.n_samples, beta)
                took, start = time.time() - start, time.time()
                print("epoch {0}/{1}, step {2}/{3}, train ELBO: {4:.2f}, val ELBO: {5:.2f}, time: {6:.2f}"
                      .format(epoch, epochs, step, total_steps, res[args.objective].numpy(), test_res[args.objective], took))
            if 0 <= args.max_steps <= step:
                done = True
                break
        if done:
            break
    model.save_weights("/tmp/iwae/{0}/final_weights".format(string))
    per_example, total = model.llh(Xtest, args.eval_samples, y=ytest)     # replaces the per-image loop, task04.py:359-373
    llh = float(total) / Ntest
    print("Test-set {0} sample log likelihood estimate: {1:.4f}".format(args.eval_samples, llh))
    return llh


if __name__ == "__main__":
    main()
