"""Driver with the reference's command line (nbip/IWAE main.py:16-25) on the B200-native model classes.

    python main.py --n_samples 5 --objective iwae_elbo --stochastic_layers 1 --batch_size 20 --epochs 1

The reference loads MNIST through keras (needs the network) - here the synthetic binarised stand-in of
iwae_b200.data is used instead (`--n_train/--n_test` shrink it for a quick run).  Everything the reference's
loop does with the model goes through the same calls: `model.train_step`, `model.val_step`, `model(x, L)`
and the Keras-style Adam with the paper's learning-rate schedule (main.py:44-56,128-133).  The k=5000
test-set log-likelihood uses the fused evaluation (`model.llh`) unless `--reference_eval_loop` asks for the
reference's per-image loop (main.py:170-184) - both give the same estimator.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "src"))
import utils    # noqa: E402
import iwae1    # noqa: E402
import iwae2    # noqa: E402
from iwae_b200 import Adam  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--stochastic_layers", type=int, default=1, choices=[1, 2], help="1 -> iwae1.IWAE, 2 -> iwae2.IWAE")
parser.add_argument("--n_samples", type=int, default=5, help="k, importance samples per image")
parser.add_argument("--batch_size", type=int, default=20, help="images per train step")
parser.add_argument("--epochs", type=int, default=-1,
                    help="-1: the 3280-epoch schedule of Burda et al. (learning rate decays at epochs 0,1,4,13,...)")
parser.add_argument("--objective", type=str, default="iwae_elbo", choices=["vae_elbo", "iwae_elbo", "iwae_eq14", "vae_elbo_kl"])
parser.add_argument("--gpu", type=str, default="0", help="CUDA_VISIBLE_DEVICES")
parser.add_argument("--n_train", type=int, default=60000)
parser.add_argument("--n_test", type=int, default=10000)
parser.add_argument("--eval_samples", type=int, default=5000)
parser.add_argument("--reference_eval_loop", action="store_true")
parser.add_argument("--eval_precision", type=str, default="auto", choices=["auto", "fp32", "tc_f16", "tc_f16_fast"],
                    help="arithmetic of the final k-sample test LLH: auto = the fused tensor-core kernel in its fast mode "
                         "(mean-LLH bias vs fp32 measured < 1e-3 nats on 1,000 images, tests/test_gpu_hardening.py), "
                         "tc_f16 = libm-accurate epilogues, fp32 = the FFMA parity kernels")
parser.add_argument("--max_steps", type=int, default=-1, help="stop early (smoke runs)")
parser.add_argument("--device_data", action="store_true",
                    help="keep the train set in HBM: per-epoch Philox re-binarisation + shuffle on the device, no H2D per step")
parser.add_argument("--fused_monitor", action="store_true",
                    help="monitor the test set through the fused evaluation kernel (iwae_elbo / vae_elbo only, beta = 1)")


def main(argv=None):
    args = parser.parse_args(argv)
    print(args)
    string = "main_{0}_{1}_{2}".format(args.objective, args.stochastic_layers, args.n_samples)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)

    np.random.seed(123)
    if args.epochs == -1:
        learning_rate_dict, epochs = utils._data.lr_schedule()
    else:
        epochs = args.epochs
        learning_rate_dict = {0: 0.0001}

    (Xtrain, ytrain), (Xtest, ytest) = utils.synthetic_mnist(args.n_train, args.n_test)
    Ntrain, Ntest = Xtrain.shape[0], Xtest.shape[0]
    n_samples, batch_size, objective = args.n_samples, args.batch_size, args.objective
    steps_pr_epoch = Ntrain // batch_size
    total_steps = steps_pr_epoch * epochs

    bias = utils.get_bias(Xtrain.mean(axis=0))
    if args.stochastic_layers == 1:
        model = iwae1.IWAE(200, 100, out_bias=bias)
    else:
        model = iwae2.IWAE([200, 100], [100, 50], out_bias=bias)
    optimizer = Adam(learning_rate_dict[0], epsilon=1e-4)
    print("Initial learning rate: ", optimizer.learning_rate.numpy())

    Xtest = utils.bernoullisample(Xtest)
    start = time.time()
    step = 0
    done = False
    device_data = utils.DeviceDataset(Xtrain, batch_size=batch_size, seed=123) if args.device_data else None
    for epoch in range(epochs):
        if device_data is not None:
            device_data.start_epoch(epoch)
        else:
            Xtrain_binarized = utils.bernoullisample(Xtrain)
            perm = np.random.permutation(Ntrain)      # tf.data shuffle(Ntrain).batch(batch_size)
        if args.epochs == -1 and epoch in learning_rate_dict:
            new_learning_rate = learning_rate_dict[epoch]
            old_learning_rate = optimizer.learning_rate.numpy()
            print("Changing learning rate from {0} to {1}".format(old_learning_rate, new_learning_rate))
            optimizer.learning_rate.assign(new_learning_rate)
        for _step in range(steps_pr_epoch):
            step = _step + steps_pr_epoch * epoch
            if device_data is not None:
                x_batch, _ = device_data.batch(_step)
            else:
                x_batch = Xtrain_binarized[perm[_step * batch_size:(_step + 1) * batch_size]]
            beta = 1.0
            res = model.train_step(x_batch, n_samples, beta, optimizer, objective=objective)
            if step % 200 == 0:
                model.write_to_tensorboard(res, step)
                if args.fused_monitor and objective in ("iwae_elbo", "vae_elbo") and beta == 1.0:
                    test_res = model.monitor(Xtest, n_samples)
                else:
                    test_res = model.val_step(Xtest, n_samples, beta)
                    model.write_to_tensorboard(test_res, step)
                took = time.time() - start
                start = time.time()
                print("epoch {0}/{1}, step {2}/{3}, train ELBO: {4:.2f}, val ELBO: {5:.2f}, time: {6:.2f}"
                      .format(epoch, epochs, step, total_steps, res[objective].numpy(), test_res[objective], took))
            if 0 <= args.max_steps <= step:
                done = True
                break
        if done:
            break

    model.save_weights("/tmp/iwae/{0}/final_weights".format(string))

    L = args.eval_samples
    if args.reference_eval_loop:
        test_elbo_metric = utils.MyMetric()
        for i, x in enumerate(Xtest):
            res = model(x[None, :].astype(np.float32), L)
            test_elbo_metric.update_state(res["iwae_elbo"][None, None])
            if i % 200 == 0:
                print("{0}/{1}".format(i, Ntest))
        test_set_llh = test_elbo_metric.result()
    else:
        per_example, total = model.llh(Xtest, L, precision=args.eval_precision)
        test_set_llh = float(total) / Ntest
    print("Test-set {0} sample log likelihood estimate: {1:.4f}".format(L, test_set_llh))
    return test_set_llh


if __name__ == "__main__":
    main()
