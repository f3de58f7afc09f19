"""Run the child of tests/test_gpu_experiments.py several times per mode and print the parameter checksums."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_experiments as t
for kind in sys.argv[1:] or ["iwae1"]:
    for env in ({}, {}, {"IWAE_MEGA": "1"}, {"IWAE_MEGA": "1"}, {"IWAE_NO_GRAPH": "1"}, {"IWAE_MEGA": "1", "IWAE_NO_GRAPH": "1"},
                {"IWAE_NO_LANES": "1"}, {"IWAE_NO_PDL": "1"}):
        r = t.run_child(kind, **env)
        print(kind, env, repr(r["sum"]), r["elbos"][-1], flush=True)
