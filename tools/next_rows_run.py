"""Dev tool: a few launches of the next-row kernels (for an ncu capture)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, bench
a = argparse.Namespace(envs=65536)
print(bench.measure_record_transition(a, 6535.7))
print(bench.measure_minibatch_gather(a, 6535.7))
