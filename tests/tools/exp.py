"""Development aid: time the control-step kernel for a list of experiments given on the command line.

    python tests/tools/exp.py iq:8192 QCART_STAGGER=0 QCART_STAGGER=4000 "QCART_STAGGER=8000 QCART_T=3"

First argument = task[:batch[:steps]] (iq | q | ih | h), every further argument = one experiment (space-separated KEY=VALUE env settings).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import gpu_time

TASK = {"iq": "inverted_quartic", "q": "quartic", "ih": "inverted_harmonic", "h": "harmonic"}
KEYS = ("QCART_PIPE_NE", "QCART_PIPE_L", "QCART_L", "QCART_T", "QCART_P", "QCART_TABS", "QCART_GC", "QCART_JACOBI", "QCART_BIN", "QCART_MAXT", "QCART_STAGGER", "QCART_DEBUG", "QCART_XFER", "QCART_COOP", "QCART_PIPE")

if __name__ == "__main__":
    spec = sys.argv[1].split(":")
    task = TASK.get(spec[0], spec[0]); B = int(spec[1]) if len(spec) > 1 else 8192; steps = int(spec[2]) if len(spec) > 2 else 5
    exps = sys.argv[2:] or [""]
    for e in exps:
        for k in KEYS:
            os.environ.pop(k, None)
        env = dict(kv.split("=") for kv in e.split()) if e.strip() else {}
        gpu_time.run(task, B, env, steps=steps)
