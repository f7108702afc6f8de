import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_time import run
task, B = sys.argv[1], int(sys.argv[2])
for spec in sys.argv[3:]:
    env = dict(kv.split("=") for kv in spec.split(",") if kv)
    run(task, B, env, steps=int(os.environ.get("STEPS", "5")))
