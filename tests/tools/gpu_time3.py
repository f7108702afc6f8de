import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_time
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
task, B = sys.argv[1], int(sys.argv[2])
over = dict(kv.split("=") for kv in sys.argv[3].split(",") if kv) if len(sys.argv) > 3 else {}
orig = configs.PRESETS[task]
configs.PRESETS[task] = lambda **kw: orig(**{**{k: (int(v) if v.lstrip("-").isdigit() else float(v)) for k, v in over.items()}, **kw})
gpu_time.run(task, B, {}, steps=int(os.environ.get("STEPS", "10")))
