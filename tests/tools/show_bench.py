"""Print the key numbers of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.4g %s  ms/step %.4f (median5 %s)  kernel_ms %.4f  frac %.3f  e2e %.4g  n_gpus %d" % (d["value"], d["unit"], d["ms_per_step"], d.get("ms_per_step_median_of_5"),
      d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["n_gpus"]))
print("check", d.get("check"), "clocks", d.get("clocks"))
for k, v in (d.get("extra") or {}).items():
    for r in (v if isinstance(v, list) else [v]):
        print("  extra.%s: %.4g/s  %.3f ms  frac %.3f  global %d | %s" % (k, r["value"], r["ms_per_step"], r["roofline"]["frac"], r["global_trajectories"], r["roofline"]["kernel"][:80]))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]; print("cpu_baseline %.4g (%d cores) per core %.4g, without reset_ab %.4g" % (c["value"], c["cores"], c.get("value_per_core", 0), c.get("value_per_core_without_reset_ab", 0)))
if d.get("closed_loop"):
    print("closed_loop %.4g" % d["closed_loop"]["value"])
