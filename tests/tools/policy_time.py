"""Times the device-resident policy step (obs -> direct_DQN -> epsilon-greedy) for a batch; CUDA events, L2 not flushed (weights are meant to stay L2-resident)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import rollout as R
from oracle import rollout_oracle as RO
for B in [int(a) for a in (sys.argv[1:] or ["1024", "8192", "65536"])]:
    pol = R.DirectDQNPolicy(20)
    pol.load_state_dict(RO.policy_state_dict(1))
    mom = torch.randn((B, 20), dtype=torch.float64, device="cuda")
    for noise in (None, "philox"):
        def once(c):
            obs = R.observation(mom, 1.0)
            out = pol.forward(obs, noise=noise, counter=c, want_q=False)
            return pol.epsilon_greedy(out["greedy"], 0.1, counter=c)
        for c in range(5): once(c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for c in range(n): once(c)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fma = B * (20 * 512 + 512 * 512 + (2 if noise else 1) * (512 * 256 + 256 * 21))
        print("B=%6d noise=%-6s %8.1f us/step  %6.2f TFLOP/s fp32 (%.1f%% of 2*128*148*1.92e9)" % (B, noise, ms * 1e3, 2 * fma / ms / 1e9, 100 * 2 * fma / (ms * 1e-3) / (2 * 128 * 148 * 1.92e9)))
