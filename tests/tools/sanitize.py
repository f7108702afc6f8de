"""One small batch through every kernel instantiation family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tests/tools/sanitize.py

Families: one-warp grid (chunk-Jacobi, per-trajectory tables), multi-warp grid generic width, large grid with its second line in global
memory, Fock harmonic / inverted harmonic (interface iteration), force-binned launch (shared table), the warp-specialised pipeline kernel,
the fused result exchange, the reset kernels, moments-only launch, and the policy / replay kernels either side of the path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, dist as qdist, rollout as R
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states

only = sys.argv[1:] or None


def case(name, params, B, n_sub, env=None, steps=1):
    if only and name not in only:
        return
    for k in list(os.environ):
        if k.startswith("QCART_") and k != "QCART_LIB":
            os.environ.pop(k)
    os.environ.update(env or {})
    params = dict(params, n_sub=n_sub)
    sim = BatchedSim(params, batch=B, seed=1)
    psi0 = initial_states(params, min(B, 16), 1)
    sim.set_state(np.tile(psi0, ((B + 15) // 16, 1))[:B])
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for _ in range(steps):
        act = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
        out = sim.step(act, want_q=True)
    sim.get_moments()
    torch.cuda.synchronize()
    print("%-22s %s  normdev %.1e" % (name, sim.kernel_info(), float((out["aux"][:, 3] - 1).abs().max())), flush=True)
    return sim


case("grid_one_warp", configs.quartic(), 9, 3)
case("grid_multi_warp", configs.inverted_quartic(), 5, 2)
case("grid_generic_257", configs.quartic_sweep(257), 3, 2)
case("grid_1025", configs.quartic_sweep(1025), 2, 2)
case("grid_big_vglobal", configs.quartic_sweep(8193), 1, 1)
case("fock_harmonic", configs.harmonic(), 9, 3)
case("fock_inv_harmonic", configs.inverted_harmonic(), 9, 3)
case("binned_one_warp", configs.quartic(), 1500, 2, {"QCART_BIN": "1"})
case("binned_fock_ih", configs.inverted_harmonic(), 1400, 2, {"QCART_BIN": "1"})
case("binned_multi_warp", configs.inverted_quartic(), 1400, 2, {"QCART_PIPE": "0"})
case("pipeline_iq", configs.inverted_quartic(), 1400, 2)
case("pipeline_one_warp", configs.quartic(), 1400, 2, {"QCART_PIPE": "2", "QCART_PIPE_NE": "4"})
if not only or "fused_gather" in only:
    params = configs.quartic(n_sub=2)
    sims = [BatchedSim(params, batch=12, seed=3, traj_offset=r * 12) for r in range(2)]
    for r, s in enumerate(sims):
        s.set_state(initial_states(params, 12, seed=10 + r))
    fgs = qdist.FusedGather.local_group(sims)
    for step in range(3):
        for s in sims:
            s.step(torch.zeros(12, dtype=torch.int32, device="cuda"))
        for fg in fgs:
            fg.wait()
    torch.cuda.synchronize()
    print("fused_gather           ok", float(fgs[0].block().abs().sum()), flush=True)
    for fg in fgs:
        fg.close()
if not only or "reset" in only:
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import QuantumCartpoleEnv
    env = QuantumCartpoleEnv("inverted_quartic", batch=8, seed=1, auto_reset=True, n_sub=2)
    env.reset()
    for _ in range(3):
        env.step(torch.full((8,), 20, dtype=torch.int64, device="cuda"))
    sim = BatchedSim(configs.quartic(n_sub=2), batch=4)
    sim.set_state(initial_states(configs.quartic(), 4, 1))
    out = sim.get_moments()
    pend = torch.ones(4, dtype=torch.uint8, device="cuda"); store = torch.zeros((4, sim.n), dtype=torch.complex128, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    sim.reset_accept(out["aux"], 7.5, pend, store, cnt)
    torch.cuda.synchronize()
    print("reset                  ok", pend.cpu().tolist(), flush=True)
if not only or "policy" in only:
    from oracle import rollout_oracle
    pol = R.DirectDQNPolicy(20, 21)
    pol.load_state_dict(rollout_oracle.policy_state_dict(1, n_in=20))
    obs = torch.randn(70, 20, device="cuda", dtype=torch.float32)
    res = pol.forward(obs, noise="philox", seed=1, traj_offset=0, counter=0)
    a, _ = pol.epsilon_greedy(res["greedy"], 0.3, seed=1, traj_offset=0, counter=0)
    ring = R.ReplayRing(20, 256)
    ring.push(obs, obs, a, torch.zeros(70, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    print("policy                 ok", ring.total(), flush=True)
