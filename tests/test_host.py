"""Host-side logic that needs no GPU: presets, the Philox restatement, sharding, the gloo all-gather (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest
from math import pi

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, philox_normals, make_config, dist as qdist, _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = list(ctr); k = list(key)
    for r in range(10):
        if r > 0:
            k[0] = (k[0] + W0) & 0xFFFFFFFF; k[1] = (k[1] + W1) & 0xFFFFFFFF
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
    return c


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_library_noise_matches_python_restatement():
    for seed, traj, step in [(0, 0, 0), (12345, 7, 80), (2 ** 40 + 3, 65535, 10 ** 7)]:
        o = philox4x32_10([traj & 0xFFFFFFFF, traj >> 32, step & 0xFFFFFFFF, step >> 32], [seed & 0xFFFFFFFF, seed >> 32])
        u1 = (((o[1] << 32 | o[0]) >> 11) + 0.5) / 2 ** 53
        u2 = (((o[3] << 32 | o[2]) >> 11) + 0.5) / 2 ** 53
        r = np.sqrt(-2 * np.log(u1))
        want = np.array([r * np.cos(2 * pi * u2), r * np.sin(2 * pi * u2)])
        got = philox_normals(seed, traj, step)
        assert np.max(np.abs(got - want)) < 1e-14


def test_noise_statistics():
    v = np.array([philox_normals(3, t, s) for t in range(40) for s in range(100)])
    assert abs(v.mean()) < 0.05 and abs(v.std() - 1) < 0.05
    assert abs(np.corrcoef(v[:, 0], v[:, 1])[0, 1]) < 0.05


def test_presets_match_reference_defaults():
    q = configs.quartic()
    assert 2 * int(q["x_max"] / q["grid_size"] + 0.5) + 1 == 171 and q["n_sub"] == 1440 // 18
    iq = configs.inverted_quartic()
    assert 2 * int(iq["x_max"] / iq["grid_size"] + 0.5) + 1 == 521 and iq["n_sub"] == 2880 // 18
    assert abs(iq["x_threshold"] - 5.0) < 1e-9                       # (F_max*pi/(4|lambda|))^(1/3), IQ/main_parallel.py:150
    assert configs.harmonic()["n_max"] + 1 == 71 and configs.inverted_harmonic()["n_max"] + 1 == 181
    for npts in (257, 513, 1025, 2049, 4097, 8193):
        s = configs.quartic_sweep(npts)
        assert 2 * int(s["x_max"] / s["grid_size"] + 0.5) + 1 == npts
        assert abs(s["dt"] / s["grid_size"] ** 2 - (1 / 2880) / 0.05 ** 2) < 1e-9
    c = make_config(q)
    assert c.variant == L.QC_QUARTIC and c.n_levels == 21 and c.moment_order == 5


def test_shard_range_partitions_exactly():
    for total in (1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [qdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys, torch
sys.path.insert(0, %r)
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import dist as qdist
rank, local_rank, world = qdist.init_process_group("gloo")
assert world == 2
B, K = 6, 20
lo, hi = qdist.shard_range(world * B, rank, world)
mom = torch.arange(lo, hi, dtype=torch.float64).unsqueeze(1).repeat(1, K) + 0.25
aux = torch.full((B, 4), float(rank))
flags = torch.full((B,), rank + 1, dtype=torch.uint8)
blk = qdist.pack_block(mom, aux, flags)
allb = qdist.all_gather_block(blk, world)
m, a, f = qdist.unpack_block(allb, K)
assert allb.shape == (world * B, K + 5)
assert torch.equal(m[:, 0], torch.arange(world * B, dtype=torch.float64) + 0.25)
assert torch.equal(f, torch.tensor([1] * B + [2] * B, dtype=torch.uint8))
assert torch.equal(a[:, 0], torch.tensor([0.0] * B + [1.0] * B, dtype=torch.float64))
torch.distributed.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_all_gather_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
