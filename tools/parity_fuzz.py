"""Parity fuzz: GPU (strict and hybrid) vs the oracle over many (M, r_obs) and ray mixes.
Prints one line per configuration; exits non-zero on any violation of the north-star bar that the
documented libm-sensitivity clause does not cover (see tests/test_gpu_binet.py)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lp_oracle as O
from light_path_tracer_b200.metrics import Schwarzschild
from test_gpu_binet import _check_batch

O.build()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    M = float(rng.choice([0.3, 1.0, 1.0, 2.5, 10.0]))
    r_obs = float(M * 10 ** rng.uniform(0.35, 3.2))
    n = 400000
    ac = float(O.alpha_crit(M, r_obs)) if r_obs > 2 * M else 0.3
    alpha = np.concatenate([rng.uniform(0, np.pi, n // 2), ac * (1 + rng.normal(0, 10 ** rng.uniform(-8, -1), n // 4)),
                            np.float64(np.float32(rng.uniform(0, min(6 * ac, np.pi), n // 4)))])
    fa_o, w_o, st_o, steps_o = O.trace_rays_batch(M, r_obs, alpha)
    d_a = torch.from_numpy(alpha).cuda()
    line = "M=%-5g r_obs=%-9.4g" % (M, r_obs)
    for flags in (0, 4):
        fa = torch.empty(alpha.size, dtype=torch.float64, device="cuda")
        w = torch.empty(alpha.size, dtype=torch.int64, device="cuda")
        Schwarzschild(M).trace_rays_batch(r_obs, d_a, fa, w, flags=flags)
        try:
            exempt, worst = _check_batch(M, r_obs, alpha, fa_o, w_o, fa.cpu().numpy(), w.cpu().numpy(), "fuzz", O)
            line += " | flags=%d ok: %d exempt, worst rel %.2e" % (flags, exempt, worst)
        except AssertionError as e:
            bad += 1
            line += " | flags=%d FAIL: %s" % (flags, str(e)[:200])
    print(line, flush=True)
print("violations:", bad)
sys.exit(1 if bad else 0)
