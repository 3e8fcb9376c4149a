import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bosip_b200 as bb
ctx = bb.Context(0)
rng = np.random.default_rng(1)
for G in (10**5, 10**6, 10**7, 10**8):
    lw = torch.as_tensor(rng.normal(0, 0.1, G), device="cuda"); a = torch.as_tensor(rng.normal(-3, 2, G), device="cuda")
    t = a + 0.5 * torch.randn(G, dtype=torch.float64, device="cuda")
    for _ in range(3): v = bb.calc_tv(lw, a, t, ctx)
    reps = 30 if G < 10**8 else 10
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): v = bb.calc_tv(lw, a, t, ctx)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"G": G, "us": dt * 1e6, "algorithmic_GBps": 24 * G / dt * 1e-9, "actual_traffic_frac_of_hbm_6451": 48 * G / dt * 1e-9 / 6451.2}), flush=True)
