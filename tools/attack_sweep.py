"""ALIE z sweep (Dist C): for every z the malicious rows 0..f-1 are replaced by mu - z*sigma on
the device, then Krum and Bulyan run; reports attack success and checks the indices against the oracle on
the same inputs.  usage: python tools/attack_sweep.py [n] [d] > profiles/r01_attack_sweep.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from attacking_federate_learning_b200 import defences as D, malicious as M, metrics
from oracle import ref_numpy as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
d = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
f = int(0.24 * n)
rng = np.random.default_rng(2026)
base = (0.1 * rng.standard_normal(d) + rng.standard_normal((n, d)) * np.exp(0.25 * rng.standard_normal((n, 1)))).astype(np.float32)
rows = []
for z in (0.25, 0.5, 1.0, 1.5, 2.0, 3.0):
    Gd = torch.from_numpy(base).cuda()
    honest_mean = Gd[f:].mean(0)
    M.DriftAttack(z).attack_rows(Gd, f)
    G = Gd.cpu().numpy()
    k = D.krum(Gd, n, f, return_index=True)
    agg, sel = D.bulyan(Gd, n, f, return_selection=True)
    t64 = orc.pairwise_distances_f64(G)
    k_ref, margin = orc.krum_select(t64, orc.visit_order(n), n, f, dtype=np.float64, with_margin=True)
    sel_ref = orc.bulyan_select(t64, n, f, dtype=np.float64)
    rows.append({"z": z, "krum_index": int(k), "krum_matches_oracle": int(k) == int(k_ref), "krum_margin": float(margin),
                 "krum_success": metrics.krum_attack_success(k, f),
                 "bulyan_matches_oracle": sel.cpu().tolist() == list(sel_ref),
                 "bulyan_malicious_fraction": metrics.bulyan_attack_success(sel.cpu().tolist(), f),
                 "bulyan_rel_deviation": metrics.relative_deviation(agg, honest_mean),
                 "krum_rel_deviation": metrics.relative_deviation(Gd[k], honest_mean)})
print(json.dumps({"n": n, "d": d, "f": f, "sweep": rows}, indent=1))
