"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports /root/reference/defences.py and malicious.py (NumPy-only modules), evaluates them on small
seeded inputs plus the edge cases listed in SURVEY.md section 4, and writes tests/golden/golden_v1.npz.
The fixtures are what pins oracle/ref_numpy.py (tests/test_oracle_golden.py) and, through it, the
CUDA path (tests/test_gpu_*.py read the same file and compare directly as well).
"""
import os
import sys

import numpy as np

REF = os.environ.get("AFL_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REF)
import defences as ref_def      # noqa: E402
import malicious as ref_mal     # noqa: E402


class _User:
    def __init__(self, g):
        self.grads = g
        self.original_params = None
        self.learning_rate = None


def hetero(rng, n, d):
    mu = 0.1 * rng.standard_normal(d)
    s = np.exp(0.25 * rng.standard_normal(n))
    return (mu[None, :] + s[:, None] * rng.standard_normal((n, d))).astype(np.float32)


def alie_rows(G, f, z):
    att = ref_mal.DriftAttack(z)
    users = [_User(G[i].copy()) for i in range(f)]
    att.attack(users)
    out = G.copy()
    for i in range(f):
        out[i] = users[i].grads
    return out


def main():
    rng = np.random.default_rng(20260922)
    cases = {}
    meta = []

    def add(name, G, f, rules=("krum", "tm", "bulyan", "mean")):
        n = G.shape[0]
        cases[f"{name}/G"] = G
        cases[f"{name}/f"] = np.int64(f)
        dist = ref_def._krum_create_distances(G)
        table = np.zeros((n, n), np.float32)
        for i in dist:
            for j in dist[i]:
                table[i, j] = dist[i][j]
        cases[f"{name}/dist"] = table
        cases[f"{name}/order"] = np.array(list(dist.keys()), np.int64)
        if "krum" in rules:
            cases[f"{name}/krum_idx"] = np.int64(ref_def.krum(G, n, f, return_index=True))
        if "tm" in rules:
            cases[f"{name}/tm"] = ref_def.trimmed_mean(G, n, f)
        if "bulyan" in rules and n >= 4 * f + 3:
            # replay bulyan's loop to record the selection sequence, then the real call for the output
            d2 = ref_def._krum_create_distances(G)
            sel = []
            while len(sel) < n - 2 * f:
                idx = ref_def.krum(G, n - len(sel), f, d2, True)
                sel.append(idx)
                d2.pop(idx)
                for u in d2:
                    d2[u].pop(idx)
            cases[f"{name}/bulyan_sel"] = np.array(sel, np.int64)
            cases[f"{name}/bulyan"] = ref_def.bulyan(G, n, f)
        if "mean" in rules:
            cases[f"{name}/mean"] = ref_def.no_defense(G, n, f)
        meta.append(name)

    # random shapes, iid and heterogeneous
    for n, d, f in [(4, 7, 0), (5, 16, 1), (7, 33, 1), (10, 64, 2), (11, 50, 2), (16, 40, 3),
                    (23, 29, 5), (32, 24, 7), (12, 96, 2)]:
        add(f"iid_n{n}_d{d}_f{f}", rng.standard_normal((n, d)).astype(np.float32), f)
        add(f"het_n{n}_d{d}_f{f}", hetero(rng, n, d), f)
    # ALIE-shaped inputs: rows 0..f-1 identical (exact ties inside Krum / Bulyan)
    for n, d, f, z in [(10, 48, 2, 1.5), (16, 40, 3, 0.5), (23, 32, 5, 1.0), (31, 20, 7, 2.0)]:
        add(f"alie_n{n}_d{d}_f{f}", alie_rows(hetero(rng, n, d), f, z), f)
    # rows 0,1,2 identical and clearly the best -> reference answers index 1
    G = 5.0 * hetero(rng, 9, 32); G[0] = G[1] = G[2] = 0.002 * G[3]
    add("tie012_win", G.astype(np.float32), 2)
    # trimmed-mean tie cases from SURVEY section 4
    add("tm_tie_a", np.array([[0.], [-1.], [1.], [5.], [-5.]], np.float32), 2, rules=("tm",))
    add("tm_tie_b", np.array([[0.], [1.], [-1.], [5.], [-5.]], np.float32), 2, rules=("tm",))
    add("tm_even", np.array([[1., 4.], [2., 3.], [3., 2.], [10., 1.]], np.float32), 1, rules=("tm", "mean"))
    add("tm_f0", np.array([[1.], [2.], [3.], [100.], [4.]], np.float32), 0, rules=("tm",))
    # bf16-representable values (many exact |x-med| ties), as the bf16 config feeds the reference
    Gb = hetero(rng, 24, 40)
    Gb = (Gb.view(np.uint32) + 0x8000 & 0xFFFF0000).astype(np.uint32).view(np.float32)
    add("bf16vals_n24", Gb, 5)
    # degenerate sizes
    add("n1", rng.standard_normal((1, 8)).astype(np.float32), 0, rules=("krum", "mean"))
    add("n2", rng.standard_normal((2, 8)).astype(np.float32), 0, rules=("krum", "mean", "tm"))
    add("n3", rng.standard_normal((3, 8)).astype(np.float32), 0, rules=("krum", "mean", "tm", "bulyan"))

    # ALIE attack itself
    for nm, f, d, z in [("a", 3, 17, 1.5), ("b", 8, 64, 0.25), ("c", 1, 9, 2.0), ("d", 5, 33, 0.0)]:
        rows = hetero(rng, f, d)
        users = [_User(rows[i].copy()) for i in range(f)]
        att = ref_mal.DriftAttack(z)
        att.attack(users)
        cases[f"alie_{nm}/rows"] = rows
        cases[f"alie_{nm}/z"] = np.float64(z)
        cases[f"alie_{nm}/mean"] = att.grads_mean
        cases[f"alie_{nm}/stdev"] = att.grads_stdev
        cases[f"alie_{nm}/grads0"] = users[0].grads
        cases[f"alie_{nm}/aliased"] = np.bool_(all(u.grads is users[0].grads for u in users))

    cases["__names__"] = np.array(meta)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes,", len(meta), "defence cases")


if __name__ == "__main__":
    main()
