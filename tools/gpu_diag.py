"""GPU diagnostic battery (development tool, run under gpurun).  Every case runs in its own
subprocess with a timeout so that a trapping kernel cannot take the others down.  Results are
appended to gpurun_out/diag.jsonl.

    python tools/gpu_diag.py            # run everything
    python tools/gpu_diag.py case NAME  # run one case in-process (used by the parent)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def hetero(n, d, seed=0, device="cuda"):
    import torch
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    mu = 0.1 * torch.randn(d, generator=g, device=device)
    s = torch.exp(0.25 * torch.randn(n, 1, generator=g, device=device))
    return (mu[None, :] + s * torch.randn(n, d, generator=g, device=device)).float().contiguous()


def d2_exact(G, chunk=1 << 18):
    """float64 squared distances of the fp32 rows, via fp64 Gram on the GPU, chunked over columns."""
    import torch
    n, d = G.shape
    acc = torch.zeros((n, n), dtype=torch.float64, device=G.device)
    for c0 in range(0, d, chunk):
        X = G[:, c0:c0 + chunk].double()
        # direct differences for accuracy when n is small, Gram otherwise
        if n <= 128:
            for i in range(n):
                diff = X[i][None, :] - X
                acc[i] += (diff * diff).sum(dim=1)
        else:
            S = X @ X.T
            dg = torch.diagonal(S)
            acc += dg[:, None] + dg[None, :] - 2 * S
    acc.fill_diagonal_(0)
    return acc


def timeit(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------------------------------------------------------------------------------------
def case_gram(n, d, flags, label, time_it=False, seed=0):
    import torch
    from attacking_federate_learning_b200 import _device as dev
    G = hetero(n, d, seed)
    ex = d2_exact(G)
    d2 = dev.sqdist_partial(G, flags)
    torch.cuda.synchronize()
    off = ~torch.eye(n, dtype=torch.bool, device=G.device)
    rel = ((d2 - ex).abs() / ex.clamp_min(1e-30))[off] if n > 1 else torch.zeros(1)
    res = {"n": n, "d": d, "flags": flags, "label": label, "max_rel": float(rel.max()), "mean_rel": float(rel.mean()),
           "signed_mean_rel": float((((d2 - ex) / ex.clamp_min(1e-30))[off]).mean()) if n > 1 else 0.0,
           "sym": float((d2 - d2.T).abs().max()), "diag": float(torch.diagonal(d2).abs().max()),
           "flush": os.environ.get("AFL_GRAM_FLUSH", "default"), "splits": os.environ.get("AFL_GRAM_SPLITS", "default")}
    if time_it:
        res["ms"] = timeit(lambda: dev.sqdist_partial(G, flags), iters=10)
        res["GBps"] = n * d * 4 / res["ms"] / 1e6
    return res


def case_gram_identical(n, d, flags):
    """rows 0..4 identical: their table rows must be bit-identical and mutual distances exactly 0."""
    import torch
    from attacking_federate_learning_b200 import _device as dev
    G = hetero(n, d, 3)
    G[1:5] = G[0]
    d2 = dev.sqdist_partial(G, flags)
    dist = dev.sqdist_to_dist(d2)
    same = all(torch.equal(torch.cat([dist[0, 5:]]), torch.cat([dist[i, 5:]])) for i in range(1, 5))
    return {"n": n, "d": d, "flags": flags, "zero_block": float(d2[:5, :5].abs().max()), "rows_identical": bool(same)}


def case_tf32_behaviour():
    """Does kind::tf32 truncate or round the low 13 mantissa bits of fp32 operands?  One row with
    value 1 + 2^-11 + 2^-12 against ones: truncation gives 1.0, round-to-nearest gives 1 + 2^-10."""
    import torch
    from attacking_federate_learning_b200 import _device as dev, _native as nat
    n, d = 16, 32
    G = torch.zeros((n, d), device="cuda")
    G[0, :] = 1.0
    G[1, :] = 1.0 + 2.0 ** -11 + 2.0 ** -12
    out = {}
    for name, flags in [("single", nat.GRAM_FORCE_TCGEN05 | nat.GRAM_SINGLE_PASS), ("split", nat.GRAM_FORCE_TCGEN05),
                        ("split_rewrite", nat.GRAM_FORCE_TCGEN05 | nat.GRAM_REWRITE_HI)]:
        d2 = dev.sqdist_partial(G, flags)
        # d2[0,1] = d * (x-1)^2 ; with truncated operand x->1 => 0
        out[name] = {"d2_01": float(d2[0, 1]), "expected": d * (2.0 ** -11 + 2.0 ** -12) ** 2, "d2_02": float(d2[0, 2])}
    return out


def case_accum_rounding():
    """Long accumulation of a constant: s_ii = sum of D copies of c^2.  RZ accumulation shows a
    negative bias growing with the chain length."""
    import torch
    from attacking_federate_learning_b200 import _device as dev, _native as nat
    n, d = 16, 1 << 20
    G = torch.zeros((n, d), device="cuda")
    torch.manual_seed(0)
    G[0] = 1.0 + torch.rand(d, device="cuda")          # positive terms -> monotone sums
    # row 1 zero: d2[0,1] = s_00
    ex = float((G[0].double() ** 2).sum())
    d2 = dev.sqdist_partial(G, nat.GRAM_FORCE_TCGEN05)
    return {"flush": os.environ.get("AFL_GRAM_FLUSH", "default"), "rel_err_s00": (float(d2[0, 1]) - ex) / ex}


def case_select(n, f, seed=0):
    import numpy as np
    import torch
    from attacking_federate_learning_b200 import _device as dev
    from oracle import ref_numpy as orc
    rng = np.random.default_rng(seed)
    d = 64
    G = (0.1 * rng.standard_normal(d) + np.exp(0.25 * rng.standard_normal((n, 1))) * rng.standard_normal((n, d))).astype(np.float32)
    if seed % 2:
        G[:max(f, 2)] = G[0]
    table = orc.pairwise_distances_f32(G)
    dist = torch.from_numpy(table).cuda()
    res = {"n": n, "f": f, "seed": seed}
    idx = int(dev.krum_select(dist, n, f).item())
    res["krum_ok"] = idx == orc.krum_select(table, orc.visit_order(n), n, f)
    if n >= 4 * f + 3:
        t0 = time.time()
        sel_ref, margins = orc.bulyan_select(table, n, f, with_margins=True)
        res["oracle_s"] = time.time() - t0
        sel = dev.bulyan_select(dist, n, f).cpu().tolist()
        res["bulyan_ok"] = sel == sel_ref
        res["min_margin"] = float(min(margins))
        if not res["bulyan_ok"]:
            first = next(i for i, (a, b) in enumerate(zip(sel, sel_ref)) if a != b)
            res["first_diff"] = [first, sel[first], sel_ref[first], margins[first]]
        res["bulyan_ms"] = timeit(lambda: dev.bulyan_select(dist, n, f), iters=5, warm=1)
    res["krum_ms"] = timeit(lambda: dev.krum_select(dist, n, f), iters=5, warm=1)
    return res


def case_tmean(n, d, f, bf16=False, time_it=False, alie=False, seed=0):
    import numpy as np
    import torch
    from attacking_federate_learning_b200 import _device as dev
    from oracle import ref_numpy as orc
    G = hetero(n, d, seed)
    if alie:
        G[:f] = (G[:f].mean(0) - 1.5 * G[:f].std(0, unbiased=False))[None, :]
    if bf16:
        G = G.bfloat16()
    out = dev.trimmed_mean(G, f)
    torch.cuda.synchronize()
    dcheck = min(d, 2048)
    Gh = G[:, :dcheck].float().cpu().numpy()
    ref = orc.trimmed_mean(Gh, n, f)
    ref64 = orc.trimmed_mean_f64(Gh, n, f)
    got = out[:dcheck].cpu().numpy()
    scale = float(np.abs(Gh).mean())
    res = {"n": n, "d": d, "f": f, "bf16": bf16, "alie": alie,
           "max_abs_vs_oracle": float(np.abs(got - ref).max()), "max_abs_vs_f64": float(np.abs(got - ref64).max()),
           "oracle_vs_f64": float(np.abs(ref - ref64).max()), "scale": scale,
           "rel_l2": float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)),
           "exact_frac": float((got == ref).mean()), "nan": int(np.isnan(got).sum())}
    if time_it:
        res["ms"] = timeit(lambda: dev.trimmed_mean(G, f), iters=5, warm=2)
        res["GBps"] = n * d * G.element_size() / res["ms"] / 1e6
    return res


def case_colstats(n, d, f):
    import numpy as np
    import torch
    from attacking_federate_learning_b200 import _device as dev
    from oracle import ref_numpy as orc
    G = hetero(n, d, 5)
    res = {"n": n, "d": d}
    m = dev.mean(G)
    dcheck = min(d, 4096)
    Gh = G[:, :dcheck].cpu().numpy()
    res["mean_bitexact"] = bool((m[:dcheck].cpu().numpy() == orc.no_defense(Gh)).all())
    crafted, mu, sigma = dev.alie(G[:f], 1.5, None, alias_mean=False)
    rc, rmu, rs = orc.alie_attack([Gh[i].copy() for i in range(f)], 1.5)
    res["alie_max_rel"] = float(np.max(np.abs(crafted[:dcheck].cpu().numpy() - rc) / (np.abs(rc) + 1e-3)))
    res["alie_sigma_rel"] = float(np.max(np.abs(sigma[:dcheck].cpu().numpy() - rs) / rs))
    res["mean_ms"] = timeit(lambda: dev.mean(G))
    res["mean_GBps"] = n * d * 4 / res["mean_ms"] / 1e6
    res["alie_ms"] = timeit(lambda: dev.alie(G[:f], 1.5, None))
    res["alie_GBps"] = f * d * 4 / res["alie_ms"] / 1e6
    return res


def case_host_krum(n, d, f):
    import numpy as np
    import torch
    from attacking_federate_learning_b200 import defences as D
    from oracle import ref_numpy as orc
    G = hetero(n, d, 2).cpu()
    Gp = G.pin_memory().numpy()
    t0 = time.time(); row = D.krum(Gp, n, f); t1 = time.time()
    row = D.krum(Gp, n, f); t2 = time.time()
    idx_ref = None
    if n * d <= 4e7:
        idx_ref = orc.krum(Gp, n, f, return_index=True, dtype=np.float64)
    got = int((row.ctypes.data - Gp.ctypes.data) // (Gp.strides[0]))
    return {"n": n, "d": d, "idx": got, "idx_ref": idx_ref, "first_s": t1 - t0, "second_s": t2 - t1,
            "GBps_h2d": n * d * 4 / (t2 - t1) / 1e9}


CASES = {}


def reg(name, fn, *a, env=None, **k):
    CASES[name] = (fn, a, k, env or {})


F = {"AUTO": 0, "SIMT": 1, "TC": 2, "SINGLE": 4, "REWRITE": 8}
reg("simt_small", case_gram, 10, 1000, F["SIMT"], "simt")
reg("tc_tiny", case_gram, 16, 64, F["TC"], "tc")
reg("tc_small", case_gram, 10, 1000, F["TC"], "tc")
reg("tc_n100", case_gram, 100, 4096, F["TC"], "tc")
reg("tc_n128", case_gram, 128, 8192, F["TC"], "tc")
reg("tc_n200", case_gram, 200, 4096, F["TC"], "tc")
reg("tc_n500", case_gram, 500, 16384, F["TC"], "tc")
reg("tc_single", case_gram, 100, 4096, F["TC"] | F["SINGLE"], "single")
reg("tc_rewrite", case_gram, 100, 4096, F["TC"] | F["REWRITE"], "rewrite")
reg("tf32_behaviour", case_tf32_behaviour)
reg("ident_tc", case_gram_identical, 40, 5000, F["TC"])
reg("ident_simt", case_gram_identical, 40, 5000, F["SIMT"])
for fl in ["1", "2", "4", "8", "32", "100000"]:
    reg(f"accum_flush{fl}", case_accum_rounding, env={"AFL_GRAM_FLUSH": fl})
    reg(f"prec_flush{fl}", case_gram, 100, 1 << 20, F["TC"], "prec", True, env={"AFL_GRAM_FLUSH": fl})
reg("prec_single", case_gram, 100, 1 << 20, F["TC"] | F["SINGLE"], "prec_single", True)
reg("prec_rewrite", case_gram, 100, 1 << 20, F["TC"] | F["REWRITE"], "prec_rewrite", True)
reg("prec_simt", case_gram, 100, 1 << 18, F["SIMT"], "prec_simt", True)
reg("c2_gram", case_gram, 100, 11_200_000, F["TC"], "c2", True)
reg("c2_gram_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], "c2_single", True)
reg("c2_gram_flush2", case_gram, 100, 11_200_000, F["TC"], "c2", True, env={"AFL_GRAM_FLUSH": "2"})
reg("c2_gram_flush32", case_gram, 100, 11_200_000, F["TC"], "c2", True, env={"AFL_GRAM_FLUSH": "32"})
reg("c2_cpasync", case_gram, 100, 11_200_000, F["TC"], "c2_cpasync", True, env={"AFL_GRAM_LOADER": "1"})
reg("c2_cpasync_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], "c2_cpasync_single", True, env={"AFL_GRAM_LOADER": "1"})
reg("c2_box104", case_gram, 100, 11_200_000, F["TC"], "c2_box104", True, env={"AFL_GRAM_BOXROWS": "104"})
reg("c2_box104_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], "c2_box104_single", True, env={"AFL_GRAM_BOXROWS": "104"})
reg("ident_cpasync", case_gram_identical, 40, 5000, F["TC"], env={"AFL_GRAM_LOADER": "1"})
reg("n200_cpasync", case_gram, 200, 4096, F["TC"], "tc", env={"AFL_GRAM_LOADER": "1"})
for kc in ["0","1","2","3"]:
    reg(f"c2_kc{kc}", case_gram, 100, 11_200_000, F["TC"], f"c2_kc{kc}", True, env={"AFL_GRAM_KCHUNK_LOG2": kc})
    reg(f"c2_kc{kc}_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], f"c2_kc{kc}_single", True, env={"AFL_GRAM_KCHUNK_LOG2": kc})
for sp in ["37","74","148"]:
    reg(f"c2_sp{sp}_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], f"c2_sp{sp}_single", True, env={"AFL_GRAM_SPLITS": sp})
    reg(f"c2_sp{sp}", case_gram, 100, 11_200_000, F["TC"], f"c2_sp{sp}", True, env={"AFL_GRAM_SPLITS": sp})
for stg in ["4","6","8"]:
    reg(f"c2_st{stg}", case_gram, 100, 11_200_000, F["TC"], f"c2_st{stg}", True, env={"AFL_GRAM_STAGES": stg})
    reg(f"c2_st{stg}_single", case_gram, 100, 11_200_000, F["TC"] | F["SINGLE"], f"c2_st{stg}_single", True, env={"AFL_GRAM_STAGES": stg})
reg("b16_c2", case_gram, 100, 11_200_000, F["TC"] | 32, "b16_c2", True)
reg("tf32_c2", case_gram, 100, 11_200_000, F["TC"] | 16, "tf32_c2", True)
reg("b16_ragged", case_gram, 100, 100_004, F["TC"], "b16_ragged")
reg("b16_n64", case_gram, 64, 65_540, F["TC"], "b16_n64")
reg("b16_n100_200k", case_gram, 100, 200_000, F["TC"] | 32, "b16_n100_200k")
reg("b16_n100_2m", case_gram, 100, 2_000_000, F["TC"] | 32, "b16_n100_2m")
reg("b16_n80_2m", case_gram, 80, 2_000_000, F["TC"] | 32, "b16_n80_2m")
reg("b16_n112", case_gram, 112, 200_000, F["TC"], "b16_n112")
reg("b16_n70", case_gram, 70, 1_000_000, F["TC"], "b16_n70", True)
reg("b16_ident", case_gram_identical, 100, 70_000, F["TC"])
reg("c2_tf32", case_gram, 100, 11_200_000, F["TC"] | 16, "c2_tf32", True)
reg("b16_ident32", case_gram_identical, 100, 70_004, F["TC"] | 32)
reg("b16_n100_200k", case_gram, 100, 200_000, F["TC"] | 32, "b16_n100_200k")
reg("b16_n100_2m", case_gram, 100, 2_000_000, F["TC"] | 32, "b16_n100_2m")
reg("b16_n80_2m", case_gram, 80, 2_000_000, F["TC"] | 32, "b16_n80_2m")
reg("b16_n112", case_gram, 112, 65_540, F["TC"] | 32, "b16_n112")
reg("b16_n97", case_gram, 97, 40_000, F["TC"] | 32, "b16_n97")
for sp in ["37","74"]:
    reg(f"b16_sp{sp}", case_gram, 100, 11_200_000, F["TC"], f"b16_sp{sp}", True, env={"AFL_GRAM_SPLITS": sp})
reg("b16_flush2", case_gram, 100, 11_200_000, F["TC"], "b16_flush2", True, env={"AFL_GRAM_FLUSH": "2"})
reg("b16_flush8", case_gram, 100, 11_200_000, F["TC"], "b16_flush8", True, env={"AFL_GRAM_FLUSH": "8"})
reg("b16_flush16", case_gram, 100, 11_200_000, F["TC"], "b16_flush16", True, env={"AFL_GRAM_FLUSH": "16"})
for pf in ["0","1","2","4"]:
    reg(f"b16_pf{pf}", case_gram, 100, 11_200_000, F["TC"], f"b16_pf{pf}", True, env={"AFL_GRAM_PREFETCH": pf})
for kn in ["1","2","4","3","6"]:
    reg(f"b16_knock{kn}", case_gram, 100, 11_200_000, F["TC"], f"b16_knock{kn}", True, env={"AFL_GRAM_KNOCK": kn})
for fl in ["8","32","512"]:
    reg(f"b16_fl{fl}", case_gram, 100, 11_200_000, F["TC"], f"b16_fl{fl}", True, env={"AFL_GRAM_FLUSH": fl})
for dbg in ["1","2"]:
    reg(f"c2_dbg{dbg}", case_gram, 100, 11_200_000, F["TC"], f"c2_dbg{dbg}", True, env={"AFL_GRAM_DBG": dbg})
for kc in ["1","2","3"]:
    reg(f"b16_kc{kc}", case_gram, 100, 11_200_000, F["TC"] | 32, f"b16_kc{kc}", True, env={"AFL_GRAM_KCHUNK_LOG2": kc})
reg("b16_ragged32", case_gram, 100, 100_004, F["TC"] | 32, "b16_ragged32")
reg("b16_n64_32", case_gram, 64, 65_540, F["TC"] | 32, "b16_n64_32")
reg("n500_gram", case_gram, 500, 1 << 20, F["TC"], "n500", True)
reg("n1000_gram", case_gram, 1000, 1 << 19, F["TC"], "n1000", True)
reg("select_10", case_select, 10, 2, 0)
reg("select_31_ties", case_select, 31, 7, 1)
reg("select_100", case_select, 100, 24, 2)
reg("select_500", case_select, 500, 100, 4)
reg("select_1000_ties", case_select, 1000, 240, 5)
reg("tm_small", case_tmean, 10, 1000, 2)
reg("tm_100", case_tmean, 100, 5000, 24)
reg("tm_300", case_tmean, 300, 5000, 100)
reg("tm_1000", case_tmean, 1000, 1 << 16, 240, False, True)
reg("tm_1000_bf16", case_tmean, 1000, 1 << 16, 240, True, True)
reg("tm_1000_alie", case_tmean, 1000, 1 << 16, 240, False, True, True)
reg("tm_1000_big", case_tmean, 1000, 1 << 20, 240, False, True)
reg("tm_1000_bf16_big", case_tmean, 1000, 1 << 21, 240, True, True)
reg("colstats", case_colstats, 1000, 1 << 20, 240)
reg("host_krum_small", case_host_krum, 10, 79510, 2)
reg("host_krum_c2", case_host_krum, 100, 11_200_000, 24)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "case":
        fn, a, k, env = CASES[sys.argv[2]]
        res = fn(*a, **k)
        print("RESULT " + json.dumps(res))
        return
    only = sys.argv[1:] or list(CASES)
    log = open(os.path.join(OUT, "diag.jsonl"), "a")
    for name in only:
        fn, a, k, env = CASES[name]
        e = dict(os.environ); e.update(env)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "case", name], env=e, capture_output=True, text=True, timeout=int(os.environ.get('DIAG_CASE_TIMEOUT', '150')))
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            rec = {"case": name, "rc": p.returncode, "s": round(time.time() - t0, 1)}
            if line:
                rec["result"] = json.loads(line[-1][7:])
            else:
                rec["stdout"] = p.stdout[-1500:]
                rec["stderr"] = p.stderr[-2500:]
        except subprocess.TimeoutExpired:
            rec = {"case": name, "rc": "timeout"}
        log.write(json.dumps(rec) + "\n"); log.flush()
        print(json.dumps(rec)[:600], flush=True)


if __name__ == "__main__":
    main()
