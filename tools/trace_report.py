"""Summarise an AFL_GRAM_TRACE dump (clock64 timestamps per role and k-block)."""
import sys
import numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.int64)
for cta in (0, 1):
    t = rows[rows[:, 0] == cta][:, 2:]
    it = rows[rows[:, 0] == cta][:, 1]
    ok = t[:, 1] > 0
    t = t[ok]; n = len(t)
    if n < 20:
        continue
    t0 = t[0, 0]
    ev = ["prod_wait_start", "tma_issue", "split_start(full)", "split_done", "mma_start", "mma_issued", "epi_start", "epi_done"]
    print(f"CTA {cta}: {n} k-blocks traced; steady-state per k-block = {(t[n-1,1]-t[20,1])/(n-21):.0f} cycles")
    sl = slice(24, min(n, 400))
    print("  tma_issue -> full (TMA latency)      :", np.median(t[sl, 2] - t[sl, 1]))
    print("  full -> split_done (split work)      :", np.median(t[sl, 3] - t[sl, 2]))
    if t.shape[1] > 10:
        print("    split: start->loads issued         :", np.median(t[sl, 8] - t[sl, 2]))
        print("    split: loads->stores issued        :", np.median(t[sl, 9] - t[sl, 8]))
        print("    split: fence.proxy.async           :", np.median(t[sl, 10] - t[sl, 9]))
        print("    split: fence->arrive               :", np.median(t[sl, 3] - t[sl, 10]))
    print("  split_done -> mma_start (MMA queue)  :", np.median(t[sl, 4] - t[sl, 3]))
    print("  mma_start -> mma_issued              :", np.median(t[sl, 5] - t[sl, 4]))
    if t.shape[1] > 13:
        print("    mma loop: issued(i-1) -> loop top(i) :", np.median(t[sl, 11][1:] - t[sl, 5][:-1]))
        print("    mma loop: wait split_bar            :", np.median(t[sl, 12] - t[sl, 11]))
        print("    mma loop: tcgen05.fence::after      :", np.median(t[sl, 13] - t[sl, 12]))
        print("    mma loop: fence -> elected          :", np.median(t[sl, 4] - t[sl, 13]))
    print("  producer wait for empty              :", np.median(t[sl, 1] - t[sl, 0]))
    print("  mma_start[i+1]-mma_start[i]          :", np.median(np.diff(t[sl, 4])))
    print("  tma_issue[i+1]-tma_issue[i]          :", np.median(np.diff(t[sl, 1])))
    # stage ring depth 6: empty wait of it = mma completion of it-6
    g = t[:, 6] > 0
    if g.sum() > 4:
        print("  epilogue drain (start->done)         :", np.median((t[g, 7] - t[g, 6])[2:]))
