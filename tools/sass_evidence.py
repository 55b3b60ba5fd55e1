"""profiles/r02_sass_tensor_kernels.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell paths.
    python tools/sass_evidence.py > profiles/r02_sass_tensor_kernels.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "attacking_federate_learning_b200/lib/libafl_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "HMMA", "HGMMA", "LDGSTS",
        "REDUX", "PRMT", "VIMNMX")
cur, stats, samples = None, collections.OrderedDict(), collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); samples[cur] = []; continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if cur is None or not m:
        continue
    ins = m.group(1).strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
    for key in KEYS:
        if op.startswith(key):
            stats[cur][key] += 1
            if key in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR") and len([s for s in samples[cur] if s.startswith(key)]) < 2:
                samples[cur].append(key + ": " + ins)
print("# cuobjdump -sass %s (sm_100a), per kernel: counts of the SASS mnemonics that prove the Blackwell paths" % lib)
print("# (B200_PROFILING.md): tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, tcgen05.commit -> UTCBAR, TMEM alloc ->")
print("# UTCATOMSWS; warp reductions -> REDUX; the trimmed-mean kernel's bf16 unpack -> PRMT, its integer-key sort -> VIMNMX.")
print("# No HMMA (legacy mma.sync) and no HGMMA (Hopper) anywhere.\n")
for fn, c in stats.items():
    if not c:
        continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(dem + "\n    " + ", ".join(f"{k}: {v}" for k, v in sorted(c.items())))
    for s in samples[fn]:
        print("      e.g. " + s)
