"""Mnemonic counts per kernel from cuobjdump -sass (committed as profiles/<tag>_sass_evidence.txt)."""
import collections, re, subprocess, sys
lib = "wgpu-3dgs-viewer_b200/lib/libsplat_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UBLKCP", "SYNCS", "VOTE", "R2P", "MATCH", "ATOMS", "MUFU.EX2", "LDGSTS", "LDG", "STG", "LDS", "STS", "FFMA2", "FMUL2", "FADD2", "FFMA", "HMMA", "UTC", "LDTM", "UTMALDG"]
fn, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0].replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        for k in keys:  # first match wins, so LDGSTS / FFMA2 are not also counted as LDG / FFMA
            if op.startswith(k):
                counts[fn][k] += 1
                break
print("# cuobjdump -sass", lib)
print("# UBLKCP = cp.async.bulk (TMA 1-D bulk copy); UTMALDG = cp.async.bulk.tensor (here: tile::gather4); SYNCS = mbarrier;")
print("# LDGSTS = cp.async (global -> shared without a register); FFMA2/FMUL2/FADD2 = packed FP32 (two IEEE f32 ops per instruction); R2P = digit bits -> predicates;")
print("# VOTE = warp ballots; ATOMS = shared atomics.  HMMA / UTC*MMA / LDTM (tensor cores, TMEM) are absent by design: no stage is a dense contraction.")
want = sys.argv[1:] or ["preprocess_kernel<0, 0>", "preprocess_kernel<1, 1>", "preprocess_kernel<2, 1>", "onesweep_kernel<8>", "onesweep_kernel<5>", "onesweep3_kernel<8>", "sort_plan_kernel", "select_brush_kernel",
                        "raster_gather4_kernel<0, 0, false, false, false>", "raster_gather4_kernel<0, 0, false, false, true>", "raster_gather4_kernel<1, 3, false, false, false>",
                        "histogram_kernel", "raster_kernel<0, 0, false, false>", "raster_kernel<1, 0, false, false>", "raster_kernel<0, 3, true, false>",
                        "dup_scan_kernel", "dup_emit_kernel", "gather_kernel", "select_rect_kernel", "sort_init_kernel", "sort_finish_kernel"]
for f in sorted(total):
    if any(w in f for w in want):
        print(f"{f:70s} total={total[f]:5d} " + " ".join(f"{k}={v}" for k, v in counts[f].items()))
