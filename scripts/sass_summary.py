"""tcgen05 / TMEM / TMA / tensor-path instruction counts per kernel from the built objects (cuobjdump -sass):
    python scripts/sass_summary.py r02b  ->  profiles/sass_summary_<tag>.txt   (run in the build container, no GPU needed)"""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02b"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eeg-gnn-ssl_b200", "lib")
OBJS = ["rnn_fwd.o", "rnn_bwd.o", "bulk_dp.o", "dw_mm16.o", "fft.o", "head.o", "optim.o", "graph.o", "seq_fwd_tc.o", "seq_bwd_tc.o", "dw_mm.o", "dw_tc.o"]
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                               ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"),
                               ("LDSM", r"\bLDSM"), ("FFMA2", r"\bFFMA2"), ("FFMA", r"\bFFMA\b"), ("DFMA", r"\bDFMA")])
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
out = [f"# SASS evidence ({tag}, commit {commit}): tcgen05 / TMEM / TMA / tensor-path instruction counts per kernel",
       "# command: cuobjdump -sass eeg-gnn-ssl_b200/lib/<file>.o, instructions counted per `Function :` block",
       "# UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = tensor-map TMA load / store,",
       "# UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA = mma.sync (warp-level tensor path of the diffusion), LDSM = ldmatrix,",
       "# FFMA2 = packed fp32 FMA (sm_100)", ""]
for o in OBJS:
    path = os.path.join(LIB, o)
    if not os.path.exists(path):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out.append(f"## {o}")
    out.append("kernel".ljust(44) + "".join(k.rjust(9) for k in PAT))
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        short = re.sub(r"^_ZN5dcgru\d*", "", name)[:42]
        counts = [len(re.findall(p, blk)) for p in PAT.values()]
        if sum(counts[:6]) + counts[7] + counts[9] == 0 and not any(k in name for k in ("fft", "head", "clip_adam", "corr")):
            continue
        out.append(short.ljust(44) + "".join(str(c).rjust(9) for c in counts))
    out.append("")
open(os.path.join(ROOT, "profiles", f"sass_summary_{tag}.txt"), "w").write("\n".join(out))
print("\n".join(out))
