"""SASS opcode histogram per kernel of libswne_nmf.so (cuobjdump -sass; runs without a GPU):
    python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
Shows which kernels issue tcgen05 MMAs (UTCHMMA = kind::f16, UTCMMA/UTCQMMA for tf32 / other kinds), TMA loads (UTMALDG),
bulk copies (UBLKCP), TMEM loads / stores (LDTM / STTM), and how much local-memory traffic (LDL / STL) the compiler left."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "swne_b200", "libswne_nmf.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "SYNCS", "ELECT", "REDG", "RED", "ATOMG", "ATOMS", "FFMA", "FFMA2", "HFMA2",
       "DFMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "LDCU", "SHFL", "BAR", "USETMAXREG"]
funcs = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0] or n
    except OSError:
        return n
want = sys.argv[1:] or ["tc_pass_kernel", "tc_wside_kernel", "tc_wfinish_kernel", "tcx_cells_kernel", "tcx_genes_kernel", "tcx_solve_kernel",
                        "tc_split_a_kernel", "csc_to_dense", "kl_", "scd_", "gram_kernel", "ica_step", "seed_"]
print("SASS opcode counts per kernel (static instruction counts, sm_100a), libswne_nmf.so")
print("%-64s %6s  %s" % ("kernel", "instrs", "  ".join("%s" % k for k in KEY)))
for name, c in funcs.items():
    d = demangle(name)
    if not any(w in d for w in want):
        continue
    tot = sum(c.values())
    cells = []
    for k in KEY:
        v = sum(n for op, n in c.items() if op == k or (k == "RED" and op == "RED"))
        cells.append("%s=%d" % (k, v) if v else "")
    print("%-64s %6d  %s" % (d[-64:], tot, " ".join(x for x in cells if x)))
