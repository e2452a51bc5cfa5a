"""profiles/rNN_sass_evidence.md: mnemonic counts of the kernels of the default path (cuobjdump -sass of the built library).
    python tools/sass_evidence.py profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("k_chain_ring<true, false>", "k_chain_ring<(bool)1, (bool)0>", "k_edges_enum", "k_edges_place", "k_xss_argsort<unsigned short, 8, 8>",
        "k_xss_argsort<int, 8, 8>", "k_np_argsort<unsigned short>", "k_tree_heads", "k_tree_walk_small", "k_tree_walk_staged",
        "k_pvalue_owned<true>", "k_pvalue_owned<(bool)1>", "k_input_stats", "k_item", "k_block_stats")
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "wgdi_b200", "libicl_b200.so")], capture_output=True, text=True).stdout
rows = []
for f in re.split(r"\n\s*Function : ", out)[1:]:
    name = f.split("\n", 1)[0].strip()
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if not any(k in short for k in KEEP):
        continue
    c = collections.Counter(re.findall(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|UTMALDG[.\w]*|UTC\w*MMA[.\w]*|LDS\.128|ATOMS[.\w]*|REDUX[.\w]*|VOTE[.\w]*|SHFL[.\w]*|HMMA[.\w]*)", f))
    rows.append((short, len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", f)), c))
with open(sys.argv[1], "w") as fh:
    fh.write("# SASS of wgdi_b200/libicl_b200.so (sm_100a): `cuobjdump -sass`, kernels of the default path\n\n"
             "`UBLKCP.S.G` + `SYNCS.*` = the 1-D TMA bulk copy (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`)\n"
             "and its mbarrier: the edge ring of the DP kernel.  There is no `UTMALDG` / `UTC*MMA` / `HMMA` anywhere: the path is\n"
             "integer dynamic programming, sorting and graph walks — nothing to contract on tensor cores (BASELINE.json says so too).\n\n"
             "| kernel | SASS instructions | TMA bulk copy, mbarrier | wide shared loads, warp collectives, shared atomics |\n|---|---|---|---|\n")
    tot = collections.Counter()
    for short, n, c in sorted(rows, key=lambda r: -r[1]):
        tma = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if k.startswith(("UBLKCP", "SYNCS", "UTMA", "UTC", "HMMA")))
        oth = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if not k.startswith(("UBLKCP", "SYNCS", "UTMA", "UTC", "HMMA")))
        fh.write(f"| `{short[:80]}` | {n} | {tma or '-'} | {oth or '-'} |\n")
        tot.update(c)
    fh.write(f"\nTotals over these kernels: UBLKCP {sum(v for k, v in tot.items() if k.startswith('UBLKCP'))}, "
             f"SYNCS {sum(v for k, v in tot.items() if k.startswith('SYNCS'))}, UTMALDG/UTC*MMA/HMMA "
             f"{sum(v for k, v in tot.items() if k.startswith(('UTMA', 'UTC', 'HMMA')))}.\n")
print(open(sys.argv[1]).read())
