"""SASS evidence per kernel of libpmgt_b200.so: counts of the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk, HMMA = mma.sync) and a short
excerpt around the first tensor-core instruction.   python tools/sass_summary.py > profiles/r2_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pmgt_b200", "libpmgt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "GATHER4", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "REDG", "ATOMS", "SHFL", "MUFU"]
kern, lines = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "")
        kern = re.sub(r"\(.*", "", kern)
        lines[kern] = []
    elif kern and "/*" in ln and ";" in ln:
        lines[kern].append(ln)
print("# SASS evidence, libpmgt_b200.so (cuobjdump -sass, sm_100a)\n")
print("| kernel | instructions | " + " | ".join(MN) + " |")
print("|---|---:|" + "---:|" * len(MN))
tot = collections.Counter()
for k, ls in lines.items():
    c = collections.Counter()
    for ln in ls:
        for mn in MN:
            if re.search(r"\b" + mn, ln):
                c[mn] += 1
    tot.update(c)
    print(f"| `{k[:90]}` | {len(ls)} | " + " | ".join(str(c[m]) for m in MN) + " |")
print(f"| **total** | {sum(len(v) for v in lines.values())} | " + " | ".join(str(tot[m]) for m in MN) + " |")
print("\n## Excerpts (first tensor-core / TMA instructions of the hot kernels)\n")
for pat in ("linear_tile_kernel<1, 1, false, 2", "ffn_fwd_kernel", "ffn_bwd_kernel", "dw_tile_kernel<4", "umma_gemm_kernel<false, false, true",
            "attn_mma_bwd_kernel<6, 128>", "sample_contexts_kernel<8, true>", "ln_bwd_stream_kernel<2, false>", "gather_proj_fwd_kernel<true", "gather_proj_dw_kernel<8, true, 16>", "umma_gemm_persist_kernel<false, false, 256>", "attn_reg_bwd_kernel<64, 3>"):
    for k, ls in lines.items():
        if pat in k:
            first = r"GATHER4" if "gather_proj" in pat else r"UTCHMMA|HMMA|UTMALDG|UBLKCP|LDG"
            idx = next((i for i, ln in enumerate(ls) if re.search(first, ln)), 0)
            print(f"### `{k[:100]}`\n```")
            for ln in ls[max(0, idx - 3): idx + 9]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip()))
            print("```\n")
            break
