#!/usr/bin/env python3
"""SASS mnemonic counts of the built library (cuobjdump -sass xfem-amie_b200/libamie_b200.so), per kernel: the evidence
for what the kernels are made of -- UBLKCP / UBLKPF (1D TMA bulk copies and L2 prefetches), SYNCS (mbarrier),
LDGSTS (cp.async gathers), LDS.64 / LDS.128, DFMA / DADD / DMUL (FP64), and the absence of tensor-core instructions
(UTMALDG, UTCMMA, HMMA): the path is HBM-bound FP64, 18 flop per 76 B block.  Runs without a GPU.
    python tools/sass_extract.py > profiles/r02_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "xfem-amie_b200", "libamie_b200.so")
WANT = ["UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "LDS.64", "LDS.128", "LDS", "DFMA", "DADD", "DMUL", "LDG", "STG", "BAR",
        "ATOMG", "UTMALDG", "UTCMMA", "HMMA"]
SHOW = ("k_spmv_s3_rt", "k_spmv_s2_rt", "k_spmv_s3<", "k_spmv_s2<", "k_spmv_gen", "k_cg_update", "k_cg_dir", "k_smooth", "k_bicg",
        "k_assemble_gather", "k_place_elements", "k_dirichlet", "k_halo_push", "k_halo_wait", "k_finalize_peer", "k_element_fields",
        "k_block_inverse", "k_repack")


def count(body, w):
    if w in ("LDS", "LDG", "STG", "BAR"):
        return len(re.findall(r"\b" + w + r"\b", body))
    return len(re.findall(re.escape(w), body))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    rows = []
    for f, d in zip(funcs, dem):
        c = {w: count(f, w) for w in WANT}
        tot.update(c)
        if any(s in d for s in SHOW):
            rows.append((d, c))
    print(f"# {os.path.relpath(SO, ROOT)}: {len(funcs)} kernels, arch {arch}")
    print("# whole library:", {w: tot[w] for w in WANT})
    print("# tensor-core / tiled-TMA instructions (UTMALDG, UTCMMA, HMMA):", tot["UTMALDG"] + tot["UTCMMA"] + tot["HMMA"])
    for d, c in sorted(rows, key=lambda t: t[0]):
        print(d[:150], {k: v for k, v in c.items() if v})


if __name__ == "__main__":
    sys.exit(main())
