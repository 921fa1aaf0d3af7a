#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library, for profiles/:

    python tools/sass_summary.py s2anet_b200/csrc/libs2a_b200.so profiles/r2_sass_summary.md
"""
import collections
import re
import subprocess
import sys

KEEP = ("conv_tc_kernel", "box_iou_rotated_kernel", "mc_mask_kernel", "nms_mask_kernel", "assign_labels_kernel", "mc_emit_kernel",
        "select_decode_kernel", "pack_weight_kernel", "wgrad_tc_kernel", "conv_tf32x3_kernel", "transpose_planes_kernel")
BLACKWELL = re.compile(r"^(UTCHMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCATOMSWS|USETMAXREG|UCGABAR|SYNCS|ELECT|FENCE\.VIEW\.ASYNC)")


def main():
    so, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(demangle(m.group(1)), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            cur[ins.split()[0]] += 1
    with open(out, "w") as f:
        f.write("# SASS summary of the shipped kernels (round 2)\n\n`cuobjdump -sass %s` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a "
                "-lineinfo -O3`), mnemonic counts per kernel.\nBlackwell-only instructions: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), "
                "`UTMALDG` / `UTMASTG` = TMA tensor load / store, `LDTM` / `STTM` = tcgen05.ld / tcgen05.st, `UTCBAR` = tcgen05.commit, "
                "`UTCATOMSWS` = tcgen05.alloc / dealloc, `USETMAXREG` = setmaxnreg, `SYNCS` = mbarrier, `UCGABAR` = cluster barrier, "
                "`ELECT` = elect.sync.  `LDL` / `STL` = local memory (the general 24-point clipper only; see DESIGN.md 5.3).\n\n" % so)
        total = collections.Counter()
        for name, c in kernels.items():
            for k, v in c.items():
                if BLACKWELL.match(k):
                    total[k.split(".")[0] + ("." + k.split(".")[1] if k.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM")) and "." in k else "")] += v
        f.write("## whole library, Blackwell-specific mnemonics\n\n```\n")
        for k, v in sorted(total.items(), key=lambda x: -x[1]):
            f.write("%6d  %s\n" % (v, k))
        f.write("```\n\n")
        for name, c in kernels.items():
            if not any(k in name for k in KEEP):
                continue
            n = sum(c.values())
            f.write("## `%s`\n\n%d instructions; Blackwell-specific / memory / math mix:\n\n```\n" % (name[:140], n))
            for k, v in sorted(c.items(), key=lambda x: -x[1]):
                if BLACKWELL.match(k) or re.match(r"^(LDS|STS|LDG|STG|LDL|STL|RED|ATOM|HFMA2|HMUL2|F2FP|FFMA|FMUL|FADD|MUFU|DFMA|VOTE|SHFL|BAR)", k):
                    f.write("%6d  %s\n" % (v, k))
            f.write("```\n\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
