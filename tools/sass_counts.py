"""Per-kernel SASS mnemonic counts of libb200vad.so (`cuobjdump -sass`): which kernels use the 5th-generation tensor cores
(UTC*MMA), tensor memory (LDTM / STTM / UTCBAR), TMA (UTMALDG / UTMASTG / UBLKCP), mbarriers (SYNCS) and cluster barriers.

    python tools/sass_counts.py [path/to/lib.so] > profiles/rNN_sass_counts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "universal-voice-activity-detection_b200", "b200vad", "lib", "libb200vad.so")
GROUPS = [("UTC*MMA", r"^UTC[A-Z]*MMA"), ("UTCBAR", r"^UTCBAR"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UTMALDG", r"^UTMALDG"),
          ("UTMASTG", r"^UTMASTG"), ("UTMAPF", r"^UTMAPF"), ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("UCGABAR", r"^UCGABAR"),
          ("HMMA", r"^HMMA"), ("MUFU", r"^MUFU"), ("F*2 (packed fp32)", r"^F(ADD|MUL|FMA)2"), ("LDL/STL", r"^(LDL|STL)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, total, name = collections.OrderedDict(), {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("b200vad::", "").replace("void ", "")
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1).split(".")[0]
            total[name] += 1
            for g, pat in GROUPS:
                if re.match(pat, op):
                    counts[name][g] += 1
    print(f"# SASS mnemonic counts per kernel (`cuobjdump -sass {os.path.relpath(LIB, ROOT)}`, default build)\n")
    print("| kernel | SASS instr | " + " | ".join(g for g, _ in GROUPS) + " |")
    print("|---|---:|" + "---:|" * len(GROUPS))
    for k in counts:
        print(f"| `{k}` | {total[k]} | " + " | ".join(str(counts[k][g]) if counts[k][g] else "" for g, _ in GROUPS) + " |")


if __name__ == "__main__":
    main()
