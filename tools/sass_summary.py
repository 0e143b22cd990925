"""Per-kernel SASS mnemonic counts of libeno.so (cuobjdump -sass) -> profiles/: which kernels issue tcgen05 MMAs
(UTC*MMA), TMEM loads (LDTM), TMA tensor loads (UTMALDG), bulk copies (UBLKCP), cluster barriers (UCGABAR), and how
many FFMA / DFMA they hold.  Usage: python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "easy-neural-ode_b200", "libeno.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UCGABAR", "SYNCS", "FFMA", "DFMA",
        "HMMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, excerpts = None, collections.OrderedDict(), {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)[:150]
            counts[kern] = collections.Counter()
            excerpts[kern] = []
            continue
        if kern is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k] += 1
                break
        if op.startswith(("UTC", "LDTM", "UTMALDG", "UBLKCP")) and len(excerpts[kern]) < 10:
            excerpts[kern].append(re.sub(r"\s+", " ", line.strip())[:170])
    with open(out, "w") as f:
        f.write("SASS mnemonic counts per kernel, libeno.so built from HEAD (cuobjdump -sass, sm_100a)\n")
        f.write("kernel | " + " ".join(KEYS) + "\n")
        tot = collections.Counter()
        for k, c in counts.items():
            tot.update(c)
            if sum(c.values()) == 0:
                continue
            f.write(f"{k}\n    " + " ".join(f"{key}={c[key]}" for key in KEYS if c[key]) + "\n")
        f.write("\nTOTAL " + " ".join(f"{key}={tot[key]}" for key in KEYS) + "\n")
        f.write("\nExcerpts (first tensor-core / TMEM / TMA / bulk-copy instructions of the kernels that have them)\n")
        for k, ex in excerpts.items():
            if ex:
                f.write(f"\n{k}\n")
                for e in ex:
                    f.write("    " + e + "\n")
    print(out, dict(tot))


if __name__ == "__main__":
    main()
