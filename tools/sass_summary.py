"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (works without a GPU):
UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit,
UCGABAR = cluster barrier, HMMA = legacy mma.sync (only the attention core uses it).

    python tools/sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "perspectivefields_b200", "libpf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[HQIO]MMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UCGABAR_ARV|HMMA|UBLKCP)((?:\.[0-9A-Za-z_]+)*)")
per, tot, fn = collections.OrderedDict(), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    for mm in pat.finditer(line.split("/*")[1] if line.count("/*") > 1 else line):
        key = mm.group(1) + (".2CTA" if "2CTA" in mm.group(2) else "")
        per.setdefault(fn, collections.Counter())[key] += 1
        tot[mm.group(1) + mm.group(2)] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass), {time.strftime('%Y-%m-%dT%H:%MZ', time.gmtime())}")
for (f, c), nm in sorted(zip(per.items(), names), key=lambda x: x[1]):
    nm = re.sub(r"\(.*$", "", nm.replace("pf::", "").replace("(int)", "").replace("void ", ""))
    print(f"{nm}: " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
print("# totals")
for k, v in tot.most_common():
    print(f"{v:6d} {k}")
