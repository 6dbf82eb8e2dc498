"""Counts the Blackwell-specific SASS mnemonics per kernel of liblrpcap.so (cuobjdump -sass): UTCHMMA (tcgen05.mma kind::f16),
UTCQMMA (kind::f8f6f4), 2CTA (instructions with the cta_group::2 suffix), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit), UTMAPF (TMA prefetch), SYNCS (mbarrier),
USETMAXREG (setmaxnreg), REDUX (warp reductions).  Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lrp_imagecaptioning_b200", "liblrpcap.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
MN = ["UTCHMMA", "UTCQMMA", "2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "USETMAXREG", "REDUX", "HMMA", "DFMA", "FFMA"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = cur.replace("(anonymous namespace)::", "").replace("lrpcap::", "")
        mm = re.match(r"^(?:void )?([\w:]+(?:<[^(]*>)?)", cur)
        cur = mm.group(1) if mm else cur
        order.append(cur)
        continue
    if cur is None:
        continue
    for mn in MN:
        if (mn == "2CTA" and ".2CTA" in line) or (mn != "2CTA" and re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line)):
            counts[cur][mn] += 1
print("library: %s" % os.path.basename(lib))
print("%-88s %s" % ("kernel", " ".join("%9s" % m for m in MN)))
tot = collections.Counter()
for k in sorted(set(order)):
    c = counts[k]
    if not any(c[m] for m in MN[:11]):
        continue
    print("%-88s %s" % (k[:88], " ".join("%9d" % c[m] for m in MN)))
    tot.update(c)
print("%-88s %s" % ("TOTAL (kernels listed above)", " ".join("%9d" % tot[m] for m in MN)))
