"""Per-kernel SASS instruction counts of the built library: python profiles/tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections, os, re, subprocess, sys

R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(R, "tacotron2-vae_b200", "libt2v_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "STTM", "UBLKCP", "SYNCS", "STAS", "FFMA", "HFMA2", "MUFU", "RED", "ATOM"]
counts, cur, i = collections.OrderedDict(), None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[i].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        cur = re.sub(r"\((bool|int)\)", "", cur)          # <(bool)0, (int)1> -> <0, 1>
        cur = re.sub(r"\(.*$", "", cur)
        i += 1
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op == o or (o in ("RED", "ATOM") and op in (o, o + "G", o + "S")) or (o == "UTCHMMA" and op.startswith("UTC") and "MMA" in op):
                counts[cur][o] += 1
print("# SASS instruction counts per kernel of libt2v_b200.so (round 2)\n")
print("`cuobjdump -sass tacotron2-vae_b200/libt2v_b200.so`, counted per `Function :` block (profiles/tools/sass_summary.py).  UTCHMMA = tcgen05.mma\n"
      "(kind::tf32 / kind::f16), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = 1-D bulk\n"
      "copy, SYNCS = mbarrier ops, STAS = st.async (DSMEM stores with mbarrier completion), UTMASTG / STTM = TMA store / tcgen05.st (not used:\n"
      "the epilogues store from registers).\n")
print("| kernel | " + " | ".join(OPS) + " |")
print("|---|" + "---:|" * len(OPS))
for k, c in sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["UTMALDG"], kv[0])):
    print("| `%s` | " % k + " | ".join(str(c[o]) for o in OPS) + " |")
