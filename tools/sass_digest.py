"""Static SASS evidence: per kernel of libvqacl_b200.so, how many tcgen05 / TMEM / TMA / mbarrier / mma.sync instructions it
holds (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
HMMA.16816 = mma.sync m16n8k16).   python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "vqacl_b200", "libvqacl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
cnt = collections.defaultdict(collections.Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"\b(UTCHMMA|UTCBAR|LDTM|UTMALDG|HMMA\.16816|SYNCS|UTCATOMSWS)", line)
    if m and fn:
        cnt[fn][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
for mangled, d in sorted(zip(cnt, dem), key=lambda x: x[1]):
    short = re.sub(r"\((CUtensorMap_st|vq::|float|__nv|long|int|unsigned|void\*).*", "", d).replace("void ", "")
    print(f"{short:70s} " + "  ".join(f"{k} {v}" for k, v in sorted(cnt[mangled].items())))
