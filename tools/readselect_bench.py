"""Read selection at chromosome scale: the reference's readselection() (oracle/_ref/pyref, build container only) timed
beside gg_readselection on the same ReadSet, and the two results compared.
usage: python tools/readselect_bench.py <variants> <reads> <variants-per-read> <max_cov>"""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pyref"))
from giggles.core import ReadSet, Read, readselection as ref_sel
from giggles_b200.readselect import readselection as my_sel, flatten
from giggles_b200 import capi
rnd = random.Random(5)
npos = int(sys.argv[1]); nreads = int(sys.argv[2]); L = int(sys.argv[3]); mc = int(sys.argv[4])
pos = sorted(rnd.sample(range(10, 200 * npos), npos))
rs = ReadSet()
for k in range(nreads):
    a = rnd.randrange(0, npos - 2); n = max(2, int(rnd.gauss(L, L * 0.2)))
    span = pos[a:a + n]
    if len(span) < 2: span = pos[-2:]
    r = Read("r%d" % k, 60, 0, 0)
    for i, p in enumerate(span):
        if i in (0, len(span) - 1) or rnd.random() >= 0.02:
            r.add_variant(p, 0, [0.0, -5.0], rnd.choice([20, 30, 40]))
    rs.add(r)
rs.sort()
t = time.time(); a = ref_sel(rs, mc); t_ref = time.time() - t
t = time.time(); fl = flatten(rs); t_fl = time.time() - t
t = time.time(); b = set(capi.readselection_flat(*fl, mc).tolist()); t_my = time.time() - t
print("reads", len(rs), "variants", npos, "selected", len(a), "same", a == b, "ref s", round(t_ref, 3), "flatten s", round(t_fl, 3), "C++ s", round(t_my, 4))
