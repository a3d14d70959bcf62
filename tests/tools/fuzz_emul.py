"""Dev tool: fuzz the CPU-stepped device rules against the reference-built oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cmp_emul import run_case
from squid_b200 import synth
import random
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    rnd = random.Random(seed)
    n = rnd.choice([300, 1000, 3000, 10000, 30000])
    d = rnd.choice([0.005, 0.02, 0.05, 0.2])
    ref = rnd.choice([synth.CHR17_LEN, [3000000, 2000000, 500000, 16569], synth.GRCH38_LEN])
    ng = rnd.choice([None, 5, 20, 100])
    res = run_case(n, seed, d, ref, "/tmp/fuzz_%d" % (seed % 8), verbose=True, n_genes=ng, fusion_support=rnd.choice([6, 20, 60]))
    if not res.startswith("OK"):
        bad += 1
        print("seed", seed, "n", n, "d", d, "nref", len(ref), "genes", ng, res)
    elif "-v" in sys.argv:
        print("seed", seed, res)
print("done, bad =", bad, "of", hi - lo)
