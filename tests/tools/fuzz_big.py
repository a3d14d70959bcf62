import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cmp_emul import run_case
from squid_b200 import synth
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rnd = random.Random(seed * 7919)
    n = rnd.choice([100000, 200000, 400000])
    d = rnd.choice([0.005, 0.02, 0.05])
    ref = rnd.choice([synth.CHR17_LEN, [30000000, 20000000, 5000000, 16569], synth.GRCH38_LEN])
    ng = rnd.choice([None, 30, 300])
    res = run_case(n, seed, d, ref, "/tmp/fuzzbig_%d" % (seed % 4), verbose=True, n_genes=ng, fusion_support=rnd.choice([10, 20, 100]))
    print("seed", seed, "n", n, "d", d, "nref", len(ref), "genes", ng, res, flush=True)
    bad += not res.startswith("OK")
print("done, bad =", bad)
