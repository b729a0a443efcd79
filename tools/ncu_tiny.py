"""Target of the ncu launch list of the molecule-size path: three evaluations of the N2 fixture (N = 10) through the single-kernel
path, then three through the 11-launch pipeline (MAMBO_NO_TINY=1).  profiles/r02_launches_tiny.csv"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mambo_b200 as mb
import csa_oracle as o
fx = np.load(os.path.join(ROOT, "tests", "golden", "molecules.npz"))
T = fx["n2_tbt"]
xs = [o.random_x(10, s) for s in range(3)]
os.environ["MAMBO_NO_GRAPH"] = "1"          # plain launches: one ncu row per kernel
for env in (None, "1"):
    if env: os.environ["MAMBO_NO_TINY"] = env
    with mb.CSAEngine(T) as e:
        for x in xs:
            c, g = e.cost_grad(x)
        print("tiny" if not env else "pipeline", e.info()["launches_per_eval"], c)
