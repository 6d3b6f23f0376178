import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from f16_jsb_b200 import F16VecEnv
for mode in ("fp32", "fp64"):
    for rings in (1, 2):
        n = 4096
        a = F16VecEnv(n, mode=mode, seed=3, host_obs="window", host_rings=rings)
        b = F16VecEnv(n, mode=mode, seed=3, host_obs="copy")
        a.seed(100); b.seed(100)
        oa, ob = a.reset(), b.reset()
        print(mode, rings, "aliased", a._win.aliased, "reset equal", np.array_equal(oa, ob))
        rng = np.random.default_rng(0)
        for k in range(3):
            act = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(n, 4)).astype(np.float32)
            oa, ra, da, ia = a.step(act)
            ob, rb, db, ib = b.step(act)
            d = oa != ob
            print(" step", k, "diff elements", int(d.sum()), "rows", np.unique(np.nonzero(d)[1]), "cols", np.unique(np.nonzero(d)[2]),
                  "envs", np.unique(np.nonzero(d)[0])[:10], "max abs", float(np.abs(oa - ob).max()), "rew eq", np.array_equal(ra, rb))
            if d.any():
                e, r, c = [x[0] for x in np.nonzero(d)]
                print("  first", e, r, c, oa[e, r, c], ob[e, r, c], oa[e, r, c].view(np.uint32) if hasattr(oa[e,r,c],'view') else '')
        a.close(); b.close()
