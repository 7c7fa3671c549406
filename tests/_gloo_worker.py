import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minilp_b200 as mb  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 1000
    rng = np.random.default_rng(3)
    d = rng.standard_normal(n)
    gam = 1.0 + rng.random(n)
    d[123] = d[877] = 5.0  # exact tie across the two blocks
    gam[123] = gam[877] = 1.0
    score = d * d / gam
    score[d > -1e-8] = -np.inf  # at_min eligibility (solver.rs:705)
    d[123] = d[877] = -5.0
    score[123] = score[877] = 25.0
    b, e = mb.shard_range(n, world, rank)
    loc = int(np.argmax(score[b:e])) + b
    mine = torch.tensor([score[loc], float(loc), float(loc)], dtype=torch.float64)
    allc = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allc, mine)
    sc = [float(t[0]) for t in allc]
    pos = [int(t[1]) for t in allc]
    var = [int(t[2]) for t in allc]
    w = mb.reduce_candidates(sc, pos, var)
    assert var[w] == 123, (w, var)
    assert var[w] == int(np.argmax(score))
    print("GLOO_OK", rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
