"""torchrun worker: the column-sharded solve over NCCL (one process per GPU) against the oracle.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/_nccl_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minilp_b200 as mb  # noqa: E402
import oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_uid():  # an ncclUniqueId serves ONE ncclCommInitRank round: every engine gets its own
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(mb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    for kind, m, n, seed in ((0, 200, 300, 1), (3, 97, 131, 9), (1, 64, 96, 2)):
        lp = mb.synth_dense(kind, m, n, seed)
        uid = fresh_uid()
        s = mb.Solver.from_dense(lp, device=local, rank=rank, world=world, comm=uid)
        kinds = [None] * world
        dist.all_gather_object(kinds, s.engine.exchange_kind())
        assert len(set(kinds)) == 1, kinds  # every rank took the same exchange path
        assert s.run()
        ref = oracle.DenseSolver(lp.direction, lp.a, lp.obj, lp.mins, lp.maxs, lp.ops, lp.rhs)  # the reference's own tie rule
        assert ref.continue_solve()
        tg, tr = s.trace(), ref.trace()
        assert ref.near_tie_pivots == 0 and s.tie_stats()["tied_pivots"] == 0
        assert tg.shape == tr.shape and np.array_equal(tg[:, :5], tr[:, :5]), "basis sequence differs from the oracle"
        assert abs(s.cur_obj_val - ref.cur_obj_val) <= 1e-8 * max(1.0, abs(ref.cur_obj_val))
        objs = [None] * world
        dist.all_gather_object(objs, s.cur_obj_val)
        assert all(o == objs[0] for o in objs), objs
        print(f"NCCL_OK rank {rank}/{world} kind {kind} via {kinds[0]}: {s.pivots_done} pivots obj {s.cur_obj_val:.12g}", flush=True)
        s.close()
    # sparse storage, column-sharded (north_star: Netlib-shaped LPs at 1/2/4/8 GPUs), through MPS text
    from minilp_b200 import mps, synth
    for family, args in (("netlib_like", (400, 400, 6.0, 2)), ("sparse_pos", (300, 500, 6.0, 1))):
        text, d = getattr(synth, family)(*args)
        p = mps.MpsFile.parse(text, d).problem
        rp, ci, va, ops, rhs = p.to_csr()
        uid = fresh_uid()
        s = mb.Solver(len(ops), len(p.obj_coeffs), device=local, rank=rank, world=world, comm=uid, csr=(rp, ci, va))
        s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
        assert s.run()
        ref = oracle.MpsFile.parse(text, d).problem.solve()
        tg, tr = s.trace(), ref.trace()
        assert ref.near_tie_pivots == 0 and s.tie_stats()["tied_pivots"] == 0
        assert tg.shape == tr.shape and np.array_equal(tg[:, :5], tr[:, :5]), "sparse sharded: basis sequence differs from the oracle"
        assert abs(s.cur_obj_val - ref.cur_obj_val) <= 1e-8 * max(1.0, abs(ref.cur_obj_val))
        print(f"NCCL_OK rank {rank}/{world} sparse {family} via {s.engine.exchange_kind()}: {s.pivots_done} pivots obj {s.cur_obj_val:.12g}", flush=True)
        s.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
