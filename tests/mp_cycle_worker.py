"""Worker of tests/test_multi_gpu.py: K closed-loop cycles of one world on WORLD_SIZE ranks (one GPU each), records
exchanged by the commit kernel's peer-to-peer stores; rank 0 writes the ring contents after every cycle to an .npz.
    torchrun --nproc-per-node W tests/mp_cycle_worker.py <cfg> <seed> <cycles> <graph 0|1> <out.npz>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from neptune_b200 import capi, config
    from neptune_b200.cycle import ReplanCycle, shard_agents
    from neptune_b200.scenes import make_scene, slice_scene
    cfg, seed, cycles, graph, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    par = config(cfg)
    mine = shard_agents(par.num_of_agents, world, rank)
    gen = capi.Solver(par, device=local)
    sc0 = make_scene(par, seed, sync=False)
    if par.num_of_static_obst:
        gen.set_static(sc0.batch.st_ptr, sc0.batch.st_xy, sc0.strep)
    # every rank generates the WHOLE world (same seed, same bytes) and plans for its rows of it
    full = make_scene(par, seed, sync=False, ent_backend=capi.DeviceEntBackend(gen))
    sc = slice_scene(full, mine)
    b = sc.batch
    cyc = ReplanCycle(par, mine, dev, static=(b.st_ptr, b.st_xy, sc.strep), world=world, rank=rank,
                      planned=np.ones(par.num_of_agents, np.uint8))
    cyc.connect()
    cyc.seed_records(cyc.records_of(sc))
    hin, hout = cyc.host_inputs(sc), cyc.host_outputs()
    cyc.upload(hin)
    rings, coeffs = [], []
    for k in range(cycles):
        cyc.align()      # device-side rank barrier (a no-op on one rank): every cycle starts together, results unchanged
        if graph and k == 1:
            cyc.capture()
        else:
            cyc.step()
        cyc.download(hout)
        cyc.stream.synchronize()
        rings.append(cyc.records("new").copy())
        coeffs.append((mine.copy(), hout["coeff_out"].copy(), hout["status"].copy(), hout["collide"].copy(), hout["entangled"].copy()))
    cyc.check_errors()
    if world > 1:
        dist.barrier()
    np.savez(out + f".rank{rank}.npz", rings=np.stack(rings), agents=mine, coeff=np.stack([c[1] for c in coeffs]),
             status=np.stack([c[2] for c in coeffs]), collide=np.stack([c[3] for c in coeffs]), entangled=np.stack([c[4] for c in coeffs]))
    cyc.close()
    gen.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
