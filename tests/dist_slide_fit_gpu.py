"""Run by hand on a 2-GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_slide_fit_gpu.py
Slide-level fit with the target tiles sharded over two GPUs (NCCL all-reduces of the statistics) against the same fit of
all tiles on one GPU and against the oracle on the concatenated image."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import stainlib_b200 as sb
    from stainlib_b200.distributed import shard_range
    from stainlib_b200.synth import synth_tile
    tiles = np.stack([synth_tile(500 + i, 256, 256) for i in range(7)])
    lo, hi = shard_range(len(tiles), rank, world)
    mine = torch.from_numpy(tiles[lo:hi]).cuda()
    solo = dist.new_group([rank]) if world > 1 else None          # a one-rank group: the unsharded fit inside this process
    groups = [dist.new_group([r]) for r in range(world)]          # (new_group is collective: every rank creates all of them)
    ok = True
    for method in ("macenko", "vahadane"):
        a = sb.ExtractiveStainNormalizer(method)
        a.fit(mine, slide=True)                                    # sharded: default group, NCCL
        b = sb.ExtractiveStainNormalizer(method)
        b.fit(torch.from_numpy(tiles).cuda(), slide=True, group=groups[rank])
        dM = float(np.abs(a.stain_matrix_target - b.stain_matrix_target).max())
        dC = float(np.abs(a.maxC_target - b.maxC_target).max())
        # fixed-point sums and integer histograms add exactly: sharded == unsharded to the last bit, for both methods
        gathered = [None] * world
        dist.all_gather_object(gathered, (a.stain_matrix_target.tolist(), a.maxC_target.tolist()))
        same = all(g == gathered[0] for g in gathered)
        if rank == 0:
            print(f"{method}: sharded vs unsharded |dM| {dM:.2e} |dmaxC| {dC:.2e}; identical on all ranks: {same}")
        ok = ok and same and dM == 0.0 and dC == 0.0
    if rank == 0:
        from oracle import stain_oracle as so
        o = so.ExtractiveStainNormalizer("macenko")
        o.fit(np.concatenate(list(tiles), axis=0))
        a = sb.ExtractiveStainNormalizer("macenko")
    a = sb.ExtractiveStainNormalizer("macenko")
    a.fit(mine, slide=True)
    if rank == 0:
        print("macenko sharded vs oracle |dM|", float(np.abs(a.stain_matrix_target - o.stain_matrix_target).max()),
              "rel dmaxC", float(np.abs(a.maxC_target / o.maxC_target - 1).max()))
        ok = ok and np.abs(a.stain_matrix_target - o.stain_matrix_target).max() < 1e-5
        print("DIST_SLIDE_FIT", "OK" if ok else "FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
