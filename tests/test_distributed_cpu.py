"""World-size-2 `gloo` test (CPU) of the N > 1 host logic of the path: the image partition across ranks
(`partition_indices`, the reference's `partition_dataset(..., shuffle=True, seed=0, even_divisible=True)[rank]`,
src/data/get_train_and_val_dataloader.py:21-31) and the single gather of the score rows (`gather_scores`, reference
src/trainers/reconstruct.py:238-242). No GPU, no compute kernels."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_IMAGES = 11  # odd on purpose: even_divisible wrap-pads one duplicate
N_T = 3


def _score_of(image_id: int, t: int):
    return float(image_id) + 0.001 * t, float(image_id) * 2.0 + 0.5 * t


def _worker(rank: int, world: int, port: int, out_path: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        from ddpm_ood_b200.data import partition_indices
        from ddpm_ood_b200.trainers.reconstruct import gather_scores

        mine = partition_indices(N_IMAGES, world, rank)
        names, rows = [], []
        for t in (10, 330, 650)[:N_T]:  # row order of the trainer: t-start outer, image inner
            for i in mine:
                pd_, mse = _score_of(int(i), t)
                names.append(f"img{int(i):03d}")
                rows.append([float(t), pd_, mse])
        scores = torch.tensor(rows, dtype=torch.float64)
        all_scores, all_names = gather_scores(scores, names, torch.device("cpu"))
        if rank == 0:
            torch.save({"scores": all_scores, "names": all_names, "mine": [int(i) for i in mine]}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_partition_and_score_gather_world_size_2(tmp_path):
    world = 2
    port = 29600 + (os.getpid() % 200)
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    scores, names = got["scores"], got["names"]
    per_rank = (N_IMAGES + world - 1) // world
    assert scores.shape == (world * per_rank * N_T, 3)
    assert len(names) == scores.shape[0]
    # every image appears, the padded duplicate is an exact copy, and every row carries its own image's scores
    seen = {}
    for r, name in enumerate(names):
        image_id = int(name[3:])
        t = int(scores[r, 0])
        want = _score_of(image_id, t)
        assert (float(scores[r, 1]), float(scores[r, 2])) == want
        seen.setdefault((image_id, t), 0)
        seen[(image_id, t)] += 1
    assert {k[0] for k in seen} == set(range(N_IMAGES))
    assert sum(v - 1 for v in seen.values()) == (world * per_rank - N_IMAGES) * N_T  # duplicates from even padding only


def test_partitions_are_disjoint_up_to_padding():
    from ddpm_ood_b200.data import partition_indices

    for n, world in [(11, 2), (256, 8), (7, 4), (8, 8)]:
        parts = [list(map(int, partition_indices(n, world, r))) for r in range(world)]
        assert len({len(p) for p in parts}) == 1  # even_divisible
        flat = [i for p in parts for i in p]
        assert set(flat) == set(range(n))
        assert len(flat) - len(set(flat)) == len(flat) - n
