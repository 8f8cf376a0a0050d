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


# ------------------------------------------------------------------------------------------------ t-start sharding
def _t_worker(rank: int, world: int, port: int, out_path: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        from ddpm_ood_b200.reconstruction import partition_t_starts
        from ddpm_ood_b200.synthetic import chain_lengths
        from ddpm_ood_b200.trainers.reconstruct import gather_t_sharded

        lens = chain_lengths(100, 4)  # BASELINE config 2 grid: 25 t-starts
        parts = partition_t_starts(lens, world)
        owner = torch.empty(len(lens), dtype=torch.long)
        for r, idxs in enumerate(parts):
            owner[idxs] = r
        n_img = 5
        full = torch.full((len(lens), n_img, 2), float("nan"), dtype=torch.float32)
        for i in parts[rank]:  # what score_batch(t_indices=parts[rank]) fills in
            for b in range(n_img):
                full[i, b, 0], full[i, b, 1] = _score_of(b, 10 + 40 * i)
        merged = gather_t_sharded(full, owner, torch.device("cpu"))
        if rank == 0:
            torch.save({"merged": merged, "parts": parts, "lens": lens}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_t_start_sharding_world_size_2(tmp_path):
    """Union of the ranks' t-start shares == the single-rank table, bit for bit (every (t, image) cell comes from
    exactly one rank; nothing is averaged or re-computed)."""
    world = 2
    port = 29850 + (os.getpid() % 100)
    out = str(tmp_path / "tshard.pt")
    mp.spawn(_t_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    merged, parts, lens = got["merged"], got["parts"], got["lens"]
    assert sorted(i for p in parts for i in p) == list(range(len(lens)))  # a partition: disjoint, complete
    want = torch.empty_like(merged)
    for i in range(len(lens)):
        for b in range(merged.shape[1]):
            want[i, b, 0], want[i, b, 1] = _score_of(b, 10 + 40 * i)
    assert torch.equal(merged, want)


def test_partition_t_starts_is_balanced():
    from ddpm_ood_b200.reconstruction import partition_t_starts
    from ddpm_ood_b200.synthetic import chain_lengths

    for steps, skip, world in [(100, 4, 8), (100, 1, 8), (100, 16, 2), (1000, 4, 8), (100, 4, 1), (100, 64, 8)]:
        lens = chain_lengths(steps, skip)
        parts = partition_t_starts(lens, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lens)))
        loads = [sum(lens[i] for i in p) for p in parts]
        ideal = sum(lens) / world
        # longest-first greedy: no rank exceeds the ideal share by more than one (longest) chain
        assert max(loads) <= ideal + max(lens)
        if len(lens) >= 3 * world:
            assert max(loads) <= 1.10 * ideal, (steps, skip, world, loads)
